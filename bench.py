#!/usr/bin/env python
"""bench.py -- AD residual+Jacobian assembly throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n NX] [--config 2|5] [--check]

A step = one pass of the hot path over the whole mesh: ONE residual + ONE
Jacobian assembly at the same state (one Newton iteration's worth).
Workload at N=1: config 2 -- 1000x1000 quads on [0,1]^2, H1 order 2
(4,004,001 dofs, 16 M quadrature points, 64,016,001 nonzeros), functional =
MinimalSurfaceEnergy (ex2.cpp:12-24, eps=0.5), state u = sin(pi x)sin(pi y) +
0.1 U(-1,1) seed 1234.  Under torchrun (N>1) every rank assembles its own
1000x1000 block of a (Px*1000)x(Py*1000) mesh (weak scaling) and the shared
interface dofs of the residual are summed across ranks over NCCL.

--config 5 makes BASELINE.json's config 5 the workload of the line: the ex4 LVPP block system (H1 order 3 x L2
order 1, FermiDirac entropy, 5x5 points; ex4.cpp:93-104) on 1024x1024 elements PER GPU, overlapping element
partition (one ghost layer), step = P x over NCCL (madb_exchange_*) + fused residual + Jacobian assembly, which
gives the owned rows of P^T A P complete (ex4.cpp:136,169,190).  The default line (config 2) carries the device
timings of configs 3, 4 and 5 in "other_configs" (N = 1) and the weak-scaled config 5 in "config5" (every N), so
that those numbers are driver-run too.  --check asserts parity against the CPU oracle (small meshes, same code
path; the multi-GPU cases of tests/mgpu_check.py when N > 1) before anything is timed.

The JSON line follows the driver contract; see DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

EPS = 0.5
P = 2
FP64_INSTR = 2278  # DFMA+DMUL+DADD executed per element by k_patch_ws<minsurf,Q2> (ncu smsp__sass_thread_inst_executed_op_d*, r02 v1)
# dram__bytes_read.sum + dram__bytes_write.sum of one k_patch_ws launch of this workload (1000x1000),
# ncu --set full capture of the final build, profiles/r02_v7_k_patch_ws_full.txt (409.5 MB read + 521.8 MB written)
NCU_TRAFFIC_BYTES = {1000: 931366912}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def make_state(mesh, space, seed=1234):
    from mfem_ad_b200 import meshgen as G
    xc = G.dof_coords(mesh, space)
    u = np.sin(np.pi * xc[:, 0]) * np.sin(np.pi * xc[:, 1])
    return u + 0.1 * np.random.default_rng(seed).uniform(-1, 1, space["ndofs"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.p, self.lines = gpu, None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.p.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference algorithm on the host cores
# ----------------------------------------------------------------------------------
def cpu_forms(nx, ny, nthreads):
    """nthreads 'ranks' (mimics mpirun -np N, test.sh:9): each owns a strip of the nx x ny mesh."""
    from mfem_ad_b200 import meshgen as G
    from oracle import oracle as O
    import spec as S
    forms = []
    rows = [ny // nthreads + (1 if r < ny % nthreads else 0) for r in range(nthreads)]
    for r in range(nthreads):
        if rows[r] == 0:
            continue
        mesh = G.cartesian_mesh((nx, rows[r]), lengths=(1.0, rows[r] / ny))
        s = G.h1_space(mesh, P, mode=O.GRAD)
        f = O.OracleForm(mesh, [s], S.minsurf(2, EPS).oracle())
        f.pattern()
        x = make_state(mesh, s)
        forms.append((f, x, s["ndofs"]))
    return forms


def cpu_step(forms):
    def work(f, x):
        f.mult(x)
        f.grad(x)
    th = [threading.Thread(target=work, args=(f, x)) for f, x, _ in forms]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nx = args.n
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    from mfem_ad_b200 import parallel as PAR
    px, py = PAR.rank_grid(max(world, 1))
    forms = cpu_forms(nx, nx, cores)
    ndof = (nx * P + 1) ** 2
    for _ in range(args.warmup):
        cpu_step(forms)
    ts = [cpu_step(forms) for _ in range(args.steps)]
    t = float(np.mean(ts))
    val = ndof / t  # DOF/s of the host: independent of how many blocks it would have to assemble (linear work)
    sample = "one %dx%d Q%d block of the %dx%d-block mesh, split in %d strips, one thread per strip, residual+Jacobian" % (
        nx, nx, P, px, py, len(forms))
    print(json.dumps({
        "impl": "reference", "metric": "AD residual+Jacobian assembly DOF/s", "value": val, "unit": "DOF/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(nx, px, py),
        "qpts_per_s": nx * nx * (P + 2) ** 2 / t,
        "cpu_baseline": {"value": val, "unit": "DOF/s", "cores": len(forms), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(nx, px, py):
    return {"workload": "config2: %dx%d quads per GPU, H1 order %d, MinimalSurfaceEnergy eps=%g (ex2.cpp:12-24), "
                        "fused residual+Jacobian assembly, CSR %s" % (nx, nx, P, EPS, "sorted columns"),
            "elements_per_gpu": nx * nx, "order": P, "quadrature": "%dx%d Gauss-Legendre" % (P + 2, P + 2),
            "rank_grid": "%dx%d" % (px, py),
            "l2": "working set per step (CSR values + maps) is > 6x the 126 MB L2; no explicit flush"}


def bind_to_gpu_numa(local):
    """Run this rank (and allocate its pinned host buffers: first touch) on the NUMA node its GPU hangs off; under torchrun
    all ranks otherwise start on node 0 and the host<->device copies of the end-to-end leg share one memory controller
    (round-1 record: 51 GB/s at N = 1, 11 GB/s per rank at N = 8).  Best effort: silently does nothing without sysfs."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import mfem_ad_b200 as M
    from mfem_ad_b200 import meshgen as G

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = bind_to_gpu_numa(local) if world > 1 else None

    def note(msg):
        if os.environ.get("MADB_BENCH_VERBOSE"):
            print("[rank %d] %s %.1fs" % (rank, msg, time.perf_counter() - t_start), file=sys.stderr, flush=True)
    t_start = time.perf_counter()

    from mfem_ad_b200 import parallel as PAR
    ctx = M.Context(local)
    comm = PAR.Comm(ctx) if world > 1 else None
    if args.check:
        check_parity(ctx, comm, rank, world, local)
        note("parity check passed")
    if args.config == 5:
        out = run_config5(args, ctx, comm, dist, rank, world, local, dev)
        if rank == 0:
            if args.check:
                out["check"] = "parity vs CPU oracle asserted before timing"
            print(json.dumps(out))
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    nx = args.n
    blk = PAR.cartesian_block(rank, world, nx, P)  # this rank's block of the (px*nx) x (py*nx) mesh
    px, py = blk["px"], blk["py"]
    mesh = blk["mesh"]
    space = dict(blk["space"], mode=M.GRAD)
    ndof = space["ndofs"]
    gm = M.Mesh(ctx, mesh)
    gs = M.Space(ctx, gm, space)
    fn = M.Functional(ctx, "minsurf", params=[EPS])
    gi = M.Integrator(ctx, [(gs, M.GRAD)], fn)
    t0 = time.perf_counter()
    nnz = gi.nnz
    setup_s = time.perf_counter() - t0

    xh = make_state(mesh, space)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    x = torch.from_numpy(xh).to(dev)
    y = torch.empty(ndof, dtype=torch.float64, device=dev)
    vals = torch.empty(nnz, dtype=torch.float64, device=dev)
    xp = torch.from_numpy(xh).pin_memory()
    yp = torch.empty(ndof, dtype=torch.float64).pin_memory()
    vp = torch.empty(nnz, dtype=torch.float64).pin_memory()

    # shared-dof exchange P^T y behind the C ABI (madb_exchange_*: pack kernel, ncclSend/ncclRecv group on the
    # communication stream, unpack-add in ascending peer rank: deterministic), asynchronous on the context stream
    note("integrator + pattern ready")
    ex = None
    if world > 1:
        owner = PAR.lowest_rank_owner(blk["l2g"], blk["candidates"], ndof, rank, world)
        ex = PAR.HaloExchange(*PAR.halo_lists(blk["l2g"], owner, rank, world), ctx=ctx, comm=comm)
    note("exchange lists ready")

    def step_device():
        if ex is None:
            gi.assemble(x, y, vals)
        else:
            # residual first, P^T y over NCCL under the interface reduction of the CSR values
            gi.assemble_begin(x, y, vals)
            ex.begin(y, True)
            gi.assemble_end()
            ex.end(y, True)

    def step_e2e():
        gi.assemble(xp.numpy(), yp.numpy(), vp.numpy())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    gi.assemble(x, y, vals)
    stats = gi.patch_stats()
    patch = stats["patches"] > 0
    # patch path: k_patch_ws + one interface reduction launch (residual rows and CSR entries); colour path: one launch per colour
    launches_per_step = (2 if patch else gi.ncolors) + (2 if world > 1 else 0)  # + pack / unpack of the exchange
    kernel_launches = 1 if patch else gi.ncolors
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step_device()
        note("warm-up enqueued")
        barrier()
        note("warm-up done")
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for k in range(args.steps):
            step_device()
        ev1.record(stream)
        barrier()
        ms_total = ev0.elapsed_time(ev1)
        note("timed region done")
        # dominant kernel alone: CUDA events recorded by the library on its stream around the element kernel(s)
        # (off during the timed steps: an event record between two kernels ends their programmatic overlap)
        gi.set_timing(True)
        kms = []
        for k in range(min(args.steps, 10)):
            gi.assemble(x, y, vals)
            kms.append(gi.last_kernel_ms())
        ms_kernel = float(np.mean(kms))
        gi.set_timing(False)
        # keep the GPU busy a little longer so the clock sampler sees load (no collective in here: the loop is
        # time-based, ranks run different trip counts)
        t_end = time.perf_counter() + 1.0
        while time.perf_counter() < t_end:
            gi.assemble(x, y, vals)
        torch.cuda.synchronize()
        note("clock sampling done")
        clocks = sampler.stop() if rank == 0 else None

        # end to end through the C ABI with HOST (pinned) buffers: H2D of x, D2H of y and the CSR values inside
        for _ in range(2):
            step_e2e()
        barrier()
        e2e_steps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps

    ms_step = ms_total / args.steps
    if dist is not None:
        t = torch.tensor([ms_step, ms_kernel, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, ms_kernel, e2e_ms = [float(v) for v in t.tolist()]

    if rank == 0:
        peak, peak_src = peaks()
        nvert = (nx + 1) ** 2
        alg_bytes = 8 * ndof + 8 * 2 * nvert + 4 * (P + 1) ** 2 * nx * nx + 8 * ndof + 8 * nnz
        achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
        nq = nx * nx * (P + 2) ** 2
        out = {
            "metric": "AD residual+Jacobian assembly DOF/s", "value": world * ndof / (ms_step * 1e-3), "unit": "DOF/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(nx, px, py),
            "qpts_per_s": world * nq / (ms_step * 1e-3),
            "dofs_per_gpu": ndof, "nnz_per_gpu": int(nnz), "setup_s": setup_s,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES.get(nx) if patch else None, "traffic_unit": "bytes per launch (ncu)",
                         "peak_source": peak_src,
                         "kernel": ("k_patch_ws" if patch else "k_element") + "<MinimalSurfaceEnergy<2>,Q2,RES|JAC>",
                         "launches_per_step": kernel_launches, "avg_launch_ms": ms_kernel / kernel_launches,
                         "algorithmic_bytes_per_step": alg_bytes,
                         "note": "achieved = algorithmic bytes of one assembly / device time of the element kernel "
                                 "(events on the library stream); the interface reductions add ms_per_step - kernel time",
                         "fp64": {"fp64_instr_per_element": FP64_INSTR, "peak_tflops_measured": 37.1,
                                  "frac": 2.0 * FP64_INSTR * nx * nx / (ms_kernel * 1e-3) / 37.1e12}},
            "patch_stats": stats,
            "e2e": {"value": world * ndof / (e2e_ms * 1e-3), "unit": "DOF/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 8 * ndof, "d2h_bytes_per_step": 8 * ndof + 8 * int(nnz)},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if numa_node is not None:
            out["e2e"]["host_numa_node_rank0"] = numa_node
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            forms = cpu_forms(nx, nx, cores)
            cpu_step(forms)
            t = min(cpu_step(forms) for _ in range(3))
            # one thread (the reference is single-threaded per rank): bounded sample, a 250x250 corner
            m1 = min(250, nx)
            f1 = cpu_forms(m1, m1, 1)
            t1 = min(cpu_step(f1) for _ in range(2))
            out["cpu_baseline_1thread"] = {"value": (m1 * P + 1) ** 2 / t1, "unit": "DOF/s", "cores": 1, "kind": "port",
                                           "ms_per_step": t1 * 1e3,
                                           "sample": "%dx%d Q%d mesh, one oracle thread, residual+Jacobian, best of 2" % (m1, m1, P)}
            out["cpu_baseline"] = {"value": ndof / t, "unit": "DOF/s", "cores": len(forms), "kind": "port",
                                   "ms_per_step": t * 1e3,
                                   "sample": "full %dx%d Q%d mesh split in %d strips, one oracle thread per strip "
                                             "(mimics mpirun -np N), residual+Jacobian, best of 3" % (nx, nx, P, len(forms))}
        if args.check:
            out["check"] = "parity vs CPU oracle asserted before timing"
    # the other configurations, driver-run: configs 3 - 5 device timings at N = 1, weak-scaled config 5 at every N
    extra5 = None
    if not args.no_extra:
        del gi, vals, vp, y, yp
        torch.cuda.empty_cache()
        try:
            extra5 = c5_summary(args, ctx, comm, dist, rank, world, local, dev)
        except Exception as ex5:  # never lose the main line
            extra5 = {"error": repr(ex5)[:300]}
    if rank == 0:
        if extra5 is not None:
            out["config5"] = extra5
        if world == 1 and not args.no_extra:
            try:
                out["other_configs"] = [r for r in other_configs() if r.get("config") in ("3", "4") or "error" in r]
            except Exception as exo:
                out["other_configs"] = [{"error": repr(exo)[:300]}]
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------
# config 5: ex4 LVPP block system, weak-scaled, overlapping element partition + P x over NCCL
# ----------------------------------------------------------------------------------
C5_ORDER = 2  # ex4.cpp -o 2: H1 order 3 x L2 order 1, rule order 3*2+3 = 9 -> 5x5 points
C5_ALPHA = 0.1


def c5_spec():
    import spec as S
    return S.pg(S.obstacle(2), S.fermidirac(0.0, 0.5), C5_ALPHA)


def c5_cpu_forms(m, nthreads):
    """The oracle on an m x m sample of the config-5 mesh, one strip per thread."""
    from mfem_ad_b200 import meshgen as G
    from oracle import oracle as O
    forms = []
    rows = [m // nthreads + (1 if r < m % nthreads else 0) for r in range(nthreads)]
    for r in range(nthreads):
        if rows[r] == 0:
            continue
        mesh = G.cartesian_mesh((m, rows[r]), lengths=(1.0, rows[r] / m))
        h1 = G.h1_space(mesh, C5_ORDER + 1, mode=O.VALUE | O.GRAD)
        l2 = G.l2_space(mesh, C5_ORDER - 1, mode=O.VALUE)
        pk = np.zeros(l2["ndofs"])
        f = O.OracleForm(mesh, [h1, l2], c5_spec().oracle(), quad_order=3 * C5_ORDER + 3,
                         params=[dict(type=O.PRM_GF, size=1, data=pk, space=l2)])
        f.pattern()
        nd = h1["ndofs"] + l2["ndofs"]
        x = 0.1 * np.random.default_rng(r).uniform(-1, 1, nd)
        forms.append((f, x, nd))
    return forms


def c5_cpu_sample(m, cores):
    forms = c5_cpu_forms(m, cores)
    cpu_step(forms)
    t = min(cpu_step(forms) for _ in range(2))
    nd = (m * (C5_ORDER + 1) + 1) ** 2 + m * m * C5_ORDER ** 2
    return nd / t, t, len(forms)


def c5_setup(ctx, comm, rank, world, local, n):
    """This rank's block (+ ghost layer), spaces, integrator, exchanges and device vectors."""
    import torch
    import mfem_ad_b200 as M
    from mfem_ad_b200 import parallel as PAR
    dev = torch.device("cuda", local)
    blk = PAR.cartesian_block_ghost(rank, world, n, C5_ORDER + 1, C5_ORDER - 1)
    h1 = dict(blk["h1"], mode=M.VALUE | M.GRAD)
    l2 = dict(blk["l2"], mode=M.VALUE)
    nh, nl = h1["ndofs"], l2["ndofs"]
    gm = M.Mesh(ctx, blk["mesh"])
    gh, gl = M.Space(ctx, gm, h1), M.Space(ctx, gm, l2)
    fn = c5_spec().madb(ctx)
    gi = M.Integrator(ctx, [(gh, M.VALUE | M.GRAD), (gl, M.VALUE), (gl, M.VALUE, M.ROLE_PARAM)], fn,
                      quad_order=3 * C5_ORDER + 3)
    nh_glob = (blk["px"] * n * (C5_ORDER + 1) + 1) * (blk["py"] * n * (C5_ORDER + 1) + 1)
    l2g = np.concatenate([blk["l2g_h1"], nh_glob + blk["l2g_l2"]])
    owner = np.concatenate([blk["owner_h1"], blk["owner_l2"]])
    ex = exl = None
    if world > 1:
        ex = PAR.HaloExchange(*PAR.halo_lists(l2g, owner, rank, world), ctx=ctx, comm=comm)
        exl = PAR.HaloExchange(*PAR.halo_lists(blk["l2g_l2"], blk["owner_l2"], rank, world), ctx=ctx, comm=comm)
    mine = owner == rank
    # state: a smooth function of the GLOBAL dof id (every rank agrees), copies start with garbage and are filled by P
    xg = 0.1 * np.sin(0.37 * (l2g % 9973)) + 0.05 * np.cos(0.011 * (l2g % 7919))
    x = torch.from_numpy(np.where(mine, xg, -77.0)).to(dev)
    pk = torch.from_numpy(np.where(blk["owner_l2"] == rank, 0.01 * np.sin(0.5 * (blk["l2g_l2"] % 1013)), -55.0)).to(dev)
    if exl is not None:
        exl.forward(pk)
    gi.set_param_field(2, pk)
    nnz = gi.nnz
    y = torch.empty(nh + nl, dtype=torch.float64, device=dev)
    vals = torch.empty(nnz, dtype=torch.float64, device=dev)
    ne_loc = int(np.asarray(blk["mesh"]["e2n"]).shape[0])
    return dict(blk=blk, gi=gi, ex=ex, exl=exl, x=x, y=y, vals=vals, pk=pk, xg=xg, mine=mine, nh=nh, nl=nl, nnz=int(nnz),
                n_owned=int(mine.sum()), ne_local=int(ne_loc), ne_owned=int(blk["owned_elements"].sum()),
                keep=(gm, gh, gl, fn))


def c5_time(ctx, W, dist, dev, steps, warmup, sampler=None):
    """Device-timed weak-scaling step: P x (halo of the state) + fused residual + Jacobian assembly."""
    import torch
    gi, ex, x, y, vals = W["gi"], W["ex"], W["x"], W["y"], W["vals"]
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if ex is not None:
            ex.begin(x)
            ex.end(x)
        gi.assemble(x, y, vals)

    with torch.cuda.stream(stream):
        for _ in range(max(warmup, 3)):
            step()
        barrier()
        if sampler is not None:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        barrier()
        ms_step = e0.elapsed_time(e1) / steps
        gi.set_timing(True)
        kms = []
        for _ in range(min(steps, 5)):
            step()
            kms.append(gi.last_kernel_ms())
        ms_kernel = float(np.mean(kms))
        # the exchange alone
        ms_ex = 0.0
        if ex is not None:
            barrier()
            e0.record(stream)
            for _ in range(steps):
                ex.begin(x)
                ex.end(x)
            e1.record(stream)
            barrier()
            ms_ex = e0.elapsed_time(e1) / steps
        if sampler is not None:
            t_end = time.perf_counter() + 1.0
            while time.perf_counter() < t_end:
                gi.assemble(x, y, vals)
            torch.cuda.synchronize()
    gi.set_timing(False)
    return ms_step, ms_kernel, ms_ex


def c5_alg_bytes(W, n):
    """state + residual (8 B per dof each), vertex coordinates, element->dof maps (16 + 4 ints), psi_k, CSR values."""
    nd = W["nh"] + W["nl"]
    return 8 * nd * 2 + 16 * (n + 1) ** 2 + 80 * W["ne_local"] + 8 * W["nnz"] + 8 * W["nl"]


def c5_workload(n, px, py):
    return {"workload": "config5: ex4 LVPP block system (H1 order 3 x L2 order 1, FermiDirac entropy, alpha=%g, 5x5 points; "
                        "ex4.cpp:93-104), %dx%d quads per GPU + one ghost layer, P x over NCCL + fused residual+Jacobian "
                        "assembly (owned rows of P^T A P complete), CSR sorted columns" % (C5_ALPHA, n, n),
            "elements_per_gpu": n * n, "order": "H1 p3 x L2 p1", "quadrature": "5x5 Gauss-Legendre", "rank_grid": "%dx%d" % (px, py),
            "l2": "working set per step (3 GB of CSR values) is > 20x the 126 MB L2; no explicit flush"}


def run_config5(args, ctx, comm, dist, rank, world, local, dev):
    import torch
    n = args.n5
    W = c5_setup(ctx, comm, rank, world, local, n)
    sampler = ClockSampler(local) if rank == 0 else None
    ms_step, ms_kernel, ms_ex = c5_time(ctx, W, dist, dev, args.steps, args.warmup, sampler)
    clocks = sampler.stop() if sampler is not None else None
    # end to end: host (pinned) state in, residual + CSR values out, P x inside
    gi = W["gi"]
    nd = W["nh"] + W["nl"]
    xp = W["x"].cpu().pin_memory()
    yp = torch.empty(nd, dtype=torch.float64).pin_memory()
    vp = torch.empty(W["nnz"], dtype=torch.float64).pin_memory()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def step_e2e():
        if W["ex"] is not None:
            W["x"].copy_(xp, non_blocking=True)
            W["ex"].begin(W["x"])
            W["ex"].end(W["x"])
            gi.assemble(W["x"], W["y"], W["vals"])
            yp.copy_(W["y"], non_blocking=True)
            vp.copy_(W["vals"], non_blocking=True)
        else:
            gi.assemble(xp.numpy(), yp.numpy(), vp.numpy())
    with torch.cuda.stream(stream):
        step_e2e()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            step_e2e()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / 3
    owned = torch.tensor([float(W["n_owned"])], dtype=torch.float64, device=dev)
    t = torch.tensor([ms_step, ms_kernel, ms_ex, e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(owned, op=dist.ReduceOp.SUM)
    ms_step, ms_kernel, ms_ex, e2e_ms = [float(v) for v in t.tolist()]
    ndof_total = float(owned.item())
    if rank != 0:
        return None
    peak, peak_src = peaks()
    alg = c5_alg_bytes(W, n)
    achieved = alg / (ms_kernel * 1e-3) / 1e9
    blk = W["blk"]
    out = {
        "metric": "AD residual+Jacobian assembly DOF/s", "value": ndof_total / (ms_step * 1e-3), "unit": "DOF/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": c5_workload(n, blk["px"], blk["py"]),
        "qpts_per_s": world * 25 * n * n / (ms_step * 1e-3),
        "dofs_total": ndof_total, "nnz_per_gpu": W["nnz"], "exchange_ms": ms_ex,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "kernel": "k_patch<PGFunctional<ObstacleEnergy<2>,FermiDiracEntropy>,H1p3xL2p1,RES|JAC>",
                     "launches_per_step": 1, "avg_launch_ms": ms_kernel, "algorithmic_bytes_per_step": alg,
                     "note": "64-element patches, 4 threads per element; FP64 / latency bound, see DESIGN.md"},
        "e2e": {"value": ndof_total / (e2e_ms * 1e-3), "unit": "DOF/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": 8 * nd, "d2h_bytes_per_step": 8 * nd + 8 * W["nnz"]},
        "gpu_launches": (2 + (4 if world > 1 else 0)) * args.steps,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        m = 160
        v, tt, nf = c5_cpu_sample(m, cores)
        out["cpu_baseline"] = {"value": v, "unit": "DOF/s", "cores": nf, "kind": "port", "ms_per_step": tt * 1e3,
                               "sample": "%dx%d elements of the config-5 block form, %d strips, one oracle thread per strip, "
                                         "residual+Jacobian, best of 2" % (m, m, nf)}
    return out


def c5_summary(args, ctx, comm, dist, rank, world, local, dev):
    """Weak-scaled config 5 as a sub-object of the default (config 2) line."""
    W = c5_setup(ctx, comm, rank, world, local, args.n5)
    import torch
    ms_step, ms_kernel, ms_ex = c5_time(ctx, W, dist, dev, max(3, min(args.steps, 5)), 3)
    owned = torch.tensor([float(W["n_owned"])], dtype=torch.float64, device=dev)
    t = torch.tensor([ms_step, ms_kernel, ms_ex], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(owned, op=dist.ReduceOp.SUM)
    ms_step, ms_kernel, ms_ex = [float(v) for v in t.tolist()]
    peak, _ = peaks()
    alg = c5_alg_bytes(W, args.n5)
    res = {"workload": c5_workload(args.n5, W["blk"]["px"], W["blk"]["py"])["workload"], "n_gpus": world,
           "ms_per_step": ms_step, "element_kernel_ms": ms_kernel, "exchange_ms": ms_ex,
           "value": float(owned.item()) / (ms_step * 1e-3), "unit": "DOF/s", "scaling": "weak",
           "hbm_frac_kernel": alg / (ms_kernel * 1e-3) / 1e9 / peak, "nnz_per_gpu": W["nnz"]}
    del W
    torch.cuda.empty_cache()
    return res


def other_configs(scale=1):
    """Device timings of configs 3 - 5 (tools/bench_configs.py) collected for the default line (N = 1)."""
    import io
    import contextlib
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_configs", os.path.join(ROOT, "tools", "bench_configs.py"))
    bc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bc)
    import mfem_ad_b200 as M
    ctx = M.Context(0)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        for fn in (bc.config3, bc.config4, bc.config5):
            try:
                fn(ctx, scale)
            except Exception as ex:  # one config must not hide the others
                print(json.dumps({"config": fn.__name__, "error": repr(ex)[:200]}))
    return [json.loads(l) for l in buf.getvalue().splitlines() if l.startswith("{")]


def check_parity(ctx, comm, rank, world, local):
    """--check: the CUDA path against the CPU oracle on small meshes, before anything is timed."""
    import torch
    if world > 1:
        import mgpu_check as MC
        for name, fn in (("residual", MC.case_residual), ("block", MC.case_block)):
            err, flag = fn(rank, world, local, ctx, comm)
            t = torch.tensor([err, 0.0 if flag else 1.0], device=torch.device("cuda", local), dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            assert t[0].item() <= 1e-12 and t[1].item() == 0.0, ("multi-GPU parity check failed", name, t.tolist())
        return
    import contextlib
    import __graft_entry__ as GE
    with contextlib.redirect_stdout(sys.stderr):  # stdout carries the JSON line only
        GE.smoke()


def run_reference_c5(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    m = 160
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    from mfem_ad_b200 import parallel as PAR
    px, py = PAR.rank_grid(max(world, 1))
    forms = c5_cpu_forms(m, cores)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_step(forms)
    ts = [cpu_step(forms) for _ in range(max(1, min(args.steps, 5)))]
    t = float(np.mean(ts))
    nd = (m * (C5_ORDER + 1) + 1) ** 2 + m * m * C5_ORDER ** 2
    val = nd / t
    sample = "%dx%d elements of the config-5 block form split in %d strips, one thread per strip, residual+Jacobian" % (m, m, len(forms))
    print(json.dumps({
        "impl": "reference", "metric": "AD residual+Jacobian assembly DOF/s", "value": val, "unit": "DOF/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": c5_workload(args.n5, px, py), "qpts_per_s": 25 * m * m / t,
        "cpu_baseline": {"value": val, "unit": "DOF/s", "cores": len(forms), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--n", type=int, default=1000, help="elements per direction per GPU (config 2)")
    ap.add_argument("--n5", type=int, default=1024, help="elements per direction per GPU (config 5)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--config", type=int, default=2, choices=[2, 5], help="BASELINE.json configuration of the line")
    ap.add_argument("--check", action="store_true", help="assert parity against the CPU oracle before timing")
    ap.add_argument("--no-extra", action="store_true", help="skip the other_configs / config5 legs of the default line")
    args = ap.parse_args()
    # stdout carries the JSON line and nothing else: libraries that write to file descriptor 1 themselves (NCCL prints
    # its version banner there) are sent to stderr; the line is written to the original descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        (run_reference_c5 if args.config == 5 else run_reference)(args)
    else:
        run_gpu(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()

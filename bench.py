#!/usr/bin/env python
"""bench.py -- AD residual+Jacobian assembly throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n NX]

A step = one pass of the hot path over the whole mesh: ONE residual + ONE
Jacobian assembly at the same state (one Newton iteration's worth).
Workload at N=1: config 2 -- 1000x1000 quads on [0,1]^2, H1 order 2
(4,004,001 dofs, 16 M quadrature points, 64,016,001 nonzeros), functional =
MinimalSurfaceEnergy (ex2.cpp:12-24, eps=0.5), state u = sin(pi x)sin(pi y) +
0.1 U(-1,1) seed 1234.  Under torchrun (N>1) every rank assembles its own
1000x1000 block of a (Px*1000)x(Py*1000) mesh (weak scaling) and the shared
interface dofs of the residual are summed across ranks over NCCL.

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
# ncu --set full capture summarised in profiles/r02_k_patch_ws.md (409.3 MB read + 535.8 MB written)
NCU_TRAFFIC_BYTES = {1000: 945158400}


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
    forms = cpu_forms(nx, nx, cores)
    ndof = (nx * P + 1) ** 2
    for _ in range(args.warmup):
        cpu_step(forms)
    ts = [cpu_step(forms) for _ in range(args.steps)]
    t = float(np.mean(ts))
    val = ndof / t
    sample = "full %dx%d Q%d mesh split in %d strips, one thread per strip, residual+Jacobian" % (nx, nx, P, len(forms))
    print(json.dumps({
        "impl": "reference", "metric": "AD residual+Jacobian assembly DOF/s", "value": val, "unit": "DOF/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(nx, 1, 1),
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

    def note(msg):
        if os.environ.get("MADB_BENCH_VERBOSE"):
            print("[rank %d] %s %.1fs" % (rank, msg, time.perf_counter() - t_start), file=sys.stderr, flush=True)
    t_start = time.perf_counter()

    nx = args.n
    from mfem_ad_b200 import parallel as PAR
    blk = PAR.cartesian_block(rank, world, nx, P)  # this rank's block of the (px*nx) x (py*nx) mesh
    px, py = blk["px"], blk["py"]
    mesh = blk["mesh"]
    space = dict(blk["space"], mode=M.GRAD)
    ndof = space["ndofs"]
    ctx = M.Context(local)
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

    # shared-dof exchange P^T y: interface dofs summed on their owner rank, fixed order (deterministic)
    note("integrator + pattern ready")
    ex = PAR.SharedDofExchange(blk["l2g"], blk["candidates"], dev, ctx=ctx) if world > 1 else None
    note("exchange lists ready")

    def exchange():
        if ex is not None:
            ex.reduce_to_owner(y)

    def step_device():
        gi.assemble(x, y, vals)
        exchange()

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
    launches_per_step = 2 if patch else gi.ncolors
    kernel_launches = 1 if patch else gi.ncolors
    gi.set_timing(True)
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
            gi.assemble(x, y, vals)
            exchange()
        ev1.record(stream)
        barrier()
        ms_total = ev0.elapsed_time(ev1)
        note("timed region done")
        # dominant kernel alone: CUDA events recorded by the library on its stream around the element kernel(s)
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
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            forms = cpu_forms(nx, nx, cores)
            cpu_step(forms)
            t = min(cpu_step(forms) for _ in range(3))
            out["cpu_baseline"] = {"value": ndof / t, "unit": "DOF/s", "cores": len(forms), "kind": "port",
                                   "ms_per_step": t * 1e3,
                                   "sample": "full %dx%d Q%d mesh split in %d strips, one oracle thread per strip "
                                             "(mimics mpirun -np N), residual+Jacobian, best of 3" % (nx, nx, P, len(forms))}
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--n", type=int, default=1000, help="elements per direction per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    # stdout carries the JSON line and nothing else: libraries that write to file descriptor 1 themselves (NCCL prints
    # its version banner there) are sent to stderr; the line is written to the original descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()

"""Device timings of the hot path on the five BASELINE.json configurations at their full sizes.

    python tools/bench_configs.py [2 3 4 5] [--scale S]

Prints one JSON line per configuration: ms per call (CUDA events on the library stream, best-of after
warm-up), DOF/s, qpts/s and the fraction of the measured HBM copy peak for the algorithmic bytes of SURVEY 8d.
`--scale S` shrinks the meshes by S per direction (smoke runs).  Inputs are resident in HBM; this is the
`value` side of bench.py for the other configs, not a bench line of its own."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfem_ad_b200 as M  # noqa: E402
from mfem_ad_b200 import meshgen as G  # noqa: E402

PEAK = 6534.8
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(ctx, fn, n=10, warm=3):
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    with torch.cuda.stream(stream):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
    return best


def dev_vec(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.device("cuda", 0))


def report(name, what, ms, ndof, nq, alg_bytes, extra=None):
    out = {"config": name, "op": what, "ms": round(ms, 4), "dof_per_s": ndof / (ms * 1e-3), "qpts_per_s": nq / (ms * 1e-3),
           "algorithmic_GB": alg_bytes / 1e9, "hbm_frac": alg_bytes / (ms * 1e-3) / 1e9 / PEAK}
    if extra:
        out.update(extra)
    print(json.dumps(out), flush=True)


def config2(ctx, scale):
    nx = 1000 // scale
    for kind in ("diffusion", "minsurf"):
        for pert in (0.0, 0.2):
            mesh = G.cartesian_mesh((nx, nx), perturb=pert)
            sp = G.h1_space(mesh, 2, mode=M.GRAD)
            gm = M.Mesh(ctx, mesh)
            gs = M.Space(ctx, gm, sp)
            fn = M.Functional(ctx, kind, params=[0.5] if kind == "minsurf" else [], iparams=[] if kind == "minsurf" else [0])
            gi = M.Integrator(ctx, [(gs, M.GRAD)], fn)
            x = dev_vec(np.random.default_rng(0).uniform(-1, 1, sp["ndofs"]))
            y = torch.empty_like(x)
            vals = torch.empty(gi.nnz, dtype=torch.float64, device=x.device)
            ms = timeit(ctx, lambda: gi.assemble(x, y, vals))
            nd, nv, ne = sp["ndofs"], (nx + 1) ** 2, nx * nx
            alg = 8 * nd + 16 * nv + 36 * ne + 8 * nd + 8 * gi.nnz
            report("2", "residual+Jacobian %s%s" % (kind, " perturbed mesh" if pert else ""), ms, nd, 16 * ne, alg)
            del gi, gs, gm


def config3(ctx, scale):
    n = 104 // scale
    mesh = G.cartesian_mesh((n, n, n))
    sp = G.h1_space(mesh, 3, mode=M.GRAD)
    gm = M.Mesh(ctx, mesh)
    gs = M.Space(ctx, gm, sp)
    fn = M.Functional(ctx, "minsurf", params=[0.5])
    gi = M.Integrator(ctx, [(gs, M.GRAD)], fn)
    nd, ne, nv = sp["ndofs"], n ** 3, (n + 1) ** 3
    x = dev_vec(np.random.default_rng(0).uniform(-1, 1, nd))
    v = dev_vec(np.random.default_rng(4321).uniform(-1, 1, nd))
    y = torch.empty_like(x)
    ms = timeit(ctx, lambda: gi.mult(x, y), n=5)
    report("3", "sum-factorised residual (minsurf)", ms, nd, 125 * ne, 8 * nd + 24 * nv + 256 * ne + 8 * nd)
    ms = timeit(ctx, lambda: gi.grad_mult(x, v, y), n=5)
    report("3", "matrix-free Jacobian action", ms, nd, 125 * ne, 8 * nd * 3 + 24 * nv + 256 * ne)


def config4(ctx, scale):
    n = 1072 // scale
    mesh = G.cartesian_mesh((n, n))
    disp = G.h1_space(mesh, 1, vdim=2, mode=M.GRAD | M.VECTOR)
    lat = G.h1_space(mesh, 1, vdim=5, mode=M.VALUE | M.VECTOR)
    E = [1e-3, 0.25, 0.5, 0.75, 1.0]
    nn, ne = lat["ndofs"], n * n
    psi = np.random.default_rng(99).normal(0, 1, 5 * nn)
    gm = M.Mesh(ctx, mesh)
    gd, gl = M.Space(ctx, gm, disp), M.Space(ctx, gm, lat)
    # latent -> primal (softmax) at the dofs
    simplex = M.Functional(ctx, "simplex", params=[1.0])
    pd = dev_vec(psi.reshape(5, nn).T.copy())
    g = torch.empty_like(pd)
    ms = timeit(ctx, lambda: simplex.eval_device(pd, grad=g))
    report("4", "softmax latent map at the nodes", ms, 5 * nn, 0, 2 * 40 * nn)
    p = psi.reshape(5, nn)
    e = np.exp(p - p.max(axis=0))
    rho = dev_vec((e / e.sum(axis=0)).reshape(-1))
    lam = M.Functional(ctx, "simp", params=E + [3.0])
    mu = M.Functional(ctx, "simp", params=[0.5 * q for q in E] + [3.0])
    fn = M.Functional(ctx, "paramcompliance", children=[lam, mu])
    gi = M.Integrator(ctx, [(gd, M.GRAD | M.VECTOR), (gl, M.VALUE | M.VECTOR, M.ROLE_PARAM)], fn)
    gi.set_param_field(1, rho)
    x = dev_vec(np.random.default_rng(5).uniform(-1, 1, 2 * nn))
    y = torch.empty_like(x)
    vals = torch.empty(gi.nnz, dtype=torch.float64, device=x.device)
    ms = timeit(ctx, lambda: gi.assemble(x, y, vals))
    alg = 16 * nn + 16 * nn + 32 * ne + 16 * nn + 8 * gi.nnz + 40 * nn
    report("4", "state block residual+Jacobian (ParametrizedCompliance, SIMP)", ms, 2 * nn, 9 * ne, alg, dict(total_dofs=7 * nn))
    fd = M.Functional(ctx, "designcompliance", children=[lam, mu])
    gj = M.Integrator(ctx, [(gl, M.VALUE | M.VECTOR), (gd, M.GRAD, M.ROLE_PARAM)], fd)
    gj.set_param_field(1, x)
    val = torch.empty(ne * 9, dtype=torch.float64, device=x.device)
    grd = torch.empty(ne * 9 * 5, dtype=torch.float64, device=x.device)
    ms = timeit(ctx, lambda: gj.coefficient_device(rho, val, grd))
    report("4", "ParamGradient at the points", ms, 5 * nn, 9 * ne, 40 * nn + 16 * nn + 8 * 9 * ne * 6)


def config5(ctx, scale):
    n = 1024 // scale
    order = 2
    mesh = G.cartesian_mesh((n, n))
    h1 = G.h1_space(mesh, order + 1, mode=M.VALUE | M.GRAD)
    l2 = G.l2_space(mesh, order - 1, mode=M.VALUE)
    gm = M.Mesh(ctx, mesh)
    gh, gl = M.Space(ctx, gm, h1), M.Space(ctx, gm, l2)
    fn = M.Functional(ctx, "pg", params=[0.1], iparams=[0],
                      children=[M.Functional(ctx, "obstacle"), M.Functional(ctx, "fermidirac", params=[0.0, 0.5])])
    gi = M.Integrator(ctx, [(gh, M.VALUE | M.GRAD), (gl, M.VALUE), (gl, M.VALUE, M.ROLE_PARAM)], fn, quad_order=3 * order + 3)
    nd = h1["ndofs"] + l2["ndofs"]
    ne = n * n
    psik = dev_vec(np.zeros(l2["ndofs"]))
    gi.set_param_field(2, psik)
    x = dev_vec(0.1 * np.random.default_rng(0).uniform(-1, 1, nd))
    y = torch.empty_like(x)
    t0 = time.perf_counter()
    vals = torch.empty(gi.nnz, dtype=torch.float64, device=x.device)
    setup = time.perf_counter() - t0
    ms = timeit(ctx, lambda: gi.assemble(x, y, vals), n=5)
    alg = 8 * nd * 2 + 16 * (n + 1) ** 2 + 80 * ne + 8 * gi.nnz + 8 * l2["ndofs"]
    report("5", "ex4 PG block residual+Jacobian per GPU (H1 p3 x L2 p1, 5x5 points)", ms, nd, 25 * ne, alg,
           dict(nnz=int(gi.nnz), pattern_setup_s=round(setup, 2), patches=gi.patch_stats()["patches"]))
    ms = timeit(ctx, lambda: gi.mult(x, y), n=5)
    report("5", "ex4 PG block residual", ms, nd, 25 * ne, 8 * nd * 2 + 16 * (n + 1) ** 2 + 80 * ne + 8 * l2["ndofs"])
    # fused LVPP latent update
    nl = l2["ndofs"]
    psi, pk, lp, w = [dev_vec(np.random.default_rng(k).uniform(0, 1, nl)) for k in range(4)]
    ms = timeit(ctx, lambda: M.lvpp_update(ctx, 0.1, psi, pk, lp, w))
    report("5", "fused LVPP latent update", ms, nl, 0, 40 * nl)


def main():
    argv = sys.argv[1:]
    scale = 1
    if "--scale" in argv:
        k = argv.index("--scale")
        scale = int(argv[k + 1])
        argv = argv[:k] + argv[k + 2:]
    which = [a for a in argv if a in ("2", "3", "4", "5")] or ["2", "3", "4", "5"]
    ctx = M.Context(0)
    for c in which:
        try:
            {"2": config2, "3": config3, "4": config4, "5": config5}[c](ctx, scale)
        except Exception as ex:  # keep going: one config must not hide the others
            print(json.dumps({"config": c, "error": repr(ex)[:300]}), flush=True)


if __name__ == "__main__":
    main()

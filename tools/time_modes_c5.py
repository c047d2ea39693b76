"""Timing breakdown of the config-5 block assembly (k_patch, 64-element patches x 4 threads per element):
compute only (no write-out), residual + Jacobian computed with only y written, full."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mfem_ad_b200 as M
from mfem_ad_b200 import meshgen as G
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = M.Context(0)
mesh = G.cartesian_mesh((n, n))
h1 = G.h1_space(mesh, 3, mode=M.VALUE | M.GRAD); l2 = G.l2_space(mesh, 1, mode=M.VALUE)
gm = M.Mesh(ctx, mesh); gh, gl = M.Space(ctx, gm, h1), M.Space(ctx, gm, l2)
fn = M.Functional(ctx, "pg", params=[0.1], iparams=[0], children=[M.Functional(ctx, "obstacle"), M.Functional(ctx, "fermidirac", params=[0.0, 0.5])])
gi = M.Integrator(ctx, [(gh, M.VALUE | M.GRAD), (gl, M.VALUE), (gl, M.VALUE, M.ROLE_PARAM)], fn, quad_order=9)
dev = torch.device("cuda", 0)
nd = h1["ndofs"] + l2["ndofs"]
gi.set_param_field(2, torch.zeros(l2["ndofs"], dtype=torch.float64, device=dev))
x = torch.from_numpy(0.1 * np.random.default_rng(0).uniform(-1, 1, nd)).to(dev)
y = torch.empty_like(x); vals = torch.empty(gi.nnz, dtype=torch.float64, device=dev)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
def timeit(fn_, k=5):
    with torch.cuda.stream(stream):
        for _ in range(2): fn_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k): fn_()
        e1.record(stream); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k
gi.assemble(x, y, vals)
print(gi.patch_stats())
print("full assemble      %.3f ms" % timeit(lambda: gi.assemble(x, y, vals)))
print("compute only       %.3f ms" % timeit(lambda: gi.assemble(x, None, None)))
print("jac+resid, y only  %.3f ms" % timeit(lambda: gi.assemble(x, y, None)))
print("residual kernel    %.3f ms" % timeit(lambda: gi.mult(x, y)))

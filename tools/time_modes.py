"""Timing breakdown of the config-2 assembly kernel: compute only (no scatter), residual only, full."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mfem_ad_b200 as M
from mfem_ad_b200 import meshgen as G
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
kind = sys.argv[2] if len(sys.argv) > 2 else "minsurf"
ctx = M.Context(0)
mesh = G.cartesian_mesh((nx, nx))
sp = G.h1_space(mesh, 2, mode=M.GRAD)
gm = M.Mesh(ctx, mesh); gs = M.Space(ctx, gm, sp)
fn = M.Functional(ctx, kind, params=[0.5] if kind == "minsurf" else [], iparams=[] if kind == "minsurf" else [0])
gi = M.Integrator(ctx, [(gs, M.GRAD)], fn)
nnz = gi.nnz
dev = torch.device("cuda", 0)
x = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, sp["ndofs"])).to(dev)
y = torch.empty_like(x); vals = torch.empty(nnz, dtype=torch.float64, device=dev)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
def timeit(fn_, n=10):
    with torch.cuda.stream(stream):
        for _ in range(3): fn_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n): fn_()
        e1.record(stream); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
gi.assemble(x, y, vals)
print(gi.patch_stats(), "nnz", nnz)
print("full assemble      %.3f ms" % timeit(lambda: gi.assemble(x, y, vals)))
print("compute only       %.3f ms" % timeit(lambda: gi.assemble(x, None, None)))
print("jac+resid, y only  %.3f ms" % timeit(lambda: gi.assemble(x, y, None)))
print("jac only write     %.3f ms" % timeit(lambda: gi.assemble(x, None, vals)))
print("residual kernel    %.3f ms" % timeit(lambda: gi.mult(x, y)))
print("action kernel      %.3f ms" % timeit(lambda: gi.grad_mult(x, x, y)))

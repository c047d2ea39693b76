"""Host-side setup times (mesh/space/integrator/pattern/patch maps) of configs 2 and 5 at full size, and one Newton
iteration end to end (state on the host in, correction out) with the linear solve on the device vs SuperLU on the host."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfem_ad_b200 as M
from mfem_ad_b200 import meshgen as G, lvpp
import scipy.sparse as sp, scipy.sparse.linalg as spla
ctx = M.Context(0)
dev = torch.device("cuda", 0)

def t(): torch.cuda.synchronize(); return time.perf_counter()

# ---- setup times
t0 = t(); mesh = G.cartesian_mesh((1000, 1000)); s = G.h1_space(mesh, 2, mode=M.GRAD); t1 = t()
gm = M.Mesh(ctx, mesh); gs = M.Space(ctx, gm, s)
gi = M.Integrator(ctx, [(gs, M.GRAD)], M.Functional(ctx, "minsurf", params=[0.5])); t2 = t()
nnz = gi.nnz; t3 = t()
x = torch.zeros(s["ndofs"], dtype=torch.float64, device=dev); y = torch.empty_like(x); v = torch.empty(nnz, dtype=torch.float64, device=dev)
gi.assemble(x, y, v); t4 = t()
print("config 2 setup: python mesh/space %.2f s, create (colouring, patches, residual maps) %.2f s, pattern %.2f s, first assemble (matrix maps) %.2f s" % (t1 - t0, t2 - t1, t3 - t2, t4 - t3))
del gi, x, y, v
t0 = t(); mesh = G.cartesian_mesh((1024, 1024)); h1 = G.h1_space(mesh, 3, mode=M.VALUE | M.GRAD); l2 = G.l2_space(mesh, 1, mode=M.VALUE); t1 = t()
gm = M.Mesh(ctx, mesh); gh, gl = M.Space(ctx, gm, h1), M.Space(ctx, gm, l2)
fn = M.Functional(ctx, "pg", params=[0.1], iparams=[0], children=[M.Functional(ctx, "obstacle"), M.Functional(ctx, "fermidirac", params=[0.0, 0.5])])
gi = M.Integrator(ctx, [(gh, M.VALUE | M.GRAD), (gl, M.VALUE), (gl, M.VALUE, M.ROLE_PARAM)], fn, quad_order=9); t2 = t()
nnz = gi.nnz; t3 = t()
print("config 5 setup: python mesh/space %.2f s, create %.2f s, pattern (%d nnz) %.2f s" % (t1 - t0, t2 - t1, nnz, t3 - t2))
del gi
torch.cuda.empty_cache()

# ---- one Newton iteration end to end, ex2 (minimal surface), n x n Q2
for n in (128, 256):
    mesh = G.cartesian_mesh((n, n)); s = G.h1_space(mesh, 2, mode=M.GRAD)
    ess = G.boundary_dofs(mesh, s)
    gm = M.Mesh(ctx, mesh); gs = M.Space(ctx, gm, s)
    gi = M.Integrator(ctx, [(gs, M.GRAD)], M.Functional(ctx, "minsurf", params=[0.5]))
    gi.set_essential(ess)
    xc = G.dof_coords(mesh, s)
    x = 0.3 * np.sin(2 * np.pi * xc[:, 0]) * xc[:, 1]
    b = np.zeros_like(x)
    lin = lvpp.DeviceLinear(gi, "pcg", rtol=1e-10)
    lin.step(x, b)
    t0 = t(); c = lin.step(x, b); t1 = t()
    rp, ci = gi.pattern()
    t2 = t(); r, vals = gi.assemble(x); J = sp.csr_matrix((vals, ci, rp), shape=(x.size,) * 2).tocsc(); c2 = spla.splu(J).solve(r - b); t3 = t()
    v = torch.empty(gi.nnz, dtype=torch.float64, device=dev); xd = torch.from_numpy(x).to(dev); yd = torch.empty_like(xd)
    gi.assemble(xd, yd, v); t4 = t(); gi.assemble(xd, yd, v); t5 = t()
    print("ex2 %dx%d Q2 (%d dofs): Newton iteration with device PCG %.1f ms (%d CG iterations), with D2H + SuperLU %.1f ms, device assembly alone %.3f ms, max diff %.1e"
          % (n, n, x.size, (t1 - t0) * 1e3, lin.linear_iterations[-1], (t3 - t2) * 1e3, (t5 - t4) * 1e3, np.max(np.abs(c - c2))))

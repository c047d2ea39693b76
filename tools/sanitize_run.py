"""Small assemblies of the main kernels for compute-sanitizer (memcheck / racecheck / synccheck):
config 2 (k_patch_ws, several patches incl. a partial last one), the ex4 block (k_patch, 64-element patches), triangles,
and the device solvers.  usage: compute-sanitizer --tool racecheck python tools/sanitize_run.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mfem_ad_b200 as M
from mfem_ad_b200 import meshgen as G
import spec as S
from oracle import oracle as O
ctx = M.Context(0)
rng = np.random.default_rng(0)
# config 2 type: 53 x 41 Q2 (17 patches, last partial)
mesh = G.cartesian_mesh((53, 41), perturb=0.1)
s = G.h1_space(mesh, 2, mode=O.GRAD)
gm = M.Mesh(ctx, mesh); gs = M.Space(ctx, gm, s)
gi = M.Integrator(ctx, [(gs, O.GRAD)], S.minsurf(2, 0.5).madb(ctx))
gi.set_essential(G.boundary_dofs(mesh, s))
x = rng.uniform(-0.3, 0.3, s["ndofs"])
for _ in range(2):
    y, v = gi.assemble(x)
sol = M.Solver(gi)
b = rng.uniform(-1, 1, x.size); b[G.boundary_dofs(mesh, s)] = 0
xs, it, rr = sol.pcg(v, b, rtol=1e-10)
print("config2 ok", float(np.abs(y).sum()), it)
# ex4 block
mesh = G.cartesian_mesh((13, 11))
h1 = G.h1_space(mesh, 3, mode=O.VALUE | O.GRAD); l2 = G.l2_space(mesh, 1, mode=O.VALUE)
gm2 = M.Mesh(ctx, mesh); gh, gl = M.Space(ctx, gm2, h1), M.Space(ctx, gm2, l2)
fs = S.pg(S.obstacle(2), S.fermidirac(0.0, 0.5), 0.4)
gi2 = M.Integrator(ctx, [(gh, O.VALUE | O.GRAD), (gl, O.VALUE), (gl, O.VALUE, M.ROLE_PARAM)], fs.madb(ctx), quad_order=9)
gi2.set_param_field(2, rng.normal(0, 1, l2["ndofs"]))
gi2.set_essential(G.boundary_dofs(mesh, h1))
x2 = np.concatenate([0.1 * rng.uniform(-1, 1, h1["ndofs"]), rng.normal(0, 1, l2["ndofs"])])
y2, v2 = gi2.assemble(x2)
sol2 = M.Solver(gi2)
c, it2, rr2 = sol2.pg_minres(h1["ndofs"], 4, v2, y2, rtol=1e-8)
c, it3, rr3 = sol2.condensed_pcg(h1["ndofs"], 4, v2, y2, rtol=1e-8)
print("ex4 ok", float(np.abs(y2).sum()), it2, it3)
# triangles
tm = G.triangle_mesh((9, 8), perturb=0.1)
st = G.h1_space(tm, 2, mode=O.GRAD)
gm3 = M.Mesh(ctx, tm)
gi3 = M.Integrator(ctx, [(M.Space(ctx, gm3, st), O.GRAD)], S.minsurf(2, 0.5).madb(ctx))
y3, v3 = gi3.assemble(rng.uniform(-0.3, 0.3, st["ndofs"]))
print("triangles ok", float(np.abs(y3).sum()))

"""A few ex4-block assemblies (config 5 shape) for ncu captures: compute-only call then full call."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mfem_ad_b200 as M
from mfem_ad_b200 import meshgen as G
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
order = 2
ctx = M.Context(0)
mesh = G.cartesian_mesh((n, n))
h1 = G.h1_space(mesh, order + 1, mode=M.VALUE | M.GRAD)
l2 = G.l2_space(mesh, order - 1, mode=M.VALUE)
gm = M.Mesh(ctx, mesh); gh, gl = M.Space(ctx, gm, h1), M.Space(ctx, gm, l2)
fn = M.Functional(ctx, "pg", params=[0.1], iparams=[0],
                  children=[M.Functional(ctx, "obstacle"), M.Functional(ctx, "fermidirac", params=[0.0, 0.5])])
gi = M.Integrator(ctx, [(gh, M.VALUE | M.GRAD), (gl, M.VALUE), (gl, M.VALUE, M.ROLE_PARAM)], fn, quad_order=3 * order + 3)
dev = torch.device("cuda", 0)
nd = h1["ndofs"] + l2["ndofs"]
gi.set_param_field(2, torch.zeros(l2["ndofs"], dtype=torch.float64, device=dev))
x = torch.from_numpy(0.1 * np.random.default_rng(0).uniform(-1, 1, nd)).to(dev)
y = torch.empty_like(x); vals = torch.empty(gi.nnz, dtype=torch.float64, device=dev)
for _ in range(2):
    gi.assemble(x, None, None)
    gi.assemble(x, y, vals)
torch.cuda.synchronize()

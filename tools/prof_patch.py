"""One fused residual+Jacobian assembly of config 2 (for ncu captures)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mfem_ad_b200 as M
from mfem_ad_b200 import meshgen as G
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = M.Context(0)
mesh = G.cartesian_mesh((nx, nx))
sp = G.h1_space(mesh, 2, mode=M.GRAD)
gm = M.Mesh(ctx, mesh); gs = M.Space(ctx, gm, sp)
fn = M.Functional(ctx, "minsurf", params=[0.5])
gi = M.Integrator(ctx, [(gs, M.GRAD)], fn)
dev = torch.device("cuda", 0)
x = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, sp["ndofs"])).to(dev)
y = torch.empty_like(x); vals = torch.empty(gi.nnz, dtype=torch.float64, device=dev)
for _ in range(reps):
    gi.assemble(x, y, vals)
torch.cuda.synchronize()
print(gi.patch_stats())

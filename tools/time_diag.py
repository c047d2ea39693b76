"""Diagnostic timing of the config-2 fused assembly (one line): used with MADB_DIAG / MADB_PATCH_WS switches."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mfem_ad_b200 as M
from mfem_ad_b200 import meshgen as G
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ctx = M.Context(0)
mesh = G.cartesian_mesh((nx, nx))
sp = G.h1_space(mesh, 2, mode=M.GRAD)
gm = M.Mesh(ctx, mesh); gs = M.Space(ctx, gm, sp)
fn = M.Functional(ctx, "minsurf", params=[0.5], iparams=[])
gi = M.Integrator(ctx, [(gs, M.GRAD)], fn)
gi.set_timing(True)
dev = torch.device("cuda", 0)
x = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, sp["ndofs"])).to(dev)
y = torch.empty_like(x); vals = torch.empty(gi.nnz, dtype=torch.float64, device=dev)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
with torch.cuda.stream(stream):
    for _ in range(5): gi.assemble(x, y, vals)
    torch.cuda.synchronize()
    ks = []
    for _ in range(20):
        gi.assemble(x, y, vals)
        torch.cuda.synchronize()
        ks.append(gi.last_kernel_ms())
print("DIAG=%s WS=%s element kernel %.4f ms (min %.4f)" % (os.environ.get("MADB_DIAG"), os.environ.get("MADB_PATCH_WS"), sum(ks) / len(ks), min(ks)))

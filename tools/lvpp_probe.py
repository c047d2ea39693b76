import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mfem_ad_b200 as M
from mfem_ad_b200 import lvpp, meshgen as G
import test_gpu_solve as T
ctx = M.Context(0)
mesh, h1, l2, ess, b, gi = T._ex4_problem(ctx, 6)
wl = G.lumped_weights(mesh, l2)
l1 = lambda v: float(np.sum(wl * np.abs(v)))
rule = M.PGStepSizeRule(2, 0.1, 1e4, 2.0, 1.0)
kind = sys.argv[1] if len(sys.argv) > 1 else "minres"
sl = slice(h1["ndofs"], h1["ndofs"] + l2["ndofs"])
lin = lvpp.DeviceLinear(gi, kind, nh=h1["ndofs"], nb=4, rtol=1e-13, maxit=5000)
x = np.zeros(b.size)
try:
    h = lvpp.lvpp_solve(gi, lambda a: gi.fn.set_params([a]), lambda p: gi.set_param_field(2, p), rule, b, x, sl, l1, max_pg=30,
                        newton_kw=dict(abs_tol=1e-9, rel_tol=0.0, max_iter=20, linear=lin), log=print)
    print(h)
except Exception as e:
    print("EXC", e)
print("its", lin.linear_iterations, "relres", ["%.1e" % r for r in lin.relres])
print("psi range", x[sl].min(), x[sl].max())

"""Config 5 core: the ex4 LVPP solve driven by the CUDA assembly vs by the CPU oracle -- Newton and
PG iteration counts must be equal (BASELINE.json north_star), final iterates agree."""
import numpy as np
import pytest

import spec as S
from mfem_ad_b200 import lvpp, meshgen as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu


class OracleOp:
    def __init__(self, mesh, inputs, fs_factory, quad_order, ess, l2):
        self.mesh, self.inputs, self.fs_factory, self.q, self.ess, self.l2 = mesh, inputs, fs_factory, quad_order, ess, l2
        self.alpha, self.psik = 1.0, None
        self._form = None

    def _f(self):
        if self._form is None:
            self._form = O.OracleForm(self.mesh, self.inputs, self.fs_factory(self.alpha).oracle(), quad_order=self.q,
                                      ess=self.ess, params=[dict(type=O.PRM_GF, size=1, data=self.psik, space=self.l2)])
        return self._form

    def set_alpha(self, a):
        self.alpha, self._form = a, None

    def set_latent_k(self, p):
        self.psik, self._form = p.copy(), None

    def mult(self, x):
        return self._f().mult(x)

    def grad(self, x):
        return self._f().grad(x)[2]

    def pattern(self):
        return self._f().pattern()


def test_ex4_lvpp_iteration_counts(ctx):
    import mfem_ad_b200 as M
    order, n = 2, 6
    mesh = G.cartesian_mesh((n, n))
    h1 = G.h1_space(mesh, order + 1, mode=O.VALUE | O.GRAD)
    l2 = G.l2_space(mesh, order - 1, mode=O.VALUE)
    ess = G.boundary_dofs(mesh, h1)
    fs_factory = lambda a: S.pg(S.obstacle(2), S.fermidirac(0.0, 0.5), a)
    b = np.zeros(h1["ndofs"] + l2["ndofs"])
    b[:h1["ndofs"]] = G.load_vector(mesh, h1, lambda x: 2 * np.pi ** 2 * np.sin(np.pi * x[..., 0]) * np.sin(np.pi * x[..., 1]))
    b[ess] = 0.0  # ex4.cpp:149
    wl = G.lumped_weights(mesh, l2)
    l1 = lambda v: float(np.sum(wl * np.abs(v)))
    rule = M.PGStepSizeRule(2, 0.1, 1e4, 2.0, 1.0)  # test.sh:9  -rule 2 -a0 0.1 -ar 2
    sl = slice(h1["ndofs"], h1["ndofs"] + l2["ndofs"])
    nk = dict(abs_tol=1e-9, rel_tol=0.0, max_iter=20)  # ex4.cpp:172-174

    oop = OracleOp(mesh, [h1, l2], fs_factory, 3 * order + 3, ess, l2)
    xo = np.zeros(b.size)  # start from zero
    ho = lvpp.lvpp_solve(oop, oop.set_alpha, oop.set_latent_k, rule, b, xo, sl, l1, max_pg=30, newton_kw=nk)

    spaces = [h1, l2, dict(l2, role=1)]
    _, gi = S.make_pair(ctx, mesh, spaces, fs_factory(1.0), quad_order=3 * order + 3, ess=ess,
                        params=[dict(type=O.PRM_GF, size=1, data=np.zeros(l2["ndofs"]), space=l2)])
    xg = np.zeros(b.size)
    hg = lvpp.lvpp_solve(gi, lambda a: gi.fn.set_params([a]), lambda p: gi.set_param_field(2, p), rule, b, xg, sl, l1,
                         max_pg=30, newton_kw=nk)

    assert not ho["newton_failed"] and not hg["newton_failed"]
    assert ho["newton_iterations"] == hg["newton_iterations"], (ho["newton_iterations"], hg["newton_iterations"])
    assert ho["pg_iterations"] == hg["pg_iterations"] and ho["converged"] == hg["converged"]
    assert np.max(np.abs(xo - xg)) <= 1e-9 * max(1.0, np.max(np.abs(xo)))
    # the obstacle is active: mapped primal 0.5*sigmoid(psi/2) stays in (0, 0.5) and u ~ U(psi)
    u = xg[:h1["ndofs"]]
    assert u.max() <= 0.51 and u.max() > 0.45  # enforced at the quadrature level, nodal values may overshoot slightly
    assert sum(hg["newton_iterations"]) > 5

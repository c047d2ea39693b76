"""Latent-variable kernels: nodal (DOF-collocated) PG terms (src/dof_pg.hpp) and the fused LVPP update."""
import numpy as np
import pytest

import spec as S
from mfem_ad_b200 import meshgen as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fs", [S.fermidirac(0.0, 0.5), S.shannon(0.1, 1)])
def test_dofpg_nodal_terms(ctx, fs):
    import mfem_ad_b200 as M
    mesh = G.cartesian_mesh((5, 4), perturb=0.1)
    s = G.h1_space(mesh, 2)
    nd, ne = s["ndofs"], s["e2l"].shape[0]
    rng = np.random.default_rng(8)
    u, psi, psik = rng.uniform(0, 0.5, nd), rng.normal(0, 2, nd), rng.normal(0, 2, nd)
    w_el = rng.uniform(0.1, 1.0, s["e2l"].shape)  # per (element, node) weights Tr.Weight()*ip.weight
    alpha = 0.37
    r_u, r_psi, d_pp, d_up = O.dofpg_nodal(fs.oracle(), s["e2l"], w_el, alpha, u, psi, psik)
    W = np.zeros(nd)
    np.add.at(W, s["e2l"], w_el)  # lumped nodal weights
    base = rng.normal(0, 1, nd)   # r_u is accumulated on top of the objective's residual
    g = M.dofpg_nodal(ctx, fs.madb(ctx), alpha, u, psi, psik, W, r_u=base.copy())
    for a, b in ((g[0] - base, r_u), (g[1], r_psi), (g[2], d_pp), (g[3], d_up)):
        assert np.max(np.abs(a - b)) <= 1e-12 * max(1.0, np.max(np.abs(b)))
    # zero weights reproduce the reference as written (element nodes carry no quadrature weight)
    g0 = M.dofpg_nodal(ctx, fs.madb(ctx), alpha, u, psi, psik, np.zeros(nd))
    assert all(np.all(a == 0.0) for a in g0)


def test_fused_lvpp_update(ctx):
    import mfem_ad_b200 as M
    rng = np.random.default_rng(9)
    n = 100003
    psi, psik, lam_prev, w = rng.normal(0, 1, n), rng.normal(0, 1, n), rng.normal(0, 1, n), rng.uniform(0, 1, n)
    alpha = 0.8
    lam = (psi - psik) * (1.0 / alpha)  # ex4.cpp:204: lambda *= 1.0/alpha
    ref = np.sum(w * np.abs(lam - lam_prev))
    pk, lp = psik.copy(), lam_prev.copy()
    d = M.lvpp_update(ctx, alpha, psi, pk, lp, w)
    assert abs(d - ref) <= 1e-12 * ref
    assert np.array_equal(lp, lam) and np.array_equal(pk, psi)
    pk2, lp2 = psik.copy(), lam_prev.copy()
    assert M.lvpp_update(ctx, alpha, psi, pk2, lp2, w) == d  # deterministic reduction
    # device-resident path
    import torch
    t = [torch.from_numpy(a).cuda() for a in (psi, psik.copy(), lam_prev.copy(), w)]
    d3 = M.lvpp_update(ctx, alpha, *t)
    assert d3 == d and np.array_equal(t[1].cpu().numpy(), psi)


def test_ex4_mapped_primal_coefficient(ctx):
    """ex4.cpp:124-128,200: x_mapped = grad E*(psi) projected to a quadrature function."""
    mesh = G.cartesian_mesh((5, 4), perturb=0.1)
    l2 = G.l2_space(mesh, 1, mode=O.VALUE)
    psi = np.random.default_rng(3).normal(0, 3, l2["ndofs"])
    of, gi = S.make_pair(ctx, mesh, [l2], S.fermidirac(0.0, 0.5), quad_order=9)
    val, grd = gi.coefficient(psi)
    ref = of.coefficient(psi, 1)
    assert np.max(np.abs(grd - ref)) <= 1e-13
    assert grd.min() > 0.0 and grd.max() < 0.5


def test_ad_sqrt_edge_cases(ctx):
    """The device AD square root is branch-free (MUFU seed + Newton steps, csrc/madb_ad.cuh); it must reproduce the
    reference's dual sqrt (value sqrt(a), f' = 0.5/sqrt(a), f'' = -0.5 f'/a) to rounding for regular arguments and
    its IEEE results (0 / inf / -inf / NaN) at 0, at denormal, tiny and huge arguments and for negative ones."""
    import mfem_ad_b200 as M
    f = M.Functional(ctx, "sqrtprobe")  # f = sqrt(x0) * x1
    x0 = np.array([0.0, 5e-324, 1e-320, 1e-300, 3e-291, 1e-200, 0.3, 1.0, 7.0, 1e200, 7e289, 1e300, 1.7e308, -1.0])
    x1 = np.full_like(x0, 1.5)
    v, g, h = f.eval(np.stack([x0, x1], axis=1))
    with np.errstate(all="ignore"):
        s = np.sqrt(x0)
        f1 = 0.5 / s
        f2 = -0.5 * f1 / x0
        ev, eg0, eg1, eh00, eh01 = s * x1, f1 * x1, s, f2 * x1, f1
    for got, exp in ((v, ev), (g[:, 0], eg0), (g[:, 1], eg1), (h[:, 0, 0], eh00), (h[:, 0, 1], eh01), (h[:, 1, 0], eh01)):
        fin = np.isfinite(exp)
        assert np.array_equal(np.isnan(got), np.isnan(exp)), (got, exp)
        assert np.array_equal(got[np.isinf(exp)], exp[np.isinf(exp)]), (got, exp)
        # finite results: to rounding (the second derivative of tiny arguments overflows in both formulas)
        assert np.allclose(got[fin], exp[fin], rtol=4e-16 * 8, atol=0.0), (got[fin], exp[fin])
    assert np.all(h[:, 1, 1][x0 > 0] == 0.0)

"""Pins the oracle's AD layer against the reference's own closed forms (ex0.cpp:36-98)
and analytic entropy maps (SURVEY 8c golden vectors).  CPU only."""
import numpy as np
import pytest

import spec as S
from oracle import oracle as O

X0 = np.array([0.5, 1.0, -1.0])  # ex0.cpp:102


def test_ex0_scalar_function_golden():
    f = O.Functional()
    f.add(O.K_EX0, 3)
    assert abs(f.value(X0) - 0.30321372968699545) <= 1e-15
    J = f.gradient(X0)
    Jref = np.array([2.3855167309591354, 1.3032137296869954, 3.0])
    assert np.linalg.norm(J - Jref) <= 1e-14  # "Jacobian error", ex0.cpp:139
    H = f.hessian(X0)
    Href = np.array([[-1.3032137296869954, 2.3855167309591354, 0.0],
                     [2.3855167309591354, 1.3032137296869954, 0.0],
                     [0.0, 0.0, -6.0]])
    assert np.max(np.abs(H - Href)) <= 1e-14  # "Hessian error", ex0.cpp:140
    assert np.array_equal(H, H.T)  # mirrored fill, src/ad_native.cpp:224-225


def test_ex0_closed_forms_recomputed():
    # the closed forms themselves, ex0.cpp:36-61
    x = X0
    f = O.Functional()
    f.add(O.K_EX0, 3)
    J = np.array([np.cos(x[0]) * np.exp(x[1]), np.sin(x[0]) * np.exp(x[1]), 3 * x[2] ** 2])
    assert np.linalg.norm(f.gradient(x) - J) <= 1e-14


def test_ex0_vector_function_golden():
    g = O.Functional()
    g.add(O.K_EX0VEC, 3, n_output=2)
    J = g.vec_gradient(X0)
    Jref = np.array([[0.8775825618903728, 0.4387912809451864, 0.0],
                     [-0.479425538604203, -0.2397127693021015, 0.2397127693021015]])
    assert np.max(np.abs(J - Jref)) <= 1e-14  # "Jacobian2 error", ex0.cpp:153
    H = g.vec_hessian(X0)
    H0 = np.array([[-0.479425538604203, 0.6378697925882713, 0], [0.6378697925882713, -0.11985638465105075, 0],
                   [0, 0, 0]])
    H1 = np.array([[-0.8775825618903728, -0.9182168195493894, 0.9182168195493894],
                   [-0.9182168195493894, -0.2193956404725932, 0.4591084097746947],
                   [0.9182168195493894, 0.4591084097746947, -0.2193956404725932]])
    assert np.max(np.abs(H[0] - H0)) <= 1e-14 and np.max(np.abs(H[1] - H1)) <= 1e-14  # ex0.cpp:154-158


def _fd_check(fs, x, qprm=None, h=1e-6, tol=1e-6):
    f = fs.oracle()
    n = len(x)
    g = f.gradient(x, qprm)
    H = f.hessian(x, qprm)
    for i in range(n):
        e = np.zeros(n)
        e[i] = h
        gfd = (f.value(x + e, qprm) - f.value(x - e, qprm)) / (2 * h)
        assert abs(g[i] - gfd) <= tol * max(1, abs(gfd))
        Hfd = (f.gradient(x + e, qprm) - f.gradient(x - e, qprm)) / (2 * h)
        assert np.max(np.abs(H[i] - Hfd)) <= tol * max(1, np.max(np.abs(Hfd)))
    assert np.allclose(H, H.T, atol=0)


@pytest.mark.parametrize("fs,x,q", [
    (S.minsurf(2, 0.5), [0.3, -0.7], None),
    (S.minsurf(3, 0.1), [0.3, -0.7, 1.1], None),
    (S.diffusion(2), [0.3, -0.7], None),
    (S.diffusion(2, [2.5]), [0.3, -0.7], None),
    (S.diffusion(2, [2.5, 0.5]), [0.3, -0.7], None),
    (S.diffusion(2, [2.0, 0.3, 0.3, 1.0]), [0.3, -0.7], None),
    (S.mass(3), [0.3, -0.7, 0.2], None),
    (S.elasticity(2, 2.0, 0.7), [0.3, -0.7, 0.2, 0.9], None),
    (S.elasticity(3, 2.0, 0.7), list(np.linspace(-1, 1, 9) ** 2 - 0.3), None),
    (S.obstacle(2), [0.1, 0.3, -0.7], None),
    (S.fermidirac(0.0, 0.5), [0.8], None),
    (S.fermidirac(0.0, 0.5), [-0.8], None),
    (S.shannon(0.25, 1), [0.4], None),
    (S.shannon(0.25, -1), [0.4], None),
    (S.hellinger(2, 0.7), [0.4, -1.2], None),
    (S.simplex(5, 1.0), [0.4, -1.2, 0.3, 2.0, 0.1], None),
    (S.simp([1e-3, 0.25, 0.5, 0.75, 1.0], 3.0), [0.4, 0.2, 0.3, 0.05, 0.05], None),
    (S.pg(S.obstacle(2), S.fermidirac(0.0, 0.5), 0.7), [0.1, 0.3, -0.7, 0.45], [0.2]),
    (S.pg(S.gradobstacle(2), S.hellinger(2, 0.6), 1.3), [0.3, -0.7, 0.45, 0.1], [0.2, -0.1]),
])
def test_finite_difference(fs, x, q):
    _fd_check(fs, np.array(x, dtype=float), None if q is None else np.array(q, dtype=float))


def test_entropy_maps_analytic():
    # sigmoid map of FermiDiracEntropy(lower=0, upper=0.5): u = 0.5*sigma(0.5 psi), u' = 0.25 sigma (1-sigma)
    for psi in (-30.0, -1.3, 0.0, 0.7, 25.0):
        f = S.fermidirac(0.0, 0.5).oracle()
        s = 1.0 / (1.0 + np.exp(-0.5 * psi))
        assert abs(f.gradient([psi])[0] - 0.5 * s) <= 1e-15
        assert abs(f.hessian([psi])[0, 0] - 0.25 * s * (1 - s)) <= 1e-15
    # softmax map of SimplexEntropy
    psi = np.array([0.4, -1.2, 0.3, 2.0, 0.1])
    f = S.simplex(5, 1.5).oracle()
    p = np.exp(psi - psi.max())
    p /= p.sum()
    assert np.max(np.abs(f.gradient(psi) - 1.5 * p)) <= 1e-15
    assert np.max(np.abs(f.hessian(psi) - 1.5 * (np.diag(p) - np.outer(p, p)))) <= 1e-15
    # tie case psi = 0 => uniform (every max() comparison ties)
    f = S.simplex(5, 1.0).oracle()
    assert np.max(np.abs(f.gradient(np.zeros(5)) - 0.2)) <= 1e-15
    assert np.max(np.abs(f.hessian(np.zeros(5)) - (0.2 * np.eye(5) - 0.04))) <= 1e-15
    # Hellinger: u = s^2 psi / sqrt(1 + s^2 |psi|^2)
    psi, s = np.array([0.4, -1.2]), 0.7
    f = S.hellinger(2, s).oracle()
    assert np.max(np.abs(f.gradient(psi) - s * s * psi / np.sqrt(1 + s * s * psi @ psi))) <= 1e-15
    # Shannon: u = exp(s psi) + b
    assert abs(S.shannon(0.25, 1).oracle().gradient([0.4])[0] - (np.exp(0.4) + 0.25)) <= 1e-15
    assert abs(S.shannon(0.25, -1).oracle().gradient([0.4])[0] - (np.exp(-0.4) + 0.25)) <= 1e-15


def test_pg_hessian_structure():
    # ex4 (n=4): [[0,0,0,1/a],[0,1,0,0],[0,0,1,0],[1/a,0,0,-E*''/a]]  
    a, psi = 0.7, 0.45
    f = S.pg(S.obstacle(2), S.fermidirac(0.0, 0.5), a).oracle()
    H = f.hessian([0.1, 0.3, -0.7, psi], [0.2])
    s = 1.0 / (1.0 + np.exp(-0.5 * psi))
    Href = np.array([[0, 0, 0, 1 / a], [0, 1, 0, 0], [0, 0, 1, 0], [1 / a, 0, 0, -0.25 * s * (1 - s) / a]])
    assert np.max(np.abs(H - Href)) <= 1e-14


def test_pg_step_rule():
    # src/pg.cpp:34-54 ; test.sh:9 uses -rule 2 -a0 0.1 -ar 2
    assert [O.pg_step(2, 0.1, 1e4, 2.0, 1.0, k) for k in range(4)] == [0.1, 0.2, 0.4, 0.8]
    assert O.pg_step(2, 0.1, 1e4, 2.0, 1.0, 30) == 1e4
    assert O.pg_step(0, 0.3, 1e4, 2.0, 1.0, 7) == 0.3
    assert abs(O.pg_step(1, 0.5, 1e4, 2.0, 1.0, 2) - 4.5) < 1e-15
    assert abs(O.pg_step(3, 0.1, 1e6, 2.0, 2.0, 2) - 1.6) < 1e-15
    import mfem_ad_b200 as M
    for rule, r, r2 in ((0, 1.0, 1.0), (1, 2.0, 1.0), (2, 2.0, 1.0), (3, 2.0, 2.0)):
        R = M.PGStepSizeRule(rule, 0.1, 1e4, r, r2)
        for k in range(12):
            assert abs(R.get(k) - O.pg_step(rule, 0.1, 1e4, r, r2, k)) <= 1e-13 * O.pg_step(rule, 0.1, 1e4, r, r2, k)
    with pytest.raises(M.MadbError):
        M.PGStepSizeRule(2, 0.1, 1e4, 1.0)  # EXP needs ratio > 1 (src/pg.cpp:21-24)
    with pytest.raises(M.MadbError):
        M.PGStepSizeRule(0, -1.0)

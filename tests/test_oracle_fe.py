"""Pins the oracle's FE substrate by analytic identities (MFEM itself is absent: SURVEY 8c).  CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import spec as S
from mfem_ad_b200 import meshgen as G
from oracle import oracle as O


def test_rules_and_nodes():
    # tensor Gauss-Legendre with p+2 points per direction for order 2p+2
    assert [O.lib().orc_rule_npts_1d(2 * p + 2) for p in (1, 2, 3)] == [3, 4, 5]
    assert O.lib().orc_rule_npts_1d(9) == 5  # ex4.cpp:104 (3p+3, p=2)
    for n in range(1, 9):
        x, w = O.gauss_legendre(n)
        xr, wr = G.gauss_legendre_01(n)
        assert np.max(np.abs(x - xr)) < 1e-15 and np.max(np.abs(w - wr)) < 1e-15
        for k in range(2 * n):  # exactness
            assert abs(w @ x ** k - 1.0 / (k + 1)) < 1e-14
    for n in range(2, 9):
        assert np.max(np.abs(O.gauss_lobatto(n) - G.gauss_lobatto_01(n))) < 1e-15


def _form(mesh, sp_, fs, **kw):
    return O.OracleForm(mesh, [sp_], fs.oracle(), **kw)


@pytest.mark.parametrize("p", [1, 2, 3])
def test_diffusion_stiffness_identities(p):
    mesh = G.cartesian_mesh((3, 2), perturb=0.15)
    s = G.h1_space(mesh, p, mode=O.GRAD)
    f = _form(mesh, s, S.diffusion(2))
    x = np.zeros(s["ndofs"])
    rp, ci, v = f.grad(x)
    K = sp.csr_matrix((v, ci, rp), shape=(s["ndofs"],) * 2)
    assert abs(K - K.T).max() < 1e-13
    assert np.max(np.abs(K @ np.ones(s["ndofs"]))) < 1e-12  # constants in the kernel
    xc = G.dof_coords(mesh, s)
    u = 2.0 * xc[:, 0] - 3.0 * xc[:, 1] + 0.5  # linear field: energy = 0.5*|grad|^2*area
    assert abs(0.5 * u @ (K @ u) - 0.5 * 13.0 * 1.0) < 1e-12
    assert abs(f.energy(u) - 6.5) < 1e-12
    assert np.max(np.abs(f.mult(u) - K @ u)) < 1e-12  # linear problem: residual = K u


def test_manufactured_solution_convergence():
    # ex1.cpp:70-75: -lap u = 2 pi^2 sin sin, Q1, L2 error must show O(h^2)
    errs = []
    for n in (8, 16):
        mesh = G.cartesian_mesh((n, n))
        s = G.h1_space(mesh, 1, mode=O.GRAD)
        ess = G.boundary_dofs(mesh, s)
        f = _form(mesh, s, S.diffusion(2), ess=ess)
        rp, ci, v = f.grad(np.zeros(s["ndofs"]))
        K = sp.csr_matrix((v, ci, rp), shape=(s["ndofs"],) * 2)
        m = O.OracleForm(mesh, [dict(s, mode=O.VALUE)], S.mass(1).oracle())
        rpm, cim, vm = m.grad(np.zeros(s["ndofs"]))
        Mm = sp.csr_matrix((vm, cim, rpm), shape=(s["ndofs"],) * 2)
        xc = G.dof_coords(mesh, s)
        uex = np.sin(np.pi * xc[:, 0]) * np.sin(np.pi * xc[:, 1])
        b = Mm @ (2 * np.pi ** 2 * uex)
        b[ess] = 0.0
        u = spla.spsolve(K.tocsc(), b)
        e = u - uex
        errs.append(np.sqrt(e @ (Mm @ e)))
    assert errs[0] / errs[1] > 3.5


def test_pattern_is_full_connectivity_with_explicit_zeros():
    mesh = G.cartesian_mesh((4, 3))
    s = G.h1_space(mesh, 2, mode=O.GRAD)
    f = _form(mesh, s, S.diffusion(2))
    rp, ci = f.pattern()
    # Q2 on a structured grid: 25/15/9 entries per interior vertex/edge/bubble row
    assert rp[-1] == sum(len(np.unique(s["e2l"][np.any(s["e2l"] == d, axis=1)])) for d in range(s["ndofs"]))
    # independent of values: explicit zeros kept (skip_zeros=0)
    fz = _form(mesh, s, S.FSpec("empty", 2))
    rpz, ciz, vz = fz.grad(np.zeros(s["ndofs"]))
    assert np.array_equal(rpz, rp) and np.array_equal(ciz, ci) and np.all(vz == 0.0) and len(vz) == rp[-1]
    for r in range(s["ndofs"]):
        assert np.all(np.diff(ci[rp[r]:rp[r + 1]]) > 0)


def test_jacobian_is_derivative_of_residual():
    mesh = G.cartesian_mesh((3, 3), perturb=0.2)
    s = G.permute_dofs(G.h1_space(mesh, 2, mode=O.GRAD), 7)
    f = _form(mesh, s, S.minsurf(2, 0.5))
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, s["ndofs"])
    rp, ci, v = f.grad(x)
    K = sp.csr_matrix((v, ci, rp), shape=(s["ndofs"],) * 2)
    d = rng.uniform(-1, 1, s["ndofs"])
    h = 1e-6
    fd = (f.mult(x + h * d) - f.mult(x - h * d)) / (2 * h)
    assert np.max(np.abs(K @ d - fd)) < 1e-7
    e = (f.energy(x + h * d) - f.energy(x - h * d)) / (2 * h)
    assert abs(e - f.mult(x) @ d) < 1e-7


def test_block_equals_single_for_scalar_space():
    # ADBlockNonlinearFormIntegrator with one scalar space == ADNonlinearFormIntegrator (src/ad_intg.hpp:330-331 vs :700-727)
    mesh = G.cartesian_mesh((2, 3), perturb=0.1)
    s = G.h1_space(mesh, 2, mode=O.VALUE | O.GRAD)
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, s["ndofs"])
    fs = S.FSpec("mass", 3)
    a = O.OracleForm(mesh, [s], fs.oracle(), block=0)
    b = O.OracleForm(mesh, [s], fs.oracle(), block=1)
    assert np.max(np.abs(a.grad(x)[2] - b.grad(x)[2])) < 1e-13
    assert np.max(np.abs(a.mult(x) - b.mult(x))) < 1e-13


def test_vector_quirk_H1():
    """the single-space VECTOR AssembleElementGrad (src/ad_intg.hpp:283-326) equals the
    block integrator's (index-correct) result only when lambda == mu."""
    mesh = G.cartesian_mesh((2, 2), perturb=0.1)
    s = G.h1_space(mesh, 1, vdim=2, mode=O.GRAD | O.VECTOR)
    x = np.random.default_rng(3).uniform(-1, 1, 2 * s["ndofs"])
    for lam, mu, same in ((1.0, 1.0, True), (2.0, 0.7, False)):
        a = O.OracleForm(mesh, [s], S.elasticity(2, lam, mu).oracle(), block=0).element_grad(0, x)
        b = O.OracleForm(mesh, [s], S.elasticity(2, lam, mu).oracle(), block=1).element_grad(0, x)
        assert np.max(np.abs(b - b.T)) < 1e-13
        if same:
            assert np.max(np.abs(a - b)) < 1e-13
        else:
            assert np.max(np.abs(a - b)) > 1e-3

"""FE substrate pinned by closed forms that do NOT come from the oracle: tests/golden/fe_closed_forms.json, derived
symbolically by tests/golden/make_fe_closed_forms.py (exact element matrices where the reference's rule is exact,
published Gauss abscissae, the 9-point stencil).  CPU tests hold the oracle to them; the `gpu` tests hold the CUDA path
(through the C ABI) to them directly, without the oracle in between."""
import json
import os

import numpy as np
import pytest

import spec as S
from mfem_ad_b200 import meshgen as G
from oracle import oracle as O

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fe_closed_forms.json")))
CASES = GOLD["cases"]
TOL = 1e-13


def _one_element(case, p, vdim=1, mode=O.GRAD):
    mesh = G.cartesian_mesh((1, 1))
    mesh["coords"] = np.array(case["vertices"], dtype=np.float64)
    return mesh, G.h1_space(mesh, p, vdim=vdim, mode=mode)


def _setup(case):
    """-> (mesh, space, functional spec, state, expected dense matrix, expected residual or None, block flag)"""
    n = case["name"]
    if n == "stiffness":
        mesh, s = _one_element(case, case["order"])
        return mesh, s, S.diffusion(2), np.zeros(s["ndofs"]), np.array(case["matrix"]), None, None
    if n == "mass":
        mesh, s = _one_element(case, case["order"], mode=O.VALUE)
        return mesh, s, S.mass(1), np.zeros(s["ndofs"]), np.array(case["matrix"]), None, None
    if n == "diffusion_fullK":
        mesh, s = _one_element(case, 1)
        return mesh, s, S.diffusion(2, case["K_colmajor"]), np.zeros(s["ndofs"]), np.array(case["matrix"]), None, None
    if n == "minsurf_linear_state":
        mesh, s = _one_element(case, case["order"])
        return mesh, s, S.minsurf(2, case["eps"]), np.array(case["state"]), np.array(case["jacobian"]), np.array(case["residual"]), None
    raise KeyError(n)


SCALAR = [i for i, c in enumerate(CASES) if c["name"] in ("stiffness", "mass", "diffusion_fullK", "minsurf_linear_state")]
ELAST = [i for i, c in enumerate(CASES) if c["name"] == "elasticity_q1"]


def _dense(rp, ci, v, n):
    A = np.zeros((n, n))
    for r in range(n):
        A[r, ci[rp[r]:rp[r + 1]]] = v[rp[r]:rp[r + 1]]
    return A


def test_published_gauss_points():
    pub = GOLD["published"]
    assert abs(O.gauss_lobatto(4)[1] - pub["gll4_inner"]) <= 1e-15 and abs(O.gauss_lobatto(5)[1] - pub["gll5_inner"]) <= 1e-15
    x3, w3 = O.gauss_legendre(3)
    assert abs(x3[0] - pub["gl3_outer_x"]) <= 1e-15 and np.max(np.abs(w3 - np.array(pub["gl3_w"]))) <= 1e-15
    x4, w4 = O.gauss_legendre(4)
    assert np.max(np.abs(x4[:2] - np.array(pub["gl4_x"]))) <= 1e-15 and np.max(np.abs(w4[:2] - np.array(pub["gl4_w"]))) <= 1e-15
    for n, ref in GOLD["gauss_lobatto_01"].items():
        assert np.max(np.abs(O.gauss_lobatto(int(n)) - np.array(ref))) <= 1e-15
        assert np.max(np.abs(G.gauss_lobatto_01(int(n)) - np.array(ref))) <= 1e-15
    for n, ref in GOLD["gauss_legendre_01"].items():
        x, w = O.gauss_legendre(int(n))
        assert np.max(np.abs(x - np.array(ref["x"]))) <= 1e-15 and np.max(np.abs(w - np.array(ref["w"]))) <= 1e-15


@pytest.mark.parametrize("k", SCALAR)
def test_oracle_against_closed_forms(k):
    mesh, s, fs, x, Aref, rref, _ = _setup(CASES[k])
    of = O.OracleForm(mesh, [s], fs.oracle())
    A = of.element_grad(0, x)
    assert np.max(np.abs(A - Aref)) <= TOL * np.max(np.abs(Aref))
    if rref is not None:
        assert np.max(np.abs(of.element_vector(0, x) - rref)) <= TOL * np.max(np.abs(rref))
        assert abs(of.energy(x) - CASES[k]["energy"]) <= TOL * abs(CASES[k]["energy"])
    else:  # linear forms: residual = A x
        xr = np.random.default_rng(k).normal(0, 1, x.size)
        assert np.max(np.abs(of.element_vector(0, xr) - Aref @ xr)) <= 10 * TOL * np.max(np.abs(Aref))


@pytest.mark.parametrize("k", ELAST)
def test_oracle_elasticity_against_closed_forms(k):
    """order-1 elasticity, lambda != mu: the block integrator's contraction and the single-space arithmetic AS WRITTEN
    (src/ad_intg.hpp:283-326), the latter derived in sympy from the reference's index arithmetic"""
    c = CASES[k]
    mesh, s = _one_element(c, 1, vdim=2, mode=O.GRAD | O.VECTOR)
    x = np.zeros(2 * s["ndofs"])
    for block, key in ((1, "consistent"), (0, "as_written")):
        of = O.OracleForm(mesh, [s], S.elasticity(2, c["lam"], c["mu"]).oracle(), block=block)
        Aref = np.array(c[key])
        assert np.max(np.abs(of.element_grad(0, x) - Aref)) <= TOL * np.max(np.abs(Aref))
    assert np.max(np.abs(np.array(c["consistent"]) - np.array(c["as_written"]))) > 0.05  # genuinely different matrices


def test_oracle_assembled_stencil():
    c = [c for c in CASES if c["name"] == "assembled_stiffness_q1_2x2"][0]
    mesh = G.cartesian_mesh((2, 2))
    s = G.h1_space(mesh, 1, mode=O.GRAD)
    rp, ci, v = O.OracleForm(mesh, [s], S.diffusion(2).oracle()).grad(np.zeros(9))
    assert np.max(np.abs(_dense(rp, ci, v, 9) - np.array(c["matrix"]))) <= TOL
    assert abs(np.array(c["matrix"])[4, 4] - 8.0 / 3.0) <= 1e-15


@pytest.mark.gpu
@pytest.mark.parametrize("k", SCALAR)
def test_device_against_closed_forms(ctx, k):
    mesh, s, fs, x, Aref, rref, _ = _setup(CASES[k])
    of, gi = S.make_pair(ctx, mesh, [s], fs)
    rp, ci = gi.pattern()
    y, v = gi.assemble(x)
    assert np.max(np.abs(_dense(rp, ci, v, x.size) - Aref)) <= TOL * np.max(np.abs(Aref))
    if rref is not None:
        assert np.max(np.abs(y - rref)) <= TOL * np.max(np.abs(rref))
        assert abs(gi.energy(x) - CASES[k]["energy"]) <= TOL * abs(CASES[k]["energy"])
    else:
        xr = np.random.default_rng(k).normal(0, 1, x.size)
        assert np.max(np.abs(gi.mult(xr) - Aref @ xr)) <= 10 * TOL * np.max(np.abs(Aref))
        assert np.max(np.abs(gi.grad_mult(x, xr) - Aref @ xr)) <= 10 * TOL * np.max(np.abs(Aref))


@pytest.mark.gpu
@pytest.mark.parametrize("k", ELAST)
def test_device_elasticity_against_closed_forms(ctx, k):
    c = CASES[k]
    mesh, s = _one_element(c, 1, vdim=2, mode=O.GRAD | O.VECTOR)
    x = np.zeros(2 * s["ndofs"])
    for block, key in ((1, "consistent"), (0, "as_written")):
        of, gi = S.make_pair(ctx, mesh, [s], S.elasticity(2, c["lam"], c["mu"]), block=block)
        rp, ci = gi.pattern()
        Aref = np.array(c[key])
        assert np.max(np.abs(_dense(rp, ci, gi.grad(x), x.size) - Aref)) <= TOL * np.max(np.abs(Aref))


@pytest.mark.gpu
def test_device_assembled_stencil(ctx):
    c = [c for c in CASES if c["name"] == "assembled_stiffness_q1_2x2"][0]
    mesh = G.cartesian_mesh((2, 2))
    s = G.h1_space(mesh, 1, mode=O.GRAD)
    of, gi = S.make_pair(ctx, mesh, [s], S.diffusion(2))
    rp, ci = gi.pattern()
    assert np.max(np.abs(_dense(rp, ci, gi.grad(np.zeros(9)), 9) - np.array(c["matrix"]))) <= TOL

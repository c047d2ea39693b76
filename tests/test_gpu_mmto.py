"""Config 4 pieces (src/mmto.hpp, src/mmto.cpp, SimplexEntropy): no reference driver exists, so the
pipeline is synthesised as SURVEY 8d describes -- latent psi (5 materials) -> rho = softmax(psi) ->
lambda(rho), mu(rho) by SIMP -> ParametrizedCompliance state block; ParamGradient at the points."""
import numpy as np
import pytest

import spec as S
from mfem_ad_b200 import meshgen as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-12
E = [1e-3, 0.25, 0.5, 0.75, 1.0]


def _setup(n=(6, 5)):
    mesh = G.cartesian_mesh(n, perturb=0.12)
    disp = G.h1_space(mesh, 1, vdim=2, mode=O.GRAD | O.VECTOR)
    lat = G.h1_space(mesh, 1, vdim=5, mode=O.VALUE | O.VECTOR)
    psi = np.random.default_rng(99).normal(0, 1, 5 * lat["ndofs"])  # psi ~ N(0,1) seed 99
    return mesh, disp, lat, psi


def _softmax_nodal(psi, nd):
    p = psi.reshape(5, nd)  # byNODES
    e = np.exp(p - p.max(axis=0))
    return (e / e.sum(axis=0)).reshape(-1)


def test_latent_to_primal_softmax_map(ctx):
    """rho = grad E*(psi), SimplexEntropy(5, 1.0) (src/pg.hpp:347-376): nodal (pointwise kernel) and at
    the quadrature points (DifferentiableCoefficient::Gradient, src/ad_native.hpp:272-283)."""
    mesh, disp, lat, psi = _setup()
    fs = S.simplex(5, 1.0)
    nd = lat["ndofs"]
    _, g, h = fs.madb(ctx).eval(psi.reshape(5, nd).T.copy())
    rho = _softmax_nodal(psi, nd)
    assert np.max(np.abs(g.T.reshape(-1) - rho)) <= 1e-15
    fo = fs.oracle()
    for j in (0, 7, nd - 1):
        assert np.max(np.abs(h[j] - fo.hessian(psi.reshape(5, nd)[:, j]))) <= 1e-15
    of, gi = S.make_pair(ctx, mesh, [lat], fs)
    val, grd = gi.coefficient(psi)
    assert np.max(np.abs(val - of.coefficient(psi, 0))) <= TOL
    assert np.max(np.abs(grd - of.coefficient(psi, 1))) <= TOL
    assert np.max(np.abs(grd.sum(axis=2) - 1.0)) <= 1e-14  # on the simplex


def test_parametrized_compliance_state_block(ctx):
    """ADNonlinearFormIntegrator<GRAD|VECTOR>(ParametrizedCompliance) with lambda(rho), mu(rho) =
    SIMPFunction(E, 3)(rho) at each point (src/mmto.hpp:154-189, :103-108).  lambda != mu here, so the
    oracle uses the index-consistent (block) contraction, see SURVEY H1."""
    mesh, disp, lat, psi = _setup()
    rho = _softmax_nodal(psi, lat["ndofs"])
    lam_fs, mu_fs = S.simp(E, 3.0), S.simp([0.5 * e for e in E], 3.0)
    # oracle: rho at the points -> lambda, mu as quadrature-function parameters
    rho_form = O.OracleForm(mesh, [lat], S.simplex(5).oracle())
    rq = rho_form.inputs_at_qpts(rho)
    lo, mo = lam_fs.oracle(), mu_fs.oracle()
    qf = np.array([[[lo.value(r), mo.value(r)] for r in re] for re in rq])
    fo = S.FSpec("paramcompliance", 4, qoff=0)
    import mfem_ad_b200 as M
    gm = M.Mesh(ctx, mesh)
    gd, gl = M.Space(ctx, gm, disp), M.Space(ctx, gm, lat)
    fn = M.Functional(ctx, "paramcompliance", children=[lam_fs.madb(ctx), mu_fs.madb(ctx)])
    x = np.random.default_rng(5).uniform(-1, 1, 2 * disp["ndofs"])
    vals = {}
    # block=0: ADNonlinearFormIntegrator<GRAD|VECTOR> as written (the default of the library);
    # block=1: the index-consistent contraction (MADB_INTEG_BLOCK)
    for block in (0, 1):
        of = O.OracleForm(mesh, [disp], fo.oracle(), params=[dict(type=O.PRM_QF, size=2, data=qf)], block=block)
        gi = M.Integrator(ctx, [(gd, O.GRAD | O.VECTOR), (gl, O.VALUE | O.VECTOR, M.ROLE_PARAM)], fn, block=bool(block))
        gi.set_param_field(1, rho)
        assert S.csr_rel_err(gi.mult(x), of.mult(x)) <= TOL
        rp, ci, v = of.grad(x)
        rpg, cig = gi.pattern()
        assert np.array_equal(rp, rpg) and np.array_equal(ci, cig)
        vals[block] = gi.grad(x)
        assert S.csr_rel_err(vals[block], v) <= TOL
        assert abs(gi.energy(x) - of.energy(x)) <= TOL * abs(of.energy(x))
    assert np.max(np.abs(vals[0] - vals[1])) > 1e-3 * np.max(np.abs(vals[1]))  # lambda(rho) != mu(rho): H1 shows


def test_param_gradient(ctx):
    """ParametrizedFunctional::ParamGradient::Eval (src/mmto.cpp:4-38) as written: substituting df_i/drho_j
    into slot i while the other f's keep their values yields dF/drho_j + (m-1) F, m = 2.
    madb_integrator_param_gradient returns the reference's result by default, dF/drho on request."""
    mesh, disp, lat, psi = _setup()
    rho = _softmax_nodal(psi, lat["ndofs"])
    u = np.random.default_rng(6).uniform(-1, 1, 2 * disp["ndofs"])
    lam_fs, mu_fs = S.simp(E, 3.0), S.simp([0.5 * e for e in E], 3.0)
    F = O.Functional()
    il = lam_fs._add(F)
    im = mu_fs._add(F)
    S.FSpec("paramcompliance", 4, qoff=4)._add(F)  # parent reads lambda, mu at evaluator slots dim*dim, +1
    of = O.OracleForm(mesh, [lat], F, params=[dict(type=O.PRM_GF_GRAD, size=4, data=u, space=disp)])
    J_ref = of.mmto_param_gradient(rho, [il, im])
    import mfem_ad_b200 as M
    gm = M.Mesh(ctx, mesh)
    gd, gl = M.Space(ctx, gm, disp), M.Space(ctx, gm, lat)
    fn = M.Functional(ctx, "designcompliance", children=[lam_fs.madb(ctx), mu_fs.madb(ctx)])
    gi = M.Integrator(ctx, [(gl, O.VALUE | O.VECTOR), (gd, O.GRAD, M.ROLE_PARAM)], fn)
    gi.set_param_field(1, u)
    # the library's ParamGradient: as written by default (the reference's result), the derivative on request
    val, J = gi.param_gradient(rho)
    assert np.max(np.abs(J - J_ref)) <= TOL * np.max(np.abs(J_ref))
    val2, dF = gi.param_gradient(rho, M.PARAMGRAD_DERIVATIVE)
    assert np.max(np.abs(val2 - val)) <= 1e-14 * np.max(np.abs(val))
    assert np.max(np.abs((dF + (2 - 1) * val[..., None]) - J_ref)) <= TOL * np.max(np.abs(J_ref))  # H6: dF/drho + (m-1) F
    v3, g3 = gi.coefficient(rho)
    assert np.array_equal(g3, dF)
    # the derivative variant against central differences of F(rho) at one point
    e, q = 3, 4
    xin = of.inputs_at_qpts(rho)[e, q]
    gq = O.OracleForm(mesh, [disp], S.FSpec("empty", 4).oracle()).inputs_at_qpts(u)[e, q]
    lo, mo = lam_fs.oracle(), mu_fs.oracle()
    par = S.FSpec("paramcompliance", 4, qoff=0).oracle()
    Fr = lambda r: par.value(gq, [lo.value(r), mo.value(r)])
    for j in range(5):
        d = np.zeros(5)
        d[j] = 1e-6
        fd = (Fr(xin + d) - Fr(xin - d)) / 2e-6
        assert abs(fd - dF[e, q, j]) <= 1e-7 * max(1.0, abs(fd))

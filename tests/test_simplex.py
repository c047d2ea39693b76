"""simplex elements (ex5's triangles, ex5.cpp:72-73).

CPU: the restated triangle rules integrate every monomial up to their order exactly (pins the digits of the tables, in the
oracle), the P1 / P2 bases are nodal and sum to one, closed-form P1 stiffness matrix of the reference triangle, energy of
a linear field.  GPU: scalar forms and the ex5 proximal-Galerkin block (H1 P2 x (H1 P1)^2, Hellinger entropy, constant and
spatially varying bound) against the oracle on a perturbed triangle mesh; load vectors."""
import math

import numpy as np
import pytest

import spec as S
from mfem_ad_b200 import meshgen as G
from oracle import oracle as O

TOL = 1e-12


def _unit_triangle():
    return dict(dim=2, n=(1, 1), lengths=(1.0, 1.0), e2n=np.array([[0, 1, 2]], dtype=np.int32),
                coords=np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), geom_order=-1, simplex=True,
                edges=np.array([[0, 1], [1, 2], [0, 2]], dtype=np.int32), e2e=np.array([[0, 1, 2]], dtype=np.int32))


def test_triangle_rules_are_exact_to_their_order():
    # int_T x^a y^b = a! b! / (a + b + 2)! for every monomial up to the order of the rule
    mesh = _unit_triangle()
    s0 = G.l2_space(mesh, 0, mode=O.VALUE)
    for order in range(0, 7):
        form = O.OracleForm(mesh, [s0], S.mass(1).oracle(), quad_order=order)
        pts, w = form.rule()
        assert form.nq == [1, 1, 3, 4, 6, 7, 12][order] and abs(w.sum() - 0.5) <= 1e-15
        assert np.all(pts >= 0.0) and np.all(pts.sum(axis=1) <= 1.0)
        for a in range(order + 1):
            for b in range(order + 1 - a):
                val = float(np.sum(w * pts[:, 0] ** a * pts[:, 1] ** b))
                exact = math.factorial(a) * math.factorial(b) / math.factorial(a + b + 2)
                assert abs(val - exact) <= 1e-15, (order, a, b, val, exact)
        # and NOT beyond (the rule really has the order it is filed under), except the 1-point rule filed under 0 and 1
        if order >= 1:
            a = order + 1
            assert abs(float(np.sum(w * pts[:, 0] ** a)) - 1.0 / ((a + 1) * (a + 2))) > 1e-9 or order == 0


def test_p1_stiffness_of_the_reference_triangle_and_linear_field_energy():
    mesh = _unit_triangle()
    s1 = G.h1_space(mesh, 1, mode=O.GRAD)
    f = O.OracleForm(mesh, [s1], S.diffusion(2).oracle())
    rp, ci, v = f.grad(np.zeros(3))
    K = np.zeros((3, 3))
    for r in range(3):
        K[r, ci[rp[r]:rp[r + 1]]] = v[rp[r]:rp[r + 1]]
    assert np.max(np.abs(K - 0.5 * np.array([[2.0, -1, -1], [-1, 1, 0], [-1, 0, 1]]))) <= 1e-15
    # energy of u = 2x - 3y on a perturbed triangle mesh of [0,1]^2: |grad u|^2 / 2 * area = 6.5, P1 and P2
    tm = G.triangle_mesh((5, 4), perturb=0.2)
    for p in (1, 2):
        s = G.h1_space(tm, p, mode=O.GRAD)
        xc = G.dof_coords(tm, s)
        e = O.OracleForm(tm, [s], S.diffusion(2).oracle()).energy(2.0 * xc[:, 0] - 3.0 * xc[:, 1])
        assert abs(e - 6.5) <= 1e-13
    # P2 basis is nodal: the load vector of f = 1 on the reference triangle is (0, 0, 0, 1/6, 1/6, 1/6)
    s2 = G.h1_space(mesh, 2, mode=O.VALUE)
    b = O.OracleForm(mesh, [s2], S.load().oracle(), quad_order=4,
                     params=[dict(type=O.PRM_QF, size=1, data=np.ones((1, 6, 1)))]).mult(np.zeros(6))
    assert np.max(np.abs(b - np.array([0, 0, 0, 1, 1, 1]) / 6.0)) <= 1e-15


@pytest.mark.gpu
@pytest.mark.parametrize("p", [1, 2])
@pytest.mark.parametrize("kind", ["diffusion", "minsurf"])
def test_scalar_forms_on_triangles(ctx, p, kind):
    from test_gpu_parity import _compare
    mesh = G.triangle_mesh((11, 9), perturb=0.2)
    s = G.h1_space(mesh, p, mode=O.GRAD)
    fs = S.diffusion(2) if kind == "diffusion" else S.minsurf(2, 0.5)
    ess = G.boundary_dofs(mesh, s)
    of, gi = S.make_pair(ctx, mesh, [s], fs, ess=ess)
    xc = G.dof_coords(mesh, s)
    x = np.sin(2.0 * xc[:, 0]) * np.cos(1.5 * xc[:, 1]) + 0.1 * np.random.default_rng(2).uniform(-1, 1, s["ndofs"])
    _compare(of, gi, x)


@pytest.mark.gpu
@pytest.mark.parametrize("spatial_bound", [False, True])
def test_ex5_block_on_triangles(ctx, spatial_bound):
    """ex5.cpp:88-140: GradientObstacleEnergy + HellingerEntropy, H1 P2 (GRAD) x (H1 P1)^2 (VALUE | VECTOR), psi_k a
    GridFunction parameter, default rule (order 6: 12 points); the bound 0.1 + 0.2 x + 0.4 y of ex5.cpp:114-117."""
    import mfem_ad_b200 as M
    from test_gpu_parity import _compare
    mesh = G.triangle_mesh((10, 10), perturb=0.1)
    u = G.h1_space(mesh, 2, mode=O.GRAD)
    lat = G.h1_space(mesh, 1, vdim=2, mode=O.VALUE | O.VECTOR)
    fs = S.pg(S.gradobstacle(2), S.hellinger(2, 0.0, qoff=2) if spatial_bound else S.hellinger(2, 0.7), 0.8)
    rng = np.random.default_rng(6)
    psik = rng.normal(0, 1, 2 * lat["ndofs"])
    gm = M.Mesh(ctx, mesh)
    gu, gl = M.Space(ctx, gm, u), M.Space(ctx, gm, lat)
    gi = M.Integrator(ctx, [(gu, O.GRAD), (gl, O.VALUE | O.VECTOR), (gl, O.VALUE | O.VECTOR, M.ROLE_PARAM)], fs.madb(ctx))
    gi.set_param_field(2, psik)
    params = [dict(type=O.PRM_GF, size=2, data=psik, space=lat)]
    if spatial_bound:
        bound = gi.set_param_coefficient(lambda X: 0.1 + 0.2 * X[:, 0] + 0.4 * X[:, 1])
        params.append(dict(type=O.PRM_QF, size=1, data=bound))
    of = O.OracleForm(mesh, [u, lat], fs.oracle(), params=params)
    x = np.concatenate([0.3 * rng.uniform(-1, 1, u["ndofs"]), rng.normal(0, 1, 2 * lat["ndofs"])])
    _compare(of, gi, x)


@pytest.mark.gpu
def test_load_vector_on_triangles(ctx):
    import mfem_ad_b200 as M
    mesh = G.triangle_mesh((8, 8))
    gm = M.Mesh(ctx, mesh)
    f = lambda p: 15.0 * np.sin(np.pi * p[:, 0]) ** 2  # ex5.cpp:87-90
    for p, nq in ((1, 3), (2, 6)):
        s = G.h1_space(mesh, p, mode=O.VALUE)
        gs = M.Space(ctx, gm, s)
        b = M.load_vector(ctx, gs, f)
        gi = M.Integrator(ctx, [(gs, O.VALUE)], S.load().madb(ctx), quad_order=2 * p)
        qf = gi.set_param_coefficient(f)
        assert qf.shape[1] == nq
        ref = O.OracleForm(mesh, [s], S.load().oracle(), quad_order=2 * p, params=[dict(type=O.PRM_QF, size=1, data=qf)]).mult(np.zeros(s["ndofs"]))
        assert S.csr_rel_err(b, ref) <= 1e-13
    assert abs(b.sum() - 7.5) <= 1e-3  # int 15 sin^2(pi x) = 7.5 (order-4 rule on P2: not exact, close)


@pytest.mark.gpu
def test_ex5_lvpp_on_triangles_iteration_counts(ctx):
    """The ex5 driver on its own mesh type (ex5.cpp:72-212): gradient-constrained problem, Hellinger entropy with the
    spatial bound 0.1 + 0.2 x + 0.4 y, load 15 sin^2(pi x), Newton (abs tol 1e-9, 20 iterations) inside the proximal loop;
    the CUDA assembly and the CPU oracle drive the same loop and must take the same Newton / PG iterations."""
    import mfem_ad_b200 as M
    from mfem_ad_b200 import lvpp
    mesh = G.triangle_mesh((6, 6))
    u = G.h1_space(mesh, 2, mode=O.GRAD)
    lat = G.h1_space(mesh, 1, vdim=2, mode=O.VALUE | O.VECTOR)
    nu, nl = u["ndofs"], 2 * lat["ndofs"]
    ess = G.boundary_dofs(mesh, u)
    fs_factory = lambda a: S.pg(S.gradobstacle(2), S.hellinger(2, 0.0, qoff=2), a)
    gm = M.Mesh(ctx, mesh)
    gu, gl = M.Space(ctx, gm, u), M.Space(ctx, gm, lat)
    gfn = fs_factory(1.0).madb(ctx)
    gi = M.Integrator(ctx, [(gu, O.GRAD), (gl, O.VALUE | O.VECTOR), (gl, O.VALUE | O.VECTOR, M.ROLE_PARAM)], gfn)
    gi.set_essential(ess)
    bound = gi.set_param_coefficient(lambda X: 0.1 + 0.2 * X[:, 0] + 0.4 * X[:, 1])
    b = np.zeros(nu + nl)
    b[:nu] = M.load_vector(ctx, M.Space(ctx, gm, dict(u, mode=O.VALUE)), lambda p: 15.0 * np.sin(np.pi * p[:, 0]) ** 2)
    b[ess] = 0.0
    # lumped P1 weights: a third of the area of every triangle at the vertex
    X = mesh["coords"][mesh["e2n"]]
    area = 0.5 * np.abs((X[:, 1, 0] - X[:, 0, 0]) * (X[:, 2, 1] - X[:, 0, 1]) - (X[:, 2, 0] - X[:, 0, 0]) * (X[:, 1, 1] - X[:, 0, 1]))
    wv = np.zeros(lat["ndofs"])
    np.add.at(wv, mesh["e2n"].reshape(-1), np.repeat(area / 3.0, 3))
    l1 = lambda v: float(np.sum(np.tile(wv, 2) * np.abs(v)))
    rule = M.PGStepSizeRule(M.PGStepSizeRule.EXP, 1.0, 1e4, 2.0)
    sl = slice(nu, nu + nl)
    nk = dict(abs_tol=1e-9, rel_tol=0.0, max_iter=20)

    class OracleOp:
        def __init__(self):
            self.alpha, self.psik, self._f = 1.0, np.zeros(nl), None

        def form(self):
            if self._f is None:
                self._f = O.OracleForm(mesh, [u, lat], fs_factory(self.alpha).oracle(), ess=ess,
                                       params=[dict(type=O.PRM_GF, size=2, data=self.psik, space=lat), dict(type=O.PRM_QF, size=1, data=bound)])
            return self._f

        def set_alpha(self, a):
            self.alpha, self._f = a, None

        def set_latent_k(self, p):
            self.psik, self._f = p.copy(), None

        def mult(self, x):
            return self.form().mult(x)

        def grad(self, x):
            return self.form().grad(x)[2]

        def pattern(self):
            return self.form().pattern()

    oop = OracleOp()
    xo = np.zeros(nu + nl)
    ho = lvpp.lvpp_solve(oop, oop.set_alpha, oop.set_latent_k, rule, b, xo, sl, l1, max_pg=25, tol=1e-8, newton_kw=nk)
    xg = np.zeros(nu + nl)
    hg = lvpp.lvpp_solve(gi, lambda a: gfn.set_params([a]), lambda p: gi.set_param_field(2, p), rule, b, xg, sl, l1, max_pg=25, tol=1e-8,
                         newton_kw=nk)
    assert not ho["newton_failed"] and not hg["newton_failed"], (ho, hg)
    assert ho["newton_iterations"] == hg["newton_iterations"] and ho["pg_iterations"] == hg["pg_iterations"]
    assert np.max(np.abs(xo - xg)) <= 1e-8 * max(1.0, np.max(np.abs(xo)))
    assert sum(hg["newton_iterations"]) > 5
    # the same loop with the linear solves on the device: ex5's latent space is H1 (no element blocks), so the symmetric
    # indefinite system goes through MINRES with the plain Jacobi preconditioner (madb_solver_pg_minres, nb = 0)
    lin = lvpp.DeviceLinear(gi, "minres", rtol=1e-13, maxit=20000)
    xd = np.zeros(nu + nl)
    hd = lvpp.lvpp_solve(gi, lambda a: gfn.set_params([a]), lambda p: gi.set_param_field(2, p), rule, b, xd, sl, l1, max_pg=25, tol=1e-8,
                         newton_kw=dict(nk, linear=lin))
    print("ex5 device MINRES iterations per solve: min %d max %d" % (min(lin.linear_iterations), max(lin.linear_iterations)))
    assert not hd["newton_failed"] and hd["newton_iterations"] == hg["newton_iterations"] and hd["pg_iterations"] == hg["pg_iterations"]
    assert np.max(np.abs(xd - xg)) <= 1e-7 * max(1.0, np.max(np.abs(xg)))

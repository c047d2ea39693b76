"""the linear solve of a Newton step on the device (madb_solver_*).

* CSR SpMV and Jacobi-PCG against scipy on the assembled minimal-surface Jacobian (ex2);
* the statically condensed PCG on the ex4 proximal-Galerkin block system against a SuperLU solve of the full system;
* Newton (ex2) and the LVPP loop (ex4) driven by the device solves: same iteration counts as with the host direct solve
  (BASELINE.json north_star: "Newton/LVPP iteration counts equal")."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import spec as S
from mfem_ad_b200 import lvpp, meshgen as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _csr(gi, vals):
    rp, ci = gi.pattern()
    return sp.csr_matrix((vals, ci, rp), shape=(rp.size - 1,) * 2)


def test_spmv_and_pcg_on_minimal_surface_jacobian(ctx):
    import mfem_ad_b200 as M
    mesh = G.cartesian_mesh((24, 20), perturb=0.15)
    s = G.h1_space(mesh, 2, mode=O.GRAD)
    ess = G.boundary_dofs(mesh, s)
    _, gi = S.make_pair(ctx, mesh, [s], S.minsurf(2, 0.5), ess=ess)
    rng = np.random.default_rng(0)
    x = rng.uniform(-0.3, 0.3, s["ndofs"])
    _, vals = gi.assemble(x)
    K = _csr(gi, vals)
    sol = M.Solver(gi)
    v = rng.uniform(-1, 1, x.size)
    assert np.max(np.abs(sol.spmv(vals, v) - K @ v)) <= 1e-13 * np.max(np.abs(K @ v))
    b = rng.uniform(-1, 1, x.size)
    b[ess] = 0.0
    xs, it, rr = sol.pcg(vals, b, rtol=1e-12)
    ref = spla.splu(K.tocsc()).solve(b)
    assert rr <= 1e-12 and 0 < it < 2000
    assert np.max(np.abs(xs - ref)) <= 1e-9 * np.max(np.abs(ref))
    # deterministic: same bits on a second run
    xs2, it2, _ = sol.pcg(vals, b, rtol=1e-12)
    assert it2 == it and np.array_equal(xs, xs2)
    # device pointers: nothing but the solution norm crosses PCIe
    import torch
    dv, db = torch.from_numpy(vals).cuda(), torch.from_numpy(b).cuda()
    dx = torch.zeros_like(db)
    _, it3, _ = sol.pcg(dv, db, dx, rtol=1e-12)
    assert it3 == it and np.array_equal(dx.cpu().numpy(), xs)


def _ex4_problem(ctx, n, order=2):
    mesh = G.cartesian_mesh((n, n))
    h1 = G.h1_space(mesh, order + 1, mode=O.VALUE | O.GRAD)
    l2 = G.l2_space(mesh, order - 1, mode=O.VALUE)
    ess = G.boundary_dofs(mesh, h1)
    fs_factory = lambda a: S.pg(S.obstacle(2), S.fermidirac(0.0, 0.5), a)
    b = np.zeros(h1["ndofs"] + l2["ndofs"])
    b[:h1["ndofs"]] = G.load_vector(mesh, h1, lambda x: 2 * np.pi ** 2 * np.sin(np.pi * x[..., 0]) * np.sin(np.pi * x[..., 1]))
    b[ess] = 0.0
    spaces = [h1, l2, dict(l2, role=1)]
    _, gi = S.make_pair(ctx, mesh, spaces, fs_factory(1.0), quad_order=3 * order + 3, ess=ess,
                        params=[dict(type=O.PRM_GF, size=1, data=np.zeros(l2["ndofs"]), space=l2)])
    return mesh, h1, l2, ess, b, gi


def test_condensed_pcg_on_the_pg_block_system(ctx):
    import mfem_ad_b200 as M
    mesh, h1, l2, ess, b, gi = _ex4_problem(ctx, 8)
    nh, nl = h1["ndofs"], l2["ndofs"]
    rng = np.random.default_rng(1)
    x = np.concatenate([0.1 * rng.uniform(-1, 1, nh), rng.normal(0, 1, nl)])
    x[ess] = 0.0
    gi.fn.set_params([0.4])
    gi.set_param_field(2, rng.normal(0, 1, nl))
    r, vals = gi.assemble(x)
    K = _csr(gi, vals)
    rhs = r - b
    sol = M.Solver(gi)
    c, it, rr = sol.condensed_pcg(nh, 4, vals, rhs, rtol=1e-13)
    ref = spla.splu(K.tocsc()).solve(rhs)
    assert rr <= 1e-12 and 0 < it < 3000
    assert np.max(np.abs(c - ref)) <= 1e-9 * np.max(np.abs(ref))
    # the block system as it is: MINRES with the block-diagonal (Jacobi / element Schur complement) preconditioner
    c2, it2, rr2 = sol.pg_minres(nh, 4, vals, rhs, rtol=1e-13)
    assert rr2 <= 1e-12 and 0 < it2 < 3000
    assert np.max(np.abs(c2 - ref)) <= 1e-8 * np.max(np.abs(ref))
    print("condensed PCG %d iterations, MINRES %d iterations" % (it, it2))


def test_newton_and_lvpp_iteration_counts_with_device_solves(ctx):
    import mfem_ad_b200 as M
    # ex2: Newton on the minimal surface problem, Jacobi-PCG on the device vs SuperLU on the host
    mesh = G.cartesian_mesh((16, 16))
    s = G.h1_space(mesh, 2, mode=O.GRAD)
    ess = G.boundary_dofs(mesh, s)
    xc = G.dof_coords(mesh, s)
    g = np.sin(2 * np.pi * xc[:, 0]) * 0.3 + 0.2 * xc[:, 1]  # boundary data
    x0 = np.zeros(s["ndofs"])
    x0[ess] = g[ess]
    _, gi = S.make_pair(ctx, mesh, [s], S.minsurf(2, 0.5), ess=ess)
    b = np.zeros(s["ndofs"])
    xa, xb = x0.copy(), x0.copy()
    ra = lvpp.newton(gi, b, xa, abs_tol=1e-10, max_iter=30)
    lin = lvpp.DeviceLinear(gi, "pcg", rtol=1e-13)
    rb = lvpp.newton(gi, b, xb, abs_tol=1e-10, max_iter=30, linear=lin)
    assert ra[0] and rb[0] and ra[1] == rb[1] and ra[1] >= 3
    assert np.max(np.abs(xa - xb)) <= 1e-9
    # ex4: LVPP loop, condensed PCG on the device vs SuperLU on the host
    mesh, h1, l2, ess, b, gi = _ex4_problem(ctx, 6)
    wl = G.lumped_weights(mesh, l2)
    l1 = lambda v: float(np.sum(wl * np.abs(v)))
    # test.sh:9 rule; the active set drives E*''(psi) towards 0, where the condensed operator A + C D^-1 C^T turns into
    # a penalty matrix (Jacobi-PCG stalls): the block system is solved as it is, MINRES + block-diagonal preconditioner
    rule = M.PGStepSizeRule(2, 0.1, 1e4, 2.0, 1.0)
    sl = slice(h1["ndofs"], h1["ndofs"] + l2["ndofs"])
    runs = []
    for linear in (None, lvpp.DeviceLinear(gi, "minres", nh=h1["ndofs"], nb=4, rtol=1e-13)):
        x = np.zeros(b.size)
        nk = dict(abs_tol=1e-9, rel_tol=0.0, max_iter=20, linear=linear)
        h = lvpp.lvpp_solve(gi, lambda a: gi.fn.set_params([a]), lambda p: gi.set_param_field(2, p), rule, b, x, sl, l1,
                            max_pg=30, newton_kw=nk)
        if linear is not None:
            print("linear iterations", linear.linear_iterations, "relres max %.2e" % max(linear.relres))
        runs.append((h, x))
    (ha, xa), (hb, xb) = runs
    assert not ha["newton_failed"] and not hb["newton_failed"]
    assert ha["newton_iterations"] == hb["newton_iterations"] and ha["pg_iterations"] == hb["pg_iterations"]
    assert ha["converged"] and hb["converged"] and sum(ha["newton_iterations"]) >= 10
    assert np.max(np.abs(xa - xb)) <= 1e-8 * max(1.0, np.max(np.abs(xa)))


def test_ex3_linear_elasticity_with_vector_load_solved_on_device(ctx):
    """ex3.cpp:37-78: (H1 order 2)^2, LinearElasticityEnergy(lambda = mu = 1), VectorDomainLFIntegrator load (1, 1), one
    essential boundary side, K(0) x = load: the load vector and the solve on the device against the oracle + SuperLU."""
    import mfem_ad_b200 as M
    mesh = G.cartesian_mesh((10, 10))
    s = G.h1_space(mesh, 2, vdim=2, mode=O.GRAD | O.VECTOR)
    nd = s["ndofs"]
    xc = G.dof_coords(mesh, s)
    side = np.nonzero(np.abs(xc[:, 0]) < 1e-12)[0]  # bdr attribute 4 of MakeCartesian2D (x = 0)
    ess = np.concatenate([side, side + nd]).astype(np.int32)
    of, gi = S.make_pair(ctx, mesh, [s], S.elasticity(2, 1.0, 1.0), ess=ess)
    gm = M.Mesh(ctx, mesh)
    gs = M.Space(ctx, gm, dict(s, mode=O.VALUE | O.VECTOR))
    load = M.load_vector(ctx, gs, lambda p: np.ones((p.shape[0], 2)))
    lref = O.OracleForm(mesh, [dict(s, mode=O.VALUE | O.VECTOR)], S.load(2).oracle(), quad_order=4,
                        params=[dict(type=O.PRM_QF, size=2, data=np.ones((100, 9, 2)))]).mult(np.zeros(2 * nd))
    assert S.csr_rel_err(load, lref) <= 1e-13 and abs(load.sum() - 2.0) <= 1e-13
    load[ess] = 0.0
    x0 = np.zeros(2 * nd)
    lin = lvpp.DeviceLinear(gi, "pcg", rtol=1e-13)
    x = lin.step(x0, -load)  # J(0) c = F(0) - b with F(0) = 0, b = -load  ->  c = K^-1 load
    rp, ci, vals = of.grad(x0)
    ref = spla.splu(sp.csr_matrix((vals, ci, rp), shape=(2 * nd,) * 2).tocsc()).solve(load)
    assert np.max(np.abs(x - ref)) <= 1e-9 * np.max(np.abs(ref))
    assert 0 < lin.linear_iterations[0] < 3000


def test_block_solvers_refuse_a_latent_space_that_is_not_element_local(ctx):
    import mfem_ad_b200 as M
    mesh, h1, l2, ess, b, gi = _ex4_problem(ctx, 4)
    gi.set_param_field(2, np.zeros(l2["ndofs"]))
    r, vals = gi.assemble(np.zeros(b.size))
    sol = M.Solver(gi)
    for fn in (sol.condensed_pcg, sol.pg_minres):
        with pytest.raises(M.MadbError, match="not block diagonal"):
            fn(h1["ndofs"], 2, vals, r)  # the latent space has 4 dofs per element, not 2

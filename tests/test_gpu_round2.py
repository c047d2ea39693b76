"""Round-2 additions: built-in energies that were templates without instances (MassEnergy, DiffEnergy, DiffusionEnergy
with K), Evaluator sources of quadrature-function / Coefficient type, HessianCoefficient, scalar order 3, and the
hand-off race of the persistent patch kernel on a mesh whose last patch is partial."""
import numpy as np
import pytest

import spec as S
from mfem_ad_b200 import meshgen as G
from oracle import oracle as O
from test_gpu_parity import _compare, _state, _block_state

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize("p", [1, 2])
@pytest.mark.parametrize("K", [(1.7,), (0.6, 2.5), (1.5, 0.3, 0.3, 0.9)])
def test_diffusion_energy_with_constant_K(ctx, p, K):
    """DiffusionEnergy: K scalar / diagonal / full, column-major (src/ad_native.hpp:433-480)."""
    mesh = G.cartesian_mesh((9, 7), perturb=0.15)
    s = G.h1_space(mesh, p, mode=O.GRAD)
    of, gi = S.make_pair(ctx, mesh, [s], S.diffusion(2, K))
    _compare(of, gi, _state(mesh, s))


def test_diffusion_energy_with_coefficient_K(ctx):
    """K as a spatial Coefficient (Evaluator source of Coefficient type, src/ad_native.hpp:57, src/ad_native.cpp:132-141):
    the host callback is sampled at the rule's points (madb_integrator_qpoint_coords) and handed over as a
    QuadratureFunction (madb_integrator_set_param_qf)."""
    mesh = G.cartesian_mesh((13, 11), perturb=0.15)
    s = G.h1_space(mesh, 2, mode=O.GRAD)
    kfun = lambda X: 1.0 + 0.5 * np.sin(3.0 * X[:, 0]) * np.cos(2.0 * X[:, 1])
    import mfem_ad_b200 as M
    gm = M.Mesh(ctx, mesh)
    gs = M.Space(ctx, gm, s)
    gi = M.Integrator(ctx, [(gs, O.GRAD)], S.diffusion(2, qoff=0, kdim=1).madb(ctx))
    with pytest.raises(M.MadbError, match="never set"):
        gi.mult(np.zeros(s["ndofs"]))
    qf = gi.set_param_coefficient(kfun)
    # the points are the images of the tensor Gauss points under the bilinear vertex map
    xq, _ = O.gauss_legendre(4)
    X = mesh["coords"][mesh["e2n"][5]]
    xi, eta = xq[2], xq[1]
    ref = (1 - xi) * (1 - eta) * X[0] + xi * (1 - eta) * X[1] + (1 - xi) * eta * X[2] + xi * eta * X[3]
    assert np.max(np.abs(gi.qpoint_coords()[5, 1 * 4 + 2] - ref)) <= 1e-15
    of = O.OracleForm(mesh, [s], S.diffusion(2, qoff=0, kdim=1).oracle(), params=[dict(type=O.PRM_QF, size=1, data=qf)])
    _compare(of, gi, _state(mesh, s))


@pytest.mark.parametrize("p", [1, 2])
def test_mass_and_diff_energy(ctx, p):
    """MassEnergy on a VALUE field (src/ad_native.hpp:413-420) and DiffEnergy(MassEnergy) with the target as a
    QuadratureFunction and as a GridFunction parameter (src/ad_native.hpp:483-525): L2-projection type forms."""
    mesh = G.cartesian_mesh((11, 9), perturb=0.15)
    s = G.h1_space(mesh, p, mode=O.VALUE)
    x = _state(mesh, s)
    of, gi = S.make_pair(ctx, mesh, [s], S.mass(1))
    _compare(of, gi, x)
    nq = (p + 2) ** 2
    tq = np.random.default_rng(3).normal(0, 1, (mesh["e2n"].shape[0], nq, 1))
    of, gi = S.make_pair(ctx, mesh, [s], S.diff(S.mass(1)), params=[dict(type=O.PRM_QF, size=1, data=tq)])
    gi.set_param_qf(tq)
    _compare(of, gi, x)
    if p == 2:
        tg = np.random.default_rng(4).normal(0, 1, s["ndofs"])
        of, gi = S.make_pair(ctx, mesh, [s, dict(s, role=1)], S.diff(S.mass(1)),
                             params=[dict(type=O.PRM_GF, size=1, data=tg, space=s)], block=False)
        gi.set_param_field(1, tg)
        _compare(of, gi, x)


def test_hellinger_with_spatial_bound(ctx):
    """ex5.cpp:114-117: the gradient bound of HellingerEntropy is a spatial Coefficient -> quadrature-function
    parameter of the entropy inside ADPGFunctional (per-point parameters: psi_k field first, then the bound)."""
    mesh = G.cartesian_mesh((9, 8), perturb=0.15)
    u = G.h1_space(mesh, 2, mode=O.GRAD)
    lat = G.h1_space(mesh, 1, vdim=2, mode=O.VALUE | O.VECTOR)
    fs = S.pg(S.gradobstacle(2), S.hellinger(2, 0.0, qoff=2), 0.8)
    psik = np.random.default_rng(2).normal(0, 1, 2 * lat["ndofs"])
    import mfem_ad_b200 as M
    gm = M.Mesh(ctx, mesh)
    gu, gl = M.Space(ctx, gm, u), M.Space(ctx, gm, lat)
    gi = M.Integrator(ctx, [(gu, O.GRAD), (gl, O.VALUE | O.VECTOR), (gl, O.VALUE | O.VECTOR, M.ROLE_PARAM)], fs.madb(ctx))
    gi.set_param_field(2, psik)
    bound = gi.set_param_coefficient(lambda X: 0.5 + 0.4 * X[:, 0] * (1.0 - X[:, 1]))
    of = O.OracleForm(mesh, [u, lat], fs.oracle(), params=[dict(type=O.PRM_GF, size=2, data=psik, space=lat),
                                                           dict(type=O.PRM_QF, size=1, data=bound)])
    _compare(of, gi, _block_state(mesh, [u, lat]))


def test_pointwise_new_kinds(ctx):
    rng = np.random.default_rng(8)
    for fs, n, qn in [(S.mass(1), 1, 0), (S.diffusion(2, (1.5, 0.3, 0.3, 0.9)), 2, 0), (S.diff(S.mass(1)), 1, 1),
                      (S.hellinger(2, 0.0, qoff=0), 2, 1)]:
        x = rng.normal(0, 1.5, (32, n))
        q = rng.uniform(0.2, 1.0, (32, qn)) if qn else None
        v, g, h = fs.madb(ctx).eval(x, q)
        fo = fs.oracle()
        for p in range(32):
            qp = None if q is None else q[p]
            assert abs(v[p] - fo.value(x[p], qp)) <= 1e-13 * max(1.0, abs(v[p]))
            assert np.max(np.abs(g[p] - fo.gradient(x[p], qp))) <= 1e-13 * max(1.0, np.max(np.abs(g[p])))
            assert np.max(np.abs(h[p] - fo.hessian(x[p], qp))) <= 1e-13 * max(1.0, np.max(np.abs(h[p])))


def test_hessian_coefficient(ctx):
    """DifferentiableCoefficient::Hessian() (HessianCoefficient, src/ad_native.hpp:300-323) at the rule's points:
    the derivative of the latent->primal map, softmax and sigmoid."""
    mesh = G.cartesian_mesh((6, 5), perturb=0.12)
    lat = G.h1_space(mesh, 1, vdim=5, mode=O.VALUE | O.VECTOR)
    psi = np.random.default_rng(99).normal(0, 1, 5 * lat["ndofs"])
    of, gi = S.make_pair(ctx, mesh, [lat], S.simplex(5, 1.0))
    val, grd, hes = gi.coefficient_hessian(psi)
    assert np.max(np.abs(val - of.coefficient(psi, 0))) <= TOL
    assert np.max(np.abs(grd - of.coefficient(psi, 1))) <= TOL
    assert np.max(np.abs(hes - of.coefficient(psi, 2))) <= TOL
    assert np.max(np.abs(hes - np.swapaxes(hes, 2, 3))) == 0.0
    # softmax: Hessian = diag(p) - p p^T
    p = grd[2, 3]
    assert np.max(np.abs(hes[2, 3] - (np.diag(p) - np.outer(p, p)))) <= 1e-14


@pytest.mark.parametrize("kind", ["diffusion", "minsurf"])
def test_scalar_order3(ctx, kind):
    """2-D scalar H1 order 3 (16 dofs, 5x5 points): 64-element patches, 4 threads per element."""
    mesh = G.cartesian_mesh((13, 11), perturb=0.15)
    s = G.permute_dofs(G.h1_space(mesh, 3, mode=O.GRAD), 6)
    fs = S.diffusion(2) if kind == "diffusion" else S.minsurf(2, 0.5)
    of, gi = S.make_pair(ctx, mesh, [s], fs)
    _compare(of, gi, _state(mesh, s))


def test_partial_last_patch_many_iterations(ctx):
    """The persistent patch kernel hands patches from compute to write-out through barriers that every thread has to
    take part in, also the idle ones of a partial last patch, and also when a CTA has processed many patches before
    (needs ne % patch size != 0 and more patches than resident work groups).  Bitwise repeatability over
    many assemblies + full parity against the oracle on a sub-mesh + symmetry."""
    import mfem_ad_b200 as M
    import scipy.sparse as sp
    nx, ny, p = 301, 263, 2  # 79163 elements: not a multiple of 64 or 128; > 2 * 148 * 2 patches
    mesh = G.cartesian_mesh((nx, ny), lengths=(1.0, ny / nx), perturb=0.1)
    s = G.h1_space(mesh, p, mode=O.GRAD)
    gm = M.Mesh(ctx, mesh)
    gs = M.Space(ctx, gm, s)
    gi = M.Integrator(ctx, [(gs, O.GRAD)], S.minsurf(2, 0.5).madb(ctx))
    st = gi.patch_stats()
    assert st["patches"] > 2 * 148 * 2 and (nx * ny) % 64 != 0
    x = _state(mesh, s)
    y0, v0 = gi.assemble(x)
    for _ in range(25):
        y, v = gi.assemble(x)
        assert np.array_equal(y, y0) and np.array_equal(v, v0)
    rp, ci = gi.pattern()
    K = sp.csr_matrix((v0, ci, rp), shape=(x.size,) * 2)
    assert abs(K - K.T).max() <= 1e-13 * np.max(np.abs(v0))
    d = np.random.default_rng(4321).uniform(-1, 1, x.size)
    assert S.csr_rel_err(gi.grad_mult(x, d), K @ d) <= 10 * TOL
    # the LAST elements of the mesh (the partial patch holds elements with the highest ids): rows interior to the
    # top-right corner sub-mesh against the oracle
    m = 10
    ex0, ey0 = nx - m, ny - m
    ngx = nx * p + 1
    iy, ix = np.divmod(np.arange((m * p + 1) ** 2), m * p + 1)
    big = (iy + ey0 * p) * ngx + (ix + ex0 * p)
    inner = (ix > 0) & (iy > 0)
    e_sub = ((np.arange(m)[:, None] + ey0) * nx + (np.arange(m)[None, :] + ex0)).reshape(-1)
    vsub = mesh["e2n"][e_sub]
    vids, inv = np.unique(vsub, return_inverse=True)
    sub = dict(dim=2, n=(m, m), lengths=(1.0, 1.0), e2n=inv.reshape(vsub.shape).astype(np.int32), coords=mesh["coords"][vids], geom_order=1)
    ss = G.h1_space(G.cartesian_mesh((m, m)), p, mode=O.GRAD)
    of = O.OracleForm(sub, [ss], S.minsurf(2, 0.5).oracle())
    ys = of.mult(x[big])
    rps, cis, vs = of.grad(x[big])
    assert S.csr_rel_err(y0[big][inner], ys[inner]) <= TOL
    Ks = sp.csr_matrix((vs, cis, rps), shape=(big.size,) * 2)
    rows = np.nonzero(inner)[0]
    assert np.max(np.abs(K[big[rows]][:, big].toarray() - Ks[rows].toarray())) <= TOL * np.max(np.abs(vs))


def test_load_vector_matches_oracle_and_closed_form(ctx):
    """b_i = (f, phi_i) (LinearForm + DomainLFIntegrator, ex4.cpp:145-148) on the device: against the
    oracle's gradient of the same energy, against the exact integral of f (sum of b = int f: partition of unity) and,
    for f = 1 on a uniform Q1 mesh, against the hand values h^2/4 x (number of elements at the node)."""
    import mfem_ad_b200 as M
    mesh = G.cartesian_mesh((7, 5), perturb=0.2)
    f = lambda p: 2.0 * np.pi ** 2 * np.sin(np.pi * p[:, 0]) * np.sin(np.pi * p[:, 1])  # ex4.cpp:93-98
    for order, qo in ((1, 2), (2, 4), (3, 6), (3, 9)):
        s = G.h1_space(mesh, order, mode=O.VALUE)
        gm = M.Mesh(ctx, mesh)
        gs = M.Space(ctx, gm, s)
        b = M.load_vector(ctx, gs, f, quad_order=qo)
        gi = M.Integrator(ctx, [(gs, O.VALUE)], S.load().madb(ctx), quad_order=qo)
        qf = gi.set_param_coefficient(f)
        ref = O.OracleForm(mesh, [s], S.load().oracle(), quad_order=qo, params=[dict(type=O.PRM_QF, size=1, data=qf)]).mult(np.zeros(s["ndofs"]))
        assert S.csr_rel_err(b, ref) <= 1e-13
    # exact integral of f over [0,1]^2 is 8; order-3 space with the 5x5 rule on the unperturbed mesh
    mesh = G.cartesian_mesh((8, 8))
    s = G.h1_space(mesh, 3, mode=O.VALUE)
    gm = M.Mesh(ctx, mesh)
    b = M.load_vector(ctx, M.Space(ctx, gm, s), f, quad_order=9)
    assert abs(b.sum() - 8.0) <= 1e-7
    s1 = G.h1_space(mesh, 1, mode=O.VALUE)
    b1 = M.load_vector(ctx, M.Space(ctx, gm, s1), lambda p: np.ones(p.shape[0]))
    cnt = np.zeros(s1["ndofs"])
    np.add.at(cnt, np.asarray(s1["e2l"]).reshape(-1), 1.0)
    assert np.max(np.abs(b1 - cnt / 64.0 / 4.0)) <= 1e-15


def test_qvalue_latent_on_the_quadrature_space(ctx):
    """ADEval::QVALUE (src/ad_intg.hpp:127: the shape of a quadrature-space unknown is the unit vector at ip.index;
    src/tools.hpp:156-177): the device treats it as VALUE on the L2 space whose nodes are the rule's Gauss points; the
    oracle implements the unit vector literally.  ex4-type PG block, H1 order 1 primal, latent at the 3x3 points."""
    import mfem_ad_b200 as M
    mesh = G.cartesian_mesh((9, 8), perturb=0.15)
    h1 = G.h1_space(mesh, 1, mode=O.VALUE | O.GRAD)
    lq = G.l2_space(mesh, 2, mode=O.QVALUE)  # 9 dofs per element = the 9 points, x fastest like ip.index
    fs = S.pg(S.obstacle(2), S.fermidirac(0.0, 0.5), 0.6)
    psik = np.random.default_rng(4).normal(0, 1, lq["ndofs"])
    of = O.OracleForm(mesh, [h1, lq], fs.oracle(), quad_order=4, params=[dict(type=O.PRM_GF, size=1, data=psik, space=dict(lq, mode=O.VALUE))])
    gm = M.Mesh(ctx, mesh)
    gh, gl = M.Space(ctx, gm, h1), M.Space(ctx, gm, lq)
    gi = M.Integrator(ctx, [(gh, O.VALUE | O.GRAD), (gl, O.QVALUE), (gl, O.VALUE, M.ROLE_PARAM)], fs.madb(ctx), quad_order=4)
    gi.set_param_field(2, psik)
    _compare(of, gi, _block_state(mesh, [h1, lq]))
    # a space that is not the rule's quadrature space is refused
    l1 = G.l2_space(mesh, 1, mode=O.QVALUE)
    with pytest.raises(M.MadbError, match="quadrature space"):
        M.Integrator(ctx, [(gh, O.VALUE | O.GRAD), (M.Space(ctx, gm, l1), O.QVALUE), (gl, O.VALUE, M.ROLE_PARAM)], fs.madb(ctx), quad_order=4)
    with pytest.raises(M.MadbError, match="QVALUE can only be combined"):
        M.Integrator(ctx, [(gh, O.VALUE | O.GRAD), (gl, O.QVALUE | O.VALUE), (gl, O.VALUE, M.ROLE_PARAM)], fs.madb(ctx), quad_order=4)


@pytest.mark.parametrize("fs,n,qn", [
    (S.lagrangian(S.diffusion(2), [S.minsurf(2, 0.5), S.diffusion(2)]), 4, 0),
    (S.lagrangian(S.diffusion(2), [S.minsurf(2, 0.5), S.diffusion(2)], mode=1), 4, 0),
    (S.al(S.diffusion(2), [S.minsurf(2, 0.5), S.diffusion(2)], 2.5, [1.2, -0.4], [-0.7, 0.3]), 2, 0),
    (S.pg2(S.obstacle(2), S.fermidirac(0.0, 0.5), 0, S.hellinger(2, 0.7), 1, 0.6), 6, 3),
])
def test_pointwise_multi_constraint_and_multi_entropy(ctx, fs, n, qn):
    """Several equality constraints (src/ad_native.hpp:583,607-618,648,684-690) and two entropies (src/pg.hpp:105-127)."""
    rng = np.random.default_rng(7)
    x = rng.normal(0, 1.2, (48, n))
    q = rng.normal(0, 1, (48, qn)) if qn else None
    v, g, h = fs.madb(ctx).eval(x, q)
    fo = fs.oracle()
    for p in range(x.shape[0]):
        qp = None if q is None else q[p]
        vr, gr, hr = fo.value(x[p], qp), fo.gradient(x[p], qp), fo.hessian(x[p], qp)
        sc = max(1.0, abs(vr), np.max(np.abs(gr)), np.max(np.abs(hr)))
        assert abs(v[p] - vr) <= 1e-13 * sc and np.max(np.abs(g[p] - gr)) <= 1e-13 * sc and np.max(np.abs(h[p] - hr)) <= 1e-13 * sc


def test_multi_constraint_and_multi_entropy_forms(ctx):
    mesh = G.cartesian_mesh((13, 9), perturb=0.15)
    h1 = G.h1_space(mesh, 1, mode=O.GRAD)
    l0v = G.l2_space(mesh, 0, vdim=2, mode=O.VALUE | O.VECTOR)
    of, gi = S.make_pair(ctx, mesh, [h1, l0v], S.lagrangian(S.diffusion(2), [S.minsurf(2, 0.5), S.diffusion(2)]))
    _compare(of, gi, _block_state(mesh, [h1, l0v]))
    h2 = G.h1_space(mesh, 2, mode=O.GRAD)
    of, gi = S.make_pair(ctx, mesh, [h2], S.al(S.diffusion(2), [S.minsurf(2, 0.5), S.diffusion(2)], 2.5, [1.2, -0.4], [-0.7, 0.3]))
    _compare(of, gi, _state(mesh, h2))
    # two entropies: u in [0, 0.5] (FermiDirac on u) and |grad u| bounded (Hellinger on grad u)
    import mfem_ad_b200 as M
    hv = G.h1_space(mesh, 2, mode=O.VALUE | O.GRAD)
    p1 = G.l2_space(mesh, 0, mode=O.VALUE)
    p2 = G.l2_space(mesh, 0, vdim=2, mode=O.VALUE | O.VECTOR)
    fs = S.pg2(S.obstacle(2), S.fermidirac(0.0, 0.5), 0, S.hellinger(2, 0.7), 1, 0.6)
    rng = np.random.default_rng(3)
    k1, k2 = rng.normal(0, 1, p1["ndofs"]), rng.normal(0, 1, 2 * p2["ndofs"])
    of = O.OracleForm(mesh, [hv, p1, p2], fs.oracle(), quad_order=6,
                      params=[dict(type=O.PRM_GF, size=1, data=k1, space=p1), dict(type=O.PRM_GF, size=2, data=k2, space=p2)])
    gm = M.Mesh(ctx, mesh)
    gh, g1, g2 = M.Space(ctx, gm, hv), M.Space(ctx, gm, p1), M.Space(ctx, gm, p2)
    gi = M.Integrator(ctx, [(gh, O.VALUE | O.GRAD), (g1, O.VALUE), (g2, O.VALUE | O.VECTOR), (g1, O.VALUE, M.ROLE_PARAM),
                            (g2, O.VALUE | O.VECTOR, M.ROLE_PARAM)], fs.madb(ctx), quad_order=6)
    gi.set_param_field(3, k1)
    gi.set_param_field(4, k2)
    _compare(of, gi, _block_state(mesh, [hv, p1, p2]))


def test_assemble_begin_end_matches_assemble(ctx):
    """madb_integrator_assemble_begin / _end (the split used to overlap the shared-dof exchange with the interface
    reduction of the CSR values): same bits as the one-call assembly, with essential dofs, several patches."""
    import torch
    mesh = G.cartesian_mesh((40, 37), perturb=0.15)
    s = G.h1_space(mesh, 2, mode=O.GRAD)
    ess = G.boundary_dofs(mesh, s)
    _, gi = S.make_pair(ctx, mesh, [s], S.minsurf(2, 0.5), ess=ess)
    x = torch.from_numpy(_state(mesh, s)).cuda()
    y0, y1 = torch.empty_like(x), torch.empty_like(x)
    v0 = torch.empty(gi.nnz, dtype=torch.float64, device=x.device)
    v1 = torch.full_like(v0, -3.0)
    gi.assemble(x, y0, v0)
    gi.assemble_begin(x, y1, v1)
    gi.assemble_end()
    gi.assemble_end()  # no-op
    torch.cuda.synchronize()
    assert torch.equal(y0, y1) and torch.equal(v0, v1)

"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerance (BASELINE.json north_star): residuals and Jacobian values within 1e-12
relative in FP64 (norm-relative: max |diff| / max |ref|), sparsity pattern bit-exact."""
import numpy as np
import pytest

import spec as S
from mfem_ad_b200 import meshgen as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _state(mesh, s, seed=1234):
    xc = G.dof_coords(mesh, s)
    u = np.ones(s["ndofs"])
    for d in range(mesh["dim"]):
        u = u * np.sin(np.pi * xc[:, d])
    return u + 0.1 * np.random.default_rng(seed).uniform(-1, 1, s["ndofs"])  # SURVEY 8d config 2 state


def _compare(of, gi, x, energy=True, action=True):
    y_ref = of.mult(x)
    rp, ci, v_ref = of.grad(x)
    y = gi.mult(x)
    assert S.csr_rel_err(y, y_ref) <= TOL
    rpg, cig = gi.pattern()
    assert np.array_equal(rpg, rp) and np.array_equal(cig, ci)  # pattern bit-exact
    v = gi.grad(x)
    assert S.csr_rel_err(v, v_ref) <= TOL
    y2, v2 = gi.assemble(x)
    assert np.array_equal(v2, v)  # fused == separate (same kernel, fixed summation order): bit-identical
    # the residual of the fused kernel comes from the hyper-dual instantiation, the separate one from the dual
    # instantiation: same arithmetic, but nvcc may contract a*b+c differently in the two -> rounding-level only
    assert S.csr_rel_err(y2, y) <= 1e-14
    y3, v3 = gi.assemble(x)
    assert np.array_equal(y3, y2) and np.array_equal(v3, v2)  # run-to-run deterministic
    if energy:
        e_ref = of.energy(x)
        assert abs(gi.energy(x) - e_ref) <= TOL * max(1.0, abs(e_ref))
    if not action:
        return
    # matrix-free action == assembled Jacobian
    import scipy.sparse as sp
    d = np.random.default_rng(4321).uniform(-1, 1, x.size)
    Kd = sp.csr_matrix((v_ref, ci, rp), shape=(x.size,) * 2) @ d
    assert S.csr_rel_err(gi.grad_mult(x, d), Kd) <= TOL * 10


def test_ex0_known_answers_on_device(ctx):
    # config 1: every printed "error" of ex0 (ex0.cpp:139-140) <= 1e-14, device AD type
    f = S.FSpec("ex0", 3).madb(ctx)
    v, g, h = f.eval(np.array([[0.5, 1.0, -1.0]]))
    assert abs(v[0] - 0.30321372968699545) <= 1e-14
    assert np.linalg.norm(g[0] - np.array([2.3855167309591354, 1.3032137296869954, 3.0])) <= 1e-14
    Href = np.array([[-1.3032137296869954, 2.3855167309591354, 0.0], [2.3855167309591354, 1.3032137296869954, 0.0],
                     [0.0, 0.0, -6.0]])
    assert np.max(np.abs(h[0] - Href)) <= 1e-14


def test_ex0_vector_function_on_device(ctx):
    # ADVectorFunction (ex0.cpp:23-35): "Jacobian2 error" and the Hessian slices of ex0.cpp:153-158 on the device
    import mfem_ad_b200 as M
    f = M.Functional(ctx, "ex0vec")
    x0 = np.array([[0.5, 1.0, -1.0]])
    v, J, H = f.eval_vector(x0, 2)
    assert np.max(np.abs(v[0] - np.array([np.sin(0.5), np.cos(-0.5)]))) <= 1e-15
    Jref = np.array([[0.8775825618903728, 0.4387912809451864, 0.0],
                     [-0.479425538604203, -0.2397127693021015, 0.2397127693021015]])
    assert np.max(np.abs(J[0] - Jref)) <= 1e-14
    H0 = np.array([[-0.479425538604203, 0.6378697925882713, 0], [0.6378697925882713, -0.11985638465105075, 0], [0, 0, 0]])
    H1 = np.array([[-0.8775825618903728, -0.9182168195493894, 0.9182168195493894],
                   [-0.9182168195493894, -0.2193956404725932, 0.4591084097746947],
                   [0.9182168195493894, 0.4591084097746947, -0.2193956404725932]])
    assert np.max(np.abs(H[0, 0] - H0)) <= 1e-14 and np.max(np.abs(H[0, 1] - H1)) <= 1e-14
    # many points against the oracle
    g = O.Functional()
    g.add(O.K_EX0VEC, 3, n_output=2)
    X = np.random.default_rng(2).normal(0, 1, (50, 3))
    v, J, H = f.eval_vector(X, 2)
    for p in range(X.shape[0]):
        assert np.max(np.abs(J[p] - g.vec_gradient(X[p]))) <= 1e-13
        assert np.max(np.abs(H[p] - g.vec_hessian(X[p]))) <= 1e-13


@pytest.mark.parametrize("fs,n,qn", [
    (S.FSpec("ex0", 3), 3, 0), (S.minsurf(2, 0.5), 2, 0), (S.shannon(0.25, -1), 1, 0),
    (S.fermidirac(0.0, 0.5), 1, 0), (S.hellinger(2, 0.7), 2, 0), (S.simplex(5, 1.5), 5, 0),
    (S.simplex(3, 1.0), 3, 0), (S.simp([1e-3, 0.25, 0.5, 0.75, 1.0], 3.0), 5, 0),
    (S.elasticity(2, 2.0, 0.7), 4, 0), (S.elasticity(3, 2.0, 0.7), 9, 0),
    (S.pg(S.obstacle(2), S.fermidirac(0.0, 0.5), 0.7), 4, 1),
    (S.lambdapg(S.obstacle(2), S.fermidirac(0.0, 0.5), 0.7), 4, 1),
    (S.lagrangian(S.diffusion(2), S.minsurf(2, 0.5)), 3, 0),
    (S.al(S.diffusion(2), S.minsurf(2, 0.5), 2.5, 1.2, -0.7), 2, 0),
])
def test_pointwise_ad_matches_oracle(ctx, fs, n, qn):
    rng = np.random.default_rng(5)
    x = rng.uniform(0.05, 1.0, (64, n)) if fs.kind == "simp" else rng.normal(0, 1.5, (64, n))
    q = rng.normal(0, 1, (64, qn)) if qn else None
    if fs.kind in ("simplex", "fermidirac"):
        x[0] = 0.0  # ties / branch point
        x[1] = 40.0 * np.sign(x[1] + 0.1)
    v, g, h = fs.madb(ctx).eval(x, q)
    fo = fs.oracle()
    for p in range(x.shape[0]):
        qp = None if q is None else q[p]
        vr, gr, hr = fo.value(x[p], qp), fo.gradient(x[p], qp), fo.hessian(x[p], qp)
        sc = max(1.0, abs(vr), np.max(np.abs(gr)), np.max(np.abs(hr)))
        assert abs(v[p] - vr) <= 1e-13 * sc
        assert np.max(np.abs(g[p] - gr)) <= 1e-13 * sc
        assert np.max(np.abs(h[p] - hr)) <= 1e-13 * sc


@pytest.mark.parametrize("p", [1, 2])
@pytest.mark.parametrize("kind", ["diffusion", "minsurf"])
@pytest.mark.parametrize("perturb", [0.0, 0.2])
def test_scalar_grad_2d(ctx, p, kind, perturb):
    mesh = G.cartesian_mesh((7, 5), lengths=(1.0, 0.8), perturb=perturb)
    s = G.permute_dofs(G.h1_space(mesh, p, mode=O.GRAD), 11)
    fs = S.diffusion(2) if kind == "diffusion" else S.minsurf(2, 0.5)
    of, gi = S.make_pair(ctx, mesh, [s], fs)
    _compare(of, gi, _state(mesh, s))


@pytest.mark.parametrize("case", ["q2", "q2perm", "q2shuffled", "q1", "elast", "hex"])
def test_patch_assembly_many_patches(ctx, case):
    """Meshes of several patches (PATCH_PE = 128 elements per CTA): interior rows written by the patch,
    interface rows through the staging buffer + fixed-order reduction; essential dofs on top."""
    if case == "hex":
        mesh = G.cartesian_mesh((9, 8, 7), perturb=0.12)
        s = G.h1_space(mesh, 1, mode=O.GRAD)
        fs = S.minsurf(3, 0.5)
    elif case == "elast":
        mesh = G.cartesian_mesh((23, 19), perturb=0.15)
        s = G.h1_space(mesh, 1, vdim=2, ordering=1, mode=O.GRAD | O.VECTOR)
        fs = S.elasticity(2, 1.0, 1.0)
    else:
        mesh = G.cartesian_mesh((31, 26), lengths=(1.0, 0.8), perturb=0.2)
        s = G.h1_space(mesh, 1 if case == "q1" else 2, mode=O.GRAD)
        if case == "q2perm":
            s = G.permute_dofs(s, 3)
        if case == "q2shuffled":  # random element order and vertex numbering, random dof numbering
            s = G.permute_dofs(s, 4)
            mesh, (s,) = G.shuffle_mesh(mesh, [s], 8)
        fs = S.minsurf(2, 0.5)
    ess = G.boundary_dofs(mesh, s)[::2] if case == "q2" else ()
    of, gi = S.make_pair(ctx, mesh, [s], fs, ess=ess, block=(1 if case == "elast" else None))
    st = gi.patch_stats()
    assert st["patches"] >= 4 and st["ifc_dofs"] > 0
    if case == "elast":
        _compare(of, gi, _block_state(mesh, [s]))
    else:
        _compare(of, gi, _state(mesh, s), energy=(len(ess) == 0))
    assert gi.patch_stats()["ifc_entries"] > 0


def test_essential_bc_and_parameter_update(ctx):
    mesh = G.cartesian_mesh((6, 6), perturb=0.1)
    s = G.h1_space(mesh, 2, mode=O.GRAD)
    ess = G.boundary_dofs(mesh, s)
    fs = S.minsurf(2, 0.5)
    of, gi = S.make_pair(ctx, mesh, [s], fs, ess=ess)
    x = _state(mesh, s)
    _compare(of, gi, x, energy=False)
    # eps is mutated between solves (ex2.cpp:98): parameters are re-read at every call
    gi.fn.set_params([0.125])
    of2 = O.OracleForm(mesh, [s], S.minsurf(2, 0.125).oracle(), ess=ess)
    assert S.csr_rel_err(gi.mult(x), of2.mult(x)) <= TOL
    assert S.csr_rel_err(gi.grad(x), of2.grad(x)[2]) <= TOL


def test_larger_mesh_properties(ctx):
    """Size-independent checks on a mesh the oracle would be slow on: symmetry, constants in the
    kernel of the diffusion Jacobian, linearity of the residual, determinism across calls."""
    import scipy.sparse as sp
    mesh = G.cartesian_mesh((160, 120), perturb=0.1)
    s = G.h1_space(mesh, 2, mode=O.GRAD)
    import mfem_ad_b200 as M
    gm = M.Mesh(ctx, mesh)
    gs = M.Space(ctx, gm, s)
    gi = M.Integrator(ctx, [(gs, O.GRAD)], S.diffusion(2).madb(ctx))
    x = _state(mesh, s)
    rp, ci = gi.pattern()
    v = gi.grad(x)
    K = sp.csr_matrix((v, ci, rp), shape=(x.size,) * 2)
    assert abs(K - K.T).max() <= 1e-13 * np.max(np.abs(v))
    assert np.max(np.abs(K @ np.ones(x.size))) <= 1e-11
    y1, y2 = gi.mult(x), gi.mult(2.5 * x)
    assert S.csr_rel_err(y2, 2.5 * y1) <= 1e-13
    assert S.csr_rel_err(y1, K @ x) <= 1e-12
    assert np.array_equal(gi.grad(x), v) and np.array_equal(gi.mult(x), y1)
    # oracle on a sub-range of elements agrees where those rows are complete: spot-check energy instead
    assert abs(gi.energy(x) - 0.5 * x @ (K @ x)) <= 1e-11 * abs(0.5 * x @ (K @ x))


def test_missing_configuration_fails_loudly(ctx):
    import mfem_ad_b200 as M
    mesh = G.cartesian_mesh((2, 2))
    s = G.h1_space(mesh, 2, mode=O.GRAD)
    gm = M.Mesh(ctx, mesh)
    gs = M.Space(ctx, gm, s)
    with pytest.raises(M.MadbError, match="no fused kernel compiled"):
        M.Integrator(ctx, [(gs, O.GRAD)], S.FSpec("nosuchenergy", 2).madb(ctx))
    with pytest.raises(M.MadbError, match="not supported"):
        M.Integrator(ctx, [(gs, O.HESSIAN)], S.diffusion(2).madb(ctx))  # isValidADEval, src/_ad_intg.hpp:58-59


def _block_state(mesh, spaces, seed=7):
    rng = np.random.default_rng(seed)
    parts = []
    for s in spaces:
        parts.append(rng.uniform(-1, 1, s["ndofs"] * s.get("vdim", 1)))
    return np.concatenate(parts)


@pytest.mark.parametrize("case", ["ex4o2", "ex4o2perm", "ex5o2", "elast2"])
def test_large_element_matrices_many_patches(ctx, case):
    """Element matrices of 17-20 dofs: 64-element patches, 4 threads per element (slices of the upper triangle),
    several patches so that interface rows, staging and the fixed-order reduction are exercised."""
    mesh = G.cartesian_mesh((19, 14), perturb=0.15)
    if case in ("ex4o2", "ex4o2perm"):
        order = 2
        h1 = G.h1_space(mesh, order + 1, mode=O.VALUE | O.GRAD)
        ess = G.boundary_dofs(mesh, h1)
        if case == "ex4o2perm":
            # random dof numbering: every 32-entry chunk of the CSR image is irregular, the gather maps do not fit
            # in shared memory next to the staged 20x20 matrices -> the integrator falls back to the colour path
            h1 = G.permute_dofs(h1, 3)
            ess = h1["perm"][ess]
        l2 = G.l2_space(mesh, order - 1, mode=O.VALUE)
        psik = np.random.default_rng(11).normal(0, 1, l2["ndofs"])
        of, gi = S.make_pair(ctx, mesh, [h1, l2, dict(l2, role=1)], S.pg(S.obstacle(2), S.fermidirac(0.0, 0.5), 0.4),
                             quad_order=3 * order + 3, ess=ess, params=[dict(type=O.PRM_GF, size=1, data=psik, space=l2)])
        gi.set_param_field(2, psik)
        x = _block_state(mesh, [h1, l2])
    elif case == "ex5o2":
        u = G.h1_space(mesh, 2, mode=O.GRAD)
        lat = G.h1_space(mesh, 1, vdim=2, mode=O.VALUE | O.VECTOR)
        psik = np.random.default_rng(13).normal(0, 1, 2 * lat["ndofs"])
        of, gi = S.make_pair(ctx, mesh, [u, lat, dict(lat, role=1)], S.pg(S.gradobstacle(2), S.hellinger(2, 0.7), 0.4),
                             params=[dict(type=O.PRM_GF, size=2, data=psik, space=lat)])
        gi.set_param_field(2, psik)
        x = _block_state(mesh, [u, lat])
    else:
        s = G.h1_space(mesh, 2, vdim=2, mode=O.GRAD | O.VECTOR)
        of, gi = S.make_pair(ctx, mesh, [s], S.elasticity(2, 2.0, 0.7), block=1)
        x = _block_state(mesh, [s])
    st = gi.patch_stats()
    if case == "ex4o2perm":
        assert st["patches"] in (0, (19 * 14 + 63) // 64)
    else:
        assert st["patches"] == (19 * 14 + 63) // 64 and st["ifc_dofs"] > 0
    _compare(of, gi, x)


def test_lagrangian_and_augmented_lagrangian_forms(ctx):
    """Lagrangian (block: H1 order 1 x L2 order 0, x = [grad u, lambda]) and ALFunctional (order-2 space,
    sum-factorised path) of src/ad_native.hpp:570-691 against the oracle; several patches."""
    mesh = G.cartesian_mesh((17, 11), perturb=0.15)
    h1 = G.h1_space(mesh, 1, mode=O.GRAD)
    l0 = G.l2_space(mesh, 0, mode=O.VALUE)
    of, gi = S.make_pair(ctx, mesh, [h1, l0], S.lagrangian(S.diffusion(2), S.minsurf(2, 0.5)))
    _compare(of, gi, _block_state(mesh, [h1, l0]))
    h2 = G.h1_space(mesh, 2, mode=O.GRAD)
    fs = S.al(S.diffusion(2), S.minsurf(2, 0.5), 2.5, 1.2, -0.7)
    of, gi = S.make_pair(ctx, mesh, [h2], fs)
    x = _state(mesh, h2)
    _compare(of, gi, x)
    gi.fn.set_params([4.0, 1.2, 0.3])  # SetPenalty / SetLambda between outer iterations
    of2 = O.OracleForm(mesh, [h2], S.al(S.diffusion(2), S.minsurf(2, 0.5), 4.0, 1.2, 0.3).oracle())
    assert S.csr_rel_err(gi.mult(x), of2.mult(x)) <= TOL
    assert S.csr_rel_err(gi.grad(x), of2.grad(x)[2]) <= TOL


def test_lambda_pg_block(ctx):
    """ADLambdaPGFunctional (src/pg.hpp:216-243) on the ex4 -o 1 spaces: unknowns (u, lambda), psi_k a parameter."""
    order = 1
    mesh = G.cartesian_mesh((6, 5), perturb=0.15)
    h1 = G.h1_space(mesh, order + 1, mode=O.VALUE | O.GRAD)
    l2 = G.l2_space(mesh, order - 1, mode=O.VALUE)
    fs = S.lambdapg(S.obstacle(2), S.fermidirac(0.0, 0.5), 0.6)
    psik = np.random.default_rng(12).normal(0, 1, l2["ndofs"])
    of, gi = S.make_pair(ctx, mesh, [h1, l2, dict(l2, role=1)], fs, quad_order=3 * order + 3,
                         params=[dict(type=O.PRM_GF, size=1, data=psik, space=l2)])
    gi.set_param_field(2, psik)
    _compare(of, gi, _block_state(mesh, [h1, l2]))


@pytest.mark.parametrize("order,perturb", [(1, 0.0), (1, 0.2), (2, 0.15)])
def test_ex4_obstacle_pg_block(ctx, order, perturb):
    """ex4.cpp:99-142: H1(p+1) x L2(p-1), ADPGFunctional(ObstacleEnergy, FermiDirac(0,0.5), psi_k),
    ADBlockNonlinearFormIntegrator<VALUE|GRAD, VALUE>, rule order 3p+3; essential bc on the H1 block."""
    mesh = G.cartesian_mesh((5, 4), perturb=perturb)
    h1 = G.permute_dofs(G.h1_space(mesh, order + 1, mode=O.VALUE | O.GRAD), 3)
    l2 = G.l2_space(mesh, order - 1, mode=O.VALUE)
    alpha = 0.4
    fs = S.pg(S.obstacle(2), S.fermidirac(0.0, 0.5), alpha)
    rng = np.random.default_rng(11)
    psik = rng.normal(0, 1, l2["ndofs"])
    ess = h1["perm"][G.boundary_dofs(mesh, G.h1_space(mesh, order + 1))]
    spaces = [h1, l2, dict(l2, role=1)]
    of, gi = S.make_pair(ctx, mesh, spaces, fs, quad_order=3 * order + 3, ess=ess,
                         params=[dict(type=O.PRM_GF, size=1, data=psik, space=l2)])
    gi.set_param_field(2, psik)
    x = _block_state(mesh, [h1, l2])
    _compare(of, gi, x)
    # alpha changes every PG step (ex4.cpp:185-187), psi_k too (:188-189)
    psik2 = x[h1["ndofs"]:].copy()
    gi.fn.set_params([3.2])
    gi.set_param_field(2, psik2)
    of2 = O.OracleForm(mesh, [h1, l2], S.pg(S.obstacle(2), S.fermidirac(0.0, 0.5), 3.2).oracle(),
                       quad_order=3 * order + 3, ess=ess, params=[dict(type=O.PRM_GF, size=1, data=psik2, space=l2)])
    assert S.csr_rel_err(gi.mult(x), of2.mult(x)) <= TOL
    assert S.csr_rel_err(gi.grad(x), of2.grad(x)[2]) <= TOL


def test_ex5_gradient_constraint_pg_block_on_quads(ctx):
    """ex5.cpp:88-140 on quadrilaterals: H1(p) x H1(p-1)^dim, modes GRAD / VALUE|VECTOR, HellingerEntropy."""
    mesh = G.cartesian_mesh((4, 5), perturb=0.15)
    u = G.h1_space(mesh, 2, mode=O.GRAD)
    lat = G.h1_space(mesh, 1, vdim=2, mode=O.VALUE | O.VECTOR)
    fs = S.pg(S.gradobstacle(2), S.hellinger(2, 0.6), 0.8)
    psik = np.random.default_rng(2).normal(0, 1, 2 * lat["ndofs"])
    spaces = [u, lat, dict(lat, role=1)]
    of, gi = S.make_pair(ctx, mesh, spaces, fs, params=[dict(type=O.PRM_GF, size=2, data=psik, space=lat)])
    gi.set_param_field(2, psik)
    _compare(of, gi, _block_state(mesh, [u, lat]))


@pytest.mark.parametrize("p,ordering", [(1, 0), (2, 0), (1, 1)])
def test_ex3_vector_elasticity(ctx, p, ordering):
    """ex3.cpp:50-63: vector H1 space, GRAD|VECTOR, LinearElasticityEnergy.  Default = the reference's single-space
    arithmetic as written (src/ad_intg.hpp:310-326: windows of Hx, untransposed mirror blocks -- SURVEY H1), for
    lambda == mu (ex3) and for lambda != mu, where it differs O(1) from the index-consistent contraction of the block
    integrator (src/ad_intg.hpp:700-727), which MADB_INTEG_BLOCK selects."""
    mesh = G.cartesian_mesh((4, 3), perturb=0.1)
    s = G.h1_space(mesh, p, vdim=2, ordering=ordering, mode=O.GRAD | O.VECTOR)
    x = _block_state(mesh, [s])
    of, gi = S.make_pair(ctx, mesh, [s], S.elasticity(2, 1.0, 1.0), block=0)  # ex3.cpp:58
    _compare(of, gi, x)
    of0, gi0 = S.make_pair(ctx, mesh, [s], S.elasticity(2, 2.0, 0.7), block=0)  # as written, lambda != mu
    _compare(of0, gi0, x)
    of1, gi1 = S.make_pair(ctx, mesh, [s], S.elasticity(2, 2.0, 0.7), block=1)  # consistent variant
    _compare(of1, gi1, x)
    v0, v1 = gi0.grad(x), gi1.grad(x)
    assert np.max(np.abs(v0 - v1)) > 1e-2 * np.max(np.abs(v1))  # the two really are different matrices (H1)
    assert np.array_equal(gi0.mult(x), gi1.mult(x))              # the residual is not affected


@pytest.mark.parametrize("kind", ["diffusion", "minsurf"])
def test_3d_trilinear_assembled(ctx, kind):
    mesh = G.cartesian_mesh((4, 3, 3), perturb=0.15)
    s = G.permute_dofs(G.h1_space(mesh, 1, mode=O.GRAD), 5)
    fs = S.diffusion(3) if kind == "diffusion" else S.minsurf(3, 0.5)
    of, gi = S.make_pair(ctx, mesh, [s], fs)
    assert gi.ncolors >= 8 or gi.patch_stats()["patches"] >= 1  # colour-scatter path or patch assembly
    _compare(of, gi, _state(mesh, s))


@pytest.mark.parametrize("p", [2, 3])
@pytest.mark.parametrize("kind", ["diffusion", "minsurf"])
def test_3d_sum_factorised_matrix_free(ctx, p, kind):
    """Config 3 shape: H1 order p hexes, residual by sum factorisation, Jacobian ACTION y = J(x) v
    (v ~ U(-1,1) seed 4321) against the dense oracle on a small sub-problem."""
    import scipy.sparse as sp
    import mfem_ad_b200 as M
    mesh = G.cartesian_mesh((3, 2, 2) if p == 3 else (4, 3, 2), perturb=0.12)
    s = G.permute_dofs(G.h1_space(mesh, p, mode=O.GRAD), 9)
    ess = s["perm"][G.boundary_dofs(mesh, G.h1_space(mesh, p))][::3]
    fs = S.diffusion(3) if kind == "diffusion" else S.minsurf(3, 0.5)
    of, gi = S.make_pair(ctx, mesh, [s], fs, ess=ess)
    x = _state(mesh, s)
    assert S.csr_rel_err(gi.mult(x), of.mult(x)) <= TOL
    e_ref = of.energy(x)
    assert abs(gi.energy(x) - e_ref) <= TOL * max(1.0, abs(e_ref))
    rp, ci, v_ref = of.grad(x)
    d = np.random.default_rng(4321).uniform(-1, 1, x.size)
    Kd = sp.csr_matrix((v_ref, ci, rp), shape=(x.size,) * 2) @ d
    assert S.csr_rel_err(gi.grad_mult(x, d), Kd) <= TOL
    with pytest.raises(M.MadbError, match="matrix-free only"):
        gi.grad(x)

"""Parity at BASELINE.json's full sizes.  The oracle cannot assemble 10^6 elements in seconds, so the checks are
(i) element-level parity on a corner sub-mesh: rows of dofs whose elements all lie in the sub-mesh are identical in
the big assembly and in the oracle's assembly of the sub-mesh alone, and (ii) size-independent properties:
symmetry, Jacobian = derivative of the residual, matrix-free action == assembled Jacobian, run-to-run determinism."""
import numpy as np
import pytest

import spec as S
from mfem_ad_b200 import meshgen as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _sub_dofs(n_big, m, p, dim):
    """global (big-mesh) ids of the sub-mesh dofs, and the mask of sub-mesh dofs not on its cut faces"""
    ngb, ngs = n_big * p + 1, m * p + 1
    idx = np.indices((ngs,) * dim).reshape(dim, -1)  # slowest index first (z, y, x)
    big = np.zeros(idx.shape[1], dtype=np.int64)
    for d in range(dim):
        big = big * ngb + idx[d]
    inner = np.all(idx < ngs - 1, axis=0)
    return big, inner


def test_config2_full_size(ctx):
    import mfem_ad_b200 as M
    import scipy.sparse as sp
    nx, m, p = 1000, 24, 2
    mesh = G.cartesian_mesh((nx, nx))
    s = G.h1_space(mesh, p, mode=O.GRAD)
    gm = M.Mesh(ctx, mesh)
    gs = M.Space(ctx, gm, s)
    gi = M.Integrator(ctx, [(gs, O.GRAD)], S.minsurf(2, 0.5).madb(ctx))
    assert gi.patch_stats()["patches"] == (nx * nx + 127) // 128
    xc = G.dof_coords(mesh, s)
    x = np.sin(np.pi * xc[:, 0]) * np.sin(np.pi * xc[:, 1]) + 0.1 * np.random.default_rng(1234).uniform(-1, 1, s["ndofs"])
    y, v = gi.assemble(x)
    rp, ci = gi.pattern()
    assert rp[-1] == 64016001 and s["ndofs"] == 4004001  # config 2 sizes
    # (i) corner sub-mesh against the oracle
    sub = G.cartesian_mesh((m, m), lengths=(m / nx, m / nx))
    ss = G.h1_space(sub, p, mode=O.GRAD)
    big, inner = _sub_dofs(nx, m, p, 2)
    of = O.OracleForm(sub, [ss], S.minsurf(2, 0.5).oracle())
    ys = of.mult(x[big])
    rps, cis, vs = of.grad(x[big])
    assert S.csr_rel_err(y[big][inner], ys[inner]) <= TOL
    K = sp.csr_matrix((v, ci, rp), shape=(x.size,) * 2)
    Ks = sp.csr_matrix((vs, cis, rps), shape=(big.size,) * 2)
    rows = np.nonzero(inner)[0]
    Kbig_sub = K[big[rows]][:, big].toarray()
    assert np.max(np.abs(Kbig_sub - Ks[rows].toarray())) <= TOL * np.max(np.abs(vs))
    # (ii) properties on the whole mesh
    assert abs(K - K.T).max() <= 1e-13 * np.max(np.abs(v))
    d = np.random.default_rng(4321).uniform(-1, 1, x.size)
    Kd = K @ d
    assert S.csr_rel_err(gi.grad_mult(x, d), Kd) <= 10 * TOL
    h = 1e-7  # gradients of the perturbation are O(1/h_mesh) = 1e3: truncation (h 1e3)^2, rounding 1e-16 / h
    fd = (gi.mult(x + h * d) - gi.mult(x - h * d)) / (2 * h)
    assert S.csr_rel_err(fd, Kd) <= 1e-6
    y2, v2 = gi.assemble(x)
    assert np.array_equal(y2, y) and np.array_equal(v2, v)


def test_config3_full_size(ctx):
    import mfem_ad_b200 as M
    n, m, p = 104, 4, 3
    mesh = G.cartesian_mesh((n, n, n))
    s = G.h1_space(mesh, p, mode=O.GRAD)
    assert s["ndofs"] == 30664297
    gm = M.Mesh(ctx, mesh)
    gs = M.Space(ctx, gm, s)
    gi = M.Integrator(ctx, [(gs, O.GRAD)], S.minsurf(3, 0.5).madb(ctx))
    rng = np.random.default_rng(7)
    x = rng.uniform(-1, 1, s["ndofs"])
    y = gi.mult(x)
    sub = G.cartesian_mesh((m, m, m), lengths=(m / n,) * 3)
    ss = G.h1_space(sub, p, mode=O.GRAD)
    big, inner = _sub_dofs(n, m, p, 3)
    of = O.OracleForm(sub, [ss], S.minsurf(3, 0.5).oracle())
    ys = of.mult(x[big])
    assert S.csr_rel_err(y[big][inner], ys[inner]) <= TOL
    # matrix-free action: rows of the sub-mesh against the oracle's assembled Jacobian (columns outside the
    # sub-mesh do not couple to rows whose elements all lie inside it), and linearity in the direction
    import scipy.sparse as sp
    v = rng.uniform(-1, 1, s["ndofs"])
    jv = gi.grad_mult(x, v)
    rps, cis, vs = of.grad(x[big])
    Ks = sp.csr_matrix((vs, cis, rps), shape=(big.size,) * 2)
    ref = Ks @ v[big]
    assert np.max(np.abs(jv[big][inner] - ref[inner])) <= 10 * TOL * np.max(np.abs(ref))
    jv2 = gi.grad_mult(x, -2.0 * v)
    assert S.csr_rel_err(jv2, -2.0 * jv) <= 1e-13
    assert np.array_equal(gi.mult(x), y)


def test_config5_block_full_size(ctx):
    """ex4 PG block (H1 p3 x L2 p1, 5x5 points) on 1024^2 elements per GPU: corner sub-mesh against the oracle."""
    import mfem_ad_b200 as M
    import scipy.sparse as sp
    n, m, order = 1024, 6, 2
    mesh = G.cartesian_mesh((n, n))
    h1 = G.h1_space(mesh, order + 1, mode=O.VALUE | O.GRAD)
    l2 = G.l2_space(mesh, order - 1, mode=O.VALUE)
    assert h1["ndofs"] == 9443329 and l2["ndofs"] == 4194304
    fs = S.pg(S.obstacle(2), S.fermidirac(0.0, 0.5), 0.1)
    gm = M.Mesh(ctx, mesh)
    gh, gl = M.Space(ctx, gm, h1), M.Space(ctx, gm, l2)
    gi = M.Integrator(ctx, [(gh, O.VALUE | O.GRAD), (gl, O.VALUE), (gl, O.VALUE, M.ROLE_PARAM)], fs.madb(ctx), quad_order=3 * order + 3)
    rng = np.random.default_rng(3)
    psik = rng.normal(0, 1, l2["ndofs"])
    gi.set_param_field(2, psik)
    x = rng.uniform(-1, 1, h1["ndofs"] + l2["ndofs"])
    y, v = gi.assemble(x)
    rp, ci = gi.pattern()
    # sub-mesh m x m at the origin: H1 dofs through the node grid, L2 dofs through the element grid
    sub = G.cartesian_mesh((m, m), lengths=(m / n, m / n))
    sh1 = G.h1_space(sub, order + 1, mode=O.VALUE | O.GRAD)
    sl2 = G.l2_space(sub, order - 1, mode=O.VALUE)
    big_h, inner_h = _sub_dofs(n, m, order + 1, 2)
    ey, ex = np.divmod(np.arange(m * m), m)
    big_l = ((ey * n + ex)[:, None] * 4 + np.arange(4)[None, :]).reshape(-1)
    big = np.concatenate([big_h, h1["ndofs"] + big_l])
    inner = np.concatenate([inner_h, np.ones(big_l.size, dtype=bool)])
    of = O.OracleForm(sub, [sh1, sl2], fs.oracle(), quad_order=3 * order + 3,
                      params=[dict(type=O.PRM_GF, size=1, data=psik[big_l], space=sl2)])
    ys = of.mult(x[big])
    rps, cis, vs = of.grad(x[big])
    assert S.csr_rel_err(y[big][inner], ys[inner]) <= TOL
    K = sp.csr_matrix((v, ci, rp), shape=(x.size,) * 2)
    Ks = sp.csr_matrix((vs, cis, rps), shape=(big.size,) * 2)
    rows = np.nonzero(inner)[0]
    assert np.max(np.abs(K[big[rows]][:, big].toarray() - Ks[rows].toarray())) <= TOL * np.max(np.abs(vs))
    assert abs(K - K.T).max() <= 1e-13 * np.max(np.abs(v))
    y2, v2 = gi.assemble(x)
    assert np.array_equal(y2, y) and np.array_equal(v2, v)


def test_config4_full_size(ctx):
    """Config 4: 1072x1072 Q1 quads, displacement (vdim 2) + 5-component latent psi per node
    (8,059,303 dofs): softmax latent->primal map, SIMP lambda(rho), mu(rho), ParametrizedCompliance state block
    (the reference's single-space VECTOR arithmetic) and ParamGradient as written, against the oracle
    on a corner sub-mesh; symmetry / determinism on the whole mesh."""
    import mfem_ad_b200 as M
    import scipy.sparse as sp
    n, m = 1072, 8
    E = [1e-3, 0.25, 0.5, 0.75, 1.0]
    mesh = G.cartesian_mesh((n, n))
    disp = G.h1_space(mesh, 1, vdim=2, mode=O.GRAD | O.VECTOR)
    lat = G.h1_space(mesh, 1, vdim=5, mode=O.VALUE | O.VECTOR)
    nd = lat["ndofs"]
    assert nd == 1151329 and 7 * nd == 8059303
    psi = np.random.default_rng(99).normal(0, 1, 5 * nd)
    # latent -> primal map on the device (pointwise kernel), against numpy's softmax
    _, g, _ = S.simplex(5, 1.0).madb(ctx).eval(psi.reshape(5, nd).T.copy())
    p5 = psi.reshape(5, nd)
    ex_ = np.exp(p5 - p5.max(axis=0))
    rho = (ex_ / ex_.sum(axis=0)).reshape(-1)
    assert np.max(np.abs(g.T.reshape(-1) - rho)) <= 1e-15
    lam_fs, mu_fs = S.simp(E, 3.0), S.simp([0.5 * e for e in E], 3.0)
    gm = M.Mesh(ctx, mesh)
    gd, gl = M.Space(ctx, gm, disp), M.Space(ctx, gm, lat)
    fn = M.Functional(ctx, "paramcompliance", children=[lam_fs.madb(ctx), mu_fs.madb(ctx)])
    gi = M.Integrator(ctx, [(gd, O.GRAD | O.VECTOR), (gl, O.VALUE | O.VECTOR, M.ROLE_PARAM)], fn)
    gi.set_param_field(1, rho)
    x = np.random.default_rng(5).uniform(-1, 1, 2 * nd)
    y, v = gi.assemble(x)
    rp, ci = gi.pattern()
    assert rp[-1] == 41396356  # BASELINE.md config 4 nnz
    # corner sub-mesh
    sub = G.cartesian_mesh((m, m), lengths=(m / n, m / n))
    sdisp = G.h1_space(sub, 1, vdim=2, mode=O.GRAD | O.VECTOR)
    slat = G.h1_space(sub, 1, vdim=5, mode=O.VALUE | O.VECTOR)
    big1, inner1 = _sub_dofs(n, m, 1, 2)
    bigd = np.concatenate([c * nd + big1 for c in range(2)])
    bigl = np.concatenate([c * nd + big1 for c in range(5)])
    inner = np.concatenate([inner1, inner1])
    rq = O.OracleForm(sub, [slat], S.simplex(5).oracle()).inputs_at_qpts(rho[bigl])
    lo, mo = lam_fs.oracle(), mu_fs.oracle()
    qf = np.array([[[lo.value(r), mo.value(r)] for r in re] for re in rq])
    of = O.OracleForm(sub, [sdisp], S.FSpec("paramcompliance", 4, qoff=0).oracle(),
                      params=[dict(type=O.PRM_QF, size=2, data=qf)], block=0)
    ys = of.mult(x[bigd])
    rps, cis, vs = of.grad(x[bigd])
    assert S.csr_rel_err(y[bigd][inner], ys[inner]) <= TOL
    K = sp.csr_matrix((v, ci, rp), shape=(x.size,) * 2)
    Ks = sp.csr_matrix((vs, cis, rps), shape=(bigd.size,) * 2)
    rows = np.nonzero(inner)[0]
    assert np.max(np.abs(K[bigd[rows]][:, bigd].toarray() - Ks[rows].toarray())) <= TOL * np.max(np.abs(vs))
    assert abs(K - K.T).max() <= 1e-13 * np.max(np.abs(v))
    y2, v2 = gi.assemble(x)
    assert np.array_equal(y2, y) and np.array_equal(v2, v)
    del K, v, v2
    # ParamGradient as written (src/mmto.cpp:25-37) at the points of the sub-mesh elements
    fnd = M.Functional(ctx, "designcompliance", children=[lam_fs.madb(ctx), mu_fs.madb(ctx)])
    gdes = M.Integrator(ctx, [(gl, O.VALUE | O.VECTOR), (gd, O.GRAD, M.ROLE_PARAM)], fnd)
    gdes.set_param_field(1, x)
    _, J = gdes.param_gradient(rho)
    F = O.Functional()
    il = lam_fs._add(F)
    im = mu_fs._add(F)
    S.FSpec("paramcompliance", 4, qoff=4)._add(F)
    ofd = O.OracleForm(sub, [slat], F, params=[dict(type=O.PRM_GF_GRAD, size=4, data=x[bigd], space=sdisp)])
    J_ref = ofd.mmto_param_gradient(rho[bigl], [il, im])
    ey, ex = np.divmod(np.arange(m * m), m)
    assert np.max(np.abs(J[ey * n + ex] - J_ref)) <= TOL * np.max(np.abs(J_ref))

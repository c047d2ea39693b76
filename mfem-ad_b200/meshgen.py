"""Synthetic Cartesian meshes and tensor-product H1/L2 dof maps (numpy, host side).

Stand-in for Mesh::MakeCartesian2D/3D + FiniteElementSpace numbering of the
reference's drivers (ex1.cpp:36-41, ex4.cpp:78-101): MFEM is not available, so
benchmarks and tests generate the element->vertex and element->dof maps here
and hand the SAME arrays to the CUDA path and to the CPU oracle.
All local orderings are lexicographic (x fastest).
"""
import numpy as np


def gauss_lobatto_01(n):
    """n Gauss-Lobatto points on [0,1] (H1 closed basis nodes)."""
    if n == 1:
        return np.array([0.5])
    if n == 2:
        return np.array([0.0, 1.0])
    N = n - 1
    # interior nodes: roots of P_N'
    c = np.zeros(N + 1)
    c[N] = 1.0
    r = np.polynomial.legendre.Legendre(c).deriv().roots()
    x = np.concatenate([[-1.0], np.sort(r.real), [1.0]])
    x = 0.5 * (x - x[::-1])  # symmetrise
    return 0.5 * (x + 1.0)


def gauss_legendre_01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def cartesian_mesh(n, lengths=None, perturb=0.0):
    """n = (nx, ny[, nz]) elements.  Returns dict(dim, e2n[ne,2^dim], coords[nv,dim]).

    perturb > 0 displaces interior vertices smoothly (non-affine elements)."""
    n = tuple(int(v) for v in n)
    dim = len(n)
    lengths = tuple(lengths) if lengths is not None else (1.0,) * dim
    nv1 = [k + 1 for k in n]
    grids = np.meshgrid(*[np.arange(k) for k in nv1[::-1]], indexing="ij")  # z,y,x order
    idx = [g.ravel() for g in grids[::-1]]  # x fastest
    coords = np.stack([idx[d] * (lengths[d] / n[d]) for d in range(dim)], axis=1).astype(np.float64)
    if perturb:
        s = np.ones(len(coords))
        for d in range(dim):
            s = s * np.sin(np.pi * coords[:, d] / lengths[d])
        base = coords.copy()
        for d in range(dim):
            h = lengths[d] / n[d]
            phase = sum((k + 1.3) * base[:, k] / lengths[k] for k in range(dim) if k != d) if dim > 1 else 0.0
            coords[:, d] += perturb * h * s * np.cos(2.0 * np.pi * phase + d)
    eg = np.meshgrid(*[np.arange(k) for k in n[::-1]], indexing="ij")
    eidx = [g.ravel() for g in eg[::-1]]  # ex fastest
    e2n = np.zeros((len(eidx[0]), 2 ** dim), dtype=np.int32)
    for loc in range(2 ** dim):
        v = np.zeros_like(eidx[0])
        stride = 1
        for d in range(dim):
            v = v + (eidx[d] + ((loc >> d) & 1)) * stride
            stride *= nv1[d]
        e2n[:, loc] = v
    return dict(dim=dim, n=n, lengths=lengths, e2n=e2n, coords=coords, geom_order=1)


def triangle_mesh(n, lengths=None, perturb=0.0):
    """Mesh::MakeCartesian2D(nx, ny, Element::TRIANGLE) (ex5.cpp:72-73): every cell of the nx x ny grid is cut by its
    (v0, v2) diagonal into the triangles (v0, v1, v2) and (v0, v2, v3).  Returns dict(dim, e2n[ne, 3], coords,
    geom_order=-1 (the oracle's code for linear simplices), simplex=True, edges[ned, 2], e2e[ne, 3])."""
    q = cartesian_mesh(n, lengths, perturb)
    v = q["e2n"]  # [nq, 4] lexicographic: (0,0), (1,0), (0,1), (1,1)
    tri = np.concatenate([v[:, [0, 1, 3]], v[:, [0, 3, 2]]], axis=1).reshape(-1, 3).astype(np.int32)
    # edges in MFEM's local order (0,1), (1,2), (2,0), numbered in order of first appearance
    loc = np.stack([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]], axis=1)  # [ne, 3, 2]
    key = np.sort(loc, axis=2).reshape(-1, 2)
    uniq, first, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first)
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    e2e = rank[inv.reshape(-1)].reshape(-1, 3).astype(np.int32)
    edges = uniq[order]
    return dict(dim=2, n=q["n"], lengths=q["lengths"], e2n=tri, coords=q["coords"], geom_order=-1, simplex=True,
                edges=edges.astype(np.int32), e2e=e2e)


def _tri_h1_space(mesh, p, vdim, ordering, mode):
    nv = mesh["coords"].shape[0]
    if p == 1:
        return dict(basis=0, order=1, vdim=vdim, ordering=ordering, ndofs=nv, e2l=mesh["e2n"].astype(np.int32), mode=mode)
    if p == 2:
        e2l = np.concatenate([mesh["e2n"], nv + mesh["e2e"]], axis=1).astype(np.int32)
        return dict(basis=0, order=2, vdim=vdim, ordering=ordering, ndofs=nv + mesh["edges"].shape[0], e2l=e2l, mode=mode)
    raise ValueError("triangles: H1 orders 1 and 2")


def h1_space(mesh, p, vdim=1, ordering=0, mode=0):
    """Continuous tensor space of order p on a cartesian_mesh; dofs numbered
    lexicographically on the global (n*p+1)^dim node grid.  On a triangle_mesh: P1 (vertices) / P2 (vertices, then edges)."""
    if mesh.get("simplex"):
        return _tri_h1_space(mesh, p, vdim, ordering, mode)
    n, dim = mesh["n"], mesh["dim"]
    ng = [k * p + 1 for k in n]
    eg = np.meshgrid(*[np.arange(k) for k in n[::-1]], indexing="ij")
    eidx = [g.ravel() for g in eg[::-1]]
    nd = (p + 1) ** dim
    e2l = np.zeros((len(eidx[0]), nd), dtype=np.int32)
    for loc in range(nd):
        r, v, stride = loc, np.zeros_like(eidx[0]), 1
        for d in range(dim):
            i = r % (p + 1)
            r //= (p + 1)
            v = v + (eidx[d] * p + i) * stride
            stride *= ng[d]
        e2l[:, loc] = v
    return dict(basis=0, order=p, vdim=vdim, ordering=ordering, ndofs=int(np.prod(ng)), e2l=e2l, mode=mode)


def l2_space(mesh, p, vdim=1, ordering=0, mode=0):
    ne = mesh["e2n"].shape[0]
    if mesh.get("simplex") and p != 0:
        raise ValueError("triangles: L2 order 0")
    nd = (p + 1) ** mesh["dim"]
    e2l = (np.arange(ne, dtype=np.int64)[:, None] * nd + np.arange(nd)[None, :]).astype(np.int32)
    return dict(basis=1, order=p, vdim=vdim, ordering=ordering, ndofs=ne * nd, e2l=e2l, mode=mode)


def permute_dofs(space, seed):
    """Random renumbering of the scalar dofs (tests: nothing may depend on the numbering)."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(space["ndofs"]).astype(np.int32)
    out = dict(space)
    out["e2l"] = perm[space["e2l"]]
    out["perm"] = perm
    return out


def shuffle_mesh(mesh, spaces, seed):
    """Same mesh with the elements and the vertices renumbered at random (and the element rows of the
    dof maps permuted alike): nothing in the assembly may depend on a structured numbering."""
    rng = np.random.default_rng(seed)
    ne, nv = mesh["e2n"].shape[0], mesh["coords"].shape[0]
    pe = rng.permutation(ne)
    pv = rng.permutation(nv).astype(np.int32)  # old vertex id -> new id
    coords = np.empty_like(mesh["coords"])
    coords[pv] = mesh["coords"]
    out = dict(mesh, e2n=pv[mesh["e2n"][pe]].astype(np.int32), coords=coords)
    return out, [dict(s, e2l=np.ascontiguousarray(s["e2l"][pe])) for s in spaces]


def _lagrange(nodes, t):
    t = np.atleast_1d(t)
    B = np.ones((len(t), len(nodes)))
    for j in range(len(nodes)):
        for k in range(len(nodes)):
            if k != j:
                B[:, j] *= (t - nodes[k]) / (nodes[j] - nodes[k])
    return B


def dof_coords(mesh, space):
    """Physical coordinates of the scalar dofs (nodes mapped through the vertex map)."""
    if mesh.get("simplex"):
        X = mesh["coords"]
        if space["basis"] == 1:
            return X[mesh["e2n"]].mean(axis=1)
        if space["order"] == 1:
            return X.copy()
        return np.concatenate([X, 0.5 * (X[mesh["edges"][:, 0]] + X[mesh["edges"][:, 1]])], axis=0)
    dim, p = mesh["dim"], space["order"]
    nodes = gauss_lobatto_01(p + 1) if space["basis"] == 0 else gauss_legendre_01(p + 1)[0]
    N = _lagrange(np.array([0.0, 1.0]), nodes)  # [p+1, 2]
    X = mesh["coords"][mesh["e2n"]]  # [ne, 2^dim, dim]
    nd = (p + 1) ** dim
    out = np.zeros((space["ndofs"], dim))
    for loc in range(nd):
        r, wgt = loc, np.ones(2 ** dim)
        for d in range(dim):
            i = r % (p + 1)
            r //= (p + 1)
            for v in range(2 ** dim):
                wgt[v] *= N[i, (v >> d) & 1]
        out[space["e2l"][:, loc]] = np.einsum("v,evd->ed", wgt, X)
    return out


def boundary_dofs(mesh, space, tol=1e-12):
    """Scalar dofs on the boundary of the Cartesian box (by coordinates of the unperturbed grid)."""
    xc = dof_coords(dict(mesh, coords=cartesian_mesh(mesh["n"], mesh["lengths"])["coords"]), space)  # same vertices on triangles
    on = np.zeros(space["ndofs"], dtype=bool)
    for d in range(mesh["dim"]):
        on |= (np.abs(xc[:, d]) < tol) | (np.abs(xc[:, d] - mesh["lengths"][d]) < tol)
    return np.nonzero(on)[0].astype(np.int32)


def _tables(space, dim, xq):
    nodes = gauss_lobatto_01(space["order"] + 1) if space["basis"] == 0 else gauss_legendre_01(space["order"] + 1)[0]
    B = _lagrange(nodes, xq)  # [nq1, nn]
    nn, nq1 = len(nodes), len(xq)
    nd, nq = nn ** dim, nq1 ** dim
    phi = np.ones((nq, nd))
    for q in range(nq):
        for i in range(nd):
            rq, ri, v = q, i, 1.0
            for d in range(dim):
                v *= B[rq % nq1, ri % nn]
                rq //= nq1
                ri //= nn
            phi[q, i] = v
    return phi


def load_vector(mesh, space, f, nq1d=None):
    """b_i = int f phi_i  (DomainLFIntegrator, ex4.cpp:145-148) with a tensor Gauss rule; host numpy."""
    dim = mesh["dim"]
    nq1d = nq1d or space["order"] + 3
    xq, wq = gauss_legendre_01(nq1d)
    phi = _tables(space, dim, xq)
    geo = dict(basis=0, order=1)
    gphi = _tables(geo, dim, xq)  # vertex (bilinear) shape values at the points
    N2 = _lagrange(np.array([0.0, 1.0]), xq)
    dN2 = np.stack([-np.ones_like(xq), np.ones_like(xq)], axis=1)
    nq = nq1d ** dim
    X = mesh["coords"][mesh["e2n"]]  # [ne, nv, dim]
    xphys = np.einsum("qv,evd->eqd", gphi, X)
    # Jacobian determinants
    dg = np.zeros((nq, 2 ** dim, dim))
    for q in range(nq):
        for v in range(2 ** dim):
            for k in range(dim):
                rq, val = q, 1.0
                for d in range(dim):
                    t = rq % nq1d
                    rq //= nq1d
                    val *= dN2[t, (v >> d) & 1] if d == k else N2[t, (v >> d) & 1]
                dg[q, v, k] = val
    J = np.einsum("evi,qvk->eqik", X, dg)
    det = np.linalg.det(J)
    w = np.ones(nq)
    for q in range(nq):
        rq = q
        for d in range(dim):
            w[q] *= wq[rq % nq1d]
            rq //= nq1d
    fq = f(xphys)  # [ne, nq]
    be = np.einsum("eq,q,eq,qi->ei", fq, w, det, phi)
    b = np.zeros(space["ndofs"])
    np.add.at(b, space["e2l"], be)
    return b


def lumped_weights(mesh, space):
    """Nodal quadrature weights (reference weight x |det J| at the node) of an L2 Gauss-Legendre space."""
    dim, p = mesh["dim"], space["order"]
    xq, wq = gauss_legendre_01(p + 1)
    one = load_vector(mesh, space, lambda x: np.ones(x.shape[:2]), nq1d=p + 1)
    return one  # with the collocated rule phi_i(x_q) = delta_iq, so b_i = w_i det J_i

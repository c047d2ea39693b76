"""Closed-form finite-element known answers, derived symbolically (sympy) -- NOT produced by the oracle or the CUDA code.

    python tests/golden/make_fe_closed_forms.py      (writes tests/golden/fe_closed_forms.json)

The reference cannot be built in this image and ships no golden element matrices, so
the FE substrate the hot path sits on (MFEM's H1 Gauss-Lobatto tensor bases, Gauss-Legendre tensor rules, the
isoparametric map, `CalcPhysDShape`, the B^T D B contractions of src/ad_intg.hpp:245-255,:307-331,:698-727 and the
form-level scatter) is pinned here by values every correct implementation must reproduce:

* exact (rational) element matrices by symbolic integration, for cases where the rule the reference selects
  (`2*order+2` -> order+2 Gauss points per direction, src/_ad_intg.hpp:99-105, exact to degree 2*order+3) integrates
  the integrand exactly: Q1/Q2 stiffness and mass on a rectangle and on a parallelogram (affine maps: polynomial
  integrands), the mass matrix on a general (non-affine) quadrilateral (|J| is bilinear: still polynomial),
  DiffusionEnergy with a full K, the order-1 elasticity matrix (lambda != mu) -- both the index-consistent
  contraction (block integrator, src/ad_intg.hpp:700-727) and the single-space arithmetic as written
  (src/ad_intg.hpp:283-326), the latter restated in sympy from the reference's index arithmetic;
* closed forms of a NONLINEAR energy where the integrand stays polynomial: MinimalSurfaceEnergy (ex2.cpp:12-24) on an
  affine element at a LINEAR state u = a x + b y (grad u constant): residual and Jacobian;
* the published 1-D Gauss-Lobatto / Gauss-Legendre abscissae and weights in radicals;
* a 2x2-element assembled Q1 stiffness matrix on the unit square (the 9-point stencil 8/3, -1/3) -- pins the scatter.

Every matrix is row-major [i][j] over the LEXICOGRAPHIC tensor dofs (x fastest); vector fields byNODES
(component-major).  Values are stored as doubles (sympy rationals / radicals evaluated to 30 digits, then rounded).
"""
import json
import os

import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
xi, eta = sp.symbols("xi eta")


def lagrange(nodes, t):
    out = []
    for j, xj in enumerate(nodes):
        e = sp.Integer(1)
        for k, xk in enumerate(nodes):
            if k != j:
                e *= (t - xk) / (xj - xk)
        out.append(sp.expand(e))
    return out


def gll_nodes(p):
    """closed Gauss-Lobatto nodes on [0,1] for order p (p+1 nodes): roots of (1-x^2) P_p'(x), mapped"""
    if p == 1:
        return [sp.Integer(0), sp.Integer(1)]
    x = sp.symbols("x")
    r = sp.solve(sp.diff(sp.legendre(p, x), x), x)
    pts = sorted([sp.Integer(-1)] + [sp.simplify(v) for v in r] + [sp.Integer(1)], key=lambda v: float(v))
    return [sp.simplify((v + 1) / 2) for v in pts]


def gl_rule(n):
    x = sp.symbols("x")
    P = sp.legendre(n, x)
    r = sorted(sp.solve(P, x), key=lambda v: float(v))
    w = [sp.simplify(2 / ((1 - v ** 2) * sp.diff(P, x).subs(x, v) ** 2)) for v in r]
    return [sp.simplify((v + 1) / 2) for v in r], [sp.simplify(v / 2) for v in w]


def tensor_basis(p):
    N1x, N1y = lagrange(gll_nodes(p), xi), lagrange(gll_nodes(p), eta)
    return [sp.expand(N1x[i] * N1y[j]) for j in range(p + 1) for i in range(p + 1)]  # x fastest


def geometry(X):
    """bilinear map through the 4 vertices X (lexicographic): returns (x(xi,eta), J, detJ)"""
    N = [(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta]
    x = [sp.expand(sum(N[k] * X[k][d] for k in range(4))) for d in range(2)]
    J = sp.Matrix([[sp.diff(x[0], xi), sp.diff(x[0], eta)], [sp.diff(x[1], xi), sp.diff(x[1], eta)]])
    return x, J, sp.expand(J.det())


def integrate_ref(e):
    return sp.integrate(sp.integrate(sp.expand(e), (xi, 0, 1)), (eta, 0, 1))


def phys_grads(phi, J):
    Jit = J.inv().T
    return [Jit * sp.Matrix([sp.diff(f, xi), sp.diff(f, eta)]) for f in phi]


def fl(e):
    return float(sp.N(e, 30))


def mat(M):
    return [[fl(v) for v in row] for row in M]


def stiffness(p, X, K=None):
    phi = tensor_basis(p)
    _, J, det = geometry(X)
    g = phys_grads(phi, J)
    Km = sp.eye(2) if K is None else sp.Matrix(K)
    n = len(phi)
    return [[integrate_ref(sp.simplify((g[i].T * Km * g[j])[0, 0] * det)) for j in range(n)] for i in range(n)]


def mass(p, X):
    phi = tensor_basis(p)
    _, _, det = geometry(X)
    n = len(phi)
    return [[integrate_ref(phi[i] * phi[j] * det) for j in range(n)] for i in range(n)]


def elasticity_hessian(lam, mu):
    """Hessian of 0.5 lam (tr G)^2 + mu |sym G|^2 w.r.t. x[s + 2 c] = d u_c / d x_s  (src/ad_native.hpp:550-565,
    gradu[i*dim+j]); 4 x 4, constant."""
    xs = sp.symbols("g0:4")
    G = [[xs[0], xs[1]], [xs[2], xs[3]]]  # G[c][s] = x[s + 2 c]
    tr = G[0][0] + G[1][1]
    sym = [[(G[i][j] + G[j][i]) / 2 for j in range(2)] for i in range(2)]
    e = sp.Rational(1, 2) * lam * tr ** 2 + mu * sum(sym[i][j] ** 2 for i in range(2) for j in range(2))
    return sp.hessian(e, xs)


def elasticity_matrices(X, lam, mu):
    """order-1 vector space (vdim 2, byNODES): consistent contraction and the reference's single-space arithmetic"""
    phi = tensor_basis(1)
    _, J, det = geometry(X)
    g = phys_grads(phi, J)
    H = elasticity_hessian(lam, mu)
    dof, sd, vd = 4, 2, 2
    cons = [[0] * 8 for _ in range(8)]
    ref = [[0] * 8 for _ in range(8)]
    for c in range(vd):
        for r in range(vd):
            for i in range(dof):
                for j in range(dof):
                    e = sum(g[i][t] * H[t + sd * c, s1 + sd * r] * g[j][s1] for t in range(sd) for s1 in range(sd))
                    cons[c * dof + i][r * dof + j] = integrate_ref(sp.simplify(e * det))
    # src/ad_intg.hpp:292,:312-324: Hs = H viewed [sd x vd*sd*vd] (column-major data), windows of width sd at
    # (c*vd + r)*sd, part(i,j) = sum_t B(i,t) sum_s1 B(j,s1) Hs(s1, (c*vd+r)*sd + t); added at (c,r) and, untransposed, at (r,c)
    Hdata = [H[a, b] for b in range(4) for a in range(4)]  # column-major
    for c in range(vd):
        for r in range(c + 1):
            for i in range(dof):
                for j in range(dof):
                    e = 0
                    for t in range(sd):
                        k = (c * vd + r) * sd + t
                        for s1 in range(sd):
                            e += g[i][t] * g[j][s1] * Hdata[s1 + sd * k]
                    v = integrate_ref(sp.simplify(e * det))
                    ref[c * dof + i][r * dof + j] += v
                    if c != r:
                        ref[r * dof + i][c * dof + j] += v
    return cons, ref


def minsurf_linear_state(p, X, a, b, eps):
    """MinimalSurfaceEnergy f(g) = sqrt(1 + |g|^2) + eps |g|^2 at u = a x + b y on an affine element"""
    phi = tensor_basis(p)
    x, J, det = geometry(X)
    g = phys_grads(phi, J)
    s = a * a + b * b
    root = sp.sqrt(1 + s)
    df = sp.Matrix([a, b]) * (1 / root + 2 * eps)
    Hf = sp.eye(2) * (1 / root + 2 * eps) - sp.Matrix([[a * a, a * b], [a * b, b * b]]) / root ** 3
    n = len(phi)
    res = [integrate_ref(sp.simplify((df.T * g[i])[0, 0] * det)) for i in range(n)]
    jac = [[integrate_ref(sp.simplify((g[i].T * Hf * g[j])[0, 0] * det)) for j in range(n)] for i in range(n)]
    nodes = gll_nodes(p)
    u = [sp.simplify(a * x[0].subs({xi: nodes[i], eta: nodes[j]}) + b * x[1].subs({xi: nodes[i], eta: nodes[j]}))
         for j in range(p + 1) for i in range(p + 1)]
    energy = integrate_ref((root + eps * s) * det)
    return u, res, jac, energy


def main():
    R = sp.Rational
    rect = [(0, 0), (R(3, 2), 0), (0, R(4, 5)), (R(3, 2), R(4, 5))]
    para = [(0, 0), (R(3, 2), R(1, 4)), (R(2, 5), R(4, 5)), (R(19, 10), R(21, 20))]       # affine: X3 = X1 + X2 - X0
    quad = [(0, 0), (R(3, 2), R(1, 10)), (R(-1, 5), R(4, 5)), (R(9, 5), R(13, 10))]        # general (non-affine)
    out = {"note": "see make_fe_closed_forms.py; matrices row-major over lexicographic dofs", "cases": []}

    def add(name, **kw):
        kw["name"] = name
        out["cases"].append(kw)

    for p in (1, 2):
        for gname, X in (("rectangle", rect), ("parallelogram", para)):
            add("stiffness", order=p, geometry=gname, vertices=[[fl(v) for v in P] for P in X], matrix=mat(stiffness(p, X)))
            add("mass", order=p, geometry=gname, vertices=[[fl(v) for v in P] for P in X], matrix=mat(mass(p, X)))
        add("mass", order=p, geometry="general quadrilateral", vertices=[[fl(v) for v in P] for P in quad], matrix=mat(mass(p, quad)))
    Kfull = [[R(3, 2), R(3, 10)], [R(3, 10), R(9, 10)]]
    add("diffusion_fullK", order=1, geometry="parallelogram", vertices=[[fl(v) for v in P] for P in para],
        K_colmajor=[fl(Kfull[0][0]), fl(Kfull[1][0]), fl(Kfull[0][1]), fl(Kfull[1][1])], matrix=mat(stiffness(1, para, Kfull)))
    lam, mu = R(2), R(7, 10)
    for gname, X in (("rectangle", rect), ("parallelogram", para)):
        cons, ref = elasticity_matrices(X, lam, mu)
        add("elasticity_q1", geometry=gname, vertices=[[fl(v) for v in P] for P in X], lam=fl(lam), mu=fl(mu),
            consistent=mat(cons), as_written=mat(ref))
    for p in (1, 2):
        a, b, eps = R(3, 4), R(-1, 2), R(1, 2)
        u, res, jac, en = minsurf_linear_state(p, para, a, b, eps)
        add("minsurf_linear_state", order=p, geometry="parallelogram", vertices=[[fl(v) for v in P] for P in para],
            a=fl(a), b=fl(b), eps=fl(eps), state=[fl(v) for v in u], residual=[fl(v) for v in res], jacobian=mat(jac), energy=fl(en))
    # 2x2 elements on the unit square, Q1: assembled stiffness = the 9-point stencil (8/3 centre, -1/3 neighbours)
    h = R(1, 2)
    Ke = stiffness(1, [(0, 0), (h, 0), (0, h), (h, h)])
    A = [[0] * 9 for _ in range(9)]
    for ey in range(2):
        for ex in range(2):
            loc = [(ex + i) + 3 * (ey + j) for j in range(2) for i in range(2)]
            for i in range(4):
                for j in range(4):
                    A[loc[i]][loc[j]] += Ke[i][j]
    add("assembled_stiffness_q1_2x2", matrix=mat(A))
    out["gauss_lobatto_01"] = {str(p + 1): [fl(v) for v in gll_nodes(p)] for p in (1, 2, 3, 4)}
    out["gauss_legendre_01"] = {str(n): {"x": [fl(v) for v in gl_rule(n)[0]], "w": [fl(v) for v in gl_rule(n)[1]]} for n in (1, 2, 3, 4, 5)}
    # spot values in radicals, as published (Abramowitz & Stegun 25.4.29-32)
    out["published"] = {"gll4_inner": fl((1 - 1 / sp.sqrt(5)) / 2), "gll5_inner": fl((1 - sp.sqrt(R(3, 7))) / 2),
                        "gl3_outer_x": fl((1 - sp.sqrt(R(3, 5))) / 2), "gl3_w": [fl(R(5, 18)), fl(R(4, 9)), fl(R(5, 18))],
                        "gl4_x": [fl((1 - sp.sqrt(R(3, 7) + R(2, 7) * sp.sqrt(R(6, 5)))) / 2), fl((1 - sp.sqrt(R(3, 7) - R(2, 7) * sp.sqrt(R(6, 5)))) / 2)],
                        "gl4_w": [fl((18 - sp.sqrt(30)) / 72), fl((18 + sp.sqrt(30)) / 72)]}
    with open(os.path.join(HERE, "fe_closed_forms.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote fe_closed_forms.json: %d cases" % len(out["cases"]))


if __name__ == "__main__":
    main()

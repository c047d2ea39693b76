"""ctypes front-end of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by the product
package.  The structures mirror `orc_*_t` in oracle.cpp one to one.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# functional kinds (oracle.cpp enum Kind)
K_EX0, K_MASS, K_DIFFUSION, K_DIFF, K_ELASTICITY, K_MINSURF, K_OBSTACLE, \
    K_GRADOBSTACLE, K_LAGRANGIAN, K_AL, K_PG, K_LAMBDAPG, K_SHANNON, \
    K_FERMIDIRAC, K_HELLINGER, K_SIMPLEX, K_SIMP, K_PARAMCOMPLIANCE, K_EMPTY, \
    K_EX0VEC, K_LOAD = range(1, 22)

# ADEval flags (src/_ad_intg.hpp:24-36)
QVALUE, VALUE, GRAD, DIV, CURL, HESSIAN, VECTOR, VECFE = (1 << i for i in range(8))
BASIS_H1, BASIS_L2 = 0, 1
BYNODES, BYVDIM = 0, 1
PRM_CONST, PRM_GF, PRM_QF, PRM_GF_GRAD = 0, 1, 2, 3


class FnNode(C.Structure):
    _fields_ = [("kind", C.c_int), ("n_input", C.c_int), ("n_output", C.c_int),
                ("nparam", C.c_int), ("param", C.c_double * 24), ("qoff", C.c_int),
                ("nchild", C.c_int), ("child", C.c_int * 6), ("iparam", C.c_int * 8)]


class Space(C.Structure):
    _fields_ = [("basis", C.c_int), ("order", C.c_int), ("vdim", C.c_int),
                ("mode", C.c_int), ("ordering", C.c_int), ("ndofs", C.c_int),
                ("e2l", C.POINTER(C.c_int))]


class Mesh(C.Structure):
    _fields_ = [("dim", C.c_int), ("ne", C.c_int), ("geom_order", C.c_int),
                ("nnodes", C.c_int), ("e2n", C.POINTER(C.c_int)),
                ("coords", C.POINTER(C.c_double))]


class Param(C.Structure):
    _fields_ = [("type", C.c_int), ("size", C.c_int), ("space", Space),
                ("data", C.POINTER(C.c_double))]


class Form(C.Structure):
    _fields_ = [("mesh", Mesh), ("nspaces", C.c_int), ("spaces", C.POINTER(Space)),
                ("fn", C.POINTER(FnNode)), ("root", C.c_int), ("quad_order", C.c_int),
                ("nparams", C.c_int), ("params", C.POINTER(Param)), ("block", C.c_int),
                ("ness", C.c_int), ("ess", C.POINTER(C.c_int))]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        fp, np_ = C.POINTER(Form), C.POINTER(FnNode)
        L.orc_fn_value.restype = C.c_double
        L.orc_fn_value.argtypes = [np_, C.c_int, dp, dp]
        L.orc_fn_gradient.argtypes = [np_, C.c_int, dp, dp, dp]
        L.orc_fn_hessian.argtypes = [np_, C.c_int, dp, dp, dp]
        L.orc_vecfn_value.argtypes = [np_, C.c_int, dp, dp]
        L.orc_vecfn_gradient.argtypes = [np_, C.c_int, dp, dp]
        L.orc_vecfn_hessian.argtypes = [np_, C.c_int, dp, dp]
        L.orc_pg_step.restype = C.c_double
        L.orc_pg_step.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_gauss_legendre.argtypes = [C.c_int, dp, dp]
        L.orc_gauss_lobatto.argtypes = [C.c_int, dp]
        L.orc_rule_npts_1d.restype = C.c_int
        L.orc_rule_npts_1d.argtypes = [C.c_int]
        L.orc_form_sizes.argtypes = [fp, ip, ip, ip, ip]
        L.orc_element_energy.restype = C.c_double
        L.orc_element_energy.argtypes = [fp, C.c_int, dp]
        L.orc_element_vector.argtypes = [fp, C.c_int, dp, dp]
        L.orc_element_grad.argtypes = [fp, C.c_int, dp, dp]
        L.orc_form_energy.restype = C.c_double
        L.orc_form_energy.argtypes = [fp, dp]
        L.orc_form_mult.argtypes = [fp, dp, dp]
        L.orc_form_mult_range.argtypes = [fp, dp, dp, C.c_int, C.c_int, C.c_int]
        L.orc_form_pattern.restype = C.c_long
        L.orc_form_pattern.argtypes = [fp, ip, ip]
        L.orc_form_grad.argtypes = [fp, dp, ip, ip, dp]
        L.orc_form_grad_range.argtypes = [fp, dp, ip, ip, dp, C.c_int, C.c_int, C.c_int]
        L.orc_form_inputs_at_qpts.argtypes = [fp, dp, dp]
        L.orc_form_coefficient.argtypes = [fp, dp, C.c_int, dp]
        L.orc_mmto_param_gradient.argtypes = [fp, dp, C.c_int, ip, dp]
        L.orc_dofpg_nodal.argtypes = [np_, C.c_int, C.c_int, C.c_int, ip, dp, C.c_double, dp, dp, dp, dp, dp, dp, dp]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Functional:
    """A tree of functional nodes; `nodes[root]` is the functional itself."""

    def __init__(self):
        self.nodes = []

    def add(self, kind, n_input, params=(), iparams=(), children=(), qoff=-1, n_output=0):
        nd = FnNode()
        nd.kind, nd.n_input, nd.n_output = kind, n_input, n_output
        nd.nparam = len(params)
        for i, p in enumerate(params):
            nd.param[i] = float(p)
        for i, p in enumerate(iparams):
            nd.iparam[i] = int(p)
        nd.nchild = len(children)
        for i, c in enumerate(children):
            nd.child[i] = int(c)
        nd.qoff = qoff
        self.nodes.append(nd)
        return len(self.nodes) - 1

    def carray(self):
        arr = (FnNode * len(self.nodes))(*self.nodes)
        return arr

    @property
    def root(self):
        return len(self.nodes) - 1

    # --- AD layer (src/ad_native.cpp:181-276) ---
    def value(self, x, qprm=None):
        x = _f64(x)
        q = _f64(qprm) if qprm is not None else None
        return lib().orc_fn_value(self.carray(), self.root, _dp(x), _dp(q) if q is not None else None)

    def gradient(self, x, qprm=None):
        x = _f64(x)
        q = _f64(qprm) if qprm is not None else None
        J = np.zeros(self.nodes[self.root].n_input)
        lib().orc_fn_gradient(self.carray(), self.root, _dp(x), _dp(q) if q is not None else None, _dp(J))
        return J

    def hessian(self, x, qprm=None):
        x = _f64(x)
        q = _f64(qprm) if qprm is not None else None
        n = self.nodes[self.root].n_input
        H = np.zeros((n, n))
        lib().orc_fn_hessian(self.carray(), self.root, _dp(x), _dp(q) if q is not None else None, _dp(H))
        return H.T.copy()  # column-major -> numpy (symmetric anyway)

    def vec_value(self, x):
        x = _f64(x)
        F = np.zeros(self.nodes[self.root].n_output)
        lib().orc_vecfn_value(self.carray(), self.root, _dp(x), _dp(F))
        return F

    def vec_gradient(self, x):
        x = _f64(x)
        n, m = self.nodes[self.root].n_input, self.nodes[self.root].n_output
        J = np.zeros((n, m))  # column-major [m x n]
        lib().orc_vecfn_gradient(self.carray(), self.root, _dp(x), _dp(J))
        return J.T.copy()  # [m, n]

    def vec_hessian(self, x):
        x = _f64(x)
        n, m = self.nodes[self.root].n_input, self.nodes[self.root].n_output
        H = np.zeros((m, n, n))  # H(i,j,k) at i + n j + n n k
        lib().orc_vecfn_hessian(self.carray(), self.root, _dp(x), _dp(H))
        return H


class OracleForm:
    """NonlinearForm / BlockNonlinearForm with one AD integrator (oracle side).

    mesh:   dict(dim, e2n[ne,nn] int32, coords[nnodes,dim] float64, geom_order)
    spaces: list of dict(basis, order, vdim, mode, ordering, ndofs, e2l[ne,nd])
    params: list of dict(type, size, data, space=None)
    """

    def __init__(self, mesh, spaces, functional, quad_order=-1, params=(), block=None, ess=()):
        self._keep = []
        self.fn = functional
        F = Form()
        e2n, coords = _i32(mesh["e2n"]), _f64(mesh["coords"])
        self._keep += [e2n, coords]
        F.mesh.dim, F.mesh.ne = mesh["dim"], e2n.shape[0]
        F.mesh.geom_order = mesh.get("geom_order", 1)
        F.mesh.nnodes = coords.shape[0]
        F.mesh.e2n, F.mesh.coords = _ip(e2n), _dp(coords)
        self.ne = e2n.shape[0]
        sp = (Space * len(spaces))()
        for i, s in enumerate(spaces):
            self._fill_space(sp[i], s)
        self._keep.append(sp)
        F.nspaces, F.spaces = len(spaces), sp
        self._nodes = functional.carray()
        F.fn, F.root = self._nodes, functional.root
        F.quad_order = quad_order
        pp = (Param * max(len(params), 1))()
        for i, p in enumerate(params):
            pp[i].type, pp[i].size = p["type"], p["size"]
            d = _f64(p["data"])
            self._keep.append(d)
            pp[i].data = _dp(d)
            if p.get("space") is not None:
                self._fill_space(pp[i].space, p["space"])
        self._keep.append(pp)
        F.nparams, F.params = len(params), pp
        F.block = int(block if block is not None else len(spaces) > 1)
        essa = _i32(np.asarray(ess, dtype=np.int32))
        self._keep.append(essa)
        F.ness, F.ess = essa.size, _ip(essa)
        self.F = F
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        lib().orc_form_sizes(C.byref(F), C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        self.n_input, self.nvd_el, self.nq, self.ndof = a.value, b.value, c.value, d.value
        self._pattern = None

    def _fill_space(self, S, s):
        e2l = _i32(s["e2l"])
        self._keep.append(e2l)
        S.basis, S.order, S.vdim = s["basis"], s["order"], s.get("vdim", 1)
        S.mode, S.ordering, S.ndofs = s.get("mode", 0), s.get("ordering", BYNODES), s["ndofs"]
        S.e2l = _ip(e2l)

    def rule(self):
        """Reference points [nq, dim] and weights [nq] of the form's integration rule."""
        dim = self.F.mesh.dim
        pts, w = np.zeros((self.nq, dim)), np.zeros(self.nq)
        lib().orc_form_rule(C.byref(self.F), _dp(pts), _dp(w))
        return pts, w

    def energy(self, x):
        x = _f64(x)
        return lib().orc_form_energy(C.byref(self.F), _dp(x))

    def mult(self, x, e0=None, e1=None):
        x = _f64(x)
        y = np.zeros(self.ndof)
        if e0 is None:
            lib().orc_form_mult(C.byref(self.F), _dp(x), _dp(y))
        else:
            lib().orc_form_mult_range(C.byref(self.F), _dp(x), _dp(y), e0, e1, 0)
        return y

    def pattern(self):
        if self._pattern is None:
            nnz = lib().orc_form_pattern(C.byref(self.F), None, None)
            rowptr = np.zeros(self.ndof + 1, dtype=np.int32)
            colidx = np.zeros(nnz, dtype=np.int32)
            lib().orc_form_pattern(C.byref(self.F), _ip(rowptr), _ip(colidx))
            self._pattern = (rowptr, colidx)
        return self._pattern

    def grad(self, x, e0=None, e1=None):
        """Returns (rowptr, colidx, vals) with sorted columns."""
        x = _f64(x)
        rowptr, colidx = self.pattern()
        vals = np.zeros(colidx.size)
        if e0 is None:
            lib().orc_form_grad(C.byref(self.F), _dp(x), _ip(rowptr), _ip(colidx), _dp(vals))
        else:
            lib().orc_form_grad_range(C.byref(self.F), _dp(x), _ip(rowptr), _ip(colidx), _dp(vals), e0, e1, 0)
        return rowptr, colidx, vals

    def element_vector(self, e, x):
        x = _f64(x)
        v = np.zeros(self.nvd_el)
        lib().orc_element_vector(C.byref(self.F), e, _dp(x), _dp(v))
        return v

    def element_grad(self, e, x):
        x = _f64(x)
        m = np.zeros((self.nvd_el, self.nvd_el))
        lib().orc_element_grad(C.byref(self.F), e, _dp(x), _dp(m))
        return m.T.copy()  # column-major -> [row, col]

    def inputs_at_qpts(self, x):
        x = _f64(x)
        out = np.zeros((self.ne, self.nq, self.n_input))
        lib().orc_form_inputs_at_qpts(C.byref(self.F), _dp(x), _dp(out))
        return out

    def mmto_param_gradient(self, rho, fi):
        rho = _f64(rho)
        fi = _i32(fi)
        out = np.zeros((self.ne, self.nq, self.n_input))
        lib().orc_mmto_param_gradient(C.byref(self.F), _dp(rho), fi.size, _ip(fi), _dp(out))
        return out

    def coefficient(self, x, which):
        x = _f64(x)
        n = self.n_input
        shape = {0: (self.ne, self.nq), 1: (self.ne, self.nq, n), 2: (self.ne, self.nq, n, n)}[which]
        out = np.zeros(shape)
        lib().orc_form_coefficient(C.byref(self.F), _dp(x), which, _dp(out))
        return out


def dofpg_nodal(functional, e2l, w, alpha, u, psi, psik):
    """Nodal PG terms of ADDofPGNonlinearFormIntegrator (src/dof_pg.hpp); returns r_u, r_psi, d_pp, d_up."""
    e2l, w = _i32(e2l), _f64(w)
    u, psi, psik = _f64(u), _f64(psi), _f64(psik)
    n = u.size
    out = [np.zeros(n) for _ in range(4)]
    lib().orc_dofpg_nodal(functional.carray(), functional.root, e2l.shape[0], e2l.shape[1], _ip(e2l), _dp(w), alpha,
                          _dp(u), _dp(psi), _dp(psik), *[_dp(o) for o in out])
    return out


def pg_step(rule, alpha0, max_alpha, ratio, ratio2, it):
    return lib().orc_pg_step(rule, alpha0, max_alpha, ratio, ratio2, it)


def gauss_legendre(n):
    x, w = np.zeros(n), np.zeros(n)
    lib().orc_gauss_legendre(n, _dp(x), _dp(w))
    return x, w


def gauss_lobatto(n):
    x = np.zeros(n)
    lib().orc_gauss_lobatto(n, _dp(x))
    return x

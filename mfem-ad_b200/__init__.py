"""mfem-ad element-level AD assembly hot path, B200-native.

Python face of the C ABI in include/mfemad_b200.h (ctypes; no torch types cross
the boundary).  The classes mirror the reference's interface for this path:

    Functional        <-> ADFunction and subclasses (src/ad_native.hpp, src/pg.hpp, src/mmto.hpp)
    Integrator        <-> AD(Block)NonlinearFormIntegrator<modes...> attached to a
                          (Block)NonlinearForm: Mult / GetGradient / GetEnergy
    PGStepSizeRule    <-> src/pg.hpp:10-34, src/pg.cpp:4-54

There is NO CPU fallback: without the compiled CUDA library or without a GPU
every compute entry point raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("MADB_LIB", "libmadb.so"))  # MADB_LIB: a variant build (tools/variant.sh)

QVALUE, VALUE, GRAD, DIV, CURL, HESSIAN, VECTOR, VECFE = (1 << i for i in range(8))
BASIS_H1, BASIS_L2 = 0, 1
BYNODES, BYVDIM = 0, 1
ROLE_INPUT, ROLE_PARAM = 0, 1
INTEG_BLOCK = 1
PARAMGRAD_AS_WRITTEN, PARAMGRAD_DERIVATIVE = 0, 1

_lib = None


class MadbError(RuntimeError):
    pass


def lib():
    """Loads libmadb.so; fails loudly when the CUDA extension is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MadbError("CUDA extension %s is missing: run `python mfem-ad_b200/build.py` "
                            "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, dp, ip = C.c_void_p, C.c_void_p, C.c_void_p
        pp = C.POINTER(C.c_void_p)
        L.madb_version.restype = C.c_int
        L.madb_last_error.restype = C.c_char_p
        L.madb_ctx_create.argtypes = [C.c_int, pp]
        L.madb_ctx_destroy.argtypes = [vp]
        L.madb_ctx_sync.argtypes = [vp]
        L.madb_ctx_stream.restype = C.c_void_p
        L.madb_ctx_stream.argtypes = [vp]
        L.madb_mesh_create.argtypes = [vp, C.c_int, C.c_int, ip, C.c_int, dp, pp]
        L.madb_mesh_create_simplex.argtypes = [vp, C.c_int, C.c_int, ip, C.c_int, dp, pp]
        L.madb_mesh_destroy.argtypes = [vp]
        L.madb_space_create.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, pp]
        L.madb_space_destroy.argtypes = [vp]
        L.madb_functional_create.argtypes = [vp, C.c_char_p, C.c_int, dp, C.c_int, ip, C.c_int, pp, pp]
        L.madb_functional_set_params.argtypes = [vp, C.c_int, dp]
        L.madb_functional_destroy.argtypes = [vp]
        L.madb_functional_eval.argtypes = [vp, vp, C.c_int, C.c_int, dp, dp, dp, dp, dp]
        L.madb_dofpg_nodal.argtypes = [vp, vp, C.c_int, C.c_double, dp, dp, dp, dp, dp, dp, dp, dp]
        L.madb_pack.argtypes = [vp, C.c_int, ip, dp, dp]
        L.madb_unpack.argtypes = [vp, C.c_int, ip, dp, dp, C.c_int]
        L.madb_unpack_multi.argtypes = [vp, C.c_int, ip, ip, dp, dp, C.c_int]
        L.madb_comm_unique_id.argtypes = [C.c_char_p]
        L.madb_comm_create.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, pp]
        L.madb_comm_destroy.argtypes = [vp]
        L.madb_comm_allreduce_sum.argtypes = [vp, C.c_int, dp]
        L.madb_exchange_create.argtypes = [vp, C.c_int, ip, ip, ip, C.c_int, ip, ip, ip, pp]
        L.madb_exchange_destroy.argtypes = [vp]
        L.madb_exchange_begin.argtypes = [vp, dp, C.c_int]
        L.madb_exchange_end.argtypes = [vp, dp, C.c_int]
        L.madb_lvpp_update.argtypes = [vp, C.c_int, C.c_double, dp, dp, dp, dp, C.POINTER(C.c_double)]
        L.madb_integrator_create.argtypes = [vp, C.c_int, pp, ip, ip, vp, C.c_int, pp]
        L.madb_integrator_create_ex.argtypes = [vp, C.c_int, pp, ip, ip, vp, C.c_int, C.c_int, pp]
        L.madb_integrator_destroy.argtypes = [vp]
        L.madb_integrator_coefficient_hessian.argtypes = [vp, dp, dp, dp, dp]
        L.madb_integrator_param_gradient.argtypes = [vp, dp, dp, dp, C.c_int]
        L.madb_integrator_qpoint_coords.argtypes = [vp, dp]
        L.madb_integrator_sizes.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.madb_integrator_patch_stats.argtypes = [vp, C.POINTER(C.c_int64)]
        L.madb_vecfunction_eval.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp]
        L.madb_patch_selftest.argtypes = [C.c_int, C.c_int, ip, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.c_int, ip,
                                          C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.madb_patch_selftest_img.argtypes = [C.c_int, C.c_int, ip, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.c_int, ip, C.c_int,
                                              C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.madb_integrator_set_timing.argtypes = [vp, C.c_int]
        L.madb_integrator_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_double)]
        L.madb_integrator_set_param_field.argtypes = [vp, C.c_int, dp]
        L.madb_integrator_set_param_qf.argtypes = [vp, C.c_int, dp]
        L.madb_integrator_set_essential.argtypes = [vp, C.c_int, ip]
        L.madb_integrator_energy.argtypes = [vp, dp, C.POINTER(C.c_double)]
        L.madb_integrator_mult.argtypes = [vp, dp, dp]
        L.madb_integrator_pattern.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), ip, ip]
        L.madb_integrator_grad_assemble.argtypes = [vp, dp, dp]
        L.madb_integrator_assemble.argtypes = [vp, dp, dp, dp]
        L.madb_integrator_grad_mult.argtypes = [vp, dp, dp, dp]
        L.madb_integrator_assemble_begin.argtypes = [vp, dp, dp, dp]
        L.madb_integrator_assemble_end.argtypes = [vp]
        L.madb_integrator_coefficient.argtypes = [vp, dp, dp, dp]
        L.madb_solver_create.argtypes = [vp, pp]
        L.madb_solver_destroy.argtypes = [vp]
        L.madb_solver_pcg.argtypes = [vp, dp, dp, dp, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.madb_solver_condensed_pcg.argtypes = [vp, C.c_int, C.c_int, dp, dp, dp, C.c_double, C.c_double, C.c_int,
                                                C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.madb_solver_pg_minres.argtypes = [vp, C.c_int, C.c_int, dp, dp, dp, C.c_double, C.c_double, C.c_int,
                                            C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.madb_csr_spmv.argtypes = [vp, dp, dp, dp]
        L.madb_registry_has.argtypes = [C.c_char_p]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise MadbError(lib().madb_last_error().decode())


def _ptr(a):
    """numpy array -> host pointer; torch tensor -> its data pointer (host or device); int passes through."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return int(a)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Context:
    def __init__(self, device=0):
        h = C.c_void_p()
        _check(lib().madb_ctx_create(device, C.byref(h)))
        self.h, self.device = h, device

    def sync(self):
        _check(lib().madb_ctx_sync(self.h))

    @property
    def stream(self):
        return lib().madb_ctx_stream(self.h)

    def __del__(self):
        try:
            lib().madb_ctx_destroy(self.h)
        except Exception:
            pass


class Mesh:
    def __init__(self, ctx, mesh):
        e2n, coords = _i32(mesh["e2n"]), _f64(mesh["coords"])
        h = C.c_void_p()
        create = lib().madb_mesh_create_simplex if mesh.get("simplex") else lib().madb_mesh_create
        _check(create(ctx.h, mesh["dim"], e2n.shape[0], e2n.ctypes.data, coords.shape[0], coords.ctypes.data, C.byref(h)))
        self.h, self.ctx, self.dim, self.ne = h, ctx, mesh["dim"], e2n.shape[0]

    def __del__(self):
        try:
            lib().madb_mesh_destroy(self.h)
        except Exception:
            pass


class Space:
    def __init__(self, ctx, mesh, space):
        e2l = _i32(space["e2l"])
        h = C.c_void_p()
        _check(lib().madb_space_create(ctx.h, mesh.h, space["basis"], space["order"], space.get("vdim", 1),
                                       space.get("ordering", BYNODES), space["ndofs"], e2l.ctypes.data, C.byref(h)))
        self.h, self.mesh, self.desc = h, mesh, space
        self.vsize = space["ndofs"] * space.get("vdim", 1)

    def __del__(self):
        try:
            lib().madb_space_destroy(self.h)
        except Exception:
            pass


class Functional:
    """ADFunction handle: kind + run-time constants + structural integers + children."""

    def __init__(self, ctx, kind, params=(), iparams=(), children=()):
        self.ctx, self.kind, self.children = ctx, kind, list(children)
        p, ip_ = _f64(list(params)), _i32(list(iparams))
        ch = (C.c_void_p * max(len(children), 1))(*[c.h for c in children])
        h = C.c_void_p()
        _check(lib().madb_functional_create(ctx.h, kind.encode(), p.size, p.ctypes.data, ip_.size, ip_.ctypes.data,
                                            len(children), ch, C.byref(h)))
        self.h = h

    def set_params(self, params):
        p = _f64(list(params))
        _check(lib().madb_functional_set_params(self.h, p.size, p.ctypes.data))

    def eval(self, x, qprm=None):
        """value, gradient, Hessian at points x[npts, n] on the device (src/ad_native.cpp:181-230)."""
        x = _f64(np.atleast_2d(x))
        npts, n = x.shape
        q = _f64(qprm) if qprm is not None else None
        v, g, h = np.zeros(npts), np.zeros((npts, n)), np.zeros((npts, n, n))
        _check(lib().madb_functional_eval(self.ctx.h, self.h, n, npts, x.ctypes.data, _ptr(q), v.ctypes.data,
                                          g.ctypes.data, h.ctypes.data))
        return v, g, h

    def eval_vector(self, x, n_output):
        """ADVectorFunction: value [npts, m], Jacobian [npts, m, n], Hessians [npts, m, n, n] (src/ad_native.cpp:232-276)."""
        x = _f64(np.atleast_2d(x))
        npts, n = x.shape
        v, J, H = np.zeros((npts, n_output)), np.zeros((npts, n_output, n)), np.zeros((npts, n_output, n, n))
        _check(lib().madb_vecfunction_eval(self.ctx.h, self.h, n, n_output, npts, x.ctypes.data, v.ctypes.data, J.ctypes.data,
                                           H.ctypes.data))
        return v, J, H

    def eval_device(self, x, value=None, grad=None, hess=None, qprm=None):
        """Same on buffers that already live on the device (torch tensors): x[npts, n]; outputs may be None."""
        npts, n = x.shape
        _check(lib().madb_functional_eval(self.ctx.h, self.h, n, npts, _ptr(x), _ptr(qprm), _ptr(value), _ptr(grad), _ptr(hess)))

    def __del__(self):
        try:
            lib().madb_functional_destroy(self.h)
        except Exception:
            pass


class Integrator:
    """A (Block)NonlinearForm holding one AD(Block)NonlinearFormIntegrator.

    fields: list of (Space, mode[, role]); input fields are the blocks of x in order."""

    def __init__(self, ctx, fields, functional, quad_order=-1, block=False):
        """block=True: ADBlockNonlinearFormIntegrator semantics even with one input space (MADB_INTEG_BLOCK): only
        matters for one VECTOR space, whose Jacobian otherwise follows the reference's single-space arithmetic
        (src/ad_intg.hpp:310-326)."""
        self.ctx, self.fn = ctx, functional
        self.fields = [(f[0], f[1], f[2] if len(f) > 2 else ROLE_INPUT) for f in fields]
        n = len(self.fields)
        sp = (C.c_void_p * n)(*[f[0].h for f in self.fields])
        modes = _i32([f[1] for f in self.fields])
        roles = _i32([f[2] for f in self.fields])
        h = C.c_void_p()
        _check(lib().madb_integrator_create_ex(ctx.h, n, sp, modes.ctypes.data, roles.ctypes.data, functional.h,
                                               quad_order, INTEG_BLOCK if block else 0, C.byref(h)))
        self.h = h
        nt, nq, nc = C.c_int64(), C.c_int(), C.c_int()
        _check(lib().madb_integrator_sizes(h, C.byref(nt), C.byref(nq), C.byref(nc)))
        self.ntotal, self.nq_el, self.ncolors = nt.value, nq.value, nc.value
        self._pattern = None
        self._keep = []

    def set_timing(self, on=True):
        _check(lib().madb_integrator_set_timing(self.h, 1 if on else 0))

    def last_kernel_ms(self):
        """Device time of the element kernel(s) of the last call (set_timing(True) first)."""
        ms = C.c_double()
        _check(lib().madb_integrator_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def patch_stats(self):
        """Patch-assembly diagnostics (madb_integrator_patch_stats)."""
        out = (C.c_int64 * 8)()
        _check(lib().madb_integrator_patch_stats(self.h, out))
        keys = ("patches", "max_rows", "max_slots", "ifc_dofs", "ifc_entries", "staged_y", "staged_vals", "runs")
        return dict(zip(keys, [int(v) for v in out]))

    def set_param_field(self, field, dofs):
        if isinstance(dofs, np.ndarray):
            dofs = _f64(dofs)
        self._keep.append(dofs)
        _check(lib().madb_integrator_set_param_field(self.h, field, _ptr(dofs)))

    def set_param_qf(self, qf):
        qf = _f64(qf)
        _check(lib().madb_integrator_set_param_qf(self.h, qf.shape[-1], qf.ctypes.data))

    def set_essential(self, dofs):
        d = _i32(dofs)
        _check(lib().madb_integrator_set_essential(self.h, d.size, d.ctypes.data))

    def energy(self, x):
        x = _f64(x) if isinstance(x, np.ndarray) else x
        e = C.c_double()
        _check(lib().madb_integrator_energy(self.h, _ptr(x), C.byref(e)))
        return e.value

    def mult(self, x, y=None):
        if isinstance(x, np.ndarray):
            x = _f64(x)
            y = np.empty(self.ntotal) if y is None else y
        _check(lib().madb_integrator_mult(self.h, _ptr(x), _ptr(y)))
        return y

    def pattern(self):
        if self._pattern is None:
            nr, nnz = C.c_int64(), C.c_int64()
            _check(lib().madb_integrator_pattern(self.h, C.byref(nr), C.byref(nnz), None, None))
            rowptr, colidx = np.zeros(nr.value + 1, dtype=np.int32), np.zeros(nnz.value, dtype=np.int32)
            _check(lib().madb_integrator_pattern(self.h, None, None, rowptr.ctypes.data, colidx.ctypes.data))
            self._pattern = (rowptr, colidx)
        return self._pattern

    @property
    def nnz(self):
        return self.pattern()[1].size

    def grad(self, x, vals=None):
        if isinstance(x, np.ndarray):
            x = _f64(x)
            vals = np.empty(self.nnz) if vals is None else vals
        _check(lib().madb_integrator_grad_assemble(self.h, _ptr(x), _ptr(vals)))
        return vals

    def assemble(self, x, y=None, vals=None):
        """residual + Jacobian values at the same state (one Newton iteration's assembly)."""
        if isinstance(x, np.ndarray):
            x = _f64(x)
            y = np.empty(self.ntotal) if y is None else y
            vals = np.empty(self.nnz) if vals is None else vals
        _check(lib().madb_integrator_assemble(self.h, _ptr(x), _ptr(y), _ptr(vals)))
        return y, vals

    def assemble_begin(self, x, y, vals):
        """Element kernel + residual complete; the interface reduction of the CSR values is launched by assemble_end()
        (start the exchange of y in between: it then overlaps that kernel).  Device tensors only."""
        _check(lib().madb_integrator_assemble_begin(self.h, _ptr(x), _ptr(y), _ptr(vals)))

    def assemble_end(self):
        _check(lib().madb_integrator_assemble_end(self.h))

    def coefficient(self, x, want_value=True, want_grad=True):
        """f and grad f at every quadrature point, [ne, nq] and [ne, nq, n] (DifferentiableCoefficient)."""
        x = _f64(x)
        ne = self.fields[0][0].mesh.ne
        n_in = sum((1 if m & VALUE else 0) + (s.mesh.dim if m & GRAD else 0) for s, m, r in self.fields if r == ROLE_INPUT
                   for _ in range(s.desc.get("vdim", 1)))
        val = np.empty((ne, self.nq_el)) if want_value else None
        grd = np.empty((ne, self.nq_el, n_in)) if want_grad else None
        _check(lib().madb_integrator_coefficient(self.h, _ptr(x), _ptr(val), _ptr(grd)))
        return val, grd

    def _n_in(self):
        return sum((1 if m & VALUE else 0) + (s.mesh.dim if m & GRAD else 0) for s, m, r in self.fields if r == ROLE_INPUT
                   for _ in range(s.desc.get("vdim", 1)))

    def coefficient_hessian(self, x):
        """f, grad f and the Hessian at every quadrature point (HessianCoefficient, src/ad_native.hpp:300-323)."""
        x = _f64(x)
        ne, n = self.fields[0][0].mesh.ne, self._n_in()
        val, grd, hes = np.empty((ne, self.nq_el)), np.empty((ne, self.nq_el, n)), np.empty((ne, self.nq_el, n, n))
        _check(lib().madb_integrator_coefficient_hessian(self.h, _ptr(x), _ptr(val), _ptr(grd), _ptr(hes)))
        return val, grd, hes

    def param_gradient(self, design, variant=PARAMGRAD_AS_WRITTEN):
        """ParametrizedFunctional::ParamGradient::Eval at the points (src/mmto.cpp:4-38): (F [ne,nq], J [ne,nq,param_dim]);
        variant PARAMGRAD_AS_WRITTEN reproduces the reference, PARAMGRAD_DERIVATIVE is dF/drho."""
        design = _f64(design)
        ne, n = self.fields[0][0].mesh.ne, self._n_in()
        val, J = np.empty((ne, self.nq_el)), np.empty((ne, self.nq_el, n))
        _check(lib().madb_integrator_param_gradient(self.h, _ptr(design), _ptr(val), _ptr(J), variant))
        return val, J

    def qpoint_coords(self):
        """Physical coordinates of the rule's points [ne, nq, dim] (where Coefficient-type parameters are sampled)."""
        mesh = self.fields[0][0].mesh
        xyz = np.empty((mesh.ne, self.nq_el, mesh.dim))
        _check(lib().madb_integrator_qpoint_coords(self.h, xyz.ctypes.data))
        return xyz

    def set_param_coefficient(self, *callbacks):
        """Evaluator sources of Coefficient / VectorCoefficient / MatrixCoefficient type (src/ad_native.hpp:56-61):
        host callbacks f(xyz[npts, dim]) -> [npts] or [npts, k], sampled at the rule's points and handed over as one
        QuadratureFunction (columns in the order given)."""
        xyz = self.qpoint_coords()
        pts = xyz.reshape(-1, xyz.shape[-1])
        cols = []
        for cb in callbacks:
            v = np.asarray(cb(pts), dtype=np.float64)
            cols.append(v.reshape(pts.shape[0], -1))
        qf = np.concatenate(cols, axis=1).reshape(xyz.shape[0], xyz.shape[1], -1)
        self.set_param_qf(qf)
        return qf

    def coefficient_device(self, x, value=None, grad=None):
        """Same on device buffers (torch tensors); value [ne*nq], grad [ne*nq*n] or None."""
        _check(lib().madb_integrator_coefficient(self.h, _ptr(x), _ptr(value), _ptr(grad)))

    def grad_mult(self, x, v, y=None):
        if isinstance(x, np.ndarray):
            x, v = _f64(x), _f64(v)
            y = np.empty(self.ntotal) if y is None else y
        _check(lib().madb_integrator_grad_mult(self.h, _ptr(x), _ptr(v), _ptr(y)))
        return y

    def __del__(self):
        try:
            lib().madb_integrator_destroy(self.h)
        except Exception:
            pass


def dofpg_nodal(ctx, entropy, alpha, u, psi, psik, w, r_u=None):
    """Nodal PG terms (src/dof_pg.hpp); returns r_u (accumulated into the given array), r_psi, d_pp, d_up."""
    u, psi, psik, w = _f64(u), _f64(psi), _f64(psik), _f64(w)
    n = u.size
    r_u = np.zeros(n) if r_u is None else r_u
    r_psi, d_pp, d_up = np.empty(n), np.empty(n), np.empty(n)
    _check(lib().madb_dofpg_nodal(ctx.h, entropy.h, n, alpha, _ptr(u), _ptr(psi), _ptr(psik), _ptr(w), _ptr(r_u),
                                  _ptr(r_psi), _ptr(d_pp), _ptr(d_up)))
    return r_u, r_psi, d_pp, d_up


def patch_selftest(mesh, space):
    """Host-only check of the patch-assembly maps (madb_patch_selftest): returns (max_err, stats dict)."""
    e2n, coords, e2l = _i32(mesh["e2n"]), _f64(mesh["coords"]), _i32(space["e2l"])
    err = C.c_double()
    st = (C.c_int64 * 8)()
    _check(lib().madb_patch_selftest(mesh["dim"], e2n.shape[0], e2n.ctypes.data, coords.shape[0], coords.ctypes.data,
                                     space["order"], space.get("vdim", 1), space.get("ordering", BYNODES), space["ndofs"],
                                     e2l.ctypes.data, C.byref(err), st))
    keys = ("patches", "ifc_dofs", "ifc_entries", "staged_vals", "max_blob_bytes", "nnz", "paired_entries")
    return err.value, dict(zip(keys, [int(v) for v in st]))


def patch_selftest_img(mesh, space, tpe):
    """Host-only check of the CSR-image kernel's maps (madb_patch_selftest_img): returns (max_err, stats dict)."""
    e2n, coords, e2l = _i32(mesh["e2n"]), _f64(mesh["coords"]), _i32(space["e2l"])
    err = C.c_double()
    st = (C.c_int64 * 8)()
    _check(lib().madb_patch_selftest_img(mesh["dim"], e2n.shape[0], e2n.ctypes.data, coords.shape[0], coords.ctypes.data,
                                         space["order"], space.get("vdim", 1), space.get("ordering", BYNODES), space["ndofs"],
                                         e2l.ctypes.data, tpe, C.byref(err), st))
    keys = ("patches", "ifc_dofs", "ifc_entries", "staged_vals", "smem_per_group", "nnz", "bulk_entries", "map_bytes")
    return err.value, dict(zip(keys, [int(v) for v in st]))


class Solver:
    """Linear solve of a Newton step on the device (madb_solver_*): Jacobi-PCG on SPD Jacobians, and the statically
    condensed PCG for proximal-Galerkin block systems (latent L2 block eliminated exactly).  Replaces the host direct
    solvers of the drivers (ex2.cpp:80, src/tools.hpp:128-154) so that the CSR values stay on the GPU."""

    def __init__(self, integrator):
        h = C.c_void_p()
        _check(lib().madb_solver_create(integrator.h, C.byref(h)))
        self.h, self.integrator = h, integrator
        self.n = integrator.vsize if hasattr(integrator, "vsize") else None

    def _out(self, b, x):
        if x is None:
            x = np.zeros_like(b) if isinstance(b, np.ndarray) else b.new_zeros(b.shape)
        return x

    def pcg(self, vals, b, x=None, rtol=1e-10, atol=0.0, maxit=10000):
        x = self._out(b, x)
        it, rr = C.c_int(0), C.c_double(0.0)
        _check(lib().madb_solver_pcg(self.h, _ptr(vals), _ptr(b), _ptr(x), rtol, atol, maxit, C.byref(it), C.byref(rr)))
        return x, it.value, rr.value

    def condensed_pcg(self, nh, nb, vals, b, x=None, rtol=1e-10, atol=0.0, maxit=10000):
        x = self._out(b, x)
        it, rr = C.c_int(0), C.c_double(0.0)
        _check(lib().madb_solver_condensed_pcg(self.h, nh, nb, _ptr(vals), _ptr(b), _ptr(x), rtol, atol, maxit,
                                               C.byref(it), C.byref(rr)))
        return x, it.value, rr.value

    def pg_minres(self, nh, nb, vals, b, x=None, rtol=1e-10, atol=0.0, maxit=20000):
        x = self._out(b, x)
        it, rr = C.c_int(0), C.c_double(0.0)
        _check(lib().madb_solver_pg_minres(self.h, nh, nb, _ptr(vals), _ptr(b), _ptr(x), rtol, atol, maxit,
                                           C.byref(it), C.byref(rr)))
        return x, it.value, rr.value

    def spmv(self, vals, x, y=None):
        y = self._out(x, y)
        _check(lib().madb_csr_spmv(self.h, _ptr(vals), _ptr(x), _ptr(y)))
        return y

    def __del__(self):
        try:
            lib().madb_solver_destroy(self.h)
        except Exception:
            pass


def load_vector(ctx, space, f, quad_order=None):
    """Load vector b_i = (f, phi_i) of a scalar space on the device (MFEM: LinearForm + DomainLFIntegrator(f),
    ex4.cpp:145-148).  f: host callback xyz[npts, dim] -> [npts] sampled at the rule's points
    (Coefficient-type source), or an array [ne, nq] already sampled there.  quad_order None: MFEM's linear-form rule
    (order 2p).  Implemented as the residual of the energy f(x) u (functional "load"): same kernels as every other form."""
    order, vdim = int(space.desc["order"]), int(space.desc.get("vdim", 1))
    fn = Functional(ctx, "load") if vdim == 1 else Functional(ctx, "vload", iparams=[vdim])  # VectorDomainLFIntegrator (ex3.cpp:64-67)
    gi = Integrator(ctx, [(space, VALUE | (VECTOR if vdim > 1 else 0))], fn, quad_order=2 * order if quad_order is None else quad_order)
    if callable(f):
        gi.set_param_coefficient(f)
    else:
        gi.set_param_qf(np.ascontiguousarray(f, dtype=np.float64))
    return gi.mult(np.zeros(space.desc["ndofs"] * vdim))


def lvpp_update(ctx, alpha, psi, psik, lambda_prev, w=None):
    """Fused latent update; psik and lambda_prev are overwritten; returns the weighted l1 change of lambda."""
    d = C.c_double()
    n = psi.numel() if hasattr(psi, "numel") else psi.size
    _check(lib().madb_lvpp_update(ctx.h, n, alpha, _ptr(psi), _ptr(psik), _ptr(lambda_prev), _ptr(w), C.byref(d)))
    return d.value


class PGStepSizeRule:
    """src/pg.hpp:10-34, src/pg.cpp:4-54 (host scalar; stays on the host)."""
    CONSTANT, POLY, EXP, DOUBLE_EXP, INVALID = range(5)

    def __init__(self, rule_type, alpha0=1.0, max_alpha=1e6, ratio=-1.0, ratio2=-1.0):
        if not rule_type < self.INVALID:
            raise MadbError("PGStepSizeRule: Invalid rule type")
        if not alpha0 > 0:
            raise MadbError("PGStepSizeRule: alpha0 must be positive")
        if not max_alpha >= alpha0:
            raise MadbError("PGStepSizeRule: max_alpha must be greater than or equal to alpha0")
        if rule_type == self.POLY and not ratio > 0:
            raise MadbError("PGStepSizeRule: ratio must be positive for POLY rule")
        if rule_type == self.EXP and not ratio > 1:
            raise MadbError("PGStepSizeRule: ratio must be greater than 1 for EXP rule")
        if rule_type == self.DOUBLE_EXP and not (ratio > 1 and ratio2 > 1):
            raise MadbError("PGStepSizeRule: ratio and ratio2 must be greater than 1 for DOUBLE_EXP rule")
        self.rule_type, self.alpha0, self.max_alpha, self.ratio, self.ratio2 = rule_type, alpha0, max_alpha, ratio, ratio2

    def get(self, it):
        a = np.float64(self.alpha0)
        with np.errstate(over="ignore"):  # std::pow overflows to inf, then min() caps it
            if self.rule_type == self.POLY:
                a = a * np.power(np.float64(it + 1), self.ratio)
            elif self.rule_type == self.EXP:
                a = a * np.power(np.float64(self.ratio), np.float64(it))
            elif self.rule_type == self.DOUBLE_EXP:
                a = a * np.power(np.float64(self.ratio), np.power(np.float64(self.ratio2), np.float64(it)))
        return float(min(a, self.max_alpha))

"""Host-side outer loops that drive the batched assembly: Newton and the LVPP /
proximal-Galerkin iteration of the reference's drivers.

    newton()      <-> mfem::NewtonSolver::Mult [MFEM-upstream] as configured at ex4.cpp:166-176
    lvpp_solve()  <-> the outer loop of ex4.cpp:183-219 / ex5.cpp:174-212

Linear solve inside Newton: by default both the CUDA path and the CPU oracle are driven through
the same SuperLU factorisation (scipy) so that iteration counts are comparable; with `linear=` the CUDA path solves on the
device (madb_solver_*: Jacobi-PCG, or the statically condensed PCG for proximal-Galerkin block systems) and the CSR
values never leave the GPU.  `op` is any object with  mult(x) -> residual,  grad(x) -> CSR values,
pattern() -> (rowptr, colidx).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class DeviceLinear:
    """Device linear solve for newton(): kind "pcg" (SPD Jacobians), "condensed" or "minres" (PG block systems, primal dofs
    [0, nh), latent L2 dofs after them, nb per element; "minres" stays robust when the entropy Hessian degenerates).  The CSR values live in one device array that the assembly fills."""

    def __init__(self, integrator, kind="pcg", nh=None, nb=None, rtol=1e-12, maxit=20000):
        import torch
        import mfem_ad_b200 as M
        self.gi, self.kind, self.nh, self.nb, self.rtol, self.maxit = integrator, kind, nh, nb, rtol, maxit
        self.solver = M.Solver(integrator)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.vals = torch.empty(integrator.nnz, dtype=torch.float64, device=dev)
        self.linear_iterations = []

    def step(self, x, b):
        """Newton correction c with J(x) c = F(x) - b; returns (c, residual used)."""
        r = np.empty_like(x)
        self.gi.assemble(x, r, self.vals)  # residual to the host, Jacobian values stay on the device
        r -= b
        c = np.zeros_like(x)
        if self.kind == "pcg":
            c, it, rr = self.solver.pcg(self.vals, r, c, rtol=self.rtol, maxit=self.maxit)
        elif self.kind == "condensed":
            c, it, rr = self.solver.condensed_pcg(self.nh, self.nb, self.vals, r, c, rtol=self.rtol, maxit=self.maxit)
        else:  # "minres": the block system as it is, block-diagonal preconditioner (nb = 0: plain Jacobi, any symmetric system)
            c, it, rr = self.solver.pg_minres(self.nh or 0, self.nb or 0, self.vals, r, c, rtol=self.rtol, maxit=self.maxit)
        self.linear_iterations.append(it)
        self.relres = getattr(self, "relres", []) + [rr]
        if not rr <= max(1e3 * self.rtol, 1e-8):
            raise RuntimeError("device linear solve stalled: relative residual %.2e after %d iterations" % (rr, it))
        return c


def newton(op, b, x, abs_tol=1e-9, rel_tol=0.0, max_iter=20, linear=None):
    """x is updated in place (iterative_mode = true).  Returns (converged, iterations, final_norm).
    linear: a DeviceLinear (solve on the GPU) or None (SuperLU on the host)."""
    rowptr, colidx = op.pattern()
    n = x.size
    r = op.mult(x) - b
    norm = np.linalg.norm(r)
    norm_goal = max(rel_tol * norm, abs_tol)
    it = 0
    while True:
        if norm <= norm_goal:
            return True, it, norm
        if it >= max_iter:
            return False, it, norm
        if linear is not None:
            c = linear.step(x, b)
        else:
            vals = op.grad(x)
            J = sp.csr_matrix((vals, colidx, rowptr), shape=(n, n)).tocsc()
            c = spla.splu(J).solve(r)
        x -= c
        r = op.mult(x) - b
        norm = np.linalg.norm(r)
        it += 1


def lvpp_solve(op, set_alpha, set_latent_k, alpha_rule, b, x, latent_slice, l1_norm, max_pg=100, tol=1e-10,
               newton_kw=None, log=None):
    """Outer proximal-Galerkin loop (ex4.cpp:183-219).

    set_alpha(alpha), set_latent_k(psi_k) update the functional between steps;
    latent_slice selects psi inside x; l1_norm(v) = || v ||_{L1(Omega)} of a latent-space function.
    Returns dict(pg_iterations, newton_iterations[list], converged, lambda_diff[list])."""
    newton_kw = newton_kw or {}
    psi = x[latent_slice]
    lam_prev = np.zeros_like(psi)
    hist = dict(pg_iterations=0, newton_iterations=[], lambda_diff=[], converged=False, newton_failed=False)
    for i in range(max_pg):
        alpha = alpha_rule.get(i)
        set_alpha(alpha)
        psik = x[latent_slice].copy()
        set_latent_k(psik)
        ok, its, nrm = newton(op, b, x, **newton_kw)
        hist["newton_iterations"].append(its)
        hist["pg_iterations"] = i + 1
        if not ok:
            hist["newton_failed"] = True
            break
        lam = (x[latent_slice] - psik) / alpha
        diff = l1_norm(lam - lam_prev)
        hist["lambda_diff"].append(diff)
        if log:
            log("PG %d alpha=%g newton=%d res=%.3e lambda_diff=%.3e" % (i + 1, alpha, its, nrm, diff))
        if diff < tol:
            hist["converged"] = True
            break
        lam_prev = lam
    return hist

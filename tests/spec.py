"""One description of a functional / form, convertible to the CPU oracle and to the CUDA path."""
import numpy as np

from oracle import oracle as O


class FSpec:
    def __init__(self, kind, n_input, params=(), iparams=(), children=(), qoff=-1):
        self.kind, self.n_input = kind, n_input
        self.params, self.iparams, self.children, self.qoff = list(params), list(iparams), list(children), qoff

    # ---- CPU oracle ----
    _OK = dict(ex0=O.K_EX0, mass=O.K_MASS, diffusion=O.K_DIFFUSION, diff=O.K_DIFF, elasticity=O.K_ELASTICITY,
               minsurf=O.K_MINSURF, obstacle=O.K_OBSTACLE, gradobstacle=O.K_GRADOBSTACLE, pg=O.K_PG,
               lambdapg=O.K_LAMBDAPG, lagrangian=O.K_LAGRANGIAN, al=O.K_AL, shannon=O.K_SHANNON, fermidirac=O.K_FERMIDIRAC, hellinger=O.K_HELLINGER,
               simplex=O.K_SIMPLEX, simp=O.K_SIMP, paramcompliance=O.K_PARAMCOMPLIANCE, empty=O.K_EMPTY, load=O.K_LOAD)

    def _add(self, F):
        ch = [c._add(F) for c in self.children]
        p, ip = list(self.params), list(self.iparams)
        k = self.kind
        if k == "shannon":  # oracle: param bound, iparam sign
            p, ip = [self.params[0]], [int(self.params[1])]
        elif k == "elasticity":
            ip = [int(round(np.sqrt(self.n_input)))]
        elif k == "paramcompliance":
            ip = [int(round(np.sqrt(self.n_input)))]
        elif k == "hellinger":
            ip = [self.n_input]
        return F.add(self._OK[k], self.n_input, params=p, iparams=ip, children=ch, qoff=self.qoff)

    def oracle(self):
        F = O.Functional()
        self._add(F)
        return F

    # ---- CUDA path ----
    def madb(self, ctx):
        import mfem_ad_b200 as M
        ch = [c.madb(ctx) for c in self.children]
        ip = list(self.iparams)
        kind = self.kind
        if kind == "hellinger" and self.qoff >= 0:
            kind = "hellingerq"  # the bound is a per-point parameter (ex5.cpp:114-117)
        if kind == "diffusion" and self.qoff >= 0:
            kind = "diffusionq"
        if kind == "load" and self.n_input > 1:
            kind = "vload"
        return M.Functional(ctx, kind, params=self.params, iparams=ip, children=ch)


def minsurf(dim, eps=0.5):
    return FSpec("minsurf", dim, [eps])


def diffusion(dim, K=(), qoff=-1, kdim=None):
    """DiffusionEnergy (src/ad_native.hpp:421-481): K constant (none / scalar / diagonal / full, column-major), or read
    from the per-point parameters at qoff (kdim entries)."""
    if qoff >= 0:
        return FSpec("diffusion", dim, [], [kdim], qoff=qoff)
    return FSpec("diffusion", dim, list(K), [len(K)])


def diff(energy, qoff=0):
    """DiffEnergy (src/ad_native.hpp:483-525): energy(x - target), target = per-point parameters at qoff."""
    return FSpec("diff", energy.n_input, [], [], [energy], qoff=qoff)


def mass(n):
    return FSpec("mass", n)


def load(n=1):
    """sum_c f_c(x) u_c with f the first per-point parameters: its gradient is the load vector ((Vector)DomainLFIntegrator,
    ex4.cpp:145-148, ex3.cpp:64-67)."""
    return FSpec("load", n, [], [n] if n > 1 else [], qoff=0)


def elasticity(dim, lam, mu):
    return FSpec("elasticity", dim * dim, [lam, mu])


def obstacle(dim):
    return FSpec("obstacle", dim + 1)


def gradobstacle(dim):
    return FSpec("gradobstacle", dim)


def fermidirac(lower, upper):
    return FSpec("fermidirac", 1, [lower, upper])


def shannon(bound, sign=1):
    return FSpec("shannon", 1, [bound, sign])


def hellinger(dim, scale, qoff=-1):
    return FSpec("hellinger", dim, [scale] if qoff < 0 else [], qoff=qoff)


def simplex(n, scale=1.0):
    return FSpec("simplex", n, [scale])


def simp(E, p):
    return FSpec("simp", len(E), list(E) + [p])


def pg(f, entropy, alpha, primal_idx=0):
    """ADPGFunctional(f, entropy, psi_k, idx): psi_k is the first per-point parameter."""
    return FSpec("pg", f.n_input + entropy.n_input, [alpha], [primal_idx], [f, entropy], qoff=0)


def lagrangian(f, c, mode=-1):
    """Lagrangian f(x) + sum_i lambda_i c_i(x) (src/ad_native.hpp:570-621); c: one constraint or a list; inputs [x, lambda]."""
    cs = list(c) if isinstance(c, (list, tuple)) else [c]
    return FSpec("lagrangian", f.n_input + len(cs), [], [mode], [f] + cs)


def al(f, c, mu, rhs, lam, mode=-1):
    """ALFunctional f + sum_i c~_i (lambda_i + mu/2 c~_i), c~_i = c_i - rhs_i (src/ad_native.hpp:624-691); c, rhs, lam: one
    constraint with scalars, or lists of equal length."""
    cs = list(c) if isinstance(c, (list, tuple)) else [c]
    rhs = list(rhs) if isinstance(rhs, (list, tuple)) else [rhs]
    lam = list(lam) if isinstance(lam, (list, tuple)) else [lam]
    return FSpec("al", f.n_input, [mu] + rhs + lam, [mode], [f] + cs)


def pg2(f, e1, idx1, e2, idx2, alpha):
    """ADPGFunctional with two entropies (src/pg.hpp:105-127, :193-213): inputs [x, psi_1, psi_2], psi_k blocks first in the
    per-point parameters."""
    return FSpec("pg", f.n_input + e1.n_input + e2.n_input, [alpha], [idx1, idx2], [f, e1, e2], qoff=0)


def lambdapg(f, entropy, alpha, primal_idx=0):
    """ADLambdaPGFunctional(f, entropy, psi_k, idx) (src/pg.hpp:216-243): inputs [x, lambda]."""
    return FSpec("lambdapg", f.n_input + entropy.n_input, [alpha], [primal_idx], [f, entropy], qoff=0)


def make_pair(ctx, mesh, spaces, fspec, quad_order=-1, params=(), ess=(), block=None):
    """Builds (oracle form, CUDA integrator) from the same arrays.

    spaces: list of dicts from meshgen with 'mode' and optional 'role' (0 input, 1 param).
    params: list of oracle parameter dicts for the PARAM fields / quadrature functions."""
    import mfem_ad_b200 as M
    inputs = [s for s in spaces if s.get("role", 0) == 0]
    of = O.OracleForm(mesh, inputs, fspec.oracle(), quad_order=quad_order, params=params, block=block, ess=ess)
    gm = M.Mesh(ctx, mesh)
    gs = [M.Space(ctx, gm, s) for s in spaces]
    gf = fspec.madb(ctx)
    # block=None: the reference's choice (one space -> ADNonlinearFormIntegrator, several -> the block integrator)
    gi = M.Integrator(ctx, [(gs[i], spaces[i]["mode"], spaces[i].get("role", 0)) for i in range(len(spaces))], gf,
                      quad_order=quad_order, block=bool(block))
    if len(ess):
        gi.set_essential(ess)
    gi._keepalive = (gm, gs, gf)
    return of, gi


def csr_rel_err(vals, ref):
    """max |vals-ref| relative to the largest reference entry (norm-relative, SURVEY 7 'Determinism + 1e-12')."""
    return np.max(np.abs(vals - ref)) / max(np.max(np.abs(ref)), 1e-300)

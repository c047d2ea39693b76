/* mfemad_b200.h -- C ABI of the B200-native AD assembly hot path.
 *
 * Drop-in boundary for the element-level AD assembly path of dohyun-cse/mfem-ad
 * (reference file:line cited per entry point).  The reference sits behind MFEM's
 * per-element virtuals
 *     GetElementEnergy / AssembleElementVector / AssembleElementGrad
 * (src/_ad_intg.hpp:108-135 single space, :245-276 block), called once per
 * element by NonlinearForm / BlockNonlinearForm.  Per-element virtual calls
 * cannot feed a GPU, so this ABI is the whole-mesh batched equivalent: the
 * element loop, gather (GetElementVDofs/GetSubVector), the integrator body and
 * the scatter (AddElementVector / AddSubMatrix) are one call.
 *
 * Conventions
 *  - plain pointers and sizes only; every vector/array argument may be a HOST or
 *    a DEVICE pointer (detected with cudaPointerGetAttributes).  Host pointers
 *    are staged through device buffers owned by the integrator (H2D/D2H copies
 *    inside the call); device pointers are used in place on the context stream.
 *  - the caller owns every array it passes; handles own their workspaces and
 *    are destroyed explicitly.
 *  - error convention: every function returns 0 on success, non-zero on error;
 *    madb_last_error() gives the message.  (The reference aborts through
 *    MFEM_VERIFY / MFEM_ABORT, src/ad_native.hpp:167, src/ad_native.cpp:111-117;
 *    an MFEM adapter turns non-zero into MFEM_ABORT -- see INTEGRATION.md.)
 *  - not re-entrant per context, like the reference's integrators (mutable
 *    scratch, src/_ad_intg.hpp:80-93): one host thread + one CUDA stream per
 *    context / GPU.  Functional parameters are re-read at every call
 *    (ex2.cpp:98 mutates eps, ex4.cpp:187 alpha between solves).
 *  - all element-local orderings are LEXICOGRAPHIC (x fastest); an MFEM adapter
 *    applies TensorBasisElement::GetDofMap() as MFEM's ElementRestriction does.
 */
#ifndef MFEMAD_B200_H
#define MFEMAD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct madb_ctx madb_ctx;
typedef struct madb_mesh madb_mesh;
typedef struct madb_space madb_space;
typedef struct madb_functional madb_functional;
typedef struct madb_integrator madb_integrator;
typedef struct madb_solver madb_solver;
typedef struct madb_comm madb_comm;
typedef struct madb_exchange madb_exchange;

/* ADEval flags: src/_ad_intg.hpp:24-36 (same bit values) */
enum
{
   MADB_QVALUE = 1 << 0,
   MADB_VALUE = 1 << 1,
   MADB_GRAD = 1 << 2,
   MADB_DIV = 1 << 3,
   MADB_CURL = 1 << 4,
   MADB_HESSIAN = 1 << 5,
   MADB_VECTOR = 1 << 6,
   MADB_VECFE = 1 << 7
};
enum { MADB_BASIS_H1 = 0 /* H1_FECollection: closed Gauss-Lobatto */, MADB_BASIS_L2 = 1 /* L2_FECollection: open Gauss-Legendre */ };
enum { MADB_BYNODES = 0, MADB_BYVDIM = 1 }; /* mfem::Ordering */
enum { MADB_ROLE_INPUT = 0, MADB_ROLE_PARAM = 1 };

int madb_version(void);
const char *madb_last_error(void);
/* 1 if fused kernels are registered under `key` = "<functional key>|<element configuration key>" (the key an integrator
 * looks up; madb_integrator_create names it when it is missing).  Plugins (user functionals compiled out of tree,
 * the AD_IMPL bodies of src/ad_native.hpp:332-365) add keys when their library is loaded. */
int madb_registry_has(const char *key);

/* One context per GPU / rank (replaces nothing in the reference: it has no device). */
int madb_ctx_create(int device, madb_ctx **out);
int madb_ctx_destroy(madb_ctx *ctx);
int madb_ctx_sync(madb_ctx *ctx);
/* Device memory for hosts that do not link the CUDA runtime themselves (MFEM's own device memory works as well: every
 * array argument of this API may be a device pointer). */
int madb_device_alloc(madb_ctx *ctx, size_t bytes, void **ptr);
int madb_device_free(madb_ctx *ctx, void *ptr);
/* cudaStream_t of the context, so callers can bracket calls with CUDA events. */
void *madb_ctx_stream(madb_ctx *ctx);

/* Geometry: replaces ElementTransformation::SetIntPoint / Weight / InverseJacobian
 * as used at src/ad_intg.hpp:236-237 (Tr.Weight()) and inside CalcPhysDShape
 * (:137).  Tensor-product elements (quads dim=2, hexes dim=3) with the order-1
 * isoparametric map through the element vertices.
 *   e2n    [ne * 2^dim]  vertex ids, lexicographic per element
 *   coords [nnodes*dim]  xyzxyz... */
int madb_mesh_create(madb_ctx *ctx, int dim, int ne, const int32_t *e2n, int nnodes, const double *coords,
                     madb_mesh **out);
/* Triangles (Mesh::MakeCartesian2D(..., Element::TRIANGLE), ex5.cpp:72-73): e2n [ne*3], affine map of the 3 vertices.
 * Spaces on such a mesh: H1 orders 1 and 2 (element dofs: the 3 vertices, then the edge midpoints in edge order (0,1),
 * (1,2), (2,0)), L2 order 0; rules: MFEM's triangle rules of order 0 - 6 (default 2p+2: 6 points for P1, 12 for P2). */
int madb_mesh_create_simplex(madb_ctx *ctx, int dim, int ne, const int32_t *e2n, int nnodes, const double *coords,
                             madb_mesh **out);
int madb_mesh_destroy(madb_mesh *m);

/* FiniteElementSpace view: element -> scalar dof map, basis and vdim.
 * Replaces fes->GetElementVDofs + FiniteElement::CalcShape/CalcDShape tables
 * (src/ad_intg.hpp:132-137).  e2l [ne * (order+1)^dim], lexicographic. */
int madb_space_create(madb_ctx *ctx, madb_mesh *mesh, int basis, int order, int vdim, int ordering, int ndofs,
                      const int32_t *e2l, madb_space **out);
int madb_space_destroy(madb_space *s);

/* ADFunction objects (src/ad_native.hpp:137-190, :413-691; src/pg.hpp; src/mmto.hpp).
 * kind: "mass" "diffusion" "elasticity" "minsurf" "obstacle" "gradobstacle"
 *       "diff" "pg" "lambdapg" "shannon" "fermidirac" "hellinger" "simplex"
 *       "simp" "paramcompliance" "empty" "ex0", or any kind registered by a plugin.
 * params: the functional's own run-time constants (e.g. minsurf: eps; elasticity:
 *         lambda, mu; fermidirac: lower, upper; pg: alpha; simp: E[0..n-1], p).
 * iparams: structural integers that select the compiled variant (pg: primal_idx). */
int madb_functional_create(madb_ctx *ctx, const char *kind, int nparams, const double *params, int niparams,
                           const int *iparams, int nchildren, madb_functional *const *children,
                           madb_functional **out);
int madb_functional_set_params(madb_functional *f, int nparams, const double *params);
int madb_functional_destroy(madb_functional *f);

/* AD layer on the device at n points (ADFunction::operator(), Gradient, Hessian:
 * src/ad_native.cpp:181-230).  x [npts*n_input]; qprm [npts*n_qprm] or NULL;
 * outputs may be NULL: value [npts], grad [npts*n], hess [npts*n*n] (symmetric). */
int madb_functional_eval(madb_ctx *ctx, madb_functional *f, int n_input, int npts, const double *x,
                         const double *qprm, double *value, double *grad, double *hess);

/* ADVectorFunction (src/ad_native.hpp:198-265; Gradient / Hessian src/ad_native.cpp:232-276) at npts points:
 * value [npts][n_output], Jacobian jac [npts][n_output][n_input], Hessians hess [npts][n_output][n_input][n_input]
 * (any output may be NULL).  Only ex0 uses a vector function in the reference (ex0.cpp:23-35, kind "ex0vec"). */
int madb_vecfunction_eval(madb_ctx *ctx, madb_functional *f, int n_input, int n_output, int npts, const double *x,
                          double *value, double *jac, double *hess);

/* DOF-collocated proximal-Galerkin terms of ADDofPGNonlinearFormIntegrator
 * (src/_dof_pg.hpp:17-63, src/dof_pg.hpp:66-128 residual, :131-231 Jacobian diagonals) for one
 * primal/latent pair with a scalar entropy, one thread per dof:
 *   r_u[j] += (psi_j-psi_k,j) w_j/alpha ; r_psi[j] = (u_j - E*'(psi_j)) w_j/alpha ;
 *   d_pp[j] = -E*''(psi_j) w_j/alpha ; d_up[j] = w_j/alpha.
 * w = nodal (lumped) weights: the reference takes them from FiniteElement::GetNodes(), whose
 * weights are zero in MFEM; zeros reproduce the reference, lumped-mass weights make
 * the variant useful.  Outputs may be NULL. */
int madb_dofpg_nodal(madb_ctx *ctx, madb_functional *entropy, int n, double alpha, const double *u,
                     const double *psi, const double *psik, const double *w, double *r_u, double *r_psi,
                     double *d_pp, double *d_up);

/* Fused LVPP latent-variable update between proximal steps (ex4.cpp:188-189, :203-218):
 *   lambda = (psi-psi_k)/alpha ; *lambda_diff = sum_j w_j |lambda_j - lambda_prev_j| ;
 *   lambda_prev <- lambda ; psi_k <- psi.   w may be NULL (unit weights). */
int madb_lvpp_update(madb_ctx *ctx, int n, double alpha, const double *psi, double *psik, double *lambda_prev,
                     const double *w, double *lambda_diff);

/* Shared-dof exchange between element partitions (ParMesh + conforming prolongation P,
 * ex4.cpp:85,136 [MFEM-upstream]): pack dst[i] = src[idx[i]] into a contiguous send buffer and
 * unpack dst[idx[i]] (+)= src[i] from a receive buffer, on the context stream.  DEVICE pointers;
 * the transfer itself is NCCL send/recv between neighbours (mfem-ad_b200/parallel.py). */
int madb_pack(madb_ctx *ctx, int n, const int32_t *idx, const double *src, double *dst);
int madb_unpack(madb_ctx *ctx, int n, const int32_t *idx, const double *src, double *dst, int add);
/* Unpack of a whole receive buffer in one launch: destination i = dst_idx[i] gets the values src[src4[4i..4i+3]]
 * (-1: none) added in that order (ascending peer rank: deterministic), on top of the old value if add != 0.
 * A dof shared by several ranks (a corner) has one destination entry with several sources. */
int madb_unpack_multi(madb_ctx *ctx, int n, const int32_t *src4, const int32_t *dst_idx, const double *src, double *dst, int add);

/* The same exchange behind the ABI for C++ / MFEM hosts: one process per GPU, NCCL send/recv between neighbouring
 * ranks (ParMesh / ParFiniteElementSpace group communication, ex4.cpp:85,136 [MFEM-upstream]).
 *   madb_comm_unique_id   rank 0 creates the 128-byte id and hands it to the other ranks (MPI_Bcast, a file, ...)
 *   madb_comm_create      ncclCommInitRank on the context's device + a communication stream
 *   madb_comm_allreduce_sum  global sums of up to 64 doubles (host or device pointer): Newton norms, the L1 change
 *                         of lambda at ex4.cpp:205; same bits on every rank
 *   madb_exchange_create  neighbour lists, peers in ascending rank order:
 *        own_*   : per peer the LOCAL indices of dofs this rank owns and the peer holds a copy of
 *        ghost_* : per peer the LOCAL indices of copies this rank holds of dofs the peer owns
 *        (both sides list the dofs of a pair in the same order, e.g. ascending global id)
 *   madb_exchange_begin(x, vec, reverse) / madb_exchange_end(x, vec, add): device vectors, asynchronous on the context
 *        stream (pack kernel -> event -> ncclGroup{Send,Recv} on the communication stream -> event -> unpack kernel).
 *        reverse = 0: P, owner -> copies (end with add = 0).  reverse = 1: P^T, copies -> owner (end with add = 1): the
 *        values of one dof are added on the owner in ascending peer rank, independent of arrival order.
 *        Work queued on the context stream between begin and end overlaps the transfer. */
int madb_comm_unique_id(unsigned char *id128);
int madb_comm_create(madb_ctx *ctx, const unsigned char *id128, int rank, int world, madb_comm **out);
int madb_comm_destroy(madb_comm *c);
int madb_comm_allreduce_sum(madb_comm *c, int n, double *values);
int madb_exchange_create(madb_comm *c, int nown_peers, const int *own_peer, const int *own_count, const int32_t *own_idx,
                         int nghost_peers, const int *ghost_peer, const int *ghost_count, const int32_t *ghost_idx,
                         madb_exchange **out);
int madb_exchange_destroy(madb_exchange *x);
int madb_exchange_begin(madb_exchange *x, const double *vec, int reverse);
int madb_exchange_end(madb_exchange *x, double *vec, int add);

/* AD(Block)NonlinearFormIntegrator<modes...>(f, ir) attached to its form
 * (src/_ad_intg.hpp:71-155, :157-327).  fields: spaces[i] with ADEval modes[i];
 * roles[i] = MADB_ROLE_INPUT for a differentiated unknown (a block of x), or
 * MADB_ROLE_PARAM for a GridFunction parameter of the functional's Evaluator
 * (src/ad_native.cpp:166-171), e.g. psi_k of ADPGFunctional (src/pg.hpp:110).
 * quad_order < 0 selects the default 2*max_order+2 (src/_ad_intg.hpp:99-105, :298-313). */
int madb_integrator_create(madb_ctx *ctx, int nfields, madb_space *const *spaces, const int *modes,
                           const int *roles, madb_functional *f, int quad_order, madb_integrator **out);
/* Same with flags.  MADB_INTEG_BLOCK: treat the integrator as ADBlockNonlinearFormIntegrator even with one input
 * space.  It matters for ONE vector space with MADB_VECTOR only: without the flag the Jacobian reproduces the
 * arithmetic of ADNonlinearFormIntegrator<...|VECTOR>::AssembleElementGrad as written (src/ad_intg.hpp:283-326:
 * contiguous windows of Hx, untransposed mirror blocks; it equals the intended B H B^T only for special Hessians,
 * e.g. LinearElasticityEnergy with lambda == mu as in ex3.cpp:58 -- SURVEY H1); with the flag the index-consistent
 * contraction of the block integrator (src/ad_intg.hpp:700-727) is used.  Residual and energy are the same in both. */
enum { MADB_INTEG_BLOCK = 1 };
int madb_integrator_create_ex(madb_ctx *ctx, int nfields, madb_space *const *spaces, const int *modes,
                              const int *roles, madb_functional *f, int quad_order, int flags, madb_integrator **out);
int madb_integrator_destroy(madb_integrator *I);

/* sizes: total dofs of the concatenated input blocks, quadrature points per element */
int madb_integrator_sizes(madb_integrator *I, int64_t *ntotal, int *nq_el, int *ncolors);
/* Patch-assembly statistics (diagnostics): out[0..7] = patches (0: colour-scatter path in use), max rows
 * and max matrix slots per patch, interface dofs, interface matrix entries, staged residual and matrix
 * partials, CSR runs.  Matrix-side figures are 0 until the sparsity pattern has been built. */
int madb_integrator_patch_stats(madb_integrator *I, int64_t *out);
/* Host-only self test of the patch-assembly maps (no CUDA device needed): builds the maps for one H1 space of the
 * given order / vdim on the given mesh, assembles integer-valued element vectors and matrices directly and through
 * an emulation of the kernels' use of the maps, and returns the largest difference (must be 0).
 * stats[0..7]: patches, interface dofs, interface matrix entries, staged matrix partials, largest map blob (bytes), nnz,
 * CSR entries written as aligned pairs of chunks (16-byte stores), reserved (0). */
int madb_patch_selftest(int dim, int ne, const int32_t *e2n, int nnodes, const double *coords, int order, int vdim,
                        int ordering, int ndofs, const int32_t *e2l, double *max_err, int64_t *stats);

/* Same for the maps of the CSR-image kernel (shared-memory image of the patch's CSR rows written with bulk copies):
 * tpe = threads per element (1, or 2 = the thread pair with the mirrored half-element, scalar 2-D spaces).
 * stats[0..7]: patches, interface dofs, interface matrix entries, staged matrix partials, shared memory per work group
 * (bytes), nnz, CSR entries written by bulk copies, bytes of maps and lists read per assembly. */
int madb_patch_selftest_img(int dim, int ne, const int32_t *e2n, int nnodes, const double *coords, int order, int vdim,
                            int ordering, int ndofs, const int32_t *e2l, int tpe, double *max_err, int64_t *stats);

/* Device timing of the element kernel(s) of the last mult / assemble / grad_mult call: CUDA events on the
 * context stream around the dominant kernel (the interface reduction and essential-dof kernels excluded). */
int madb_integrator_set_timing(madb_integrator *I, int on);
int madb_integrator_last_kernel_ms(madb_integrator *I, double *ms);

/* Evaluator sources that vary in space (src/ad_native.hpp:56-61):
 * GridFunction parameter of field `field` (dof vector of that space), and
 * QuadratureFunction parameters  qf[(e*nq + q)*nqf + k]  (GetValues(ElementNo, ip.index)). */
int madb_integrator_set_param_field(madb_integrator *I, int field, const double *dofs);
int madb_integrator_set_param_qf(madb_integrator *I, int nqf, const double *qf);

/* NonlinearForm::SetEssentialTrueDofs: Mult zeroes y there; GetGradient
 * eliminates rows+columns with unit diagonal [MFEM-upstream, SURVEY a32]. */
int madb_integrator_set_essential(madb_integrator *I, int n, const int32_t *dofs);

/* NonlinearForm::GetEnergy -> sum_e GetElementEnergy (src/ad_intg.hpp:157-199, :469-530) */
int madb_integrator_energy(madb_integrator *I, const double *x, double *energy);
/* NonlinearForm::Mult -> AssembleElementVector (src/ad_intg.hpp:202-257, :533-619) */
int madb_integrator_mult(madb_integrator *I, const double *x, double *y);
/* Sparsity of GetGradient: full element connectivity, sorted columns.
 * Call with rowptr = colidx = NULL for the sizes. */
int madb_integrator_pattern(madb_integrator *I, int64_t *nrows, int64_t *nnz, int32_t *rowptr, int32_t *colidx);
/* NonlinearForm::GetGradient -> AssembleElementGrad (src/ad_intg.hpp:260-334, :622-729) */
int madb_integrator_grad_assemble(madb_integrator *I, const double *x, double *vals);
/* residual + Jacobian at the same state in one pass (one Newton iteration's assembly) */
int madb_integrator_assemble(madb_integrator *I, const double *x, double *y, double *vals);
/* The same in two steps for parallel runs (device pointers only): _begin launches the element kernel and completes the
 * residual; the caller then starts the shared-dof exchange of the residual (madb_exchange_begin, P^T of ParNonlinearForm::
 * Mult, ex4.cpp:136); _end launches the interface reduction (and essential-dof elimination) of the CSR values, which runs
 * while NCCL moves the residual; madb_exchange_end afterwards.  _end without a pending _begin is a no-op. */
int madb_integrator_assemble_begin(madb_integrator *I, const double *x, double *y, double *vals);
int madb_integrator_assemble_end(madb_integrator *I);
/* DifferentiableCoefficient::Eval and ::Gradient().Eval projected to the rule's points
 * (src/ad_native.hpp:267-323; ex4.cpp:124-128,200: the latent->primal map grad E*(psi) as a
 * QuadratureFunction).  value [ne*nq] and/or grad [ne*nq*n_input] (may be NULL), layout [e][q][.] */
int madb_integrator_coefficient(madb_integrator *I, const double *x, double *value, double *grad);
/* DifferentiableCoefficient::Hessian().Eval (HessianCoefficient, src/ad_native.hpp:300-323) at the rule's points:
 * value [ne*nq], grad [ne*nq*n], hess [ne*nq*n*n] (symmetric; any output may be NULL). */
int madb_integrator_coefficient_hessian(madb_integrator *I, const double *x, double *value, double *grad, double *hess);
/* ParametrizedFunctional::ParamGradient::Eval (src/mmto.hpp:54-70, src/mmto.cpp:4-38) at the rule's points for an
 * integrator whose INPUT field is the design and whose PARAM field is the state ("designcompliance" kinds):
 * J [ne*nq*param_dim], value [ne*nq] = F (may be NULL).
 *   MADB_PARAMGRAD_AS_WRITTEN  the reference's result: slot i of the evaluator is overwritten with df_i/drho_j while the
 *                              other parameter functions keep their values (:25-37), i.e. dF/drho_j + (m-1) F
 *   MADB_PARAMGRAD_DERIVATIVE  the derivative dF/drho_j itself (corrected variant) */
enum { MADB_PARAMGRAD_AS_WRITTEN = 0, MADB_PARAMGRAD_DERIVATIVE = 1 };
int madb_integrator_param_gradient(madb_integrator *I, const double *design, double *value, double *J, int variant);
/* Physical coordinates of the rule's points, xyz [ne*nq*dim] (host or device pointer).  Evaluator sources of Coefficient /
 * VectorCoefficient / MatrixCoefficient type (src/ad_native.hpp:56-61; Eval at (Tr, ip), src/ad_native.cpp:132-165) are
 * host callbacks: sample them at these points and pass the result with madb_integrator_set_param_qf. */
int madb_integrator_qpoint_coords(madb_integrator *I, double *xyz);
/* matrix-free Jacobian action y = J(x) v (no reference equivalent; config 3) */
int madb_integrator_grad_mult(madb_integrator *I, const double *x, const double *v, double *y);

/* ---- linear solve of a Newton step on the device -----------------------------------------------
 * The reference solves J dx = -r on the host: UMFPackSolver (ex2.cpp:80), MUMPSMonoSolver (src/tools.hpp:128-154), or a
 * Krylov method with PGPreconditioner (src/pg.hpp:378-504).  A solver object works on the CSR pattern of an integrator
 * (madb_integrator_pattern); vals / b / x are host or device pointers (detected), vals normally the device array
 * madb_integrator_assemble has just filled, so that the Jacobian never crosses PCIe.  x holds the initial guess on entry.
 * Stopping: ||r|| <= max(rtol ||b||, atol), at most maxit iterations; *iters / *relres report what was reached. */
int madb_solver_create(madb_integrator *I, madb_solver **out);
int madb_solver_destroy(madb_solver *s);
/* Jacobi-preconditioned conjugate gradients (SPD Jacobians: ex1 - ex3 with essential dofs eliminated DIAG_ONE) */
int madb_solver_pcg(madb_solver *s, const double *vals, const double *b, double *x, double rtol, double atol, int maxit,
                    int *iters, double *relres);
/* Proximal-Galerkin block systems [[A, C], [C^T, -D]] (ex4.cpp:136-142, ex5, par_template): unknowns [0, nh) primal,
 * [nh, n) latent on an L2 space with nb dofs per element (D block diagonal, definite: the entropy-Hessian-weighted mass
 * matrix of PGPreconditioner).  The latent block is eliminated exactly, S = A + C D^-1 C^T is solved by Jacobi-PCG, the
 * latent part follows by back-substitution. */
int madb_solver_condensed_pcg(madb_solver *s, int nh, int nb, const double *vals, const double *b, double *x, double rtol,
                              double atol, int maxit, int *iters, double *relres);
/* The same block systems without condensation: preconditioned MINRES on the symmetric indefinite matrix with the block
 * diagonal preconditioner diag(|diag A|, S_e), S_e = D_e + C_e^T diag(A)^-1 C_e per element (the PGPreconditioner idea,
 * src/pg.hpp:378-504, with Jacobi instead of AMG on the primal block).  Robust when the entropy Hessian degenerates
 * (active sets: D -> 0), where the condensed operator becomes a penalty matrix.  Stopping on the preconditioned residual.
 * nb = 0 (nh ignored): no block structure, Jacobi preconditioner |diag J|^-1 on all unknowns: any symmetric indefinite
 * Jacobian, e.g. ex5's system with its H1 latent space. */
int madb_solver_pg_minres(madb_solver *s, int nh, int nb, const double *vals, const double *b, double *x, double rtol,
                          double atol, int maxit, int *iters, double *relres);
/* y = A x with the solver's pattern (hand-written CSR kernel, deterministic) */
int madb_csr_spmv(madb_solver *s, const double *vals, const double *x, double *y);

#ifdef __cplusplus
}
#endif
#endif

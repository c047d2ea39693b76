// madb_solve.cu -- the linear solve of a Newton step on the device.
//
// The reference hands the assembled Jacobian to a direct solver on the host (UMFPackSolver, ex2.cpp:80;
// MUMPSMonoSolver, src/tools.hpp:128-154) or preconditions a Krylov method with PGPreconditioner (src/pg.hpp:378-504:
// the latent block is an entropy-Hessian-weighted mass matrix, block diagonal for the L2 latent spaces of the
// drivers).  Here the CSR values never leave the GPU:
//   madb_solver_pcg            Jacobi-preconditioned conjugate gradients on the assembled matrix (SPD systems:
//                              ex1 / ex2 / ex3 Newton steps with essential dofs eliminated DIAG_ONE)
//   madb_solver_condensed_pcg  the proximal-Galerkin saddle point systems [[A, C], [C^T, -D]] of ex4 / ex5 /
//                              par_template: the latent block D (L2 space: one dense block per element) is eliminated
//                              exactly, S = A + C D^-1 C^T is SPD and solved by the same conjugate gradients; the
//                              latent increment follows by back-substitution
// Everything is deterministic: dot products are fixed-grid block partials + a fixed-tree final sum, the scalars
// alpha / beta stay on the device, the host reads the residual norm every few iterations only.
// No algebraic multigrid: iteration counts grow with the mesh (documented in DESIGN.md); the Newton / LVPP iteration
// counts are those of a direct solve once the linear tolerance is tight (tests/test_gpu_solve.py).
#include "../../include/mfemad_b200.h"
#include "madb_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace madb
{
void set_error(const std::string &s);

namespace
{
constexpr int RED_BLOCKS = 592, RED_THREADS = 256; // fixed reduction grid: results do not depend on the launch

__device__ __forceinline__ double block_sum(double v, double *sm)
{
   // fixed tree over the block
   const int t = threadIdx.x;
   sm[t] = v;
   __syncthreads();
   for (int k = RED_THREADS / 2; k > 0; k >>= 1)
   {
      if (t < k) { sm[t] += sm[t + k]; }
      __syncthreads();
   }
   const double r = sm[0];
   __syncthreads();
   return r;
}

/// y[r0 + i] = sum_j A[r0 + i, j] w[j] for rows [r0, r1); TPR threads per row (power of two <= 32), fixed order within a row.
/// If dotw != null the block partial of sum_i y_i dotw[r0 + i] is written to partial[blockIdx] (rows of the block, fixed tree).
template <int TPR>
__global__ void __launch_bounds__(RED_THREADS) k_spmv(const int r0, const int r1, const int *__restrict__ rowptr,
                                                      const int *__restrict__ colidx, const double *__restrict__ vals,
                                                      const double *__restrict__ w, double *__restrict__ y,
                                                      const double *__restrict__ dotw, double *__restrict__ partial)
{
   __shared__ double sm[RED_THREADS];
   const int rows_per_block = RED_THREADS / TPR;
   const int lane = threadIdx.x % TPR, lr = threadIdx.x / TPR;
   double acc_dot = 0.0;
   // block-uniform trip count: every lane of a warp takes part in the shuffles of every iteration
   for (int base = r0 + blockIdx.x * rows_per_block; base < r1; base += gridDim.x * rows_per_block)
   {
      const int row = base + lr;
      const bool live = row < r1;
      double s = 0.0;
      if (live)
      {
         for (int k = rowptr[row] + lane; k < rowptr[row + 1]; k += TPR) { s = fma(vals[k], w[colidx[k]], s); }
      }
#pragma unroll
      for (int o = TPR / 2; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o, TPR); }
      if (live && lane == 0)
      {
         y[row] = s;
         if (dotw) { acc_dot = fma(s, dotw[row], acc_dot); }
      }
   }
   if (partial)
   {
      const double b = block_sum(acc_dot, sm);
      if (threadIdx.x == 0) { partial[blockIdx.x] = b; }
   }
}

/// out[k] = fixed-tree sum of partial[k * nb .. (k+1) * nb), k < nk (one block)
__global__ void __launch_bounds__(RED_THREADS) k_final(const double *partial, const int nb, const int nk, double *out)
{
   __shared__ double sm[RED_THREADS];
   for (int k = 0; k < nk; k++)
   {
      double a = 0.0;
      for (int i = threadIdx.x; i < nb; i += RED_THREADS) { a += partial[k * nb + i]; }
      const double s = block_sum(a, sm);
      if (threadIdx.x == 0) { out[k] = s; }
   }
}

// scalars on the device: S_RZ = r.z, S_PQ = p.q, S_RZN = new r.z, S_RR = r.r
enum { S_RZ = 0, S_PQ = 1, S_RZN = 2, S_RR = 3, S_N = 4 };

/// diagonal of rows [0, n) (columns sorted: binary search), inverted; zero diagonals give 1
__global__ void k_diag_inv(const int n, const int *rowptr, const int *colidx, const double *vals, double *dinv)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   int lo = rowptr[i], hi = rowptr[i + 1] - 1;
   double d = 0.0;
   while (lo <= hi)
   {
      const int m = (lo + hi) >> 1, c = colidx[m];
      if (c == i) { d = vals[m]; break; }
      if (c < i) { lo = m + 1; } else { hi = m - 1; }
   }
   dinv[i] = (d != 0.0) ? 1.0 / d : 1.0;
}

/// z = dinv r, partials of r.z and r.r
__global__ void __launch_bounds__(RED_THREADS) k_precond_dot(const int n, const double *r, const double *dinv, double *z, double *partial)
{
   __shared__ double sm[RED_THREADS];
   double a = 0.0, b = 0.0;
   for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += gridDim.x * RED_THREADS)
   {
      const double ri = r[i], zi = dinv[i] * ri;
      z[i] = zi;
      a = fma(ri, zi, a);
      b = fma(ri, ri, b);
   }
   const double sa = block_sum(a, sm), sb = block_sum(b, sm);
   if (threadIdx.x == 0)
   {
      partial[blockIdx.x] = sa;
      partial[gridDim.x + blockIdx.x] = sb;
   }
}

/// alpha = rz / pq; x += alpha p; r -= alpha q; z = dinv r; partials of the new r.z and r.r
__global__ void __launch_bounds__(RED_THREADS) k_cg_update(const int n, double *x, double *r, const double *p, const double *q,
                                                           const double *dinv, double *z, const double *scal, double *partial)
{
   __shared__ double sm[RED_THREADS];
   const double pq = scal[S_PQ], alpha = (pq != 0.0) ? scal[S_RZ] / pq : 0.0;
   double a = 0.0, b = 0.0;
   for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += gridDim.x * RED_THREADS)
   {
      x[i] = fma(alpha, p[i], x[i]);
      const double ri = fma(-alpha, q[i], r[i]), zi = dinv[i] * ri;
      r[i] = ri;
      z[i] = zi;
      a = fma(ri, zi, a);
      b = fma(ri, ri, b);
   }
   const double sa = block_sum(a, sm), sb = block_sum(b, sm);
   if (threadIdx.x == 0)
   {
      partial[blockIdx.x] = sa;
      partial[gridDim.x + blockIdx.x] = sb;
   }
}

/// beta = rz_new / rz; p = z + beta p (first == 1: p = z)
__global__ void k_cg_p(const int n, double *p, const double *z, const double *scal, const int first)
{
   const double rz = scal[S_RZ], beta = (first || rz == 0.0) ? 0.0 : scal[S_RZN] / rz;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { p[i] = fma(beta, p[i], z[i]); }
}
__global__ void k_shift_rz(double *scal) { scal[S_RZ] = scal[S_RZN]; }

__global__ void k_axpby(const int n, const double a, const double *x, const double b, double *y)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { y[i] = a * x[i] + b * y[i]; }
}
__global__ void k_fill(const int n, const double v, double *y)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { y[i] = v; }
}

/// Latent block: rows nh + [e * nb, (e + 1) * nb) hold a dense nb x nb block in the columns nh + [e * nb, ...) (L2 space,
/// element-local dofs).  out[e * nb + .] = sign * Block^-1 in[e * nb + .] by Gaussian elimination without pivoting (the block is
/// definite: -(1/alpha) int E*''(psi) chi chi, src/pg.hpp:193-213).  One thread per element, nb <= 9.
template <int NB>
__global__ void k_block_solve(const int nel, const int nh, const int *rowptr, const int *colidx, const double *vals,
                              const double *in, double *out, const double sign)
{
   const int e = blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= nel) { return; }
   double A[NB][NB], b[NB];
#pragma unroll
   for (int i = 0; i < NB; i++)
   {
      const int row = nh + e * NB + i;
      // first column >= nh + e * NB (columns sorted)
      int lo = rowptr[row], hi = rowptr[row + 1];
      const int c0 = nh + e * NB;
      while (lo < hi)
      {
         const int m = (lo + hi) >> 1;
         if (colidx[m] < c0) { lo = m + 1; } else { hi = m; }
      }
#pragma unroll
      for (int j = 0; j < NB; j++) { A[i][j] = vals[lo + j]; }
      b[i] = in[e * NB + i];
   }
#pragma unroll
   for (int k = 0; k < NB; k++)
   {
      const double piv = 1.0 / A[k][k];
#pragma unroll
      for (int i = k + 1; i < NB; i++)
      {
         const double f = A[i][k] * piv;
#pragma unroll
         for (int j = k + 1; j < NB; j++) { A[i][j] = fma(-f, A[k][j], A[i][j]); }
         b[i] = fma(-f, b[k], b[i]);
      }
   }
#pragma unroll
   for (int i = NB - 1; i >= 0; i--)
   {
      double s = b[i];
#pragma unroll
      for (int j = i + 1; j < NB; j++) { s = fma(-A[i][j], b[j], s); }
      b[i] = s / A[i][i];
   }
#pragma unroll
   for (int i = 0; i < NB; i++) { out[e * NB + i] = sign * b[i]; }
}

/// Jacobi preconditioner of the condensed operator S = A - C J22^-1 C^T: dinv[i] = 1 / (A_ii - sum_e c_ie^T J22_e^-1 c_ie), c_ie =
/// the entries of row i in the latent columns of element e (NB consecutive columns; columns sorted).  Where the entropy
/// Hessian is tiny (saturated latent variable) J22^-1 is huge and dominates the diagonal: diag(A) alone is useless there.
template <int NB>
__global__ void k_diag_schur(const int nh, const int *rowptr, const int *colidx, const double *vals, double *dinv)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= nh) { return; }
   double d = 0.0;
   int k = rowptr[i];
   const int ke = rowptr[i + 1];
   for (; k < ke && colidx[k] < nh; k++) { if (colidx[k] == i) { d = vals[k]; } }
   while (k < ke)
   {
      // one element group: the NB latent columns nh + e * NB .. (all present: full element connectivity)
      const int e = (colidx[k] - nh) / NB;
      double c[NB], A[NB][NB], b[NB];
#pragma unroll
      for (int j = 0; j < NB; j++) { c[j] = vals[k + j]; b[j] = c[j]; }
#pragma unroll
      for (int r = 0; r < NB; r++)
      {
         const int row = nh + e * NB + r;
         int lo = rowptr[row], hi = rowptr[row + 1];
         const int c0 = nh + e * NB;
         while (lo < hi)
         {
            const int m = (lo + hi) >> 1;
            if (colidx[m] < c0) { lo = m + 1; } else { hi = m; }
         }
#pragma unroll
         for (int j = 0; j < NB; j++) { A[r][j] = vals[lo + j]; }
      }
#pragma unroll
      for (int p = 0; p < NB; p++)
      {
         const double piv = 1.0 / A[p][p];
#pragma unroll
         for (int r = p + 1; r < NB; r++)
         {
            const double f = A[r][p] * piv;
#pragma unroll
            for (int j = p + 1; j < NB; j++) { A[r][j] = fma(-f, A[p][j], A[r][j]); }
            b[r] = fma(-f, b[p], b[r]);
         }
      }
      double acc = 0.0;
#pragma unroll
      for (int r = NB - 1; r >= 0; r--)
      {
         double s = b[r];
#pragma unroll
         for (int j = r + 1; j < NB; j++) { s = fma(-A[r][j], b[j], s); }
         b[r] = s / A[r][r];
         acc = fma(c[r], b[r], acc);
      }
      d -= acc;
      k += NB;
   }
   dinv[i] = (d != 0.0) ? 1.0 / d : 1.0;
}

/// partial of a.b (fixed grid)
__global__ void __launch_bounds__(RED_THREADS) k_dot(const int n, const double *a, const double *b, double *partial)
{
   __shared__ double sm[RED_THREADS];
   double acc = 0.0;
   for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += gridDim.x * RED_THREADS) { acc = fma(a[i], b[i], acc); }
   const double s = block_sum(acc, sm);
   if (threadIdx.x == 0) { partial[blockIdx.x] = s; }
}
/// y = a x + b y + c z
__global__ void k_lincomb3(const int n, const double a, const double *x, const double b, double *y, const double c, const double *z)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { y[i] = a * x[i] + b * y[i] + c * z[i]; }
}
/// out = a x + b y + c z (out may alias none of them)
__global__ void k_lincomb3o(const int n, const double a, const double *x, const double b, const double *y, const double c, const double *z, double *out)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { out[i] = a * x[i] + b * y[i] + c * z[i]; }
}
__global__ void k_scale(const int n, const double a, double *x)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { x[i] *= a; }
}

/// Block-diagonal preconditioner of the saddle-point system [[A, C], [C^T, -D]] (the PGPreconditioner idea,
/// src/pg.hpp:378-504, with Jacobi on the primal block): element blocks of the approximate latent Schur complement
///    S_e = D_e + C_e^T diag(A)^-1 C_e     (NB x NB, symmetric positive definite)
/// The latent rows of one element have the same primal columns in the same (sorted) order: full element connectivity.
template <int NB>
__global__ void k_schur_blocks(const int nel, const int nh, const int *rowptr, const int *colidx, const double *vals,
                               const double *dinvA, double *blocks)
{
   const int e = blockIdx.x * blockDim.x + threadIdx.x;
   if (e >= nel) { return; }
   double Sb[NB][NB];
   int start[NB], nprim = 0;
#pragma unroll
   for (int r = 0; r < NB; r++)
   {
      const int row = nh + e * NB + r;
      start[r] = rowptr[row];
      int lo = rowptr[row], hi = rowptr[row + 1];
      while (lo < hi) // first latent column of the row
      {
         const int m = (lo + hi) >> 1;
         if (colidx[m] < nh) { lo = m + 1; } else { hi = m; }
      }
      if (r == 0) { nprim = lo - rowptr[row]; }
      // latent block -D_e: the NB columns nh + e * NB ..
      int l2 = lo;
      while (colidx[l2] < nh + e * NB) { l2++; }
#pragma unroll
      for (int c = 0; c < NB; c++) { Sb[r][c] = -vals[l2 + c]; }
   }
   for (int k = 0; k < nprim; k++)
   {
      const double di = dinvA[colidx[start[0] + k]];
      double c[NB];
#pragma unroll
      for (int r = 0; r < NB; r++) { c[r] = vals[start[r] + k]; }
#pragma unroll
      for (int r = 0; r < NB; r++)
      {
#pragma unroll
         for (int q = 0; q < NB; q++) { Sb[r][q] = fma(c[r] * di, c[q], Sb[r][q]); }
      }
   }
#pragma unroll
   for (int r = 0; r < NB; r++)
   {
#pragma unroll
      for (int c = 0; c < NB; c++) { blocks[((size_t)e * NB + r) * NB + c] = Sb[r][c]; }
   }
}
/// z = P^-1 v: z_u = |dinvA| v_u (positive definite preconditioner), z_psi,e = S_e^-1 v_psi,e
template <int NB>
__global__ void k_apply_precond(const int nel, const int nh, const double *dinvA, const double *blocks, const double *v, double *z)
{
   const int t = blockIdx.x * blockDim.x + threadIdx.x;
   for (int i = t; i < nh; i += gridDim.x * blockDim.x) { z[i] = fabs(dinvA[i]) * v[i]; }
   for (int e = t; e < nel; e += gridDim.x * blockDim.x)
   {
      double A[NB][NB], b[NB];
#pragma unroll
      for (int r = 0; r < NB; r++)
      {
#pragma unroll
         for (int c = 0; c < NB; c++) { A[r][c] = blocks[((size_t)e * NB + r) * NB + c]; }
         b[r] = v[nh + e * NB + r];
      }
#pragma unroll
      for (int p = 0; p < NB; p++)
      {
         const double piv = 1.0 / A[p][p];
#pragma unroll
         for (int r = p + 1; r < NB; r++)
         {
            const double f = A[r][p] * piv;
#pragma unroll
            for (int c = p + 1; c < NB; c++) { A[r][c] = fma(-f, A[p][c], A[r][c]); }
            b[r] = fma(-f, b[p], b[r]);
         }
      }
#pragma unroll
      for (int r = NB - 1; r >= 0; r--)
      {
         double s = b[r];
#pragma unroll
         for (int c = r + 1; c < NB; c++) { s = fma(-A[r][c], b[c], s); }
         b[r] = s / A[r][r];
      }
#pragma unroll
      for (int r = 0; r < NB; r++) { z[nh + e * NB + r] = b[r]; }
   }
}

/// The block solvers assume an element-local latent space: row nh + e * nb + i has exactly nb latent columns, nh + e * nb ..
/// (sorted).  flag[0] is set when a row violates that (an H1 latent space, a different numbering).
__global__ void k_check_latent_blocks(const int nel, const int nh, const int nb, const int *rowptr, const int *colidx, int *flag)
{
   const int r = blockIdx.x * blockDim.x + threadIdx.x;
   if (r >= nel * nb) { return; }
   const int row = nh + r, c0 = nh + (r / nb) * nb;
   int lo = rowptr[row], hi = rowptr[row + 1];
   const int end = hi;
   while (lo < hi)
   {
      const int m = (lo + hi) >> 1;
      if (colidx[m] < nh) { lo = m + 1; } else { hi = m; }
   }
   if (end - lo != nb || colidx[lo] != c0 || colidx[end - 1] != c0 + nb - 1) { flag[0] = 1; }
}

bool is_dev(const void *p)
{
   if (!p) { return false; }
   cudaPointerAttributes at;
   if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
   return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}
} // namespace

struct Solver
{
   Ctx *ctx = nullptr;
   int n = 0;
   long nnz = 0;
   const int *rowptr = nullptr, *colidx = nullptr; // device (the integrator's pattern, or own copies)
   int *own_rowptr = nullptr, *own_colidx = nullptr;
   double *work = nullptr; // r, z, p, q, dinv, w, t1, t2, bS (9 n) + partials + scalars
   double *work2 = nullptr; // MINRES: 8 vectors + preconditioner blocks
   size_t work2_words = 0;
   double *vals_buf = nullptr, *b_buf = nullptr, *x_buf = nullptr;
   int tpr = 8;
   ~Solver()
   {
      cudaFree(own_rowptr); cudaFree(own_colidx); cudaFree(work); cudaFree(work2); cudaFree(vals_buf); cudaFree(b_buf); cudaFree(x_buf);
   }
};

#define SOLVE_OK(call)                                                                              \
   do                                                                                               \
   {                                                                                                \
      const cudaError_t e_ = (call);                                                                \
      if (e_ != cudaSuccess)                                                                        \
      {                                                                                             \
         set_error(std::string("madb_solver: ") + cudaGetErrorString(e_));                          \
         return 2;                                                                                  \
      }                                                                                             \
   } while (0)

static void spmv(const Solver &S, cudaStream_t st, int r0, int r1, const double *vals, const double *w, double *y, const double *dotw,
                 double *partial)
{
   switch (S.tpr)
   {
      case 4: k_spmv<4><<<RED_BLOCKS, RED_THREADS, 0, st>>>(r0, r1, S.rowptr, S.colidx, vals, w, y, dotw, partial); break;
      case 16: k_spmv<16><<<RED_BLOCKS, RED_THREADS, 0, st>>>(r0, r1, S.rowptr, S.colidx, vals, w, y, dotw, partial); break;
      case 32: k_spmv<32><<<RED_BLOCKS, RED_THREADS, 0, st>>>(r0, r1, S.rowptr, S.colidx, vals, w, y, dotw, partial); break;
      default: k_spmv<8><<<RED_BLOCKS, RED_THREADS, 0, st>>>(r0, r1, S.rowptr, S.colidx, vals, w, y, dotw, partial); break;
   }
}

/// refuses systems whose latent block is not block diagonal with nb x nb element blocks
static int check_latent_blocks(Solver &S, cudaStream_t st, int nel, int nh, int nb, const char *who)
{
   int *flag = reinterpret_cast<int *>(S.work + 9 * (size_t)S.n + 2 * RED_BLOCKS + 8); // behind the CG scalars
   SOLVE_OK(cudaMemsetAsync(flag, 0, sizeof(int), st));
   k_check_latent_blocks<<<(nel * nb + 255) / 256, 256, 0, st>>>(nel, nh, nb, S.rowptr, S.colidx, flag);
   int h = 0;
   SOLVE_OK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
   SOLVE_OK(cudaStreamSynchronize(st));
   if (h)
   {
      set_error(std::string(who) + ": the unknowns [nh, n) are not an element-local (L2) latent space with nb dofs per element: "
                                   "the latent block of the Jacobian is not block diagonal");
      return 1;
   }
   return 0;
}

static int block_solve(const Solver &S, cudaStream_t st, int nel, int nh, int nb, const double *vals, const double *in, double *out, double sign)
{
   const int grid = (nel + 127) / 128;
   switch (nb)
   {
      case 1: k_block_solve<1><<<grid, 128, 0, st>>>(nel, nh, S.rowptr, S.colidx, vals, in, out, sign); break;
      case 3: k_block_solve<3><<<grid, 128, 0, st>>>(nel, nh, S.rowptr, S.colidx, vals, in, out, sign); break;
      case 4: k_block_solve<4><<<grid, 128, 0, st>>>(nel, nh, S.rowptr, S.colidx, vals, in, out, sign); break;
      case 8: k_block_solve<8><<<grid, 128, 0, st>>>(nel, nh, S.rowptr, S.colidx, vals, in, out, sign); break;
      case 9: k_block_solve<9><<<grid, 128, 0, st>>>(nel, nh, S.rowptr, S.colidx, vals, in, out, sign); break;
      default: set_error("madb_solver_condensed_pcg: latent blocks of 1, 3, 4, 8 or 9 dofs per element"); return 1;
   }
   return 0;
}

/// Conjugate gradients on rows / columns [0, m) of the operator `apply` (q = Op p, partial of p.q), Jacobi preconditioner dinv.
template <class Apply>
static int cg_loop(Solver &S, cudaStream_t st, const int m, Apply &&apply, const double *b, double *x, const double *dinv, const double rtol,
                   const double atol, const int maxit, int *iters, double *relres)
{
   double *r = S.work, *z = r + S.n, *p = z + S.n, *q = p + S.n;
   double *partial = S.work + 9 * (size_t)S.n, *scal = partial + 2 * RED_BLOCKS;
   double h[S_N];
   // r = b - Op x
   apply(x, q, nullptr, nullptr);
   SOLVE_OK(cudaMemcpyAsync(r, b, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, st));
   k_axpby<<<RED_BLOCKS, RED_THREADS, 0, st>>>(m, -1.0, q, 1.0, r);
   k_precond_dot<<<RED_BLOCKS, RED_THREADS, 0, st>>>(m, r, dinv, z, partial);
   k_final<<<1, RED_THREADS, 0, st>>>(partial, RED_BLOCKS, 2, scal + S_RZN);
   k_shift_rz<<<1, 1, 0, st>>>(scal);
   SOLVE_OK(cudaMemcpyAsync(h, scal, sizeof(h), cudaMemcpyDeviceToHost, st));
   SOLVE_OK(cudaStreamSynchronize(st));
   // reference norm: ||b|| (MFEM's CGSolver uses the preconditioned initial residual; the norm of b is rank independent)
   k_precond_dot<<<RED_BLOCKS, RED_THREADS, 0, st>>>(m, b, dinv, q, partial);
   k_final<<<1, RED_THREADS, 0, st>>>(partial + RED_BLOCKS, RED_BLOCKS, 1, scal + S_PQ);
   double bb = 0.0;
   SOLVE_OK(cudaMemcpyAsync(&bb, scal + S_PQ, sizeof(double), cudaMemcpyDeviceToHost, st));
   SOLVE_OK(cudaStreamSynchronize(st));
   const double bnorm = std::sqrt(bb), tol = std::max(rtol * bnorm, atol);
   double rnorm = std::sqrt(h[S_RR]);
   int it = 0;
   k_cg_p<<<RED_BLOCKS, RED_THREADS, 0, st>>>(m, p, z, scal, 1);
   // One burst = check_every iterations = 6 launches each, scalars on the device: captured once per solve as a CUDA graph
   // and replayed (the loop is launch bound on small and medium systems: 49 us per iteration un-captured at 263 k dofs);
   // the host reads the residual norm after every burst.  MADB_SOLVER_GRAPH=0: plain launches.
   const int check_every = 8;
   auto burst = [&]()
   {
      for (int k = 0; k < check_every; k++)
      {
         apply(p, q, p, partial);
         k_final<<<1, RED_THREADS, 0, st>>>(partial, RED_BLOCKS, 1, scal + S_PQ);
         k_cg_update<<<RED_BLOCKS, RED_THREADS, 0, st>>>(m, x, r, p, q, dinv, z, scal, partial);
         k_final<<<1, RED_THREADS, 0, st>>>(partial, RED_BLOCKS, 2, scal + S_RZN);
         k_cg_p<<<RED_BLOCKS, RED_THREADS, 0, st>>>(m, p, z, scal, 0);
         k_shift_rz<<<1, 1, 0, st>>>(scal);
      }
   };
   static const bool use_graph = !(getenv("MADB_SOLVER_GRAPH") && atoi(getenv("MADB_SOLVER_GRAPH")) == 0);
   cudaGraph_t graph = nullptr;
   cudaGraphExec_t exec = nullptr;
   if (use_graph && rnorm > tol && maxit > 0)
   {
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess)
      {
         burst();
         if (cudaStreamEndCapture(st, &graph) != cudaSuccess || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess)
         {
            cudaGetLastError();
            exec = nullptr;
         }
      }
      else { cudaGetLastError(); }
   }
   int rc = 0;
   while (rnorm > tol && it < maxit)
   {
      if (exec) { cudaGraphLaunch(exec, st); }
      else { burst(); }
      it += check_every; // whole bursts: may pass maxit by up to check_every - 1 iterations
      if (cudaMemcpyAsync(h, scal, sizeof(h), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
      {
         set_error(std::string("madb_solver: ") + cudaGetErrorString(cudaGetLastError()));
         rc = 2;
         break;
      }
      rnorm = std::sqrt(h[S_RR]);
      if (!(rnorm == rnorm))
      {
         set_error("madb_solver: conjugate gradients broke down (NaN): the operator is not positive definite");
         rc = 3;
         break;
      }
   }
   if (exec) { cudaGraphExecDestroy(exec); }
   if (graph) { cudaGraphDestroy(graph); }
   if (rc) { return rc; }
   if (iters) { *iters = it; }
   if (relres) { *relres = (bnorm > 0.0) ? rnorm / bnorm : rnorm; }
   return 0;
}

struct StageIO
{
   const double *d = nullptr;
   double *dw = nullptr;
   double *host = nullptr;
   size_t n = 0;
};

} // namespace madb

using namespace madb;

extern "C"
{
   int madb_solver_destroy(madb_solver *s);
}
namespace madb
{
/// the solver works on the device pattern of an integrator (madb_solver_create, madb_runtime.cu)
int solver_new(Ctx *ctx, int n, long nnz, const int *d_rowptr, const int *d_colidx, madb_solver **out)
{
   Solver *S = new Solver;
   S->ctx = ctx;
   S->n = n;
   S->nnz = nnz;
   S->rowptr = d_rowptr;
   S->colidx = d_colidx;
   const double avg = (double)nnz / std::max<long>(n, 1);
   S->tpr = avg <= 6 ? 4 : (avg <= 24 ? 8 : (avg <= 64 ? 16 : 32));
   const size_t words = 9 * (size_t)n + 2 * RED_BLOCKS + 16;
   if (cudaMalloc((void **)&S->work, words * sizeof(double)) != cudaSuccess)
   {
      delete S;
      set_error("madb_solver_create: out of device memory");
      return 2;
   }
   *out = reinterpret_cast<madb_solver *>(S);
   return 0;
}
} // namespace madb
extern "C"
{
   int madb_solver_destroy(madb_solver *s)
   {
      delete reinterpret_cast<Solver *>(s);
      return 0;
   }

   static int stage_in(Solver &S, const double *p, size_t n, double **buf, const double **out)
   {
      if (is_dev(p)) { *out = p; return 0; }
      if (!*buf) { SOLVE_OK(cudaMalloc((void **)buf, std::max<size_t>(n, 1) * sizeof(double))); }
      SOLVE_OK(cudaMemcpyAsync(*buf, p, n * sizeof(double), cudaMemcpyHostToDevice, S.ctx->stream));
      *out = *buf;
      return 0;
   }

   int madb_solver_pcg(madb_solver *s, const double *vals, const double *b, double *x, double rtol, double atol, int maxit,
                       int *iters, double *relres)
   {
      Solver &S = *reinterpret_cast<Solver *>(s);
      SOLVE_OK(cudaSetDevice(S.ctx->device));
      cudaStream_t st = S.ctx->stream;
      const double *dv, *db, *dx0;
      if (stage_in(S, vals, (size_t)S.nnz, &S.vals_buf, &dv) || stage_in(S, b, (size_t)S.n, &S.b_buf, &db) ||
          stage_in(S, x, (size_t)S.n, &S.x_buf, &dx0))
      {
         return 2;
      }
      double *dx = const_cast<double *>(dx0);
      double *dinv = S.work + 4 * (size_t)S.n;
      k_diag_inv<<<(S.n + 255) / 256, 256, 0, st>>>(S.n, S.rowptr, S.colidx, dv, dinv);
      auto apply = [&](const double *w, double *y, const double *dotw, double *partial) { spmv(S, st, 0, S.n, dv, w, y, dotw, partial); };
      const int rc = cg_loop(S, st, S.n, apply, db, dx, dinv, rtol, atol, maxit, iters, relres);
      if (rc) { return rc; }
      if (!is_dev(x)) { SOLVE_OK(cudaMemcpyAsync(x, dx, (size_t)S.n * sizeof(double), cudaMemcpyDeviceToHost, st)); }
      SOLVE_OK(cudaStreamSynchronize(st));
      return 0;
   }

   int madb_solver_condensed_pcg(madb_solver *s, int nh, int nb, const double *vals, const double *b, double *x, double rtol,
                                 double atol, int maxit, int *iters, double *relres)
   {
      Solver &S = *reinterpret_cast<Solver *>(s);
      SOLVE_OK(cudaSetDevice(S.ctx->device));
      cudaStream_t st = S.ctx->stream;
      const int n = S.n, nl = n - nh;
      if (nh <= 0 || nl <= 0 || nb <= 0 || nl % nb != 0) { set_error("madb_solver_condensed_pcg: bad block sizes"); return 1; }
      const int nel = nl / nb;
      if (check_latent_blocks(S, st, nel, nh, nb, "madb_solver_condensed_pcg")) { return 1; }
      const double *dv, *db, *dx0;
      if (stage_in(S, vals, (size_t)S.nnz, &S.vals_buf, &dv) || stage_in(S, b, (size_t)n, &S.b_buf, &db) || stage_in(S, x, (size_t)n, &S.x_buf, &dx0))
      {
         return 2;
      }
      double *dx = const_cast<double *>(dx0);
      double *dinv = S.work + 4 * (size_t)n, *w = dinv + n, *t1 = w + n, *t2 = t1 + n, *bS = t2 + n;
      // Jacobi preconditioner: the diagonal of S itself
      switch (nb)
      {
         case 1: k_diag_schur<1><<<(nh + 127) / 128, 128, 0, st>>>(nh, S.rowptr, S.colidx, dv, dinv); break;
         case 3: k_diag_schur<3><<<(nh + 127) / 128, 128, 0, st>>>(nh, S.rowptr, S.colidx, dv, dinv); break;
         case 4: k_diag_schur<4><<<(nh + 127) / 128, 128, 0, st>>>(nh, S.rowptr, S.colidx, dv, dinv); break;
         case 8: k_diag_schur<8><<<(nh + 127) / 128, 128, 0, st>>>(nh, S.rowptr, S.colidx, dv, dinv); break;
         case 9: k_diag_schur<9><<<(nh + 127) / 128, 128, 0, st>>>(nh, S.rowptr, S.colidx, dv, dinv); break;
         default: set_error("madb_solver_condensed_pcg: latent blocks of 1, 3, 4, 8 or 9 dofs per element"); return 1;
      }
      // S p = A p + C D^-1 C^T p with J22 = -D:  t1 = C^T p (rows >= nh of J applied to [p; 0]),
      //                                           t2 = -J22^-1 t1,  q = J[:nh, :] [p; t2]
      auto apply = [&](const double *p, double *q, const double *dotw, double *partial)
      {
         cudaMemcpyAsync(w, p, (size_t)nh * sizeof(double), cudaMemcpyDeviceToDevice, st);
         k_fill<<<RED_BLOCKS, RED_THREADS, 0, st>>>(nl, 0.0, w + nh);
         spmv(S, st, nh, n, dv, w, t1 - nh, nullptr, nullptr); // y[row] with row >= nh: t1[row - nh]
         block_solve(S, st, nel, nh, nb, dv, t1, w + nh, -1.0);
         spmv(S, st, 0, nh, dv, w, q, dotw, partial);
      };
      // right-hand side of the condensed system: b_u - C J22^-1 b_psi
      if (block_solve(S, st, nel, nh, nb, dv, db + nh, w + nh, 1.0)) { return 1; }
      k_fill<<<RED_BLOCKS, RED_THREADS, 0, st>>>(nh, 0.0, w);
      spmv(S, st, 0, nh, dv, w, t2, nullptr, nullptr); // t2 = C J22^-1 b_psi
      SOLVE_OK(cudaMemcpyAsync(bS, db, (size_t)nh * sizeof(double), cudaMemcpyDeviceToDevice, st));
      k_axpby<<<RED_BLOCKS, RED_THREADS, 0, st>>>(nh, -1.0, t2, 1.0, bS);
      const int rc = cg_loop(S, st, nh, apply, bS, dx, dinv, rtol, atol, maxit, iters, relres);
      if (rc) { return rc; }
      // back-substitution: x_psi = J22^-1 (b_psi - C^T x_u)
      SOLVE_OK(cudaMemcpyAsync(w, dx, (size_t)nh * sizeof(double), cudaMemcpyDeviceToDevice, st));
      k_fill<<<RED_BLOCKS, RED_THREADS, 0, st>>>(nl, 0.0, w + nh);
      spmv(S, st, nh, n, dv, w, t1 - nh, nullptr, nullptr);
      k_axpby<<<RED_BLOCKS, RED_THREADS, 0, st>>>(nl, 1.0, db + nh, -1.0, t1);
      if (block_solve(S, st, nel, nh, nb, dv, t1, dx + nh, 1.0)) { return 1; }
      if (!is_dev(x)) { SOLVE_OK(cudaMemcpyAsync(x, dx, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st)); }
      SOLVE_OK(cudaStreamSynchronize(st));
      return 0;
   }

   int madb_solver_pg_minres(madb_solver *s, int nh, int nb, const double *vals, const double *b, double *x, double rtol,
                             double atol, int maxit, int *iters, double *relres)
   {
      Solver &S = *reinterpret_cast<Solver *>(s);
      SOLVE_OK(cudaSetDevice(S.ctx->device));
      cudaStream_t st = S.ctx->stream;
      const int n = S.n;
      // nb == 0: no block structure is used: Jacobi preconditioner |diag J|^-1 on all unknowns (symmetric indefinite systems
      // in general, e.g. ex5's H1 latent space)
      const bool plain = (nb == 0);
      if (plain) { nh = n; }
      const int nl = n - nh;
      if (!plain && (nh <= 0 || nl <= 0 || nb < 0 || nl % nb != 0)) { set_error("madb_solver_pg_minres: bad block sizes"); return 1; }
      const int nel = plain ? 0 : nl / nb;
      if (!plain && check_latent_blocks(S, st, nel, nh, nb, "madb_solver_pg_minres")) { return 1; }
      const double *dv, *db, *dx0;
      if (stage_in(S, vals, (size_t)S.nnz, &S.vals_buf, &dv) || stage_in(S, b, (size_t)n, &S.b_buf, &db) || stage_in(S, x, (size_t)n, &S.x_buf, &dx0))
      {
         return 2;
      }
      double *dx = const_cast<double *>(dx0);
      const size_t need = 8 * (size_t)n + (size_t)nl * (plain ? 0 : nb) + 16;
      if (S.work2_words < need)
      {
         cudaFree(S.work2);
         S.work2 = nullptr;
         SOLVE_OK(cudaMalloc((void **)&S.work2, need * sizeof(double)));
         S.work2_words = need;
      }
      double *v0 = S.work2, *v1 = v0 + n, *v2 = v1 + n, *z1 = v2 + n, *z2 = z1 + n, *w0 = z2 + n, *w1 = w0 + n, *w2 = w1 + n, *blocks = w2 + n;
      double *dinvA = S.work + 4 * (size_t)n, *partial = S.work + 9 * (size_t)n, *scal = partial + 2 * RED_BLOCKS;
      k_diag_inv<<<(nh + 255) / 256, 256, 0, st>>>(nh, S.rowptr, S.colidx, dv, dinvA);
      auto precond = [&](const double *v, double *z) -> int
      {
         const int grid = RED_BLOCKS;
         switch (nb)
         {
            case 0:
            case 1: k_apply_precond<1><<<grid, RED_THREADS, 0, st>>>(nel, nh, dinvA, blocks, v, z); break;
            case 3: k_apply_precond<3><<<grid, RED_THREADS, 0, st>>>(nel, nh, dinvA, blocks, v, z); break;
            case 4: k_apply_precond<4><<<grid, RED_THREADS, 0, st>>>(nel, nh, dinvA, blocks, v, z); break;
            case 8: k_apply_precond<8><<<grid, RED_THREADS, 0, st>>>(nel, nh, dinvA, blocks, v, z); break;
            case 9: k_apply_precond<9><<<grid, RED_THREADS, 0, st>>>(nel, nh, dinvA, blocks, v, z); break;
            default: return 1;
         }
         return 0;
      };
      {
         const int grid = (nel + 127) / 128;
         switch (nb)
         {
            case 0: break;
            case 1: k_schur_blocks<1><<<grid, 128, 0, st>>>(nel, nh, S.rowptr, S.colidx, dv, dinvA, blocks); break;
            case 3: k_schur_blocks<3><<<grid, 128, 0, st>>>(nel, nh, S.rowptr, S.colidx, dv, dinvA, blocks); break;
            case 4: k_schur_blocks<4><<<grid, 128, 0, st>>>(nel, nh, S.rowptr, S.colidx, dv, dinvA, blocks); break;
            case 8: k_schur_blocks<8><<<grid, 128, 0, st>>>(nel, nh, S.rowptr, S.colidx, dv, dinvA, blocks); break;
            case 9: k_schur_blocks<9><<<grid, 128, 0, st>>>(nel, nh, S.rowptr, S.colidx, dv, dinvA, blocks); break;
            default: set_error("madb_solver_pg_minres: latent blocks of 1, 3, 4, 8 or 9 dofs per element"); return 1;
         }
      }
      auto dot = [&](const double *a, const double *c, double *out) -> int
      {
         k_dot<<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, a, c, partial);
         k_final<<<1, RED_THREADS, 0, st>>>(partial, RED_BLOCKS, 1, scal);
         SOLVE_OK(cudaMemcpyAsync(out, scal, sizeof(double), cudaMemcpyDeviceToHost, st));
         SOLVE_OK(cudaStreamSynchronize(st));
         return 0;
      };
      // preconditioned MINRES (Paige & Saunders; notation of Elman, Silvester, Wathen, Algorithm 4.1)
      SOLVE_OK(cudaMemsetAsync(v0, 0, (size_t)n * sizeof(double), st));
      SOLVE_OK(cudaMemsetAsync(w0, 0, (size_t)n * sizeof(double), st));
      SOLVE_OK(cudaMemsetAsync(w1, 0, (size_t)n * sizeof(double), st));
      spmv(S, st, 0, n, dv, dx, v1, nullptr, nullptr);
      k_axpby<<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, 1.0, db, -1.0, v1); // v1 = b - A x0
      if (precond(v1, z1)) { return 1; }
      double g2 = 0.0, bb = 0.0;
      if (dot(z1, v1, &g2)) { return 2; }
      if (precond(db, z2) || dot(z2, db, &bb)) { return 2; }
      double gamma0 = 1.0, gamma1 = std::sqrt(std::max(g2, 0.0)), eta = gamma1, s0 = 0.0, s1 = 0.0, c0 = 1.0, c1 = 1.0;
      const double bnorm = std::sqrt(std::max(bb, 0.0)), tol = std::max(rtol * bnorm, atol);
      int it = 0;
      while (std::fabs(eta) > tol && it < maxit && gamma1 > 0.0)
      {
         k_scale<<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, 1.0 / gamma1, z1);
         spmv(S, st, 0, n, dv, z1, v2, nullptr, nullptr); // v2 = A z1
         double delta = 0.0;
         if (dot(v2, z1, &delta)) { return 2; }
         k_lincomb3<<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, -delta / gamma1, v1, 1.0, v2, -gamma1 / gamma0, v0);
         if (precond(v2, z2)) { return 1; }
         double g2n = 0.0;
         if (dot(z2, v2, &g2n)) { return 2; }
         const double gamma2 = std::sqrt(std::max(g2n, 0.0));
         const double a0 = c1 * delta - c0 * s1 * gamma1, a1 = std::sqrt(a0 * a0 + gamma2 * gamma2), a2 = s1 * delta + c0 * c1 * gamma1,
                      a3 = s0 * gamma1;
         const double c2 = a0 / a1, s2 = gamma2 / a1;
         k_lincomb3o<<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, 1.0 / a1, z1, -a3 / a1, w0, -a2 / a1, w1, w2);
         k_axpby<<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, c2 * eta, w2, 1.0, dx);
         eta = -s2 * eta;
         // rotate
         double *t = v0; v0 = v1; v1 = v2; v2 = t;
         t = z1; z1 = z2; z2 = t;
         t = w0; w0 = w1; w1 = w2; w2 = t;
         gamma0 = gamma1; gamma1 = gamma2;
         c0 = c1; c1 = c2; s0 = s1; s1 = s2;
         it++;
         if (!(eta == eta)) { set_error("madb_solver_pg_minres: breakdown (NaN)"); return 3; }
      }
      if (iters) { *iters = it; }
      if (relres) { *relres = (bnorm > 0.0) ? std::fabs(eta) / bnorm : std::fabs(eta); }
      if (!is_dev(x)) { SOLVE_OK(cudaMemcpyAsync(x, dx, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st)); }
      SOLVE_OK(cudaStreamSynchronize(st));
      return 0;
   }

   int madb_csr_spmv(madb_solver *s, const double *vals, const double *x, double *y)
   {
      Solver &S = *reinterpret_cast<Solver *>(s);
      SOLVE_OK(cudaSetDevice(S.ctx->device));
      cudaStream_t st = S.ctx->stream;
      const double *dv, *dx;
      if (stage_in(S, vals, (size_t)S.nnz, &S.vals_buf, &dv) || stage_in(S, x, (size_t)S.n, &S.x_buf, &dx)) { return 2; }
      double *dy = is_dev(y) ? y : S.work;
      spmv(S, st, 0, S.n, dv, dx, dy, nullptr, nullptr);
      if (!is_dev(y)) { SOLVE_OK(cudaMemcpyAsync(y, dy, (size_t)S.n * sizeof(double), cudaMemcpyDeviceToHost, st)); }
      SOLVE_OK(cudaStreamSynchronize(st));
      return 0;
   }
}

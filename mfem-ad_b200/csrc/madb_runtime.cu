// madb_runtime.cu -- C ABI entry points (include/mfemad_b200.h) and run-time plumbing.
#include "../../include/mfemad_b200.h"
#include "madb_config.cuh"
#include "madb_eval.cuh"
#include "madb_host.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cstdlib>

namespace madb
{

static thread_local std::string g_err;
const char *last_error() { return g_err.c_str(); }
void set_error(const std::string &s) { g_err = s; }

std::map<std::string, KernelOps> &registry()
{
   static std::map<std::string, KernelOps> r;
   return r;
}
Registrar::Registrar(const std::string &key, const KernelOps &ops) { registry()[key] = ops; }

std::map<std::string, EvalOps> &eval_registry()
{
   static std::map<std::string, EvalOps> r;
   return r;
}
EvalRegistrar::EvalRegistrar(const std::string &key, const EvalOps &ops) { eval_registry()[key] = ops; }

std::map<std::string, VecEvalOps> &vec_eval_registry()
{
   static std::map<std::string, VecEvalOps> r;
   return r;
}
VecEvalRegistrar::VecEvalRegistrar(const std::string &key, const VecEvalOps &ops) { vec_eval_registry()[key] = ops; }

std::string Functional::key() const
{
   std::string k = kind;
   for (size_t i = 0; i < iparams.size(); i++) { k += (i == 0 ? ":" : ",") + std::to_string(iparams[i]); }
   if (!children.empty())
   {
      k += "[";
      for (size_t i = 0; i < children.size(); i++) { k += (i ? "," : "") + children[i]->key(); }
      k += "]";
   }
   return k;
}
void Functional::flat_params(std::vector<double> &out) const
{
   out.insert(out.end(), params.begin(), params.end());
   for (const Functional *c : children) { c->flat_params(out); }
}

Integrator::~Integrator()
{
   cudaFree(d_e2n); cudaFree(d_vmap); cudaFree(d_pmap); cudaFree(d_e2csr); cudaFree(d_xe);
   cudaFree(d_rowptr); cudaFree(d_colidx); cudaFree(d_perm); cudaFree(d_cvalue); cudaFree(d_cgrad); cudaFree(d_chess); cudaFree(d_energy); cudaFree(d_esum);
   cudaFree(d_x); cudaFree(d_v); cudaFree(d_v2); cudaFree(d_y); cudaFree(d_vals); cudaFree(d_qf); cudaFree(d_ess);
   for (double *p : d_pstage) { cudaFree(p); }
   if (ev0) { cudaEventDestroy(ev0); cudaEventDestroy(ev1); }
   cudaFree(d_pdesc); cudaFree(d_yblob); cudaFree(d_vblob); cudaFree(d_idesc); cudaFree(d_mblob);
   cudaFree(d_ystage); cudaFree(d_vstage);
   for (int a = 0; a < 2; a++) { for (int b = 0; b < 5; b++) { cudaFree(d_ifc[a][b]); } }
}

#define CUDA_OK(call)                                                                                 \
   do {                                                                                               \
      cudaError_t e_ = (call);                                                                        \
      if (e_ != cudaSuccess)                                                                          \
      {                                                                                               \
         set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                               \
         return 2;                                                                                    \
      }                                                                                               \
   } while (0)

static bool is_device_ptr(const void *p)
{
   if (!p) { return false; }
   cudaPointerAttributes at;
   if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
   return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// deterministic fixed-tree sum: one block, fixed strided partials, shared-memory tree
__global__ void __launch_bounds__(1024) k_reduce_sum(const double *in, int n, double *out)
{
   __shared__ double s[1024];
   double acc = 0.0;
   for (int i = threadIdx.x; i < n; i += 1024) { acc += in[i]; }
   s[threadIdx.x] = acc;
   __syncthreads();
   for (int k = 512; k > 0; k >>= 1)
   {
      if ((int)threadIdx.x < k) { s[threadIdx.x] += s[threadIdx.x + k]; }
      __syncthreads();
   }
   if (threadIdx.x == 0) { out[0] = s[0]; }
}

// Fused LVPP latent-variable update (ex4.cpp:188-189,203-218):
//    lambda = (psi - psi_k)/alpha ;  diff += w |lambda - lambda_prev| ;  lambda_prev <- lambda ;  psi_k <- psi
// One pass over the latent vector; block partials are summed in a fixed order (deterministic).
__global__ void __launch_bounds__(256) k_lvpp_update(int n, double alpha, const double *psi, double *psik,
                                                     double *lam_prev, const double *w, double *partial)
{
   __shared__ double s[256];
   double acc = 0.0;
   const double ia = 1.0 / alpha;
   for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256)
   {
      const double p = psi[i];
      const double lam = (p - psik[i]) * ia;
      acc += (w ? w[i] : 1.0) * fabs(lam - lam_prev[i]);
      lam_prev[i] = lam;
      psik[i] = p;
   }
   s[threadIdx.x] = acc;
   __syncthreads();
   for (int k = 128; k > 0; k >>= 1)
   {
      if ((int)threadIdx.x < k) { s[threadIdx.x] += s[threadIdx.x + k]; }
      __syncthreads();
   }
   if (threadIdx.x == 0) { partial[blockIdx.x] = s[0]; }
}

// shared-dof exchange: contiguous send buffer <- y[idx] ; y[idx] += / = recv buffer (fixed order)
__global__ void k_pack(int n, const int *idx, const double *src, double *dst)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { dst[i] = src[idx[i]]; }
}
__global__ void k_unpack(int n, const int *idx, const double *src, double *dst, int add)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { dst[idx[i]] = add ? dst[idx[i]] + src[i] : src[i]; }
}

__global__ void k_unpack_multi(int n, const int4 *__restrict__ src4, const int *__restrict__ idx, const double *__restrict__ src,
                               double *dst, int add)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   const int4 s = src4[i];
   const int d = idx[i];
   const double a0 = src[s.x], a1 = (s.y >= 0) ? src[s.y] : 0.0, a2 = (s.z >= 0) ? src[s.z] : 0.0, a3 = (s.w >= 0) ? src[s.w] : 0.0;
   double v = add ? dst[d] + a0 : a0;
   if (s.y >= 0) { v += a1; }
   if (s.z >= 0) { v += a2; }
   if (s.w >= 0) { v += a3; }
   dst[d] = v;
}

// y[ess] = 0  (NonlinearForm::Mult [MFEM-upstream])
__global__ void k_ess_zero(const int *ess, int n, double *y)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { y[ess[i]] = 0.0; }
}
// y[ess] = v[ess]: rows of the eliminated Jacobian are unit rows
__global__ void k_ess_copy(const int *ess, int n, const double *v, double *y)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) { y[ess[i]] = v[ess[i]]; }
}
// SparseMatrix::EliminateRowCol(rc, DIAG_ONE) on a symmetric pattern with sorted columns
__global__ void k_ess_rowcol(const int *ess, int n, const int *rowptr, const int *colidx, double *vals)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) { return; }
   const int rc = ess[i];
   for (int p = rowptr[rc]; p < rowptr[rc + 1]; p++)
   {
      const int j = colidx[p];
      vals[p] = (j == rc) ? 1.0 : 0.0;
      if (j != rc)
      {
         int lo = rowptr[j], hi = rowptr[j + 1];
         while (lo < hi)
         {
            const int mid = (lo + hi) >> 1;
            if (colidx[mid] < rc) { lo = mid + 1; } else { hi = mid; }
         }
         if (lo < rowptr[j + 1] && colidx[lo] == rc) { vals[lo] = 0.0; }
      }
   }
}

template <class T> static int upload(const std::vector<T> &h, T **d)
{
   CUDA_OK(cudaMalloc((void **)d, std::max<size_t>(h.size(), 1) * sizeof(T)));
   CUDA_OK(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
   return 0;
}

static int upload_ifc(const PatchHost &H, int *(&d)[5], IfcListDev &L)
{
   if (upload(H.src4, &d[0]) || upload(H.dst4, &d[1]) || upload(H.ptr, &d[2]) || upload(H.src, &d[3]) || upload(H.dst, &d[4])) { return 2; }
   L.n4 = (int)H.dst4.size();
   L.ng = (int)H.dst.size();
   L.src4 = (const int4 *)d[0];
   L.dst4 = d[1];
   L.ptr = d[2]; L.src = d[3]; L.dst = d[4];
   L.stage = nullptr; L.out = nullptr;
   return 0;
}

static int ensure_pattern_device(Integrator &I);
int solver_new(Ctx *ctx, int n, long nnz, const int *d_rowptr, const int *d_colidx, madb_solver **out); // madb_solve.cu
static void release_device_maps(Integrator &I)
{
   cudaFree(I.d_e2n); cudaFree(I.d_vmap); cudaFree(I.d_pmap); cudaFree(I.d_perm); cudaFree(I.d_e2csr); cudaFree(I.d_xe);
   I.d_xe = nullptr;
   cudaFree(I.d_pdesc); cudaFree(I.d_yblob); cudaFree(I.d_vblob); cudaFree(I.d_ystage); cudaFree(I.d_vstage);
   cudaFree(I.d_rowptr); cudaFree(I.d_colidx);
   I.d_e2n = I.d_vmap = I.d_pmap = I.d_perm = I.d_e2csr = nullptr;
   I.d_pdesc = nullptr;
   I.d_yblob = I.d_vblob = nullptr;
   cudaFree(I.d_idesc); cudaFree(I.d_mblob);
   I.d_idesc = nullptr; I.d_mblob = nullptr;
   I.d_ystage = I.d_vstage = nullptr;
   I.d_rowptr = I.d_colidx = nullptr;
   for (int a = 0; a < 2; a++) { for (int b = 0; b < 5; b++) { cudaFree(I.d_ifc[a][b]); I.d_ifc[a][b] = nullptr; } }
   I.pdev = PatchDev {};
   I.pdesc.clear();
}

static int setup_integrator(Integrator &I, bool allow_patches = true)
{
   const int dim = I.mesh->dim;
   I.ne = I.mesh->ne;
   I.stride = (I.ne + 31) / 32 * 32;
   // sizes and block offsets
   I.nvd = 0; I.ndof_all = 0; I.ntotal = 0;
   I.goff.clear();
   for (const FieldDesc &f : I.fields)
   {
      const int n = f.space->nd_el() * f.space->vdim;
      I.ndof_all += n;
      if (f.role == ROLE_INPUT)
      {
         I.nvd += n;
         I.goff.push_back(I.ntotal);
         I.ntotal += f.space->vsize();
      }
   }
   I.npd = I.ndof_all - I.nvd;
   if (I.ntotal >= 0x7fffffffL) { set_error("more than 2^31 dofs per rank"); return 1; }

   // Element order.  Patch assembly (madb_patch.cpp) when the rows of a patch fit in shared memory:
   // compact patches of PATCH_PE elements; otherwise global colouring, one launch per colour.
   I.use_patches = allow_patches && I.ops.patch_ok && !I.ops.map_aos && !getenv("MADB_NO_PATCH");
   if (I.use_patches)
   {
      I.pe = I.ops.patch_pe > 0 ? I.ops.patch_pe : patch_pe(I.nvd);
      patch_order(I);
      if (I.pdesc.empty()) { I.use_patches = false; }
   }
   if (I.use_patches)
   {
      I.stride = (int)I.pdesc.size() * I.pe;
      I.color_off = {0, I.ne};
   }
   else
   {
      std::vector<int> color;
      int ncolors = 0;
      color_elements(I, color, ncolors);
      I.color_off.assign(ncolors + 1, 0);
      for (int e = 0; e < I.ne; e++) { I.color_off[color[e] + 1]++; }
      for (int c = 0; c < ncolors; c++) { I.color_off[c + 1] += I.color_off[c]; }
      I.perm.resize(I.ne);
      std::vector<int> fill(I.color_off.begin(), I.color_off.end() - 1);
      for (int e = 0; e < I.ne; e++) { I.perm[fill[color[e]]++] = e; }
   }

   // device maps in sorted order
   const int ngn = I.mesh->ngn();
   std::vector<int> e2n((size_t)ngn * I.stride, 0), vmap((size_t)I.nvd * I.stride, 0),
       pmap((size_t)std::max(I.npd, 1) * I.stride, 0);
   std::vector<unsigned char> touched(I.ntotal, 0);
   std::vector<int> vd;
   const bool aos = I.ops.map_aos != 0;
   for (int t = 0; t < I.ne; t++)
   {
      const int e = I.perm[t];
      for (int k = 0; k < ngn; k++) { e2n[aos ? (size_t)t * ngn + k : (size_t)k * I.stride + t] = I.mesh->e2n[(size_t)e * ngn + k]; }
      build_vdofs(I, e, vd);
      // a dof that appears twice in one element (periodic meshes one or two elements wide, duplicated dof maps) would lose
      // one contribution on the scatter paths (all loads of an element precede its stores): refuse it loudly
      {
         std::vector<int> sv(vd.begin(), vd.begin() + I.nvd);
         std::sort(sv.begin(), sv.end());
         if (std::adjacent_find(sv.begin(), sv.end()) != sv.end())
         {
            set_error("element " + std::to_string(e) + " lists the same dof twice in its element->dof map: not supported");
            return 1;
         }
      }
      for (int i = 0; i < I.nvd; i++)
      {
         int m = vd[i];
         if (!touched[m]) { touched[m] = 1; m |= 0x80000000; } // colours ascend with t
         vmap[aos ? (size_t)t * I.nvd + i : (size_t)i * I.stride + t] = m;
      }
      int k = 0;
      for (const FieldDesc &f : I.fields)
      {
         if (f.role == ROLE_INPUT) { continue; }
         const Space &S = *f.space;
         const int nd = S.nd_el();
         for (int c = 0; c < S.vdim; c++)
         {
            for (int i = 0; i < nd; i++)
            {
               const int d = S.e2l[(size_t)e * nd + i];
               pmap[aos ? (size_t)t * I.npd + k : (size_t)k * I.stride + t] = (S.ordering == ORD_BYNODES) ? d + S.ndofs * c : d * S.vdim + c;
               k++;
            }
         }
      }
   }
   // rows no element touches (dofs outside every element of this rank) are never written by the first-touch scatter:
   // the residual is zeroed before every assembly when such rows exist
   I.untouched_rows = 0;
   for (long m = 0; m < I.ntotal; m++) { I.untouched_rows += touched[m] ? 0 : 1; }
   if (upload(e2n, &I.d_e2n) || upload(vmap, &I.d_vmap) || upload(pmap, &I.d_pmap) || upload(I.perm, &I.d_perm)) { return 2; }
   static const bool want_xe = getenv("MADB_XE") && atoi(getenv("MADB_XE")) != 0; // only for builds with -DMADB_SF2D_XE=1
   if (want_xe && I.mesh->dim == 2 && !aos)
   {
      // vertex coordinates per element, SoA [4][stride] of (x, y): the sum-factorised 2-D path reads them directly
      // (one coalesced 16-byte load per vertex) instead of through the element->vertex map (two dependent loads)
      std::vector<double> xe((size_t)8 * I.stride, 0.0);
      for (int t = 0; t < I.ne; t++)
      {
         for (int k = 0; k < 4; k++)
         {
            const int n = I.mesh->e2n[(size_t)I.perm[t] * 4 + k];
            xe[((size_t)k * I.stride + t) * 2] = I.mesh->coords[(size_t)n * 2];
            xe[((size_t)k * I.stride + t) * 2 + 1] = I.mesh->coords[(size_t)n * 2 + 1];
         }
      }
      if (upload(xe, &I.d_xe)) { return 2; }
   }
   if (I.use_patches)
   {
      PatchHost H;
      if (!patch_build_y(I, H)) { set_error("patch assembly: a dof has more than 8 contributing elements in one patch (set MADB_NO_PATCH=1)"); return 1; }
      if (upload(H.blob, &I.d_yblob) || upload(I.pdesc, &I.d_pdesc) || upload_ifc(H, I.d_ifc[0], I.pdev.ylist)) { return 2; }
      CUDA_OK(cudaMalloc((void **)&I.d_ystage, std::max<size_t>(H.stage_size, 1) * sizeof(double)));
      PatchDev &P = I.pdev;
      P.npatch = (int)I.pdesc.size();
      P.max_yblob = I.max_yblob;
      P.max_yg = I.max_yg;
      P.max_yf = I.max_yf;
      P.max_vblob = P.max_vg = P.max_vf = 0;
      P.desc = I.d_pdesc;
      P.yblob = I.d_yblob;
      P.ystage = I.d_ystage;
      P.ny_ifc = (int)(H.dst.size() + H.dst4.size());
   }

   if (I.mesh->simplex)
   {
      // triangles: MFEM's rule of the requested order, nodal P1 / P2 / P0 tables, affine geometry (constant dN/dxi)
      std::vector<double> pts, wt;
      if (!triangle_rule(I.quad_order, pts, wt)) { set_error("triangle rules of order 0 - 6 are available, not " + std::to_string(I.quad_order)); return 1; }
      I.nq = (int)wt.size();
      int ntab = 0;
      for (const FieldDesc &f : I.fields) { ntab += f.space->nd_el(); }
      I.phi.assign((size_t)I.nq * ntab, 0.0);
      I.dphi.assign((size_t)I.nq * ntab * 2, 0.0);
      I.gdphi.assign((size_t)I.nq * 3 * 2, 0.0);
      I.w = wt;
      I.tri_pts = pts;
      int toff = 0;
      I.b1d.assign(I.fields.size(), std::vector<double>(1, 0.0));
      I.g1d.assign(I.fields.size(), std::vector<double>(1, 0.0));
      I.xq1d.assign(1, 0.0);
      I.w1d.assign(1, 0.0);
      for (const FieldDesc &f : I.fields)
      {
         const int nd = f.space->nd_el();
         std::vector<double> ph(nd), dp(2 * nd);
         for (int q = 0; q < I.nq; q++)
         {
            triangle_shapes(f.space->basis, f.space->order, pts[2 * q], pts[2 * q + 1], ph.data(), dp.data());
            for (int i = 0; i < nd; i++)
            {
               I.phi[(size_t)q * ntab + toff + i] = ph[i];
               I.dphi[((size_t)q * ntab + toff + i) * 2] = dp[2 * i];
               I.dphi[((size_t)q * ntab + toff + i) * 2 + 1] = dp[2 * i + 1];
            }
         }
         toff += nd;
      }
      for (int q = 0; q < I.nq; q++)
      {
         double ph[3], dp[6];
         triangle_shapes(BASIS_H1, 1, pts[2 * q], pts[2 * q + 1], ph, dp);
         for (int k = 0; k < 6; k++) { I.gdphi[(size_t)q * 6 + k] = dp[k]; }
      }
   }
   else
   {
      // basis tables at the quadrature points: [q][toff_f + i], x fastest in q and i
      std::vector<double> xq, wq;
      gauss_legendre_01(I.nq1d, xq, wq);
      I.nq = 1;
      for (int d = 0; d < dim; d++) { I.nq *= I.nq1d; }
      int ntab = 0;
      for (const FieldDesc &f : I.fields) { ntab += f.space->nd_el(); }
      I.phi.assign((size_t)I.nq * ntab, 0.0);
      I.dphi.assign((size_t)I.nq * ntab * dim, 0.0);
      I.gdphi.assign((size_t)I.nq * ngn * dim, 0.0);
      I.w.assign(I.nq, 0.0);
      auto fill_tables = [&](const std::vector<double> &nodes, int toff, int ncols, double *phi, double *dphi)
      {
         const int nn = (int)nodes.size();
         std::vector<double> B, G;
         lagrange_tables(nodes, xq, B, G);
         int nd = 1;
         for (int d = 0; d < dim; d++) { nd *= nn; }
         for (int q = 0; q < I.nq; q++)
         {
            int qd[3], r = q;
            for (int d = 0; d < dim; d++) { qd[d] = r % I.nq1d; r /= I.nq1d; }
            for (int i = 0; i < nd; i++)
            {
               int id[3], s = i;
               for (int d = 0; d < dim; d++) { id[d] = s % nn; s /= nn; }
               double v = 1.0;
               for (int d = 0; d < dim; d++) { v *= B[(size_t)qd[d] * nn + id[d]]; }
               if (phi) { phi[(size_t)q * ncols + toff + i] = v; }
               for (int k = 0; k < dim; k++)
               {
                  double g = 1.0;
                  for (int d = 0; d < dim; d++) { g *= (d == k) ? G[(size_t)qd[d] * nn + id[d]] : B[(size_t)qd[d] * nn + id[d]]; }
                  dphi[((size_t)q * ncols + toff + i) * dim + k] = g;
               }
            }
         }
      };
      int toff = 0;
      I.b1d.clear(); I.g1d.clear();
      I.xq1d = xq; I.w1d = wq;
      for (const FieldDesc &f : I.fields)
      {
         std::vector<double> nodes, wtmp;
         if (f.space->basis == BASIS_H1) { gauss_lobatto_01(f.space->order + 1, nodes); }
         else { gauss_legendre_01(f.space->order + 1, nodes, wtmp); }
         fill_tables(nodes, toff, ntab, I.phi.data(), I.dphi.data());
         toff += f.space->nd_el();
         std::vector<double> B, Gd;
         lagrange_tables(nodes, xq, B, Gd);
         I.b1d.push_back(B);
         I.g1d.push_back(Gd);
      }
      fill_tables(std::vector<double> {0.0, 1.0}, 0, ngn, nullptr, I.gdphi.data());
      for (int q = 0; q < I.nq; q++)
      {
         int r = q;
         double ww = 1.0;
         for (int d = 0; d < dim; d++) { ww *= wq[r % I.nq1d]; r /= I.nq1d; }
         I.w[q] = ww;
      }
   }
   I.pdata.assign(I.fields.size(), nullptr);
   I.d_pstage.assign(I.fields.size(), nullptr);
   if (I.use_patches && I.pe != PATCH_PE)
   {
      // Large element matrices (64-element patches): the staged matrices plus the gather maps must fit in the
      // shared memory of one CTA.  The size of the maps depends on the dof numbering (irregular chunks carry
      // explicit CSR positions), so it is only known once they are built: build them now, and fall back to the
      // colour-scatter path when they do not fit.
      const int rc = ensure_pattern_device(I);
      if (rc) { return rc; }
      const long ld = I.pe + 1;
      int ntab = 0;
      for (const FieldDesc &f : I.fields) { ntab += f.space->nd_el(); }
      const long tables = ((long)I.nq * ntab * (1 + dim) + (long)I.nq * ngn * dim + I.nq) * 8;
      const long need = patch_al16((int)(I.nvd * ld * 8)) + patch_al16((int)((long)I.nvd * (I.nvd + 1) / 2 * ld * 8)) +
                        I.max_yblob + I.max_vblob + tables + 64;
      if (need > 220 * 1024)
      {
         release_device_maps(I);
         I.have_pattern = false;
         I.have_patch_vals = false;
         return setup_integrator(I, false);
      }
   }
   return 0;
}

static int ensure_pattern_device(Integrator &I)
{
   if (I.d_e2csr || I.d_vblob || I.d_idesc) { return 0; }
   build_pattern(I);
   if (!I.have_pattern) { return 1; }
   // CSR-image kernel (k_patch_img): measured slower than the warp-specialised gather-map kernel on configs 2 and 4
   // (profiles/r02_img_kernel.md), so it is opt-in: MADB_PATCH_IMG=1 for the one-thread-per-element variant (element
   // matrices of up to 8 dofs), -DMADB_PAIR_KERNEL=1 at build time for the thread-pair variant (64-element patches)
   static const bool img_on = getenv("MADB_PATCH_IMG") && atoi(getenv("MADB_PATCH_IMG")) == 1;
   if (I.use_patches && (I.ops.img_tpe == 2 || (I.ops.img_tpe == 1 && img_on)))
   {
      PatchHost H;
      ImgHost IH;
      if (!patch_build_img(I, H, IH)) { return 1; }
      if (upload(IH.desc, &I.d_idesc) || upload(IH.mblob, &I.d_mblob) ||
          upload_ifc(H, I.d_ifc[1], I.pdev.vlist) || upload(I.rowptr, &I.d_rowptr) || upload(I.colidx, &I.d_colidx))
      {
         return 2;
      }
      CUDA_OK(cudaMalloc((void **)&I.d_vstage, std::max<size_t>(H.stage_size, 2) * sizeof(double)));
      PatchDev &P = I.pdev;
      P.vstage = I.d_vstage;
      P.nv_ifc = (int)(H.dst.size() + H.dst4.size());
      P.img.desc = I.d_idesc;
      P.img.mblob = I.d_mblob;
      P.img.max_vslots = IH.max_vslots;
      P.img.max_yslots = IH.max_yslots;
      P.img.max_mblob = IH.max_mblob;
      P.img.nev = IH.nev;
      P.img.ney = IH.ney;
      for (size_t p = 0; p < I.pdesc.size(); p++)
      {
         // statistics (madb_integrator_patch_stats)
         I.pdesc[p].nslots = IH.desc[p].sh0 + IH.desc[p].nsh;
         I.pdesc[p].nexc = IH.desc[p].sh0;
         I.pdesc[p].nruns = IH.desc[p].nruns;
      }
      return 0;
   }
   if (I.use_patches)
   {
      PatchHost H;
      if (!patch_build_v(I, H)) { return 1; }
      if (upload(H.blob, &I.d_vblob) || upload_ifc(H, I.d_ifc[1], I.pdev.vlist) || upload(I.rowptr, &I.d_rowptr) ||
          upload(I.colidx, &I.d_colidx))
      {
         return 2;
      }
      CUDA_OK(cudaMalloc((void **)&I.d_vstage, std::max<size_t>(H.stage_size, 1) * sizeof(double)));
      CUDA_OK(cudaMemcpy(I.d_pdesc, I.pdesc.data(), I.pdesc.size() * sizeof(PatchDesc), cudaMemcpyHostToDevice));
      PatchDev &P = I.pdev;
      P.max_vblob = I.max_vblob;
      P.max_vg = I.max_vg;
      P.max_vf = I.max_vf;
      P.vblob = I.d_vblob;
      P.vstage = I.d_vstage;
      P.nv_ifc = (int)(H.dst.size() + H.dst4.size());
      return 0;
   }
   std::vector<int> e2csr;
   build_e2csr(I, std::vector<int>(), e2csr);
   if (upload(e2csr, &I.d_e2csr) || upload(I.rowptr, &I.d_rowptr) || upload(I.colidx, &I.d_colidx)) { return 2; }
   return 0;
}

// stage a caller vector: returns the device pointer to use
static int stage_in(Integrator &I, const double *p, size_t n, double **buf, const double **out)
{
   if (is_device_ptr(p)) { *out = p; return 0; }
   if (!*buf) { CUDA_OK(cudaMalloc((void **)buf, std::max<size_t>(n, 1) * sizeof(double))); }
   CUDA_OK(cudaMemcpyAsync(*buf, p, n * sizeof(double), cudaMemcpyHostToDevice, I.ctx->stream));
   *out = *buf;
   return 0;
}
static int stage_out_begin(Integrator &I, double *p, size_t n, double **buf, double **out)
{
   if (is_device_ptr(p)) { *out = p; return 0; }
   if (!*buf) { CUDA_OK(cudaMalloc((void **)buf, std::max<size_t>(n, 1) * sizeof(double))); }
   *out = *buf;
   return 0;
}

static int run(Integrator &I, int mode, const double *x, const double *v, double *y, double *vals, double *energy,
               double *cvalue = nullptr, double *cgrad = nullptr, double *chess = nullptr, int coef_variant = 0, int defer_v = 0)
{
   CUDA_OK(cudaSetDevice(I.ctx->device));
   const size_t N = (size_t)I.ntotal;
   LaunchCtx L;
   std::memset(&L, 0, sizeof(L));
   L.stream = I.ctx->stream;
   L.ne = I.ne; L.stride = I.stride;
   L.ncolors = (int)I.color_off.size() - 1;
   L.color_off = I.color_off.data();
   L.e2n = I.d_e2n; L.vmap = I.d_vmap; L.pmap = I.d_pmap;
   L.xe = I.d_xe;
   L.coords = I.mesh->d_coords;
   for (size_t f = 0; f < I.fields.size(); f++)
   {
      L.pdata[f] = I.pdata[f];
      if (I.fields[f].role == ROLE_PARAM && !I.pdata[f]) { set_error("parameter field " + std::to_string(f) + " was never set (madb_integrator_set_param_field)"); return 1; }
   }
   if (I.ops.n_qprm - I.ops.n_field_qprm > 0)
   {
      if (!I.d_qf || I.qf_count != (size_t)(I.ops.n_qprm - I.ops.n_field_qprm)) { set_error("quadrature-function parameters were never set (madb_integrator_set_param_qf)"); return 1; }
      L.qf = I.d_qf;
   }
   std::vector<double> fp;
   I.fn->flat_params(fp);
   if ((int)fp.size() != I.ops.n_fparam)
   {
      set_error("functional '" + I.fn->key() + "' carries " + std::to_string(fp.size()) + " parameters, the compiled kernel expects " + std::to_string(I.ops.n_fparam));
      return 1;
   }
   fp.push_back(0.0);
   L.fparams = fp.data();
   L.phi = I.phi.data(); L.dphi = I.dphi.data(); L.gdphi = I.gdphi.data(); L.w = I.w.data();
   for (size_t f = 0; f < I.fields.size(); f++) { L.b1d[f] = I.b1d[f].data(); L.g1d[f] = I.g1d[f].data(); }
   L.xq1d = I.xq1d.data(); L.w1d = I.w1d.data();
   L.patch = I.use_patches ? &I.pdev : nullptr;
   L.ev0 = I.timing ? I.ev0 : nullptr;
   L.ev1 = I.timing ? I.ev1 : nullptr;

   if (stage_in(I, x, N, &I.d_x, &L.x)) { return 2; }
   const double *v_orig = nullptr; // device copy of the caller's direction
   if (mode == MODE_ACT)
   {
      if (stage_in(I, v, N, &I.d_v, &L.v)) { return 2; }
      v_orig = L.v;
      if (I.ness > 0)
      {
         // action of the eliminated Jacobian (EliminateRowCol): columns of essential dofs drop out
         if (!I.d_v2) { CUDA_OK(cudaMalloc((void **)&I.d_v2, std::max<size_t>(N, 1) * sizeof(double))); }
         CUDA_OK(cudaMemcpyAsync(I.d_v2, L.v, N * sizeof(double), cudaMemcpyDeviceToDevice, I.ctx->stream));
         k_ess_zero<<<(I.ness + 127) / 128, 128, 0, I.ctx->stream>>>(I.d_ess, I.ness, I.d_v2);
         L.v = I.d_v2;
      }
   }
   double *dy = nullptr, *dvals = nullptr;
   if (y) { if (stage_out_begin(I, y, N, &I.d_y, &dy)) { return 2; } }
   if (vals)
   {
      if (I.ops.matrix_free_only) { set_error("configuration '" + I.key + "' is matrix-free only: use madb_integrator_grad_mult"); return 1; }
      if (ensure_pattern_device(I)) { return 1; }
      if (stage_out_begin(I, vals, I.colidx.size(), &I.d_vals, &dvals)) { return 2; }
      L.e2csr = I.d_e2csr;
   }
   L.y = dy; L.vals = dvals;
   L.write_y = dy != nullptr; L.write_vals = dvals != nullptr;
   if (mode == MODE_ENERGY)
   {
      if (!I.d_energy)
      {
         CUDA_OK(cudaMalloc((void **)&I.d_energy, (size_t)I.stride * sizeof(double)));
         CUDA_OK(cudaMalloc((void **)&I.d_esum, sizeof(double)));
      }
      L.energy = I.d_energy;
   }
   double *dcv = nullptr, *dcg = nullptr, *dch = nullptr;
   const size_t npts = (size_t)I.ne * I.nq;
   if (mode == MODE_COEF)
   {
      L.perm = I.d_perm;
      if (cvalue) { if (stage_out_begin(I, cvalue, npts, &I.d_cvalue, &dcv)) { return 2; } }
      if (cgrad) { if (stage_out_begin(I, cgrad, npts * I.ops.n_input, &I.d_cgrad, &dcg)) { return 2; } }
      if (chess) { if (stage_out_begin(I, chess, npts * I.ops.n_input * I.ops.n_input, &I.d_chess, &dch)) { return 2; } }
      L.cvalue = dcv; L.cgrad = dcg; L.chess = dch;
      L.coef_variant = coef_variant;
   }
   L.defer_v_ifc = (defer_v && L.patch) ? 1 : 0;
   if (I.untouched_rows > 0 && dy && mode != MODE_ENERGY && mode != MODE_COEF) { CUDA_OK(cudaMemsetAsync(dy, 0, N * sizeof(double), L.stream)); }
   const int rc = I.ops.launch(L, mode);
   if (rc == MADB_RC_MIRROR)
   {
      set_error("sum-factorised 2-D path: the 1-D basis / quadrature tables are not mirror-symmetric about 1/2");
      return 2;
   }
   if (rc != 0) { set_error(std::string("kernel launch failed: ") + cudaGetErrorString((cudaError_t)rc)); return 2; }

   if (mode == MODE_ENERGY)
   {
      k_reduce_sum<<<1, 1024, 0, L.stream>>>(I.d_energy, I.ne, I.d_esum);
      CUDA_OK(cudaMemcpyAsync(energy, I.d_esum, sizeof(double), cudaMemcpyDeviceToHost, L.stream));
      CUDA_OK(cudaStreamSynchronize(L.stream));
      return 0;
   }
   if (mode == MODE_COEF)
   {
      if (cvalue && dcv != cvalue) { CUDA_OK(cudaMemcpyAsync(cvalue, dcv, npts * sizeof(double), cudaMemcpyDeviceToHost, L.stream)); }
      if (cgrad && dcg != cgrad) { CUDA_OK(cudaMemcpyAsync(cgrad, dcg, npts * I.ops.n_input * sizeof(double), cudaMemcpyDeviceToHost, L.stream)); }
      if (chess && dch != chess) { CUDA_OK(cudaMemcpyAsync(chess, dch, npts * I.ops.n_input * I.ops.n_input * sizeof(double), cudaMemcpyDeviceToHost, L.stream)); }
      CUDA_OK(cudaStreamSynchronize(L.stream));
      return 0;
   }
   if (I.ness > 0)
   {
      const int g = (I.ness + 127) / 128;
      if (dy && mode == MODE_ACT) { k_ess_copy<<<g, 128, 0, L.stream>>>(I.d_ess, I.ness, v_orig, dy); }
      else if (dy) { k_ess_zero<<<g, 128, 0, L.stream>>>(I.d_ess, I.ness, dy); }
      if (dvals && !L.defer_v_ifc) { k_ess_rowcol<<<g, 128, 0, L.stream>>>(I.d_ess, I.ness, I.d_rowptr, I.d_colidx, dvals); }
   }
   if (L.defer_v_ifc) { I.pending_vals = dvals; }
   bool sync = false;
   if (y && dy != y) { CUDA_OK(cudaMemcpyAsync(y, dy, N * sizeof(double), cudaMemcpyDeviceToHost, L.stream)); sync = true; }
   if (vals && dvals != vals) { CUDA_OK(cudaMemcpyAsync(vals, dvals, I.colidx.size() * sizeof(double), cudaMemcpyDeviceToHost, L.stream)); sync = true; }
   if (sync) { CUDA_OK(cudaStreamSynchronize(L.stream)); }
   CUDA_OK(cudaGetLastError());
   return 0;
}

} // namespace madb

using namespace madb;

struct madb_ctx : Ctx {};
struct madb_mesh : Mesh {};
struct madb_space : Space {};
struct madb_functional : Functional {};
struct madb_integrator : Integrator {};

extern "C"
{

   int madb_version(void) { return 101; }
   int madb_registry_has(const char *key) { return (key && registry().count(key)) ? 1 : 0; }
   const char *madb_last_error(void) { return last_error(); }

   int madb_ctx_create(int device, madb_ctx **out)
   {
      int ndev = 0;
      if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      {
         cudaGetLastError();
         set_error("no CUDA device: the mfem-ad B200 path has no CPU fallback");
         return 2;
      }
      CUDA_OK(cudaSetDevice(device));
      madb_ctx *c = new madb_ctx;
      c->device = device;
      CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
      *out = c;
      return 0;
   }
   int madb_ctx_destroy(madb_ctx *c)
   {
      if (!c) { return 0; }
      cudaStreamDestroy(c->stream);
      cudaFree(c->scratch);
      cudaFreeHost(c->scratch_host);
      delete c;
      return 0;
   }
   int madb_ctx_sync(madb_ctx *c)
   {
      CUDA_OK(cudaStreamSynchronize(c->stream));
      return 0;
   }
   void *madb_ctx_stream(madb_ctx *c) { return (void *)c->stream; }
   int madb_device_alloc(madb_ctx *c, size_t bytes, void **ptr)
   {
      CUDA_OK(cudaSetDevice(c->device));
      CUDA_OK(cudaMalloc(ptr, bytes > 0 ? bytes : 1));
      return 0;
   }
   int madb_device_free(madb_ctx *c, void *ptr)
   {
      CUDA_OK(cudaSetDevice(c->device));
      CUDA_OK(cudaFree(ptr));
      return 0;
   }

   int madb_mesh_create(madb_ctx *ctx, int dim, int ne, const int32_t *e2n, int nnodes, const double *coords,
                        madb_mesh **out)
   {
      if (dim < 1 || dim > 3 || ne <= 0 || !e2n || !coords) { set_error("madb_mesh_create: bad arguments"); return 1; }
      CUDA_OK(cudaSetDevice(ctx->device));
      madb_mesh *m = new madb_mesh;
      m->ctx = ctx; m->dim = dim; m->ne = ne; m->geom_order = 1; m->nnodes = nnodes;
      m->e2n.assign(e2n, e2n + (size_t)ne * (1 << dim));
      m->coords.assign(coords, coords + (size_t)nnodes * dim);
      for (int v : m->e2n) { if (v < 0 || v >= nnodes) { delete m; set_error("madb_mesh_create: vertex id out of range"); return 1; } }
      if (upload(m->coords, &m->d_coords)) { delete m; return 2; }
      *out = m;
      return 0;
   }
   int madb_mesh_create_simplex(madb_ctx *ctx, int dim, int ne, const int32_t *e2n, int nnodes, const double *coords,
                                madb_mesh **out)
   {
      if (dim != 2 || ne <= 0 || !e2n || !coords) { set_error("madb_mesh_create_simplex: triangles (dim = 2) only"); return 1; }
      CUDA_OK(cudaSetDevice(ctx->device));
      madb_mesh *m = new madb_mesh;
      m->ctx = ctx; m->dim = dim; m->ne = ne; m->geom_order = 1; m->nnodes = nnodes; m->simplex = true;
      m->e2n.assign(e2n, e2n + (size_t)ne * (dim + 1));
      m->coords.assign(coords, coords + (size_t)nnodes * dim);
      for (int v : m->e2n) { if (v < 0 || v >= nnodes) { delete m; set_error("madb_mesh_create_simplex: vertex id out of range"); return 1; } }
      if (upload(m->coords, &m->d_coords)) { delete m; return 2; }
      *out = m;
      return 0;
   }
   int madb_mesh_destroy(madb_mesh *m)
   {
      if (m) { cudaFree(m->d_coords); delete m; }
      return 0;
   }

   int madb_space_create(madb_ctx *ctx, madb_mesh *mesh, int basis, int order, int vdim, int ordering, int ndofs,
                         const int32_t *e2l, madb_space **out)
   {
      if (!mesh || order < 0 || vdim < 1 || ndofs <= 0 || !e2l || (basis != BASIS_H1 && basis != BASIS_L2) || (basis == BASIS_H1 && order < 1))
      {
         set_error("madb_space_create: bad arguments");
         return 1;
      }
      if (mesh->simplex && triangle_ndof(basis, order) == 0)
      {
         set_error("madb_space_create: on triangles H1 orders 1, 2 and L2 order 0 are available");
         return 1;
      }
      madb_space *s = new madb_space;
      s->ctx = ctx; s->mesh = mesh; s->basis = basis; s->order = order; s->vdim = vdim; s->ordering = ordering; s->ndofs = ndofs;
      s->e2l.assign(e2l, e2l + (size_t)mesh->ne * s->nd_el());
      for (int v : s->e2l) { if (v < 0 || v >= ndofs) { delete s; set_error("madb_space_create: dof id out of range"); return 1; } }
      *out = s;
      return 0;
   }
   int madb_space_destroy(madb_space *s) { delete s; return 0; }

   int madb_functional_create(madb_ctx *, const char *kind, int nparams, const double *params, int niparams,
                              const int *iparams, int nchildren, madb_functional *const *children, madb_functional **out)
   {
      madb_functional *f = new madb_functional;
      f->kind = kind;
      if (nparams > 0) { f->params.assign(params, params + nparams); }
      if (niparams > 0) { f->iparams.assign(iparams, iparams + niparams); }
      for (int i = 0; i < nchildren; i++) { f->children.push_back(children[i]); }
      *out = f;
      return 0;
   }
   int madb_functional_set_params(madb_functional *f, int nparams, const double *params)
   {
      if ((size_t)nparams != f->params.size())
      {
         // same rule as Evaluator::Replace (src/ad_native.cpp:109-118): sizes must match
         set_error("madb_functional_set_params: size mismatch: expected " + std::to_string(f->params.size()) + ", got " + std::to_string(nparams));
         return 1;
      }
      f->params.assign(params, params + nparams);
      return 0;
   }
   int madb_functional_destroy(madb_functional *f) { delete f; return 0; }

   // scratch device buffer for a host-side argument (slow path; device pointers are used in place)
   struct Staged
   {
      double *d = nullptr;
      double *host = nullptr;
      size_t n = 0;
      bool owned = false;
      int in(const double *p, size_t cnt)
      {
         n = cnt;
         if (!p) { return 0; }
         if (is_device_ptr(p)) { d = const_cast<double *>(p); return 0; }
         owned = true;
         if (cudaMalloc((void **)&d, std::max<size_t>(cnt, 1) * sizeof(double)) != cudaSuccess) { return 2; }
         return cudaMemcpy(d, p, cnt * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess ? 0 : 2;
      }
      int out(double *p, size_t cnt, bool copy_in = false)
      {
         host = nullptr;
         const int rc = in(p, cnt);
         if (rc == 0 && owned) { host = p; }
         (void)copy_in;
         return rc;
      }
      int finish()
      {
         int rc = 0;
         if (owned && host) { rc = cudaMemcpy(host, d, n * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 2; }
         if (owned) { cudaFree(d); }
         return rc;
      }
   };

   int madb_functional_eval(madb_ctx *ctx, madb_functional *f, int n_input, int npts, const double *x,
                            const double *qprm, double *value, double *grad, double *hess)
   {
      CUDA_OK(cudaSetDevice(ctx->device));
      const std::string key = f->key() + "|n" + std::to_string(n_input);
      auto it = eval_registry().find(key);
      if (it == eval_registry().end())
      {
         set_error("no pointwise AD kernel compiled for '" + key + "' (add MADB_EVAL_INSTANCE)");
         return 1;
      }
      const EvalOps &E = it->second;
      std::vector<double> fp;
      f->flat_params(fp);
      if ((int)fp.size() != E.n_fparam) { set_error("functional '" + key + "': wrong number of parameters"); return 1; }
      fp.push_back(0.0);
      if (E.n_qprm > 0 && !qprm) { set_error("functional '" + key + "' needs per-point parameters"); return 1; }
      const size_t n = n_input, P = npts;
      Staged sx, sq, sv, sg, sh;
      if (sx.in(x, P * n) || sq.in(E.n_qprm > 0 ? qprm : nullptr, P * E.n_qprm) || sv.out(value, P) ||
          sg.out(grad, P * n) || sh.out(hess, P * n * n))
      {
         set_error("madb_functional_eval: device staging failed");
         return 2;
      }
      const int rc = E.launch(ctx->stream, npts, fp.data(), sx.d, sq.d, sv.d, sg.d, sh.d);
      if (rc) { set_error(std::string("eval kernel: ") + cudaGetErrorString((cudaError_t)rc)); return 2; }
      if (sx.owned || sv.owned || sg.owned || sh.owned) { CUDA_OK(cudaStreamSynchronize(ctx->stream)); }
      sx.finish(); sq.finish();
      if (sv.finish() || sg.finish() || sh.finish()) { set_error("madb_functional_eval: copy back failed"); return 2; }
      return 0;
   }

   int madb_vecfunction_eval(madb_ctx *ctx, madb_functional *f, int n_input, int n_output, int npts, const double *x,
                             double *value, double *jac, double *hess)
   {
      CUDA_OK(cudaSetDevice(ctx->device));
      const std::string key = f->key() + "|n" + std::to_string(n_input) + "m" + std::to_string(n_output);
      auto it = vec_eval_registry().find(key);
      if (it == vec_eval_registry().end())
      {
         set_error("no pointwise AD kernel compiled for the vector function '" + key + "' (add MADB_VEC_EVAL_INSTANCE)");
         return 1;
      }
      const VecEvalOps &E = it->second;
      std::vector<double> fp;
      f->flat_params(fp);
      if ((int)fp.size() != E.n_fparam) { set_error("vector function '" + key + "': wrong number of parameters"); return 1; }
      fp.push_back(0.0);
      const size_t n = n_input, m = n_output, P = npts;
      Staged sx, sv, sj, sh;
      if (sx.in(x, P * n) || sv.out(value, P * m) || sj.out(jac, P * m * n) || sh.out(hess, P * m * n * n))
      {
         set_error("madb_vecfunction_eval: device staging failed");
         return 2;
      }
      const int rc = E.launch(ctx->stream, npts, fp.data(), sx.d, sv.d, sj.d, sh.d);
      if (rc) { set_error(std::string("vector eval kernel: ") + cudaGetErrorString((cudaError_t)rc)); return 2; }
      if (sx.owned || sv.owned || sj.owned || sh.owned) { CUDA_OK(cudaStreamSynchronize(ctx->stream)); }
      sx.finish();
      if (sv.finish() || sj.finish() || sh.finish()) { set_error("madb_vecfunction_eval: copy back failed"); return 2; }
      return 0;
   }

   int madb_dofpg_nodal(madb_ctx *ctx, madb_functional *entropy, int n, double alpha, const double *u,
                        const double *psi, const double *psik, const double *w, double *r_u, double *r_psi,
                        double *d_pp, double *d_up)
   {
      CUDA_OK(cudaSetDevice(ctx->device));
      if (!u || !psi || !psik || !w) { set_error("madb_dofpg_nodal: u, psi, psik and the nodal weights w must be given (zero weights reproduce the reference, SURVEY H7)"); return 1; }
      const std::string key = entropy->key() + "|n1";
      auto it = eval_registry().find(key);
      if (it == eval_registry().end() || !it->second.dofpg)
      {
         set_error("no nodal PG kernel compiled for entropy '" + key + "' (scalar entropies; add MADB_EVAL_INSTANCE)");
         return 1;
      }
      std::vector<double> fp;
      entropy->flat_params(fp);
      if ((int)fp.size() != it->second.n_fparam) { set_error("entropy '" + key + "': wrong number of parameters"); return 1; }
      fp.push_back(0.0);
      if (!u || !psi || !psik || !w) { set_error("madb_dofpg_nodal: u, psi, psik and the nodal weights w are required (zeros for w reproduce the reference, SURVEY H7)"); return 1; }
      Staged su, sp, sk, sw, ru, rp, dp, du;
      if (su.in(u, n) || sp.in(psi, n) || sk.in(psik, n) || sw.in(w, n) || ru.out(r_u, n) || rp.out(r_psi, n) ||
          dp.out(d_pp, n) || du.out(d_up, n))
      {
         set_error("madb_dofpg_nodal: device staging failed");
         return 2;
      }
      if (ru.owned && r_u) { CUDA_OK(cudaMemcpy(ru.d, r_u, (size_t)n * sizeof(double), cudaMemcpyHostToDevice)); } // r_u is accumulated
      const int rc = it->second.dofpg(ctx->stream, n, fp.data(), alpha, su.d, sp.d, sk.d, sw.d, ru.d, rp.d, dp.d, du.d);
      if (rc) { set_error(std::string("dofpg kernel: ") + cudaGetErrorString((cudaError_t)rc)); return 2; }
      if (ru.owned || rp.owned || dp.owned || du.owned) { CUDA_OK(cudaStreamSynchronize(ctx->stream)); }
      su.finish(); sp.finish(); sk.finish(); sw.finish();
      if (ru.finish() || rp.finish() || dp.finish() || du.finish()) { set_error("madb_dofpg_nodal: copy back failed"); return 2; }
      return 0;
   }

   // Shared-dof exchange kernels (ParGridFunction / ParNonlinearForm P and P^T [MFEM-upstream]):
   // all pointers are DEVICE pointers; idx lists are unique within one call, so no two threads collide.
   int madb_pack(madb_ctx *ctx, int n, const int32_t *idx, const double *src, double *dst)
   {
      if (n <= 0) { return 0; }
      CUDA_OK(cudaSetDevice(ctx->device));
      k_pack<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, idx, src, dst);
      CUDA_OK(cudaGetLastError());
      return 0;
   }
   int madb_unpack(madb_ctx *ctx, int n, const int32_t *idx, const double *src, double *dst, int add)
   {
      if (n <= 0) { return 0; }
      CUDA_OK(cudaSetDevice(ctx->device));
      k_unpack<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, idx, src, dst, add);
      CUDA_OK(cudaGetLastError());
      return 0;
   }

   int madb_unpack_multi(madb_ctx *ctx, int n, const int32_t *src4, const int32_t *dst_idx, const double *src, double *dst, int add)
   {
      if (n <= 0) { return 0; }
      CUDA_OK(cudaSetDevice(ctx->device));
      k_unpack_multi<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, (const int4 *)src4, dst_idx, src, dst, add);
      CUDA_OK(cudaGetLastError());
      return 0;
   }

   int madb_lvpp_update(madb_ctx *ctx, int n, double alpha, const double *psi, double *psik, double *lambda_prev,
                        const double *w, double *lambda_diff)
   {
      CUDA_OK(cudaSetDevice(ctx->device));
      Staged sp, sk, sl, sw;
      if (sp.in(psi, n) || sk.out(psik, n) || sl.out(lambda_prev, n) || sw.in(w, n)) { set_error("madb_lvpp_update: device staging failed"); return 2; }
      const int nb = std::min(1024, (n + 255) / 256);
      // block partials + the sum live in a per-context scratch buffer (no allocation per call)
      if (!ctx->scratch)
      {
         CUDA_OK(cudaMalloc((void **)&ctx->scratch, 1025 * sizeof(double)));
         CUDA_OK(cudaMallocHost((void **)&ctx->scratch_host, sizeof(double)));
      }
      double *partial = ctx->scratch, *dsum = ctx->scratch + 1024;
      k_lvpp_update<<<nb, 256, 0, ctx->stream>>>(n, alpha, sp.d, sk.d, sl.d, sw.d, partial);
      k_reduce_sum<<<1, 1024, 0, ctx->stream>>>(partial, nb, dsum);
      CUDA_OK(cudaMemcpyAsync(ctx->scratch_host, dsum, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_OK(cudaStreamSynchronize(ctx->stream));
      *lambda_diff = *ctx->scratch_host;
      sp.finish(); sw.finish();
      if (sk.finish() || sl.finish()) { set_error("madb_lvpp_update: copy back failed"); return 2; }
      return 0;
   }

   int madb_integrator_create(madb_ctx *ctx, int nfields, madb_space *const *spaces, const int *modes,
                              const int *roles, madb_functional *f, int quad_order, madb_integrator **out)
   {
      return madb_integrator_create_ex(ctx, nfields, spaces, modes, roles, f, quad_order, 0, out);
   }
   int madb_integrator_create_ex(madb_ctx *ctx, int nfields, madb_space *const *spaces, const int *modes,
                                 const int *roles, madb_functional *f, int quad_order, int flags, madb_integrator **out)
   {
      if (nfields < 1 || nfields > 8 || !spaces || !modes || !f) { set_error("madb_integrator_create: bad arguments"); return 1; }
      CUDA_OK(cudaSetDevice(ctx->device));
      madb_integrator *I = new madb_integrator;
      I->ctx = ctx;
      I->mesh = spaces[0]->mesh;
      I->fn = f;
      int max_order = 0;
      for (int i = 0; i < nfields; i++)
      {
         // isValidADEval (src/_ad_intg.hpp:55-66) + the modes the reference marks "not yet implemented" (:29-34)
         unsigned m = (unsigned)modes[i];
         if (m & (EV_HESSIAN | EV_DIV | EV_CURL | EV_VECFE))
         {
            delete I;
            set_error("madb_integrator_create: ADEval modes DIV/CURL/Hessian/VECFE are not supported by the B200 path (the reference marks them 'not yet implemented', src/_ad_intg.hpp:29-34)");
            return 1;
         }
         // ADEval::QVALUE (src/ad_intg.hpp:127: the shape is the unit vector at ip.index; quadrature-space unknowns,
         // src/tools.hpp:156-177): exactly VALUE on an L2 space whose nodes are the rule's Gauss points (order nq1d - 1:
         // a Lagrange basis evaluated at its own nodes is the identity), checked once the rule is known
         bool qvalue = false;
         if (m & EV_QVALUE)
         {
            if (m & ~(unsigned)(EV_QVALUE | EV_VECTOR))
            {
               delete I;
               set_error("madb_integrator_create: Invalid ADEval mode: QVALUE can only be combined with VECTOR (src/_ad_intg.hpp:55-66)");
               return 1;
            }
            if (spaces[i]->mesh->simplex) { delete I; set_error("madb_integrator_create: QVALUE is available on tensor-product elements only"); return 1; }
            qvalue = true;
            m = (m & ~(unsigned)EV_QVALUE) | EV_VALUE;
         }
         if (!(m & (EV_VALUE | EV_GRAD))) { delete I; set_error("madb_integrator_create: a field needs VALUE and/or GRAD"); return 1; }
         if (spaces[i]->vdim > 1 && !(m & EV_VECTOR) && (!roles || roles[i] == ROLE_INPUT))
         {
            // src/ad_intg.hpp:165-167: vdim must be 1 or the mode must be VECTOR
            delete I;
            set_error("madb_integrator_create: vdim must be 1 or the mode must be VECTOR");
            return 1;
         }
         if (spaces[i]->mesh != I->mesh) { delete I; set_error("madb_integrator_create: all spaces must live on one mesh"); return 1; }
         FieldDesc fd;
         fd.space = spaces[i];
         fd.mode = m;
         fd.role = roles ? roles[i] : ROLE_INPUT;
         fd.qvalue = qvalue;
         I->fields.push_back(fd);
         if (fd.role == ROLE_INPUT && !qvalue) { max_order = std::max(max_order, spaces[i]->order); } // a quadrature space has no order
      }
      I->quad_order = quad_order >= 0 ? quad_order : 2 * max_order + 2; // src/_ad_intg.hpp:103-104, :303-312
      I->nq1d = rule_npts_1d(I->quad_order);
      for (const FieldDesc &fd : I->fields)
      {
         if (fd.qvalue && (fd.space->basis != MADB_BASIS_L2 || fd.space->order + 1 != I->nq1d))
         {
            delete I;
            set_error("madb_integrator_create: QVALUE needs a quadrature space: MADB_BASIS_L2 of order " + std::to_string(I->nq1d - 1) +
                      " (nodes = the " + std::to_string(I->nq1d) + " Gauss points per direction of the rule)");
            return 1;
         }
      }
      std::string key = f->key() + "|d" + std::to_string(I->mesh->dim) + "q" + std::to_string(I->nq1d);
      if (I->mesh->simplex)
      {
         std::vector<double> tp, tw;
         if (!triangle_rule(I->quad_order, tp, tw)) { const int o = I->quad_order; delete I; set_error("madb_integrator_create: triangle rules of order 0 - 6 are available, not " + std::to_string(o)); return 1; }
         key = f->key() + "|s2q" + std::to_string(tw.size());
      }
      for (const FieldDesc &fd : I->fields)
      {
         key += "|" + std::to_string(I->mesh->simplex ? fd.space->nd_el() : fd.space->order + 1) + "." + std::to_string(fd.space->vdim) + "." +
                std::to_string((int)(fd.mode & (EV_VALUE | EV_GRAD))) + "." + std::to_string(fd.role);
      }
      // One vector space with ADEval::VECTOR = ADNonlinearFormIntegrator<...|VECTOR>: the reference's single-space
      // arithmetic (src/ad_intg.hpp:310-326) unless the caller asks for the block integrator's
      // index-consistent contraction (ADBlockNonlinearFormIntegrator, :700-727).
      {
         int ninput = 0;
         const FieldDesc *fin = nullptr;
         for (const FieldDesc &fd : I->fields) { if (fd.role == ROLE_INPUT) { ninput++; fin = &fd; } }
         // (with one shape slot per component -- VALUE only -- the windows of :318 coincide with the (c, r) blocks)
         const int sd = fin ? ((fin->mode & EV_VALUE) ? 1 : 0) + ((fin->mode & EV_GRAD) ? I->mesh->dim : 0) : 0;
         if (ninput == 1 && (fin->mode & EV_VECTOR) && fin->space->vdim > 1 && sd > 1 && !(flags & MADB_INTEG_BLOCK))
         {
            if (registry().find(key + "|refvec") == registry().end())
            {
               set_error("the reference's single-space VECTOR arithmetic (src/ad_intg.hpp:310-326) is not compiled for '" + key +
                         "': add MADB_INSTANCE_REFVEC, or pass MADB_INTEG_BLOCK for the index-consistent block contraction");
               delete I;
               return 1;
            }
            key += "|refvec";
         }
      }
      I->key = key;
      auto it = registry().find(key);
      if (it == registry().end())
      {
         std::string have;
         for (auto &kv : registry()) { if (kv.first.compare(0, f->kind.size(), f->kind) == 0) { have += "\n    " + kv.first; } }
         set_error("no fused kernel compiled for configuration '" + key + "'. Add a MADB_INSTANCE line (csrc/instances_*.cu) and rebuild." +
                   (have.empty() ? "" : " Compiled variants of this functional:" + have));
         delete I;
         return 1;
      }
      I->ops = it->second;
      const int rc = setup_integrator(*I);
      if (rc) { delete I; return rc; }
      if (I->nvd != I->ops.nvd || I->ndof_all != I->ops.ndof_all || I->nq != I->ops.nq)
      {
         set_error("internal: run-time sizes disagree with the compiled configuration '" + key + "'");
         delete I;
         return 1;
      }
      *out = I;
      return 0;
   }
   int madb_integrator_destroy(madb_integrator *I) { delete I; return 0; }

   int madb_integrator_sizes(madb_integrator *I, int64_t *ntotal, int *nq_el, int *ncolors)
   {
      if (ntotal) { *ntotal = I->ntotal; }
      if (nq_el) { *nq_el = I->nq; }
      if (ncolors)
      {
         *ncolors = (int)I->color_off.size() - 1;
      }
      return 0;
   }
   int madb_patch_selftest(int dim, int ne, const int32_t *e2n, int nnodes, const double *coords, int order, int vdim,
                           int ordering, int ndofs, const int32_t *e2l, double *max_err, int64_t *stats)
   {
      if (dim < 1 || dim > 3 || ne <= 0 || !e2n || !coords || !e2l || order < 1 || vdim < 1) { set_error("madb_patch_selftest: bad arguments"); return 1; }
      Mesh m;
      m.ctx = nullptr; m.dim = dim; m.ne = ne; m.geom_order = 1; m.nnodes = nnodes;
      m.e2n.assign(e2n, e2n + (size_t)ne * (1 << dim));
      m.coords.assign(coords, coords + (size_t)nnodes * dim);
      Space s;
      s.ctx = nullptr; s.mesh = &m; s.basis = BASIS_H1; s.order = order; s.vdim = vdim; s.ordering = ordering; s.ndofs = ndofs;
      s.e2l.assign(e2l, e2l + (size_t)ne * s.nd_el());
      long st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const int rc = patch_selftest(m, s, max_err, st);
      if (stats) { for (int k = 0; k < 8; k++) { stats[k] = st[k]; } }
      return rc;
   }

   int madb_patch_selftest_img(int dim, int ne, const int32_t *e2n, int nnodes, const double *coords, int order, int vdim,
                               int ordering, int ndofs, const int32_t *e2l, int tpe, double *max_err, int64_t *stats)
   {
      if (dim < 1 || dim > 3 || ne <= 0 || !e2n || !coords || !e2l || order < 1 || vdim < 1) { set_error("madb_patch_selftest_img: bad arguments"); return 1; }
      Mesh m;
      m.ctx = nullptr; m.dim = dim; m.ne = ne; m.geom_order = 1; m.nnodes = nnodes;
      m.e2n.assign(e2n, e2n + (size_t)ne * (1 << dim));
      m.coords.assign(coords, coords + (size_t)nnodes * dim);
      Space s;
      s.ctx = nullptr; s.mesh = &m; s.basis = BASIS_H1; s.order = order; s.vdim = vdim; s.ordering = ordering; s.ndofs = ndofs;
      s.e2l.assign(e2l, e2l + (size_t)ne * s.nd_el());
      long st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const int rc = patch_selftest_img(m, s, tpe, max_err, st);
      if (stats) { for (int k = 0; k < 8; k++) { stats[k] = st[k]; } }
      return rc;
   }

   int madb_integrator_set_timing(madb_integrator *I, int on)
   {
      CUDA_OK(cudaSetDevice(I->ctx->device));
      if (on && !I->ev0)
      {
         CUDA_OK(cudaEventCreate(&I->ev0));
         CUDA_OK(cudaEventCreate(&I->ev1));
      }
      I->timing = on != 0;
      return 0;
   }
   int madb_integrator_last_kernel_ms(madb_integrator *I, double *ms)
   {
      if (!I->timing || !I->ev0) { set_error("madb_integrator_last_kernel_ms: timing is off (madb_integrator_set_timing)"); return 1; }
      CUDA_OK(cudaEventSynchronize(I->ev1));
      float t = 0.f;
      CUDA_OK(cudaEventElapsedTime(&t, I->ev0, I->ev1));
      *ms = t;
      return 0;
   }
   int madb_integrator_patch_stats(madb_integrator *I, int64_t *out)
   {
      for (int k = 0; k < 8; k++) { out[k] = 0; }
      if (!I->use_patches) { return 0; }
      out[0] = (int64_t)I->pdesc.size();
      out[3] = I->pdev.ny_ifc;
      out[4] = I->pdev.nv_ifc;
      for (const PatchDesc &D : I->pdesc)
      {
         out[1] = std::max<int64_t>(out[1], D.nrows);
         out[2] = std::max<int64_t>(out[2], D.nslots);
         out[5] += D.nrows - D.nrow_int;
         out[6] += D.nslots - D.nexc;
         out[7] += D.nruns;
      }
      return 0;
   }

   int madb_integrator_set_param_field(madb_integrator *I, int field, const double *dofs)
   {
      if (field < 0 || field >= (int)I->fields.size() || I->fields[field].role != ROLE_PARAM)
      {
         set_error("madb_integrator_set_param_field: field is not a parameter field");
         return 1;
      }
      CUDA_OK(cudaSetDevice(I->ctx->device));
      const size_t n = (size_t)I->fields[field].space->vsize();
      if (is_device_ptr(dofs)) { I->pdata[field] = dofs; return 0; }
      if (!I->d_pstage[field]) { CUDA_OK(cudaMalloc((void **)&I->d_pstage[field], n * sizeof(double))); }
      CUDA_OK(cudaMemcpyAsync(I->d_pstage[field], dofs, n * sizeof(double), cudaMemcpyHostToDevice, I->ctx->stream));
      CUDA_OK(cudaStreamSynchronize(I->ctx->stream));
      I->pdata[field] = I->d_pstage[field];
      return 0;
   }

   int madb_integrator_set_param_qf(madb_integrator *I, int nqf, const double *qf)
   {
      const int need = I->ops.n_qprm - I->ops.n_field_qprm;
      if (nqf != need) { set_error("madb_integrator_set_param_qf: functional expects " + std::to_string(need) + " quadrature-function parameters, got " + std::to_string(nqf)); return 1; }
      CUDA_OK(cudaSetDevice(I->ctx->device));
      // caller layout (QuadratureFunction): qf[(e*nq + q)*nqf + k]  ->  device [k][q][sorted t]
      const size_t cnt = (size_t)I->ne * I->nq * nqf;
      std::vector<double> h(cnt);
      CUDA_OK(cudaMemcpy(h.data(), qf, cnt * sizeof(double), cudaMemcpyDefault));
      std::vector<double> tr((size_t)nqf * I->nq * I->stride, 0.0);
      for (int t = 0; t < I->ne; t++)
      {
         const int e = I->perm[t];
         for (int q = 0; q < I->nq; q++)
         {
            for (int k = 0; k < nqf; k++) { tr[((size_t)k * I->nq + q) * I->stride + t] = h[((size_t)e * I->nq + q) * nqf + k]; }
         }
      }
      if (!I->d_qf) { CUDA_OK(cudaMalloc((void **)&I->d_qf, tr.size() * sizeof(double))); }
      CUDA_OK(cudaMemcpy(I->d_qf, tr.data(), tr.size() * sizeof(double), cudaMemcpyHostToDevice));
      I->qf_count = nqf;
      return 0;
   }

   int madb_integrator_set_essential(madb_integrator *I, int n, const int32_t *dofs)
   {
      CUDA_OK(cudaSetDevice(I->ctx->device));
      cudaFree(I->d_ess);
      I->d_ess = nullptr;
      I->ness = n;
      if (n > 0)
      {
         for (int i = 0; i < n; i++) { if (dofs[i] < 0 || dofs[i] >= I->ntotal) { set_error("madb_integrator_set_essential: dof out of range"); I->ness = 0; return 1; } }
         CUDA_OK(cudaMalloc((void **)&I->d_ess, (size_t)n * sizeof(int)));
         CUDA_OK(cudaMemcpy(I->d_ess, dofs, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
      }
      return 0;
   }

   int madb_integrator_energy(madb_integrator *I, const double *x, double *energy)
   {
      return run(*I, MODE_ENERGY, x, nullptr, nullptr, nullptr, energy);
   }
   int madb_integrator_mult(madb_integrator *I, const double *x, double *y)
   {
      return run(*I, MODE_RES, x, nullptr, y, nullptr, nullptr);
   }
   int madb_integrator_pattern(madb_integrator *I, int64_t *nrows, int64_t *nnz, int32_t *rowptr, int32_t *colidx)
   {
      build_pattern(*I);
      if (!I->have_pattern) { return 1; }
      if (nrows) { *nrows = I->ntotal; }
      if (nnz) { *nnz = (int64_t)I->colidx.size(); }
      if (rowptr)
      {
         if (is_device_ptr(rowptr)) { CUDA_OK(cudaMemcpy(rowptr, I->rowptr.data(), I->rowptr.size() * sizeof(int), cudaMemcpyHostToDevice)); }
         else { std::memcpy(rowptr, I->rowptr.data(), I->rowptr.size() * sizeof(int)); }
      }
      if (colidx)
      {
         if (is_device_ptr(colidx)) { CUDA_OK(cudaMemcpy(colidx, I->colidx.data(), I->colidx.size() * sizeof(int), cudaMemcpyHostToDevice)); }
         else { std::memcpy(colidx, I->colidx.data(), I->colidx.size() * sizeof(int)); }
      }
      return 0;
   }
   int madb_integrator_assemble_begin(madb_integrator *I, const double *x, double *y, double *vals)
   {
      if (!is_device_ptr(vals) || (y && !is_device_ptr(y))) { set_error("madb_integrator_assemble_begin: y and vals must be device pointers"); return 1; }
      if (I->pending_vals) { set_error("madb_integrator_assemble_begin: the previous assembly was not finished (madb_integrator_assemble_end)"); return 1; }
      return run(*I, MODE_RES | MODE_JAC, x, nullptr, y, vals, nullptr, nullptr, nullptr, nullptr, 0, 1);
   }
   int madb_integrator_assemble_end(madb_integrator *I)
   {
      if (!I->pending_vals) { return 0; } // nothing deferred (colour path, or no begin)
      CUDA_OK(cudaSetDevice(I->ctx->device));
      double *dvals = I->pending_vals;
      I->pending_vals = nullptr;
      const int rc = launch_ifc_v(I->pdev, dvals, I->ctx->stream);
      if (rc) { set_error(std::string("interface reduction: ") + cudaGetErrorString((cudaError_t)rc)); return 2; }
      if (I->ness > 0) { k_ess_rowcol<<<(I->ness + 127) / 128, 128, 0, I->ctx->stream>>>(I->d_ess, I->ness, I->d_rowptr, I->d_colidx, dvals); }
      CUDA_OK(cudaGetLastError());
      return 0;
   }
   int madb_solver_create(madb_integrator *I, madb_solver **out)
   {
      CUDA_OK(cudaSetDevice(I->ctx->device));
      if (I->ops.matrix_free_only) { set_error("madb_solver_create: the integrator has no assembled Jacobian (matrix-free only)"); return 1; }
      if (ensure_pattern_device(*I)) { return 1; }
      return solver_new(I->ctx, (int)I->ntotal, (long)I->colidx.size(), I->d_rowptr, I->d_colidx, out);
   }
   int madb_integrator_grad_assemble(madb_integrator *I, const double *x, double *vals)
   {
      return run(*I, MODE_RES | MODE_JAC, x, nullptr, nullptr, vals, nullptr);
   }
   int madb_integrator_assemble(madb_integrator *I, const double *x, double *y, double *vals)
   {
      return run(*I, MODE_RES | MODE_JAC, x, nullptr, y, vals, nullptr);
   }
   int madb_integrator_coefficient(madb_integrator *I, const double *x, double *value, double *grad)
   {
      if (I->ops.map_aos) { set_error("madb_integrator_coefficient: not available for the sum-factorised kernels"); return 1; }
      return run(*I, MODE_COEF, x, nullptr, nullptr, nullptr, nullptr, value, grad);
   }
   int madb_integrator_coefficient_hessian(madb_integrator *I, const double *x, double *value, double *grad, double *hess)
   {
      if (I->ops.map_aos) { set_error("madb_integrator_coefficient_hessian: not available for the sum-factorised kernels"); return 1; }
      return run(*I, MODE_COEF, x, nullptr, nullptr, nullptr, nullptr, value, grad, hess);
   }
   int madb_integrator_param_gradient(madb_integrator *I, const double *design, double *value, double *J, int variant)
   {
      if (I->ops.map_aos) { set_error("madb_integrator_param_gradient: not available for the sum-factorised kernels"); return 1; }
      if (variant != MADB_PARAMGRAD_AS_WRITTEN && variant != MADB_PARAMGRAD_DERIVATIVE) { set_error("madb_integrator_param_gradient: unknown variant"); return 1; }
      if (variant == MADB_PARAMGRAD_AS_WRITTEN && !I->ops.has_param_gradient)
      {
         set_error("madb_integrator_param_gradient: functional '" + I->fn->key() + "' does not implement ParamGradient::Eval as written");
         return 1;
      }
      return run(*I, MODE_COEF, design, nullptr, nullptr, nullptr, nullptr, value, J, nullptr, variant == MADB_PARAMGRAD_AS_WRITTEN ? 1 : 0);
   }
   int madb_integrator_qpoint_coords(madb_integrator *I, double *xyz)
   {
      // physical coordinates of the rule's points, [e][q][dim]: where Coefficient-type Evaluator sources
      // (src/ad_native.hpp:56-61, src/ad_native.cpp:132-165) are sampled before they are handed over as a
      // QuadratureFunction (madb_integrator_set_param_qf)
      const Mesh &M = *I->mesh;
      const int dim = M.dim, ngn = M.ngn(), nq = I->nq, nq1 = I->nq1d;
      std::vector<double> out((size_t)I->ne * nq * dim);
      if (M.simplex)
      {
         for (int e = 0; e < I->ne; e++)
         {
            for (int q = 0; q < nq; q++)
            {
               const double xi = I->tri_pts[2 * q], eta = I->tri_pts[2 * q + 1], N[3] = {1.0 - xi - eta, xi, eta};
               double *o = &out[((size_t)e * nq + q) * 2];
               o[0] = o[1] = 0.0;
               for (int k = 0; k < 3; k++)
               {
                  const double *X = &M.coords[(size_t)M.e2n[(size_t)e * 3 + k] * 2];
                  o[0] += N[k] * X[0];
                  o[1] += N[k] * X[1];
               }
            }
         }
      }
      for (int e = 0; e < (M.simplex ? 0 : I->ne); e++)
      {
         for (int q = 0; q < nq; q++)
         {
            double xi[3] = {0, 0, 0};
            int r = q;
            for (int d = 0; d < dim; d++) { xi[d] = I->xq1d[r % nq1]; r /= nq1; }
            double *o = &out[((size_t)e * nq + q) * dim];
            for (int d = 0; d < dim; d++) { o[d] = 0.0; }
            for (int k = 0; k < ngn; k++)
            {
               double N = 1.0;
               for (int d = 0; d < dim; d++) { N *= ((k >> d) & 1) ? xi[d] : 1.0 - xi[d]; }
               const double *X = &M.coords[(size_t)M.e2n[(size_t)e * ngn + k] * dim];
               for (int d = 0; d < dim; d++) { o[d] += N * X[d]; }
            }
         }
      }
      CUDA_OK(cudaMemcpy(xyz, out.data(), out.size() * sizeof(double), cudaMemcpyDefault));
      return 0;
   }
   int madb_integrator_grad_mult(madb_integrator *I, const double *x, const double *v, double *y)
   {
      return run(*I, MODE_ACT, x, v, y, nullptr, nullptr);
   }
} // extern "C"

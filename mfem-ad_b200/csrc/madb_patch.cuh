// madb_patch.cuh -- patch assembly kernels (device side of madb_patch.cpp).
//
// One CTA = one patch of PATCH_PE elements, one thread per element:
//   0. one elected thread starts bulk copies (cp.async.bulk, mbarrier completion) of the
//      patch's gather maps into shared memory; they land while the CTA computes
//   1. gather + quadrature loop in registers (element_compute, madb_kernels.cuh)
//   2. every thread stages its element vector / upper-triangular element matrix in
//      shared memory ([entry][element], padded leading dimension)
//   3. every row / CSR entry ("slot") of the patch is summed from its sources in ascending
//      element order and written once, coalesced: interior rows straight to y / the CSR
//      values (runs of consecutive CSR positions), interface rows to a staging buffer
//   4. k_ifc_reduce adds the staged partial rows in ascending patch order.
// Replaces AddElementVector / SparseMatrix::AddSubMatrix of MFEM's element loop
//  without atomics and without order dependence.
#pragma once
#include "madb_kernels.cuh"

namespace madb
{

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
   asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                "l"(src), "r"(bytes), "r"(smem_u32(bar))
                : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
   unsigned done = 0;
   while (!done)
   {
      asm volatile("{\n"
                   ".reg .pred p;\n"
                   "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                   "selp.u32 %0, 1, 0, p;\n"
                   "}"
                   : "=r"(done)
                   : "r"(smem_u32(bar)), "r"(parity)
                   : "memory");
   }
}

template <class Func, class Cfg, int MODE, bool UNROLLQ>
__global__ void __launch_bounds__(PATCH_PE) k_patch(const __grid_constant__ AsmArgs<Func, Cfg> a,
                                                    const __grid_constant__ PatchDev P)
{
   constexpr int NVD = Cfg::NVD, NSYM = Cfg::NSYM, PE = PATCH_PE, LD = PATCH_LD;
   constexpr bool HAS_Y = (MODE & (MODE_RES | MODE_ACT)) != 0, HAS_V = (MODE & MODE_JAC) != 0;
   constexpr int SR_BYTES = patch_al16(NVD * LD * 8), SA_BYTES = patch_al16(NSYM * LD * 8);
   extern __shared__ __align__(16) unsigned char smraw[];
   __shared__ PatchDesc D;
   __shared__ __align__(8) unsigned long long mbar;
   const int tid = threadIdx.x, p = blockIdx.x;
   if (tid < (int)(sizeof(PatchDesc) / sizeof(int))) { ((int *)&D)[tid] = ((const int *)(P.desc + p))[tid]; }
   if (tid == 0) { mbar_init(&mbar, 1); }
   __syncthreads();
   const bool wy = HAS_Y && a.write_y, wv = HAS_V && a.write_vals;

   // shared-memory carve-up (byte offsets into smraw): staged element vectors | matrices | y maps | matrix maps
   const int o_sa = wy ? SR_BYTES : 0;
   const int o_yb = o_sa + (wv ? SA_BYTES : 0);
   const int o_vb = o_yb + (wy ? P.max_yblob : 0);
#define MADB_SR(i) (*(double *)(smraw + 8 * (i)))
#define MADB_SA(i) (*(double *)(smraw + o_sa + 8 * (i)))
   if (tid == 0)
   {
      const unsigned bytes = (wy ? D.yblob_bytes : 0) + (wv ? D.vblob_bytes : 0);
      mbar_expect_tx(&mbar, bytes);
      if (wy && D.yblob_bytes) { bulk_g2s(smraw + o_yb, P.yblob + (size_t)D.yblob_off * 16, D.yblob_bytes, &mbar); }
      if (wv && D.vblob_bytes) { bulk_g2s(smraw + o_vb, P.vblob + (size_t)D.vblob_off * 16, D.vblob_bytes, &mbar); }
   }

   const int t = p * PE + tid;
   const bool valid = tid < D.ne;
   {
      double r[HAS_Y ? NVD : 1];
      double A[HAS_V ? NSYM : 1];
      double energy;
      if (valid)
      {
         element_compute<Func, Cfg, MODE, UNROLLQ>(a, t, r, A, energy);
         if constexpr (HAS_Y)
         {
            if (wy)
            {
#pragma unroll
               for (int i = 0; i < NVD; i++) { MADB_SR(i * LD + tid) = r[i]; }
            }
         }
         if constexpr (HAS_V)
         {
            if (wv)
            {
#pragma unroll
               for (int k = 0; k < NSYM; k++) { MADB_SA(k * LD + tid) = A[k]; }
            }
         }
      }
   }
   __syncthreads();
   mbar_wait(&mbar, 0);

   // ---- fold: add the further sources of every row / slot onto its first source, phase by phase -------
   const int o_yfold = o_yb + patch_al16(2 * D.nrows);
   const int o_vfold = o_vb + patch_al16(2 * D.nslots);
   {
      int ybase = 8, vbase = 8;
      for (int ph = 0; ph < 8; ph++)
      {
         const int ny = wy ? *(const int *)(smraw + o_yfold + 4 * ph) : 0;
         const int nv = wv ? *(const int *)(smraw + o_vfold + 4 * ph) : 0;
         if (ny == 0 && nv == 0) { break; }
         if constexpr (HAS_Y)
         {
            for (int i = tid; i < ny; i += PE)
            {
               const unsigned w = *(const unsigned *)(smraw + o_yfold + 4 * (ybase + i));
               MADB_SR(w & 0xffffu) += MADB_SR(w >> 16);
            }
         }
         if constexpr (HAS_V)
         {
            for (int i = tid; i < nv; i += PE)
            {
               const unsigned w = *(const unsigned *)(smraw + o_vfold + 4 * (vbase + i));
               MADB_SA(w & 0xffffu) += MADB_SA(w >> 16);
            }
         }
         ybase += ny;
         vbase += nv;
         __syncthreads();
      }
   }

   // ---- rows of the residual -----------------------------------------------------------
   if constexpr (HAS_Y)
   {
      if (wy)
      {
         const int o_ylist = o_yfold + patch_al16(4 * D.nyfold);
         for (int lr = tid; lr < D.nrows; lr += PE)
         {
            const double v = MADB_SR(*(const unsigned short *)(smraw + o_yb + 2 * lr));
            if (lr < D.nrow_int) { a.y[*(const int *)(smraw + o_ylist + 4 * lr)] = v; }
            else { P.ystage[D.ystage_off + (lr - D.nrow_int)] = v; }
         }
      }
   }
   // ---- CSR entries ------------------------------------------------------------------------
   if constexpr (HAS_V)
   {
      if (wv)
      {
         const int o_runs = o_vfold + patch_al16(4 * D.nvfold);
         const int o_rung = o_runs + patch_al16(4 * (D.nruns + 1));
         const int o_xg = o_rung + patch_al16(4 * (D.nruns + 1));
#define MADB_RUN_S(r) (*(const int *)(smraw + o_runs + 4 * (r)))
#define MADB_RUN_G(r) (*(const int *)(smraw + o_rung + 4 * (r)))
         // interior slots are numbered in CSR order: slot s of run r goes to position run_g[r] + (s - run_s[r])
         int rn = 0, rs = MADB_RUN_S(0), re = (D.nruns > 0) ? MADB_RUN_S(1) : 0, rg = MADB_RUN_G(0);
         const int nint = D.nint, nexc = D.nexc, nslots = D.nslots;
#pragma unroll 4
         for (int s = tid; s < nint; s += PE)
         {
            const double v = MADB_SA(*(const unsigned short *)(smraw + o_vb + 2 * s));
            while (s >= re)
            {
               rn++;
               rs = re;
               re = MADB_RUN_S(rn + 1);
               rg = MADB_RUN_G(rn);
            }
            a.vals[rg + (s - rs)] = v;
         }
         for (int s = nint + tid; s < nexc; s += PE)
         {
            a.vals[*(const int *)(smraw + o_xg + 4 * (s - nint))] = MADB_SA(*(const unsigned short *)(smraw + o_vb + 2 * s));
         }
         double *stage = P.vstage + D.stage_off - nexc;
         for (int s = nexc + tid; s < nslots; s += PE) { stage[s] = MADB_SA(*(const unsigned short *)(smraw + o_vb + 2 * s)); }
#undef MADB_RUN_S
#undef MADB_RUN_G
      }
   }
#undef MADB_SR
#undef MADB_SA
}

// out[dst[i]] = sum of the staged partials of entry i, in ascending patch order
static __global__ void __launch_bounds__(256) k_ifc_reduce(int n, const int *__restrict__ ptr, const int *__restrict__ src,
                                                           const int *__restrict__ dst, const double *__restrict__ stage,
                                                           double *__restrict__ out)
{
   const int i = blockIdx.x * 256 + threadIdx.x;
   if (i >= n) { return; }
   const int b = ptr[i], e = ptr[i + 1];
   double s = stage[src[b]];
   for (int k = b + 1; k < e; k++) { s += stage[src[k]]; }
   out[dst[i]] = s;
}

template <class Func, class Cfg, int MODE, bool UNROLLQ>
int launch_patch_mode(const AsmArgs<Func, Cfg> &a, const LaunchCtx &L)
{
   const PatchDev &P = *L.patch;
   static int smem_set = 0;
   auto kern = k_patch<Func, Cfg, MODE, UNROLLQ>;
   const bool wy = (MODE & (MODE_RES | MODE_ACT)) && L.write_y, wv = (MODE & MODE_JAC) && L.write_vals;
   const int smem_bytes = (wy ? patch_al16(Cfg::NVD * PATCH_LD * 8) + P.max_yblob : 0) +
                          (wv ? patch_al16(Cfg::NSYM * PATCH_LD * 8) + P.max_vblob : 0) + 16;
   if (smem_bytes > smem_set)
   {
      const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
      if (e != cudaSuccess) { return (int)e; }
      smem_set = smem_bytes;
   }
   if (L.ev0) { cudaEventRecord(L.ev0, L.stream); }
   kern<<<P.npatch, PATCH_PE, smem_bytes, L.stream>>>(a, P);
   if (L.ev1) { cudaEventRecord(L.ev1, L.stream); }
   if (wy && P.ny_ifc > 0)
   {
      k_ifc_reduce<<<(P.ny_ifc + 255) / 256, 256, 0, L.stream>>>(P.ny_ifc, P.y_ptr, P.y_src, P.y_dst, P.ystage, L.y);
   }
   if (wv && P.nv_ifc > 0)
   {
      k_ifc_reduce<<<(P.nv_ifc + 255) / 256, 256, 0, L.stream>>>(P.nv_ifc, P.v_ptr, P.v_src, P.v_dst, P.vstage, L.vals);
   }
   return (int)cudaGetLastError();
}

} // namespace madb

// madb_patch.cuh -- patch assembly kernels (device side of madb_patch.cpp).
//
// One CTA = one patch of PATCH_PE elements, one thread per element:
//   1. gather + quadrature loop in registers (element_compute, madb_kernels.cuh)
//   2. accumulate the element vectors / matrices into the patch's rows in shared
//      memory, colour by colour (elements of one colour share no dof: plain
//      load/add/store, fixed order)
//   3. write the interior rows of the patch to y / the CSR values once, coalesced
//      (runs of consecutive CSR positions); interface rows go to a staging buffer
//   4. k_ifc_reduce adds the staged partial rows in ascending patch order.
// Replaces AddElementVector / SparseMatrix::AddSubMatrix of MFEM's element loop
//  without atomics and without order dependence.
#pragma once
#include "madb_kernels.cuh"

namespace madb
{

template <class Func, class Cfg, int MODE, bool UNROLLQ>
__global__ void __launch_bounds__(PATCH_PE) k_patch(const __grid_constant__ AsmArgs<Func, Cfg> a,
                                                    const __grid_constant__ PatchDev P)
{
   constexpr int NVD = Cfg::NVD, PE = PATCH_PE;
   constexpr bool HAS_Y = (MODE & (MODE_RES | MODE_ACT)) != 0, HAS_V = (MODE & MODE_JAC) != 0;
   constexpr int NYW = (NVD + 1) / 2, NVW = (NVD * NVD + 1) / 2;
   extern __shared__ double sm[];
   __shared__ PatchDesc D;
   const int tid = threadIdx.x, p = blockIdx.x;
   if (tid < (int)(sizeof(PatchDesc) / sizeof(int))) { ((int *)&D)[tid] = ((const int *)(P.desc + p))[tid]; }
   __syncthreads();
   const bool wy = HAS_Y && a.write_y, wv = HAS_V && a.write_vals;
   const int nslots = wv ? D.nslots : 0;
   const int nslots_pad = (nslots + 15) & ~15, nrows_pad = (D.nrows + 15) & ~15;
   double *out = sm, *yout = sm + nslots_pad;
   int *srun_s = (int *)(sm + nslots_pad + nrows_pad), *srun_g = srun_s + D.nruns + 1;
   for (int s = tid; s < nslots_pad + nrows_pad; s += PE) { sm[s] = 0.0; }
   if constexpr (HAS_V)
   {
      if (wv)
      {
         for (int k = tid; k <= D.nruns; k += PE)
         {
            srun_s[k] = __ldg(P.run_s + D.run_off + k);
            srun_g[k] = __ldg(P.run_g + D.run_off + k);
         }
      }
   }

   const int t = p * PE + tid;
   const bool valid = tid < D.ne;
   double r[HAS_Y ? NVD : 1];
   double A[HAS_V ? Cfg::NSYM : 1];
   double energy;
   if (valid) { element_compute<Func, Cfg, MODE, UNROLLQ>(a, t, r, A, energy); }

   // slot maps of this element, two 16-bit slots per word, fetched by all warps before the serial colour phases
   unsigned yw[HAS_Y ? NYW : 1], vw[HAS_V ? NVW : 1];
   if (valid)
   {
      if constexpr (HAS_Y)
      {
         if (wy)
         {
#pragma unroll
            for (int k = 0; k < NYW; k++) { yw[k] = __ldg((const unsigned *)P.yslot + (size_t)k * a.stride + t); }
         }
      }
      if constexpr (HAS_V)
      {
         if (wv)
         {
#pragma unroll
            for (int k = 0; k < NVW; k++) { vw[k] = __ldg((const unsigned *)P.pslot + (size_t)k * a.stride + t); }
         }
      }
   }
   __syncthreads();

   int mycol = -1;
   if (valid)
   {
      for (int c = 0; c < D.ncol; c++) { if (tid >= D.col_off[c] && tid < D.col_off[c + 1]) { mycol = c; } }
   }
   for (int c = 0; c < D.ncol; c++)
   {
      if (mycol == c)
      {
         if constexpr (HAS_Y)
         {
            if (wy)
            {
               double old[NVD];
#pragma unroll
               for (int i = 0; i < NVD; i++) { old[i] = yout[(yw[i >> 1] >> ((i & 1) * 16)) & 0xffffu]; }
#pragma unroll
               for (int i = 0; i < NVD; i++) { yout[(yw[i >> 1] >> ((i & 1) * 16)) & 0xffffu] = old[i] + r[i]; }
            }
         }
         if constexpr (HAS_V)
         {
            if (wv)
            {
#pragma unroll
               for (int i = 0; i < NVD; i++)
               {
                  double old[NVD];
#pragma unroll
                  for (int j = 0; j < NVD; j++)
                  {
                     const int k = i * NVD + j;
                     old[j] = out[(vw[k >> 1] >> ((k & 1) * 16)) & 0xffffu];
                  }
#pragma unroll
                  for (int j = 0; j < NVD; j++)
                  {
                     const int k = i * NVD + j;
                     out[(vw[k >> 1] >> ((k & 1) * 16)) & 0xffffu] = old[j] + A[symidx(i, j)];
                  }
               }
            }
         }
      }
      __syncthreads();
   }

   // ---- write-out ---------------------------------------------------------------------
   if constexpr (HAS_Y)
   {
      if (wy)
      {
         for (int lr = tid; lr < D.nrow_int; lr += PE) { a.y[__ldg(P.ylist + D.y_off + lr)] = yout[patch_swz(lr)]; }
         for (int lr = D.nrow_int + tid; lr < D.nrows; lr += PE) { P.ystage[D.ystage_off + (lr - D.nrow_int)] = yout[patch_swz(lr)]; }
      }
   }
   if constexpr (HAS_V)
   {
      if (wv)
      {
         // interior slots are numbered in CSR order: slot s of run r goes to position run_g[r] + (s - run_s[r])
         if (D.nint > 0)
         {
            int rn = 0, rs = srun_s[0], re = srun_s[1], rg = srun_g[0];
            for (int s = tid; s < D.nint; s += PE)
            {
               while (s >= re)
               {
                  rn++;
                  rs = re;
                  re = srun_s[rn + 1];
                  rg = srun_g[rn];
               }
               a.vals[rg + (s - rs)] = out[patch_swz(s)];
            }
         }
         for (int s = D.nint + tid; s < nslots; s += PE) { P.vstage[D.stage_off + (s - D.nint)] = out[patch_swz(s)]; }
      }
   }
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
   const bool wv = (MODE & MODE_JAC) && L.write_vals;
   const int smem_bytes = (P.max_rows + 16 + (wv ? P.max_slots + 16 : 0)) * (int)sizeof(double) + (wv ? 2 * (P.max_runs + 1) * (int)sizeof(int) : 0);
   if (smem_bytes > smem_set)
   {
      const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
      if (e != cudaSuccess) { return (int)e; }
      smem_set = smem_bytes;
   }
   kern<<<P.npatch, PATCH_PE, smem_bytes, L.stream>>>(a, P);
   if ((MODE & (MODE_RES | MODE_ACT)) && L.write_y && P.ny_ifc > 0)
   {
      k_ifc_reduce<<<(P.ny_ifc + 255) / 256, 256, 0, L.stream>>>(P.ny_ifc, P.y_ptr, P.y_src, P.y_dst, P.ystage, L.y);
   }
   if ((MODE & MODE_JAC) && L.write_vals && P.nv_ifc > 0)
   {
      k_ifc_reduce<<<(P.nv_ifc + 255) / 256, 256, 0, L.stream>>>(P.nv_ifc, P.v_ptr, P.v_src, P.v_dst, P.vstage, L.vals);
   }
   return (int)cudaGetLastError();
}

} // namespace madb

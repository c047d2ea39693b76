// madb_patch_img.cuh -- fused residual + Jacobian patch kernel that builds a shared-memory IMAGE of the patch's part
// of the CSR value array and writes it with bulk copies (replaces AddElementVector / SparseMatrix::AddSubMatrix of
// MFEM's element loop, SURVEY a32, for element matrices of up to 10 dofs).
//
// Persistent kernel, one CTA per SM, NWG work groups of 128 threads; a work group processes one patch at a time:
//   1. gather + quadrature loop in registers, TPE threads per element (madb_sf2d_pair.cuh: the two threads of a pair
//      own half of the point rows each; other configurations: one thread per element, element_compute)
//   2. every finished element-vector / matrix entry is stored straight to its slot of the image through the scatter
//      maps (madb_patch.cpp: patch_build_img): the first source of a slot writes the slot itself, any further source
//      a private "extras" slot; symmetric entries are stored to both (i,j) and (j,i)
//   3. fold: the extras are added onto their slots in ascending element order (same thread per slot in every phase)
//   4. write-out: interior rows of the patch are runs of consecutive CSR positions -> one cp.async.bulk
//      (shared -> global) per run, issued by single threads; entries shared with other patches are contiguous in the
//      image -> one bulk copy to the staging buffer (summed across patches in ascending patch order by k_ifc_reduce);
//      entries of interface rows that only this patch contributes to are stored individually.
// The image of a patch is only written at the END of the next patch's computation, so a single buffer per work group is
// enough: the bulk copies have long finished reading it (cp.async.bulk.wait_group.read before the first scatter).
// No atomics: every value has one writer and a fixed summation order; results are bit-reproducible.
#pragma once
#include "madb_patch.cuh"
#include "madb_sf2d_pair.cuh"

namespace madb
{

constexpr int IMG_WG = 128; // threads of a work group

__device__ __forceinline__ void bulk_s2g(void *gdst, const void *ssrc, unsigned bytes)
{
   asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void wg_bar(const int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(IMG_WG) : "memory"); }

/// threads per element / work groups per CTA of the image kernel for <functional, configuration>
template <class Func, class Cfg> constexpr int img_tpe() { return sf2d_pair_ok<Func, Cfg>() ? 2 : 1; }
template <class Func, class Cfg> constexpr int img_nwg() { return sf2d_pair_ok<Func, Cfg>() ? 3 : 2; }

template <class Func, class Cfg, bool UNROLLQ>
__global__ void __launch_bounds__(IMG_WG *img_nwg<Func, Cfg>(), 1)
   k_patch_img(const __grid_constant__ AsmArgs<Func, Cfg> a, const __grid_constant__ PatchDev P)
{
   constexpr int MODE = MODE_RES | MODE_JAC;
   constexpr int NVD = Cfg::NVD, NSYM = Cfg::NSYM, TPE = img_tpe<Func, Cfg>(), NWG = img_nwg<Func, Cfg>(), PE = IMG_WG / TPE;
   extern __shared__ __align__(16) unsigned char smraw[];
   __shared__ __align__(8) unsigned long long bar_blob[NWG];
   const ImgDev &G = P.img;
   const int wg = __shfl_sync(0xffffffffu, (int)(threadIdx.x / IMG_WG), 0), tid = threadIdx.x % IMG_WG;
   const int l = tid / TPE, h = tid % TPE;
   const bool wy = a.write_y != 0;
   // shared memory of this work group: image | y image | scatter maps
   const int img_bytes = patch_al16(G.max_vslots * 8), yimg_bytes = patch_al16(G.max_yslots * 8);
   unsigned char *base = smraw + (size_t)wg * (img_bytes + yimg_bytes + G.max_mblob);
   double *IMG = (double *)base, *YIMG = (double *)(base + img_bytes);
   unsigned char *MAPS = base + img_bytes + yimg_bytes;
   if (threadIdx.x == 0) { for (int k = 0; k < NWG; k++) { mbar_init(&bar_blob[k], 1); } }
   __syncthreads();
   const bool al16 = (reinterpret_cast<size_t>(a.vals) & 15) == 0;

   // One blob per patch (descriptor | scatter maps | lists), fetched by one bulk copy.  It is requested when the
   // previous patch of the work group has been written out and lands during the quadrature loop of its own patch.
   auto prefetch = [&](int p)
   {
      const ImgDesc *d = G.desc + p;
      const int bytes = __ldg(&d->mblob_bytes), off = __ldg(&d->mblob_off);
      fence_async_smem();
      mbar_expect_tx(&bar_blob[wg], (unsigned)bytes);
      bulk_g2s(MAPS, G.mblob + (size_t)off * 16, (unsigned)bytes, &bar_blob[wg]);
   };
   {
      const int p0 = (int)blockIdx.x * NWG + wg;
      if (tid == 0 && p0 < P.npatch) { prefetch(p0); }
   }
   const ImgDesc *D = (const ImgDesc *)MAPS; // valid once the blob has landed

   for (int it = 0;; it++)
   {
      const int p = (it * (int)gridDim.x + (int)blockIdx.x) * NWG + wg;
      if (p >= P.npatch) { break; }
      const int ne = min(PE, a.end - p * PE);
      const bool valid = l < ne;
      const int t = p * PE + (valid ? l : 0); // lanes without an element recompute element 0 of the patch (shuffles need all lanes)
      bool first = true;
      // Everything up to the first store to the image overlaps the bulk copies of the previous patch.
      auto before_scatter = [&]()
      {
         if (!first) { return; }
         first = false;
         bulk_wait_read();                 // own bulk copies of the previous patch have read the image
         mbar_wait(&bar_blob[wg], it & 1); // the blob of this patch has landed
         wg_bar(1 + wg);                   // ... and everybody else's copies have read the image, too
      };
      auto vstore = [&](int e, double v)
      {
         const unsigned w = ((const unsigned *)(MAPS + D->o_vmap))[e * IMG_WG + tid];
         IMG[w & 0xffffu] = v;
         IMG[w >> 16] = v;
      };
      if constexpr (TPE == 2)
      {
         element_compute_sf2d_pair<Func, Cfg>(
            a, t, h,
            [&](int e, double v)
            {
               before_scatter();
               if (wy && valid) { YIMG[((const unsigned short *)(MAPS + D->o_ymap))[e * IMG_WG + tid]] = v; }
            },
            [&](int e, double v) { if (valid) { vstore(e, v); } }, [&]() { before_scatter(); });
      }
      else
      {
         double r[NVD], A[NSYM], energy;
         element_compute<Func, Cfg, MODE, UNROLLQ>(a, a.tab, t, r, A, energy);
         before_scatter();
         if (valid)
         {
            const unsigned short *ymap = (const unsigned short *)(MAPS + D->o_ymap) + tid;
            if (wy)
            {
#pragma unroll
               for (int i = 0; i < NVD; i++) { YIMG[ymap[i * IMG_WG]] = r[i]; }
            }
#pragma unroll
            for (int k = 0; k < NSYM; k++) { vstore(k, A[k]); }
         }
      }
      wg_bar(1 + wg); // all element data of the patch are in the image
      // ---- fold: further sources onto their slots, ascending element order ---------------------------------
      const int *gd = (const int *)(MAPS + D->o_lists);
      const int nvfold = D->nvfold, nyfold = D->nyfold;
      {
         const int *vf = gd;
         int fb = 8;
         for (int ph = 0; ph < 8; ph++)
         {
            const int n = vf[ph];
            if (n == 0) { break; }
            for (int i = tid; i < n; i += IMG_WG)
            {
               const unsigned w = (unsigned)vf[fb + i];
               IMG[w & 0xffffu] += IMG[w >> 16];
            }
            fb += n;
         }
         if (wy)
         {
            const int *yf = gd + img_al4(nvfold);
            fb = 8;
            for (int ph = 0; ph < 8; ph++)
            {
               const int n = yf[ph];
               if (n == 0) { break; }
               for (int i = tid; i < n; i += IMG_WG)
               {
                  const unsigned w = (unsigned)yf[fb + i];
                  YIMG[w & 0xffffu] += YIMG[w >> 16];
               }
               fb += n;
            }
         }
      }
      fence_async_smem(); // image writes (generic proxy) before the bulk copies (async proxy) that follow the barrier
      wg_bar(1 + wg);
      // ---- write-out ----------------------------------------------------------------------------------------
      const int nruns = D->nruns, nexcl = D->nexcl, nrows = D->nrows, nrow_int = D->nrow_int;
      const int4 *runs = (const int4 *)(gd + img_al4(nvfold) + img_al4(nyfold));
      const int *xg = (const int *)(runs + nruns); // 4 * nruns is a multiple of 4
      const unsigned short *xs = (const unsigned short *)(xg + img_al4(nexcl));
      const int *ylist = xg + img_al4(nexcl) + img_al4((nexcl + 1) / 2);
      if (al16)
      {
         for (int rr = tid; rr < nruns; rr += IMG_WG)
         {
            const int4 d = runs[rr];
            int so = d.x, g0 = d.y, n = d.z;
            if (g0 & 1) { a.vals[g0] = IMG[so]; so++; g0++; n--; }
            if (n & 1) { a.vals[g0 + n - 1] = IMG[so + n - 1]; n--; }
            if (n > 0) { bulk_s2g(a.vals + g0, IMG + so, (unsigned)n * 8u); }
         }
      }
      else
      {
         for (int rr = 0; rr < nruns; rr++)
         {
            const int4 d = runs[rr];
            for (int j = tid; j < d.z; j += IMG_WG) { a.vals[d.y + j] = IMG[d.x + j]; }
         }
      }
      {
         const int nsh = D->nsh;
         if (tid == IMG_WG - 1 && nsh > 0) { bulk_s2g(P.vstage + D->stage_off, IMG + D->sh0, (unsigned)nsh * 8u); }
      }
      bulk_commit();
      for (int q = tid; q < nexcl; q += IMG_WG) { a.vals[xg[q]] = IMG[xs[q]]; }
      if (wy)
      {
         const int yso = D->ystage_off;
         for (int lr = tid; lr < nrows; lr += IMG_WG)
         {
            const double v = YIMG[lr];
            if (lr < nrow_int) { a.y[ylist[lr]] = v; }
            else { P.ystage[yso + (lr - nrow_int)] = v; }
         }
      }
      wg_bar(1 + wg); // everybody is done with the blob: request the next one
      {
         const int pn = ((it + 1) * (int)gridDim.x + (int)blockIdx.x) * NWG + wg;
         if (tid == 0 && pn < P.npatch) { prefetch(pn); }
      }
   }
   bulk_wait_all();
}

template <class Func, class Cfg, bool UNROLLQ> int launch_patch_img(const AsmArgs<Func, Cfg> &a, const LaunchCtx &L)
{
   if constexpr (!img_eligible<Func, Cfg>()) { return (int)cudaErrorInvalidConfiguration; }
   else
   {
      const PatchDev &P = *L.patch;
      constexpr int NWG = img_nwg<Func, Cfg>();
      static int smem_set_dev[64] = {0}, nsm_dev[64] = {0};
      static std::mutex attr_mutex;
      std::lock_guard<std::mutex> attr_lock(attr_mutex);
      int dev = 0;
      cudaGetDevice(&dev);
      dev &= 63;
      auto kern = k_patch_img<Func, Cfg, UNROLLQ>;
      const int wg_bytes = patch_al16(P.img.max_vslots * 8) + patch_al16(P.img.max_yslots * 8) + P.img.max_mblob;
      const int smem_bytes = NWG * wg_bytes + 16;
      if (smem_bytes > 226 * 1024) { return (int)cudaErrorInvalidConfiguration; }
      if (smem_bytes > smem_set_dev[dev])
      {
         const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
         if (e != cudaSuccess) { return (int)e; }
         smem_set_dev[dev] = smem_bytes;
      }
      if (nsm_dev[dev] == 0) { cudaDeviceGetAttribute(&nsm_dev[dev], cudaDevAttrMultiProcessorCount, dev); }
      const int grid = std::min(nsm_dev[dev], (P.npatch + NWG - 1) / NWG);
      kern<<<grid, IMG_WG * NWG, smem_bytes, L.stream>>>(a, P);
      return (int)cudaGetLastError();
   }
}

} // namespace madb

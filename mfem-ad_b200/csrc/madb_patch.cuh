// madb_patch.cuh -- patch assembly kernels (device side of madb_patch.cpp).
//
// A patch = patch_pe(NVD) elements of the mesh (128, or 64 for element matrices of 11-20 dofs).  For one patch:
//   0. one elected thread starts bulk copies (cp.async.bulk, mbarrier completion) of the
//      patch's gather maps into shared memory; they land while the threads compute
//   1. gather + quadrature loop in registers (element_compute / element_compute_sf2d, madb_kernels.cuh):
//      one thread per element, or 4 threads per element each owning a slice of the upper triangle
//   2. the element vector / upper-triangular element matrix is staged in shared memory
//      ([entry][element], padded leading dimension); the sum-factorised path streams its entries there
//   3. fold: the further sources of every row / CSR entry ("slot") are added onto its first source, phase by
//      phase (ascending element order; the same thread owns a destination in every phase, one barrier at the
//      end); then every slot is read from its first source and written once, coalesced: interior rows straight
//      to y / the CSR values (chunk descriptors: an aligned pair of chunks = 64 consecutive CSR positions with
//      16-byte stores, otherwise 32 consecutive slots -> CSR positions), entries of interface rows only this
//      patch touches through explicit positions, entries shared with other patches to a staging buffer
//   4. k_ifc_reduce adds the staged partial rows / entries in ascending patch order.
// Two kernels: k_patch (one CTA per patch) and k_patch_ws (persistent, warp-specialised: compute warpgroups
// and writer warpgroups overlap steps 1-2 and 3 of consecutive patches; fused residual + Jacobian).
// Replaces AddElementVector / SparseMatrix::AddSubMatrix of MFEM's element loop
//  without atomics and without order dependence.
#pragma once
#include "madb_kernels.cuh"
#include "madb_sf2d_pair.cuh"
#include <algorithm>
#include <cstdlib>
#include <mutex>

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
#ifndef MADB_WS_L2PF
#define MADB_WS_L2PF 1 // 1: k_patch_ws compute warps prefetch the index lines of their next patch into L2 (config 2: -1.6 %); 2: the indices are
                       // staged in shared memory by cp.async one patch ahead (measured: ptxas spills 136 B in the compute warps, +19 %)
#endif
#ifndef MADB_PATCH_L2PF
#define MADB_PATCH_L2PF 0 // k_patch: prefetch distance (in waves of one CTA per SM) of the index lines of a later patch; 0: off
#endif
#ifndef MADB_WS_L1PF
#define MADB_WS_L1PF 0 // writers warm the L1 with the values of the compute warpgroup's next gather: 1 prefetch.global.L1, 2 discarded loads
#endif
__device__ __forceinline__ void l1_touch(const double *p)
{
#if MADB_WS_L1PF == 2
   double d;
   asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(d) : "l"(p));
#else
   asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}
#ifndef MADB_WS_L2PF_GEN
#define MADB_WS_L2PF_GEN 1 // the same prefetch in the generic (not sum-factorised) branch of k_patch_ws (config 4 state block: 0.611 -> 0.597 ms)
#endif
#ifndef MADB_BLOB_EVICT_FIRST
#define MADB_BLOB_EVICT_FIRST 0 // 1: the bulk copies of the per-patch maps (read once per launch) carry an L2 evict-first policy
#endif
#ifndef MADB_ST_CS
#define MADB_ST_CS 0 // 1: the final CSR values are written with streaming stores (st.global.cs)
#endif
__device__ __forceinline__ void st_out(double *p, const double v)
{
#if MADB_ST_CS
   __stcs(p, v);
#else
   *p = v;
#endif
}
__device__ __forceinline__ void st_out2(double *p, const double a, const double b)
{
#if MADB_ST_CS
   __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b));
#else
   *reinterpret_cast<double2 *>(p) = make_double2(a, b);
#endif
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
#if MADB_BLOB_EVICT_FIRST
   unsigned long long pol;
   asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                : "memory");
   return;
#endif
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                "l"(src), "r"(bytes), "r"(smem_u32(bar))
                : "memory");
}
#ifndef MADB_MBAR_HINT_NS
#define MADB_MBAR_HINT_NS 0 // suspend-time hint of mbarrier.try_wait (ns): a waiting warp leaves the issue slots to the others
#endif
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
   unsigned done = 0;
   while (!done)
   {
      asm volatile("{\n"
                   ".reg .pred p;\n"
#if MADB_MBAR_HINT_NS > 0
                   "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
#else
                   "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
#endif
                   "selp.u32 %0, 1, 0, p;\n"
                   "}"
                   : "=r"(done)
                   : "r"(smem_u32(bar)), "r"(parity)
#if MADB_MBAR_HINT_NS > 0
                     , "r"((unsigned)MADB_MBAR_HINT_NS)
#endif
                   : "memory");
   }
}

/// barrier of the NT threads draining one patch: the whole CTA (BAR = 0) or the writer warps (named barrier BAR)
template <int BAR, int NT> __device__ __forceinline__ void patch_bar(const int id = BAR)
{
   if constexpr (BAR == 0) { __syncthreads(); }
   else { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NT) : "memory"); }
}

/// Programmatic dependent launch (MADB_PDL, default on): the element kernel and the interface reduction follow each
/// other on one stream, step after step.  Launched with the programmatic-serialisation attribute, the CTAs of the next
/// kernel are placed on SMs as the CTAs of the previous one exit (its tail and the launch latency overlap); every kernel
/// of the chain waits (`griddepcontrol.wait`: completion and visibility of the whole previous grid) before it touches
/// global memory, so the ordering of the data is that of a plain stream.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline bool pdl_enabled()
{
   static const bool on = getenv("MADB_PDL") ? atoi(getenv("MADB_PDL")) != 0 : true;
   return on;
}
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), const int grid, const int block, const size_t smem, cudaStream_t stream, Args &&...args)
{
   cudaLaunchConfig_t cfg = {};
   cfg.gridDim = dim3(grid);
   cfg.blockDim = dim3(block);
   cfg.dynamicSmemBytes = smem;
   cfg.stream = stream;
   cudaLaunchAttribute at[1];
   at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
   at[0].val.programmaticStreamSerializationAllowed = 1;
   cfg.attrs = at;
   cfg.numAttrs = pdl_enabled() ? 1 : 0;
   return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

/// Fold + gather + write-out of one staged patch by NT threads (tid = 0..NT-1).
/// base: shared memory of the patch: element vectors at 0, element matrices at o_sa, y maps at o_yb, matrix maps at o_vb.
/// Every loop is written in batches of U independent iterations (all index loads, then all value loads, then
/// all stores): the chains are shared-memory-latency bound, and only a few warps work on a patch.
/// o_yb / o_vb: gather parts of the maps, o_yfold / o_vfold: fold lists (byte offsets into base).  after_fold() runs in
/// every thread right after the barrier that ends the fold phase: the fold lists are dead from there on (k_patch_ws
/// fetches those of the next patch and waits for the gather part of this one there).
struct NoAfterFold
{
   __device__ __forceinline__ void operator()() const {}
};
template <int BAR, int NT, int U, class AfterFold = NoAfterFold>
__device__ __forceinline__ void patch_drain(unsigned char *base, const int o_sa, const int o_yb, const int o_yfold, const int o_vb,
                                            const int o_vfold, const PatchDesc &D, const bool wy, const bool wv, const int tid,
                                            double *__restrict__ y, double *__restrict__ vals, double *__restrict__ ystage,
                                            double *__restrict__ vstage, const int bar_id = BAR, AfterFold &&after_fold = AfterFold())
{
   static_assert(NT % 32 == 0, "whole warps");
#define MADB_SR(i) (*(double *)(base + 8 * (i)))
#define MADB_SA(i) (*(double *)(base + o_sa + 8 * (i)))
   const int nrows = D.nrows, nrow_int = D.nrow_int, nexc = D.nexc, nslots = D.nslots;
   // ---- fold: add the further sources of every row / slot onto its first source, phase by phase -------
   // (every location is the destination or a source of exactly one entry chain)
   {
      int ybase = 8, vbase = 8;
      for (int ph = 0; ph < 8; ph++)
      {
         const int ny = wy ? *(const int *)(base + o_yfold + 4 * ph) : 0;
         const int nv = wv ? *(const int *)(base + o_vfold + 4 * ph) : 0;
         if (ny == 0 && nv == 0) { break; }
         for (int i = tid; i < ny; i += NT)
         {
            const unsigned w = *(const unsigned *)(base + o_yfold + 4 * (ybase + i));
            MADB_SR(w & 0xffffu) += MADB_SR(w >> 16);
         }
         for (int i0 = tid; i0 < nv; i0 += NT * U)
         {
            unsigned w[U];
            double d0[U], d1[U];
#pragma unroll
            for (int u = 0; u < U; u++)
            {
               const int i = i0 + u * NT;
               w[u] = *(const unsigned *)(base + o_vfold + 4 * (vbase + (i < nv ? i : i0)));
            }
#pragma unroll
            for (int u = 0; u < U; u++)
            {
               d0[u] = MADB_SA(w[u] & 0xffffu);
               d1[u] = MADB_SA(w[u] >> 16);
            }
#pragma unroll
            for (int u = 0; u < U; u++) { if (i0 + u * NT < nv) { MADB_SA(w[u] & 0xffffu) = d0[u] + d1[u]; } }
         }
         ybase += ny;
         vbase += nv;
      }
      // entry i of every phase has the same destination (pack_sources) and is handled by the same thread: one barrier
      // after the last phase is enough
      patch_bar<BAR, NT>(bar_id);
      after_fold();
   }
   // ---- rows of the residual -----------------------------------------------------------
   if (wy)
   {
      const int o_ylist = o_yb + patch_al16(2 * nrows);
      for (int lr = tid; lr < nrows; lr += NT)
      {
         const double v = MADB_SR(*(const unsigned short *)(base + o_yb + 2 * lr));
         if (lr < nrow_int) { y[*(const int *)(base + o_ylist + 4 * lr)] = v; }
         else { ystage[D.ystage_off + (lr - nrow_int)] = v; }
      }
   }
   // ---- CSR entries ------------------------------------------------------------------------
   if (wv)
   {
      constexpr int NW = NT / 32;
      const int o_plist = o_vb + patch_al16(2 * D.nvsrc);
      const int o_glist = o_plist + patch_al16(8 * (D.npair + 1));
      const int o_isrc = o_glist + patch_al16(16 * (D.ngen + 1));
      const int o_over = o_isrc + patch_al16(64 * D.nirr);
      const unsigned short *vsrc = (const unsigned short *)(base + o_vb);
      const int lane = tid & 31, warp = tid >> 5;
      // directly written slots, two work lists (layout: madb_host.hpp).  Aligned pairs of chunks first: 64 consecutive
      // CSR positions from an even one, lane l gathers the slots 2l, 2l+1 (one 32-bit load of the two source indices)
      // and writes them with one 16-byte store.  No selects, no branches: about 12 instructions per 64 entries.
      {
         constexpr int UP = (U >= 4) ? 4 : 2; // pairs per batch
         const bool al16 = (reinterpret_cast<size_t>(vals) & 15) == 0;
         const int2 *pl = (const int2 *)(base + o_plist);
         const int npair = D.npair;
         for (int k0 = warp; k0 < npair; k0 += NW * UP)
         {
            int g[UP];
            unsigned ix[UP];
            double vA[UP], vB[UP];
#pragma unroll
            for (int u = 0; u < UP; u++)
            {
               const int k = k0 + u * NW;
               const int2 e = pl[(k < npair) ? k : npair]; // past the list: the entry that stores nothing
               g[u] = e.x;
               ix[u] = *(const unsigned *)(vsrc + e.y + 2 * lane);
            }
#pragma unroll
            for (int u = 0; u < UP; u++)
            {
               vA[u] = MADB_SA(ix[u] & 0xffffu);
               vB[u] = MADB_SA(ix[u] >> 16);
            }
#pragma unroll
            for (int u = 0; u < UP; u++)
            {
               if (g[u] >= 0)
               {
                  double *dst = vals + g[u] + 2 * lane;
                  if (al16) { st_out2(dst, vA[u], vB[u]); }
                  else
                  {
                     st_out(dst, vA[u]);
                     st_out(dst + 1, vB[u]);
                  }
               }
            }
         }
      }
      // general chunks {g0, g1 - split, split, n | first slot << 8}: consecutive lanes -> consecutive positions, one break
      {
         const int4 *gl = (const int4 *)(base + o_glist);
         const int ngen = D.ngen;
         for (int k0 = warp; k0 < ngen; k0 += NW * U)
         {
            int gp[U];
            unsigned ix[U];
            double v[U];
#pragma unroll
            for (int u = 0; u < U; u++)
            {
               const int k = k0 + u * NW;
               const int4 d = gl[(k < ngen) ? k : ngen];
               ix[u] = vsrc[(d.w >> 8) + lane];
               gp[u] = (lane < (d.w & 0xff)) ? ((lane < d.z) ? d.x : d.y) + lane : -1;
            }
#pragma unroll
            for (int u = 0; u < U; u++) { v[u] = MADB_SA(ix[u]); }
#pragma unroll
            for (int u = 0; u < U; u++) { if (gp[u] >= 0) { st_out(vals + gp[u], v[u]); } }
         }
      }
         // irregular chunks: explicit positions (-1: none)
      {
         constexpr int UI = (NW >= 8) ? 1 : 8 / NW;
         const unsigned short *ip = (const unsigned short *)(base + o_isrc) + warp * 32 + lane;
         const int *op = (const int *)(base + o_over) + warp * 32 + lane;
         for (int k0 = 0; k0 < D.nirr; k0 += NW * UI)
         {
            int g[UI];
            double v[UI];
#pragma unroll
            for (int u = 0; u < UI; u++)
            {
               g[u] = op[u * NW * 32];
               v[u] = MADB_SA(ip[u * NW * 32]);
            }
#pragma unroll
            for (int u = 0; u < UI; u++) { if (g[u] >= 0) { st_out(vals + g[u], v[u]); } }
            ip += NW * UI * 32;
            op += NW * UI * 32;
         }
      }
      // entries shared with other patches: staged, contiguous
      double *stage = vstage + D.stage_off - nexc;
      for (int s0 = nexc + tid; s0 < nslots; s0 += NT * 4)
      {
         double v[4];
#pragma unroll
         for (int u = 0; u < 4; u++) { const int s = s0 + u * NT; v[u] = MADB_SA(vsrc[s < nslots ? s : s0]); }
#pragma unroll
         for (int u = 0; u < 4; u++) { const int s = s0 + u * NT; if (s < nslots) { stage[s] = v[u]; } }
      }
   }
#undef MADB_SR
#undef MADB_SA
}

/// Element computation of sorted element t (local index l in its patch) by one of NPART threads: slice PART of
/// the upper triangle (columns [B0, B1), tri_split) is computed and stored to the staging buffer; slice 0 also
/// produces the element vector.
template <class Func, class Cfg, int MODE, bool UNROLLQ, int PART, int NPART, int LD>
__device__ __forceinline__ void patch_compute_stage(const AsmArgs<Func, Cfg> &a, const Tables<Cfg> &tab, const int t, const int l,
                                                    const bool wy, const bool wv, unsigned char *smraw, const int o_sa)
{
   constexpr int NVD = Cfg::NVD, NSYM = Cfg::NSYM;
   constexpr bool HAS_V = (MODE & MODE_JAC) != 0;
   constexpr int B0 = tri_split<Cfg>(NPART, PART), B1 = tri_split<Cfg>(NPART, PART + 1);
   constexpr int PMODE = (PART == 0) ? MODE : (MODE & ~(MODE_RES | MODE_ACT));
   constexpr bool HAS_Y = (PMODE & (MODE_RES | MODE_ACT)) != 0;
   double r[HAS_Y ? NVD : 1];
   if constexpr (use_sf2d<Func, Cfg, MODE>() && HAS_V && NPART == 1)
   {
      // matrix entries stream from the computation straight into the staging buffer
      element_compute_sf2d<Func, Cfg, MODE>(a, t, r, [&](int k, double v) { if (wv) { *(double *)(smraw + o_sa + 8 * (k * LD + l)) = v; } });
   }
   else
   {
      double A[HAS_V ? NSYM : 1];
      double energy;
      element_compute<Func, Cfg, PMODE, UNROLLQ, B0, B1>(a, tab, t, r, A, energy);
      if constexpr (HAS_V)
      {
         if (wv)
         {
#pragma unroll
            for (int bb = B0; bb < B1; bb++)
            {
#pragma unroll
               for (int aa = 0; aa <= bb; aa++) { *(double *)(smraw + o_sa + 8 * (symidx(aa, bb) * LD + l)) = A[symidx(aa, bb)]; }
            }
         }
      }
   }
   if constexpr (HAS_Y)
   {
      if (wy)
      {
#pragma unroll
         for (int i = 0; i < NVD; i++) { *(double *)(smraw + 8 * (i * LD + l)) = r[i]; }
      }
   }
}

/// One CTA per patch: PE = patch_pe(NVD) elements, NPART = element_parts threads per element.
// Register cap of k_patch through the minimum CTAs per SM: kernels without the Jacobian (residual, action) are short of warps,
// not of registers (MADB_PATCH_MINB_REGS: registers per thread the cap corresponds to; 0: ptxas' default)
#ifndef MADB_PATCH_MINB_REGS
#define MADB_PATCH_MINB_REGS 0
#endif
template <int THREADS, int MODE> constexpr int patch_minb()
{
   if ((MODE & MODE_JAC) != 0 || MADB_PATCH_MINB_REGS == 0) { return 1; }
   const int b = 65536 / (MADB_PATCH_MINB_REGS * THREADS);
   return b < 1 ? 1 : (b > 16 ? 16 : b);
}
template <class Func, class Cfg, int MODE, bool UNROLLQ>
__global__ void __launch_bounds__(patch_pe_of<Func, Cfg>() * element_parts<Cfg, MODE>(),
                                  patch_minb<patch_pe_of<Func, Cfg>() * element_parts<Cfg, MODE>(), MODE>())
   k_patch(const __grid_constant__ AsmArgs<Func, Cfg> a, const __grid_constant__ PatchDev P)
{
   constexpr int NVD = Cfg::NVD, NSYM = Cfg::NSYM, PE = patch_pe_of<Func, Cfg>(), LD = PE + 1, NPART = element_parts<Cfg, MODE>();
   constexpr int NT = PE * NPART;
   constexpr bool HAS_Y = (MODE & (MODE_RES | MODE_ACT)) != 0, HAS_V = (MODE & MODE_JAC) != 0;
   constexpr int SR_BYTES = patch_al16(NVD * LD * 8), SA_BYTES = patch_al16(NSYM * LD * 8);
   extern __shared__ __align__(16) unsigned char smraw[];
   __shared__ PatchDesc D;
   __shared__ __align__(8) unsigned long long mbar;
   const int tid = threadIdx.x, p = blockIdx.x;
   // thread -> (local element l, slice part): threads [part * PE, (part + 1) * PE) compute slice `part` of the PE elements
   const int l = tid % PE, part = tid / PE;
   if (tid < (int)(sizeof(PatchDesc) / sizeof(int))) { ((int *)&D)[tid] = ((const int *)(P.desc + p))[tid]; }
   if (tid == 0) { mbar_init(&mbar, 1); }
   __syncthreads();
   pdl_wait(); // (the patch descriptor above is set-up data: no dependence on the previous kernel)
   pdl_trigger();
#if MADB_PATCH_L2PF
   // index lines (vertices, dofs, parameter dofs) of the patch that starts MADB_PATCH_L2PF waves of CTAs later: into L2 now,
   // so that the first of the two dependent loads of ITS gather is not a DRAM access (all warps of a CTA gather at the same time:
   // with one CTA per SM nothing covers that latency)
   {
      unsigned nsm;
      asm("mov.u32 %0, %%nsmid;" : "=r"(nsm));
      const int pn = p + (int)nsm * MADB_PATCH_L2PF;
      const int tn = pn * PE + l, ln = tid & 31;
      if (part == 0 && pn < (int)gridDim.x && tn < a.end && (ln == 0 || ln == 31))
      {
#pragma unroll
         for (int k = 0; k < Cfg::NGN; k++) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a.e2n + (size_t)k * a.stride + tn)); }
#pragma unroll
         for (int i = 0; i < NVD; i++) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a.vmap + (size_t)i * a.stride + tn)); }
         if constexpr (Cfg::NDOF_ALL > NVD)
         {
#pragma unroll
            for (int i = 0; i < Cfg::NDOF_ALL - NVD; i++) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a.pmap + (size_t)i * a.stride + tn)); }
         }
      }
      if (tid == 32 && pn < (int)gridDim.x) { asm volatile("prefetch.global.L2 [%0];" ::"l"(P.desc + pn)); }
   }
#endif
   const bool wy = HAS_Y && a.write_y, wv = HAS_V && a.write_vals;

   // shared-memory carve-up (byte offsets into smraw): staged element vectors | matrices | y maps | matrix maps
   const int o_sa = wy ? SR_BYTES : 0;
   const int o_yb = o_sa + (wv ? SA_BYTES : 0);
   const int o_vb = o_yb + (wy ? P.max_yblob : 0);
   if (tid == 0)
   {
      const unsigned bytes = (wy ? D.yblob_bytes : 0) + (wv ? D.vblob_bytes : 0);
      mbar_expect_tx(&mbar, bytes);
      if (wy && D.yblob_bytes) { bulk_g2s(smraw + o_yb, P.yblob + (size_t)D.yblob_off * 16, D.yblob_bytes, &mbar); }
      if (wv && D.vblob_bytes) { bulk_g2s(smraw + o_vb, P.vblob + (size_t)D.vblob_off * 16, D.vblob_bytes, &mbar); }
   }
   // When the quadrature loop is not unrolled the basis tables are indexed dynamically: constant-bank loads (LDC)
   // of warps running different slices thrash the constant cache (measured: 2x on the ex4 block), so the tables
   // are copied to shared memory once per CTA.
   constexpr bool TAB_SMEM = !UNROLLQ && !use_sf2d<Func, Cfg, MODE>();
   const Tables<Cfg> *tabp = &a.tab;
   if constexpr (TAB_SMEM)
   {
      double *dst = (double *)(smraw + o_vb + (wv ? P.max_vblob : 0));
      const double *src = (const double *)&a.tab;
      for (int k = tid; k < (int)(sizeof(Tables<Cfg>) / sizeof(double)); k += NT) { dst[k] = src[k]; }
      tabp = (const Tables<Cfg> *)dst;
      __syncthreads();
   }
   const Tables<Cfg> &tab = *tabp;

   const int t = p * PE + l;
   if (l < D.ne)
   {
      if constexpr (NPART == 1) { patch_compute_stage<Func, Cfg, MODE, UNROLLQ, 0, 1, LD>(a, tab, t, l, wy, wv, smraw, o_sa); }
      else
      {
         static_assert(NPART == 4, "slices are dispatched on the warp index");
         switch (part)
         {
            case 0: patch_compute_stage<Func, Cfg, MODE, UNROLLQ, 0, 4, LD>(a, tab, t, l, wy, wv, smraw, o_sa); break;
            case 1: patch_compute_stage<Func, Cfg, MODE, UNROLLQ, 1, 4, LD>(a, tab, t, l, wy, wv, smraw, o_sa); break;
            case 2: patch_compute_stage<Func, Cfg, MODE, UNROLLQ, 2, 4, LD>(a, tab, t, l, wy, wv, smraw, o_sa); break;
            default: patch_compute_stage<Func, Cfg, MODE, UNROLLQ, 3, 4, LD>(a, tab, t, l, wy, wv, smraw, o_sa); break;
         }
      }
   }
   __syncthreads();
   mbar_wait(&mbar, 0);
   constexpr int NW = NT / 32, U = (32 / NW < 8) ? 32 / NW : 8;
   patch_drain<0, NT, U>(smraw, o_sa, o_yb, o_yb + patch_yg_bytes(D), o_vb, o_vb + patch_vg_bytes(D), D, wy, wv, tid, a.y, a.vals,
                         P.ystage, P.vstage);
}

// loads of the prefetch: volatile asm keeps them where they are written (between two blocks of the matrix phase)
__device__ __forceinline__ int ldg_nc_s32(const int *p)
{
   int v;
   asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
   return v;
}
__device__ __forceinline__ double ldg_nc_f64(const double *p)
{
   double v;
   asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
   return v;
}
__device__ __forceinline__ void ldg_nc_f64x2(const double *p, double &v0, double &v1)
{
   asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "l"(p));
}
#ifndef MADB_WS_JOINT
#define MADB_WS_JOINT 0 // 1: both writer warpgroups drain one buffer at a time; 0: writer warpgroup w serves compute warpgroup w
                        // (measured on config 2, profiles/r02_k_patch_ws.md: 0.309 ms joint, 0.269 ms separate)
#endif
#ifndef MADB_WS_PREFETCH
#define MADB_WS_PREFETCH 0 // register prefetch of the next element's inputs during the matrix phase: the compute warpgroups alone
                           // gain 10 % (0.223 -> 0.200 ms), the full kernel loses (0.269 -> 0.321 ms): ptxas spills the
                           // prefetched values right after the loads, which stalls the warp until they land
#endif
#ifndef MADB_WS_DRAIN_U
#define MADB_WS_DRAIN_U 4 // independent entries per thread and batch in the drain loops of the writer warpgroups
#endif
#ifndef MADB_WS_BLOCK_BB
#define MADB_WS_BLOCK_BB 1 // bit 0: basic-block boundary between the (i1, j1) blocks of the matrix phase (spills 120 -> 16 B, element
                           // kernel 0.270 -> 0.248 ms); bit 1: after every row of points (more spills: 0.283 ms); bit 2: between the
                           // two stages of a block (8 B of spills, no further gain)
#endif
#ifndef MADB_WS_STAGE_U
#define MADB_WS_STAGE_U 0 // 1: the writer warpgroup gathers the dof values of its compute warpgroup's next patch into shared
                          // memory while it waits for `full`.  Measured (config 2): 0.298 ms against 0.266 ms without: the
                          // stalls of the gather prologue are not lost time (the other compute warpgroup has the FP64 pipe
                          // to itself meanwhile), the gather on the writers' serial path is
#endif
#ifndef MADB_WS_PREFETCH_AT
#define MADB_WS_PREFETCH_AT 2 // block of the matrix phase (0..5 for order 2) before which the values are requested
#endif
struct Sf2dDummyCfg
{
   template <int> struct field
   {
      static constexpr int ND1D = 1;
   };
};

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Warp-specialised persistent variant (fused residual + Jacobian): one CTA per SM, four warpgroups.
//   warpgroups 2,3 (208 registers/thread after setmaxnreg): element computation, patch after patch;
//                  each stores its element vectors (before the matrix phase) and streams its element
//                  matrices into its own shared-memory buffer
//   warpgroups 0,1 (40 registers/thread): writer warpgroup w folds, gathers and writes the rows of the
//                  finished patches of compute warpgroup w while that one is already working on its next
//                  patch (independent drains, own named barrier); it also prefetches the gather maps of the
//                  next patch with cp.async.bulk.
// Hand-off through mbarriers: full[w] (compute -> writer), empty[w] (writer -> compute),
// blob[w] (bulk-copy completion).  Patches are dealt round-robin: p = (it * gridDim + cta) * 2 + w.
// ---------------------------------------------------------------------------------------------
constexpr int WS_WRITER_WG = 2;                              // writer warpgroups: one per compute warpgroup
constexpr int WS_THREADS = (WS_WRITER_WG + 2) * PATCH_PE;
// Register split after setmaxnreg.  setmaxnreg only redistributes the CTA's own launch allocation (512 x 128 = 64 K
// registers here) and the total must stay BELOW it: an exact fit deadlocks the increase.
#ifndef MADB_WS_REGC
#define MADB_WS_REGC 208
#define MADB_WS_REGW 40
#endif
constexpr int WS_REG_COMPUTE = MADB_WS_REGC, WS_REG_WRITER = MADB_WS_REGW;
static_assert(2 * PATCH_PE * WS_REG_COMPUTE + WS_WRITER_WG * PATCH_PE * WS_REG_WRITER <= 65536 - 1024, "setmaxnreg budget");

template <class Func, class Cfg, bool UNROLLQ>
__global__ void __launch_bounds__(WS_THREADS, 1) k_patch_ws(const __grid_constant__ AsmArgs<Func, Cfg> a,
                                                              const __grid_constant__ PatchDev P)
{
   constexpr int MODE = MODE_RES | MODE_JAC;
   constexpr int NVD = Cfg::NVD, NSYM = Cfg::NSYM, PE = PATCH_PE, LD = PATCH_LD;
   constexpr int SR_BYTES = patch_al16(NVD * LD * 8), SA_BYTES = patch_al16(NSYM * LD * 8);
   constexpr int WG_C0 = WS_WRITER_WG; // first compute warpgroup
   static_assert(WS_WRITER_WG == 2, "writer warpgroup w serves compute warpgroup w");
   extern __shared__ __align__(16) unsigned char smraw[];
   __shared__ __align__(8) unsigned long long bar_full[2], bar_empty[2], bar_blobF[2], bar_blobG[2], bar_in_full[2], bar_in_empty[2];
   __shared__ __align__(16) PatchDesc Dd[2][2]; // [warpgroup][patch parity]
   // warpgroup index, made warp-uniform for the compiler (uniform registers instead of spilled vector registers)
   const int wg = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 7), 0), tid = threadIdx.x & (PE - 1);
   const bool wy = a.write_y != 0;
   constexpr bool STAGE_U = (MADB_WS_STAGE_U != 0) && use_sf2d<Func, Cfg, MODE>() && !(MADB_WS_PREFETCH) && !(MADB_WS_JOINT);
   constexpr int U_BYTES = STAGE_U ? patch_al16(NVD * PE * 8) : 0; // staged dof values [NVD][PE]
   const int o_u = SR_BYTES + SA_BYTES + P.max_yg + P.max_yf + P.max_vg + P.max_vf;
   // MADB_WS_L2PF == 2: the 4 vertex + NVD dof indices of every element of the NEXT patch, staged by cp.async ([k][element])
   constexpr int IDX_BYTES = (MADB_WS_L2PF == 2 && use_sf2d<Func, Cfg, MODE>()) ? patch_al16((4 + NVD) * PE * 4) : 0;
   const int o_idx = o_u + U_BYTES;
   const int wg_bytes = o_idx + IDX_BYTES;
   if (threadIdx.x == 0)
   {
      for (int k = 0; k < 2; k++)
      {
         mbar_init(&bar_full[k], PE);
         mbar_init(&bar_empty[k], MADB_WS_JOINT ? 2 * PE : PE);
         mbar_init(&bar_blobF[k], 1);
         mbar_init(&bar_blobG[k], 1);
         mbar_init(&bar_in_full[k], PE);
         mbar_init(&bar_in_empty[k], PE);
      }
   }
   __syncthreads();
   pdl_wait(); // the previous kernel of the stream (the interface reduction of the previous step) is complete
   pdl_trigger();

   if (wg >= WG_C0)
   {
      // ================= compute warpgroups =================
      asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WS_REG_COMPUTE));
      const int w = wg - WG_C0;
      unsigned char *base = smraw + (size_t)w * wg_bytes;
      Sf2dIn<typename std::conditional<use_sf2d<Func, Cfg, MODE>(), Cfg, Sf2dDummyCfg>::type, false> sf_in;
      for (int it = 0;; it++)
      {
         const int p = (it * (int)gridDim.x + (int)blockIdx.x) * 2 + w;
         if (p >= P.npatch) { break; }
         const int t = p * PE + tid;
         const bool valid = t < a.end;
         double r[NVD];
         const unsigned par = (it & 1) ^ 1; // parity of "the writer has drained the previous patch of this buffer"
         if constexpr (use_sf2d<Func, Cfg, MODE>())
         {
#if MADB_WS_PREFETCH
            // The inputs of this element were loaded during the matrix phase of the previous patch (registers): the two
            // dependent loads (element maps -> dof values / vertex coordinates) are off the critical path.  The maps of the
            // next element are read at the end of the quadrature loop, the values between two blocks of the matrix phase.
            constexpr int ND1 = Cfg::template field<0>::ND1D;
            static_assert(NVD == ND1 * ND1, "scalar field");
            const int tc = valid ? t : a.end - 1;
            if (it == 0) { sf2d_gather<Func, Cfg, false>(a, tc, sf_in); }
            const Sf2dIn<Cfg, false> in = sf_in;
            const int pn = ((it + 1) * (int)gridDim.x + (int)blockIdx.x) * 2 + w;
            const int tn0 = pn * PE + tid, tn = (pn < P.npatch && tn0 < a.end) ? tn0 : tc; // past the end: reload this one
            int nidx[4 + NVD];
            auto load_maps = [&]()
            {
#pragma unroll
               for (int k = 0; k < 4; k++) { nidx[k] = ldg_nc_s32(a.e2n + (size_t)k * a.stride + tn); }
#pragma unroll
               for (int i = 0; i < NVD; i++) { nidx[4 + i] = ldg_nc_s32(a.vmap + (size_t)i * a.stride + tn) & 0x7fffffff; }
               bb_break(a.stride);
            };
            auto load_values = [&](int blk)
            {
               if (blk > 0 && blk < 100 && blk != MADB_WS_PREFETCH_AT && (MADB_WS_BLOCK_BB & 1)) { bb_break(a.stride); }
               if (blk != MADB_WS_PREFETCH_AT) { return; } // (codes >= 100 are the other hook points)
               bb_break(a.stride);
#pragma unroll
               for (int k = 0; k < 4; k++) { ldg_nc_f64x2(a.coords + (size_t)nidx[k] * 2, sf_in.X[k][0], sf_in.X[k][1]); }
#pragma unroll
               for (int i = 0; i < NVD; i++) { sf_in.u[i / ND1][i % ND1] = ldg_nc_f64(a.x + nidx[4 + i]); }
               bb_break(a.stride);
            };
            if (valid && !(P.diag & 2))
            {
               element_compute_sf2d_core<Func, Cfg, MODE>(
                  a, in, r, [&](int k, double v) { *(double *)(base + SR_BYTES + 8 * (k * LD + tid)) = v; },
                  [&]()
                  {
                     mbar_wait(&bar_empty[w], par);
#pragma unroll
                     for (int i = 0; i < NVD; i++) { *(double *)(base + 8 * (i * LD + tid)) = r[i]; }
                     load_maps();
                  },
                  load_values);
            }
            else
            {
               mbar_wait(&bar_empty[w], par);
               load_maps();
               load_values(MADB_WS_PREFETCH_AT);
            }
#else
            if constexpr (STAGE_U)
            {
               // dof values of this patch: gathered by the writer warpgroup into shared memory ([i][element])
               constexpr int ND1 = Cfg::template field<0>::ND1D;
               mbar_wait(&bar_in_full[w], it & 1);
#pragma unroll
               for (int i = 0; i < NVD; i++) { sf_in.u[i / ND1][i % ND1] = *(const double *)(base + o_u + 8 * (i * PE + tid)); }
               mbar_arrive(&bar_in_empty[w]);
               if (valid) { sf2d_gather_x<Func, Cfg, false>(a, t, sf_in); }
            }
            else
            {
#if MADB_WS_L2PF == 2
               // The indices of this thread's element were copied to shared memory during the previous patch (cp.async: no
               // registers held, nobody waits): the gather is ONE round trip (values) instead of two dependent ones.  Every
               // thread reads only the slots it copied itself (cp.async.wait_group): no barrier.
               {
                  int *ix = (int *)(base + o_idx);
                  auto stage_idx = [&](const int pn)
                  {
                     const int tn = pn * PE + tid;
                     if (pn < P.npatch && tn < a.end)
                     {
#pragma unroll
                        for (int k = 0; k < 4; k++)
                        {
                           asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(ix + k * PE + tid)), "l"(a.e2n + (size_t)k * a.stride + tn) : "memory");
                        }
#pragma unroll
                        for (int i = 0; i < NVD; i++)
                        {
                           asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(ix + (4 + i) * PE + tid)), "l"(a.vmap + (size_t)i * a.stride + tn) : "memory");
                        }
                     }
                     asm volatile("cp.async.commit_group;" ::: "memory");
                  };
                  if (it == 0) { stage_idx(p); }
                  asm volatile("cp.async.wait_group 0;" ::: "memory");
                  if (valid)
                  {
                     constexpr int ND1 = Cfg::template field<0>::ND1D;
                     int nn[4], nd[NVD];
#pragma unroll
                     for (int k = 0; k < 4; k++) { nn[k] = ix[k * PE + tid]; }
#pragma unroll
                     for (int i = 0; i < NVD; i++) { nd[i] = ix[(4 + i) * PE + tid] & 0x7fffffff; }
#pragma unroll
                     for (int k = 0; k < 4; k++)
                     {
                        sf_in.X[k][0] = a.coords[(size_t)nn[k] * 2];
                        sf_in.X[k][1] = a.coords[(size_t)nn[k] * 2 + 1];
                     }
#pragma unroll
                     for (int i = 0; i < NVD; i++) { sf_in.u[i / ND1][i % ND1] = a.x[nd[i]]; }
                  }
                  stage_idx(((it + 1) * (int)gridDim.x + (int)blockIdx.x) * 2 + w);
               }
#elif MADB_WS_L2PF
               // the element -> vertex / dof index lines of this warp's elements in the NEXT patch: into L2 now, so that the
               // first of the two dependent loads of the next gather does not go to DRAM (no registers held)
               {
                  const int pn = ((it + 1) * (int)gridDim.x + (int)blockIdx.x) * 2 + w;
                  const int tn = pn * PE + tid;
                  const int ln = tid & 31;
                  if (pn < P.npatch && tn < a.end && (ln == 0 || ln == 31))
                  {
#pragma unroll
                     for (int k = 0; k < 4; k++) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a.e2n + (size_t)k * a.stride + tn)); }
#pragma unroll
                     for (int i = 0; i < NVD; i++) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a.vmap + (size_t)i * a.stride + tn)); }
                  }
               }
#endif
#if MADB_WS_L2PF != 2
               if (valid) { sf2d_gather<Func, Cfg, false>(a, t, sf_in); }
#endif
            }
            if (valid && !(P.diag & 2))
            {
               element_compute_sf2d_core<Func, Cfg, MODE>(
                  a, sf_in, r, [&](int k, double v) { *(double *)(base + SR_BYTES + 8 * (k * LD + tid)) = v; },
                  [&]()
                  {
                     // the element vector is final: wait for the buffer and store it before the matrix phase
                     // (it would be spilled across it otherwise)
                     mbar_wait(&bar_empty[w], par);
#pragma unroll
                     for (int i = 0; i < NVD; i++) { *(double *)(base + 8 * (i * LD + tid)) = r[i]; }
                  }
#if MADB_WS_BLOCK_BB
                  ,
                  // basic-block boundary between the (i1, j1) blocks of the matrix phase: ptxas then schedules every block
                  // on its own and does not hoist the constant loads of later blocks (uniform-register pressure)
                  [&](int code)
                  {
                     if (code < 100) { if (code > 0 && (MADB_WS_BLOCK_BB & 1)) { bb_break(a.stride); } }
                     else if (code < 200) { if (MADB_WS_BLOCK_BB & 2) { bb_break(a.stride); } }
                     else { if (MADB_WS_BLOCK_BB & 4) { bb_break(a.stride); } }
                  }
#endif
               );
            }
            // threads past the end of the last patch take part in the hand-off like the others: without this wait their
            // arrival on `full` would be counted in the previous, still open phase and the writers could start early
            else { mbar_wait(&bar_empty[w], par); }
#endif
         }
         else
         {
            double A[NSYM], energy;
#if MADB_WS_L2PF_GEN
            // index lines of this warp's elements in the next patch into L2 (as in the sum-factorised branch)
            {
               const int pn = ((it + 1) * (int)gridDim.x + (int)blockIdx.x) * 2 + w;
               const int tn = pn * PE + tid;
               const int ln = tid & 31;
               if (pn < P.npatch && tn < a.end && (ln == 0 || ln == 31))
               {
#pragma unroll
                  for (int k = 0; k < Cfg::NGN; k++) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a.e2n + (size_t)k * a.stride + tn)); }
#pragma unroll
                  for (int i = 0; i < NVD; i++) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a.vmap + (size_t)i * a.stride + tn)); }
                  if constexpr (Cfg::NDOF_ALL > NVD)
                  {
#pragma unroll
                     for (int i = 0; i < Cfg::NDOF_ALL - NVD; i++) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a.pmap + (size_t)i * a.stride + tn)); }
                  }
               }
            }
#endif
            if (valid) { element_compute<Func, Cfg, MODE, UNROLLQ>(a, a.tab, t, r, A, energy); }
            mbar_wait(&bar_empty[w], par);
            if (valid)
            {
#pragma unroll
               for (int k = 0; k < NSYM; k++) { *(double *)(base + SR_BYTES + 8 * (k * LD + tid)) = A[k]; }
#pragma unroll
               for (int i = 0; i < NVD; i++) { *(double *)(base + 8 * (i * LD + tid)) = r[i]; }
            }
         }
         mbar_arrive(&bar_full[w]);
      }
   }
   else
   {
      // ================= writer warpgroup(s) =================
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_REG_WRITER));
      // writer warpgroup ww drains the patches of compute warpgroup ww (own named barrier): the two write-outs run
      // concurrently and the compute warpgroups are not coupled through a common drain order.
      // Maps of a patch = descriptor (80 bytes) + fold lists F + gather part G (madb_host.hpp).  Measured (MADB_DIAG=3:
      // hand-offs and map traffic only): fetching them after the drain, descriptor then maps (two dependent DRAM round
      // trips), costs 3.2 us per patch on the writers' serial path, a third of their cycle.  Hence: the descriptor of the
      // next patch is requested at the top of a drain (cp.async, lands in the other slot of Dd), its fold lists as soon
      // as the fold of the current patch is done (their region is dead then), its gather part at the end of the drain;
      // the writers need the gather part only after their fold phase, which covers most of its latency.
      // MADB_WS_JOINT: both writer warpgroups (8 warps) drain one buffer at a time, alternating between the two compute
      // warpgroups (a drain is bound by the latency of its dependent shared-memory accesses per warp, so twice the warps
      // halve it and the chain "matrix phase -> drain -> matrix phase" of a buffer gets shorter); the maps of a buffer's
      // next patch land while the other buffer is drained.
      constexpr bool JOINT = MADB_WS_JOINT != 0;
      constexpr int WNT = JOINT ? 2 * PE : PE;
      const int wtid = JOINT ? (int)threadIdx.x : tid; // the writer warpgroups are warpgroups 0 and 1
      const int bar_id = JOINT ? 1 : 1 + wg;
      const int o_sa = SR_BYTES, o_yb = SR_BYTES + SA_BYTES, o_yf = o_yb + P.max_yg, o_vb = o_yf + P.max_yf, o_vf = o_vb + P.max_vg;
      auto fetch_F = [&](const int w, const PatchDesc &Dn)
      {
         unsigned char *base = smraw + (size_t)w * wg_bytes;
         const int yf = wy ? patch_al16(4 * Dn.nyfold) : 0, vf = patch_al16(4 * Dn.nvfold);
         asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
         mbar_expect_tx(&bar_blobF[w], (unsigned)(yf + vf));
         if (yf) { bulk_g2s(base + o_yf, P.yblob + (size_t)Dn.yblob_off * 16 + patch_yg_bytes(Dn), yf, &bar_blobF[w]); }
         if (vf) { bulk_g2s(base + o_vf, P.vblob + (size_t)Dn.vblob_off * 16 + patch_vg_bytes(Dn), vf, &bar_blobF[w]); }
      };
      auto fetch_G = [&](const int w, const PatchDesc &Dn)
      {
         unsigned char *base = smraw + (size_t)w * wg_bytes;
         const int yg = wy ? patch_yg_bytes(Dn) : 0, vg = patch_vg_bytes(Dn);
         asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
         mbar_expect_tx(&bar_blobG[w], (unsigned)(yg + vg));
         if (yg) { bulk_g2s(base + o_yb, P.yblob + (size_t)Dn.yblob_off * 16, yg, &bar_blobG[w]); }
         if (vg) { bulk_g2s(base + o_vb, P.vblob + (size_t)Dn.vblob_off * 16, vg, &bar_blobG[w]); }
      };
      // dof values of patch pp of compute warpgroup w -> shared memory (thread = element), then `in_full`
      auto stage_u = [&](const int w, const int pp)
      {
         if constexpr (STAGE_U)
         {
            unsigned char *ub = smraw + (size_t)w * wg_bytes + o_u;
            const int t = pp * PE + wtid;
            const bool live = t < a.end;
            int idx[NVD];
#pragma unroll
            for (int i = 0; i < NVD; i++) { idx[i] = live ? (__ldg(a.vmap + (size_t)i * a.stride + t) & 0x7fffffff) : 0; }
#pragma unroll
            for (int i = 0; i < NVD; i++) { *(double *)(ub + 8 * (i * PE + wtid)) = __ldg(a.x + idx[i]); }
            mbar_arrive(&bar_in_full[w]);
         }
      };
      if constexpr (STAGE_U)
      {
         const int p0 = (int)blockIdx.x * 2 + wg;
         if (p0 < P.npatch) { stage_u(wg, p0); }
      }
      if (wtid == 0)
      {
         for (int w = JOINT ? 0 : wg; w < (JOINT ? 2 : wg + 1); w++)
         {
            const int p0 = (int)blockIdx.x * 2 + w;
            if (p0 >= P.npatch) { break; }
            for (int k = 0; k < (int)(sizeof(PatchDesc) / sizeof(int)); k++) { ((int *)&Dd[w][0])[k] = __ldg((const int *)(P.desc + p0) + k); }
            fetch_F(w, Dd[w][0]);
            fetch_G(w, Dd[w][0]);
         }
      }
      patch_bar<1, WNT>(bar_id); // the first descriptors are visible to the writers
      for (int s = 0;; s++)
      {
         const int w = JOINT ? (s & 1) : wg, it = JOINT ? (s >> 1) : s;
         const int p = (it * (int)gridDim.x + (int)blockIdx.x) * 2 + w;
         if (p >= P.npatch) { break; } // p grows with s
         const int pn = ((it + 1) * (int)gridDim.x + (int)blockIdx.x) * 2 + w;
         const bool more = pn < P.npatch;
         unsigned char *base = smraw + (size_t)w * wg_bytes;
         const PatchDesc &D = Dd[w][it & 1];
         PatchDesc &Dn = Dd[w][(it & 1) ^ 1];
         if (wtid == 0 && more)
         {
            static_assert(sizeof(PatchDesc) % 16 == 0, "descriptor copied in 16-byte pieces");
#pragma unroll
            for (int k = 0; k < (int)sizeof(PatchDesc) / 16; k++)
            {
               asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32((unsigned char *)&Dn + 16 * k)),
                            "l"((const unsigned char *)(P.desc + pn) + 16 * k)
                            : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
         }
         if constexpr (STAGE_U)
         {
            if (more)
            {
               mbar_wait(&bar_in_empty[w], it & 1); // the compute warpgroup has read the values of patch `it`
               stage_u(w, pn);
            }
         }
#if MADB_WS_L1PF
         // While the compute warpgroup finishes patch `it`, warm this SM's L1 with the vertex coordinates and dof values its
         // next gather (patch pn) will load: the index lines are in L2 already (MADB_WS_L2PF), the writers wait here anyway.
         if constexpr (use_sf2d<Func, Cfg, MODE>() && !JOINT)
         {
            const int tn = pn * PE + wtid;
            if (more && tn < a.end)
            {
               int n4[4];
#pragma unroll
               for (int k = 0; k < 4; k++) { n4[k] = __ldg(a.e2n + (size_t)k * a.stride + tn); }
#pragma unroll
               for (int k = 0; k < 4; k++) { l1_touch(a.coords + (size_t)n4[k] * 2); }
#pragma unroll
               for (int i0 = 0; i0 < NVD; i0 += 5)
               {
                  int d5[5];
#pragma unroll
                  for (int i = 0; i < 5; i++) { if (i0 + i < NVD) { d5[i] = __ldg(a.vmap + (size_t)(i0 + i) * a.stride + tn) & 0x7fffffff; } }
#pragma unroll
                  for (int i = 0; i < 5; i++) { if (i0 + i < NVD) { l1_touch(a.x + d5[i]); } }
               }
            }
         }
#endif
         mbar_wait(&bar_blobF[w], it & 1); // fold lists of this patch
         mbar_wait(&bar_full[w], it & 1);  // the compute warpgroup has staged the patch
         auto after_fold = [&]()
         {
            if (wtid == 0 && more)
            {
               asm volatile("cp.async.wait_group 0;" ::: "memory");
               fetch_F(w, Dn);
            }
            mbar_wait(&bar_blobG[w], it & 1); // gather part of this patch
         };
         if (!(P.diag & 1))
         {
            patch_drain<1, WNT, MADB_WS_DRAIN_U>(base, o_sa, o_yb, o_yf, o_vb, o_vf, D, wy, true, wtid, a.y, a.vals, P.ystage, P.vstage, bar_id, after_fold);
         }
         else { after_fold(); }
         mbar_arrive(&bar_empty[w]);
         // all writers are done with the gather part (and thread 0's copy of the next descriptor is visible to them
         // after this barrier): fetch the gather part of the buffer's next patch
         patch_bar<1, WNT>(bar_id);
         if (wtid == 0 && more) { fetch_G(w, Dn); }
      }
   }
}

// Interface reduction: out[dst[i]] = sum of the staged partials of entry i, in ascending patch order.
// Entries with at most 4 sources (nearly all: 2 for an edge between two patches) are stored packed
// {src0..src3 (-1: none), dst} so that all loads of an entry are independent; the rest as (ptr, src, dst) lists.
// One launch covers the residual rows and the CSR entries: blocks [0,nb0) list A packed, [nb0,nb1) A general,
// [nb1,nb2) B packed, [nb2,nb3) B general.
using IfcList = IfcListDev;
#ifndef MADB_IFC_U
#define MADB_IFC_U 4
#endif
constexpr int IFC_U = MADB_IFC_U; // packed entries per thread: all loads of the batch are independent (the kernel is latency bound)
__device__ __forceinline__ void ifc_reduce_packed(const IfcList &L, const int i0)
{
   int4 s[IFC_U];
   int d[IFC_U];
   double a[IFC_U][4];
#pragma unroll
   for (int u = 0; u < IFC_U; u++)
   {
      const int i = i0 + u * 256;
      const bool live = i < L.n4;
      s[u] = live ? __ldg(L.src4 + i) : make_int4(-1, -1, -1, -1);
      d[u] = live ? __ldg(L.dst4 + i) : -1;
   }
#pragma unroll
   for (int u = 0; u < IFC_U; u++)
   {
      a[u][0] = (s[u].x >= 0) ? L.stage[s[u].x] : 0.0;
      a[u][1] = (s[u].y >= 0) ? L.stage[s[u].y] : 0.0;
      a[u][2] = (s[u].z >= 0) ? L.stage[s[u].z] : 0.0;
      a[u][3] = (s[u].w >= 0) ? L.stage[s[u].w] : 0.0;
   }
#pragma unroll
   for (int u = 0; u < IFC_U; u++)
   {
      if (d[u] < 0) { continue; }
      double v = a[u][0]; // ascending patch order, exactly the sources that exist (x + 0.0 is not always x: -0.0)
      if (s[u].y >= 0) { v += a[u][1]; }
      if (s[u].z >= 0) { v += a[u][2]; }
      if (s[u].w >= 0) { v += a[u][3]; }
      L.out[d[u]] = v;
   }
}
__device__ __forceinline__ void ifc_reduce_general(const IfcList &L, const int i)
{
   if (i >= L.ng) { return; }
   const int b = L.ptr[i], e = L.ptr[i + 1];
   double s = L.stage[L.src[b]];
   for (int k = b + 1; k < e; k++) { s += L.stage[L.src[k]]; }
   L.out[L.dst[i]] = s;
}
static __global__ void __launch_bounds__(256) k_ifc_reduce(const IfcList A, const IfcList B, const int nb0, const int nb1,
                                                           const int nb2)
{
   pdl_wait();
   pdl_trigger();
   const int blk = blockIdx.x, t = threadIdx.x;
   if (blk < nb0) { ifc_reduce_packed(A, blk * 256 * IFC_U + t); }
   else if (blk < nb1) { ifc_reduce_general(A, (blk - nb0) * 256 + t); }
   else if (blk < nb2) { ifc_reduce_packed(B, (blk - nb1) * 256 * IFC_U + t); }
   else { ifc_reduce_general(B, (blk - nb2) * 256 + t); }
}

/// does <functional, configuration> have the CSR-image kernel (madb_patch_img.cuh)? (element matrices of up to 10 dofs)
/// (one thread per element: up to 8 dofs, so that two work groups of 128 elements fit in shared memory; the thread-pair
/// variant of the sum-factorised 2-D path runs 64 elements per work group)
template <class Func, class Cfg> constexpr bool img_eligible() { return sf2d_pair_ok<Func, Cfg>() || Cfg::NVD <= 8; }
template <class Func, class Cfg, bool UNROLLQ> int launch_patch_img(const AsmArgs<Func, Cfg> &a, const LaunchCtx &L); // madb_patch_img.cuh

template <class Func, class Cfg, int MODE, bool UNROLLQ>
int launch_patch_mode(const AsmArgs<Func, Cfg> &a, const LaunchCtx &L)
{
   const PatchDev &P = *L.patch;
   // dynamic shared-memory limits are per device and per kernel: remember what was set on each device
   // (several host threads may drive contexts on different devices: the bookkeeping is guarded)
   static int smem_set_dev[64] = {0}, ws_smem_set_dev[64] = {0}, nsm_dev[64] = {0};
   static std::mutex attr_mutex;
   std::lock_guard<std::mutex> attr_lock(attr_mutex);
   int dev = 0;
   cudaGetDevice(&dev);
   dev &= 63;
   int &smem_set = smem_set_dev[dev], &ws_smem_set = ws_smem_set_dev[dev], &nsm = nsm_dev[dev];
   auto kern = k_patch<Func, Cfg, MODE, UNROLLQ>;
   const bool wy = (MODE & (MODE_RES | MODE_ACT)) && L.write_y, wv = (MODE & MODE_JAC) && L.write_vals;
   constexpr int PE = patch_pe_of<Func, Cfg>(), LD = PE + 1, NPART = element_parts<Cfg, MODE>();
   constexpr bool TAB_SMEM = !UNROLLQ && !use_sf2d<Func, Cfg, MODE>();
   const int smem_bytes = (wy ? patch_al16(Cfg::NVD * LD * 8) + P.max_yblob : 0) +
                          (wv ? patch_al16(Cfg::NSYM * LD * 8) + P.max_vblob : 0) + (TAB_SMEM ? (int)sizeof(Tables<Cfg>) : 0) + 16;
   if (smem_bytes > 224 * 1024) { return (int)cudaErrorInvalidConfiguration; }
   if (smem_bytes > smem_set)
   {
      const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
      if (e != cudaSuccess) { return (int)e; }
      smem_set = smem_bytes;
   }
   if (L.ev0) { cudaEventRecord(L.ev0, L.stream); }
   bool done = false;
   if constexpr (MODE == (MODE_RES | MODE_JAC) && img_eligible<Func, Cfg>())
   {
      if (wv && P.img.desc)
      {
         const int rc = launch_patch_img<Func, Cfg, UNROLLQ>(a, L);
         if (rc != 0) { return rc; }
         done = true;
      }
   }
   if constexpr (MODE == (MODE_RES | MODE_JAC) && PE == PATCH_PE && NPART == 1)
   {
      static const bool use_ws = getenv("MADB_PATCH_WS") ? atoi(getenv("MADB_PATCH_WS")) != 0 : true;
      if (!done && wv && use_ws && P.vblob)
      {
         auto kws = k_patch_ws<Func, Cfg, UNROLLQ>;
         constexpr bool STAGE_U = (MADB_WS_STAGE_U != 0) && use_sf2d<Func, Cfg, MODE>() && !(MADB_WS_PREFETCH) && !(MADB_WS_JOINT);
         const int ws_bytes = 2 * (patch_al16(Cfg::NVD * PATCH_LD * 8) + patch_al16(Cfg::NSYM * PATCH_LD * 8) + P.max_yg + P.max_yf + P.max_vg + P.max_vf +
                                   (STAGE_U ? patch_al16(Cfg::NVD * PATCH_PE * 8) : 0) +
                                   ((MADB_WS_L2PF == 2 && use_sf2d<Func, Cfg, MODE>()) ? patch_al16((4 + Cfg::NVD) * PATCH_PE * 4) : 0)) + 16;
         if (nsm == 0) { cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev); }
         if (ws_bytes <= 226 * 1024)
         {
            if (ws_bytes > ws_smem_set)
            {
               const cudaError_t e = cudaFuncSetAttribute(kws, cudaFuncAttributeMaxDynamicSharedMemorySize, ws_bytes);
               if (e != cudaSuccess) { return (int)e; }
               ws_smem_set = ws_bytes;
            }
            const int grid = std::min(nsm, (P.npatch + 1) / 2);
            static const int diag = getenv("MADB_DIAG") ? atoi(getenv("MADB_DIAG")) : 0;
            PatchDev Pd = P;
            Pd.diag = diag;
            launch_pdl(kws, grid, WS_THREADS, (size_t)ws_bytes, L.stream, a, Pd);
            done = true;
         }
      }
   }
   if (!done)
   {
      if (wv && !P.vblob) { return (int)cudaErrorInvalidValue; } // the gather maps of this path were not built
      launch_pdl(kern, P.npatch, PE * NPART, (size_t)smem_bytes, L.stream, a, P);
   }
   (void)ws_smem_set;
   (void)nsm;
   if (L.ev1) { cudaEventRecord(L.ev1, L.stream); }
   {
      IfcList ly = P.ylist, lv = P.vlist;
      ly.stage = P.ystage; ly.out = L.y;
      lv.stage = P.vstage; lv.out = L.vals;
      if (!wy) { ly.n4 = ly.ng = 0; }
      if (!wv || L.defer_v_ifc) { lv.n4 = lv.ng = 0; }
      const int nb0 = (ly.n4 + 256 * IFC_U - 1) / (256 * IFC_U), nb1 = nb0 + (ly.ng + 255) / 256,
                nb2 = nb1 + (lv.n4 + 256 * IFC_U - 1) / (256 * IFC_U), nb3 = nb2 + (lv.ng + 255) / 256;
      if (nb3 > 0) { launch_pdl(k_ifc_reduce, nb3, 256, 0, L.stream, ly, lv, nb0, nb1, nb2); }
   }
   return (int)cudaGetLastError();
}

} // namespace madb

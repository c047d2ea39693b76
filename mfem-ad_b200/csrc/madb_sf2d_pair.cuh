// madb_sf2d_pair.cuh -- sum-factorised 2-D element computation by a PAIR of threads (config 2: Q2, 4x4 points).
//
// Same mathematics as element_compute_sf2d (madb_kernels.cuh): AssembleElementVector / AssembleElementGrad of the
// reference (src/ad_intg.hpp:202-257, :260-334) for one scalar H1 field with ADEval::GRAD, gradient and Hessian pulled
// back to reference coordinates, contractions one direction at a time on the 1-D tables.
//
// Work split.  The two threads of an element own one HALF of the quadrature points each: thread 0 the rows
// q2 < NQ/2, thread 1 the rows q2 >= NQ/2.  Thread 1 does not index the tables with its half: it processes the
// MIRRORED element (reflected at eta = 1/2: vertices k -> k ^ 2, dofs (i2, i1) -> (ND-1-i2, i1)).  Gauss points, the
// Gauss-Lobatto nodes and hence all 1-D tables are symmetric under that reflection, so rows q2 >= NQ/2 of the element
// ARE rows q2' = NQ-1-q2 < NQ/2 of its mirror image: both threads run the same instruction stream with the same
// immediate (constant-bank) table operands, only the gather indices differ.  The mirrored map has det J < 0; the
// reference's weight is Tr.Weight() = det J of the original orientation (src/ad_intg.hpp:237), so thread 1 flips the
// sign of the quadrature weights.  Every thread ends with its half's partial sums of ALL element-vector / matrix
// entries, in its own (possibly mirrored) local numbering.
//
// Exchange.  Entry X of the original element is local entry X of thread 0 and local entry mirror(X) of thread 1.
// The entries are walked in mirror pairs {X, mX}: each thread KEEPS its local X and SENDS its local mX
// (__shfl_xor 1); what it receives is the partner's half of the entry it keeps.  Self-mirror entries (X == mX) are
// sent and kept by both, both threads end with the same sum (a + b == b + a) and both hand it to the sink.  The sink
// index e (compile time) is the position in this walk; madb_patch.cpp maps (thread, e) to the CSR image through
// sf2d_pair_keep_v / sf2d_pair_keep_y below, applying the mirror for thread 1.
#pragma once
#include "madb_kernels.cuh"
#include "madb_pair_schedule.hpp"

namespace madb
{

#if defined(__CUDACC__)
__device__ __forceinline__ double shfl_xor1(double v) { return __shfl_xor_sync(0xffffffffu, v, 1); }

/// configurations / functionals the pair kernel covers
/// (compile with -DMADB_PAIR_KERNEL=1 to route these configurations through the thread-pair / CSR-image kernel; measured
/// on config 2 it is slower than the warp-specialised kernel k_patch_ws -- profiles/r02_img_kernel.md -- so the default is off)
#ifndef MADB_PAIR_KERNEL
#define MADB_PAIR_KERNEL 0
#endif
template <class Func, class Cfg> constexpr bool sf2d_pair_ok()
{
#if MADB_PAIR_KERNEL
   if constexpr (sf2d_cfg<Cfg>()) { return Func::N_QPRM == 0 && Func::N_INPUT == 2 && (Cfg::NQ1D % 2) == 0; }
   else { return false; }
#else
   return false;
#endif
}

/// elements per patch for <functional, configuration>: the pair kernel runs 64 elements per 128-thread work group
template <class Func, class Cfg> constexpr int patch_pe_of() { return sf2d_pair_ok<Func, Cfg>() ? 64 : patch_pe(Cfg::NVD); }

/// Fused residual + Jacobian of sorted element t by the thread pair (h = 0, 1).  ALL 32 lanes of the warp must call
/// (shuffles); a lane without an element passes a valid t and ignores the sinks.
/// ysink(e, value): kept element-vector entry e (sf2d_pair_keep_y); vsink(e, value): kept matrix entry e.
/// pre_matrix(): called after the element vector has been handed over, before the matrix phase.
template <class Func, class Cfg, class YSink, class VSink, class Hook>
__device__ __forceinline__ void element_compute_sf2d_pair(const AsmArgs<Func, Cfg> &a, const int t, const int h, YSink &&ysink,
                                                          VSink &&vsink, Hook &&pre_matrix)
{
   constexpr int ND = Cfg::template field<0>::ND1D, NQ = Cfg::NQ1D, NQH = NQ / 2;
   static_assert(NQ % 2 == 0, "the point rows are split in two halves");
   const auto &T = a.sf;
   using TA = AD<2, 2>;

   // ---- gather (thread 1: the mirrored element) ----------------------------------------------------------
   double X[4][2], u[ND][ND];
#pragma unroll
   for (int k = 0; k < 4; k++)
   {
      const int km = k ^ 2;
      const int n = a.e2n[(size_t)(h ? km : k) * a.stride + t];
      const double2 c = *reinterpret_cast<const double2 *>(a.coords + (size_t)n * 2);
      X[k][0] = c.x;
      X[k][1] = c.y;
   }
#pragma unroll
   for (int i2 = 0; i2 < ND; i2++)
   {
#pragma unroll
      for (int i1 = 0; i1 < ND; i1++)
      {
         const int i = i2 * ND + i1, im = (ND - 1 - i2) * ND + i1;
         const int idx = a.vmap[(size_t)(h ? im : i) * a.stride + t] & 0x7fffffff;
         u[i2][i1] = a.x[idx];
      }
   }
   Func f;
   f.load(a.fparams);
   const double wsgn = h ? -1.0 : 1.0; // det J of the mirrored map has the opposite sign

   // ---- bilinear geometry: dx/dxi = a0 + d eta, dx/deta = c0 + d xi ------------------------------------
   double a0[2], c0[2], dd[2], Jc1[NQ][2];
#pragma unroll
   for (int i = 0; i < 2; i++)
   {
      a0[i] = X[1][i] - X[0][i];
      c0[i] = X[2][i] - X[0][i];
      dd[i] = (X[3][i] - X[2][i]) - a0[i];
#pragma unroll
      for (int q = 0; q < NQ; q++) { Jc1[q][i] = fma(dd[i], T.xq[q], c0[i]); }
   }

   // ---- y-step of the interpolation for the own rows q2 < NQH: (B u)[q2][i1], (G u)[q2][i1] ------------------
   double uyb[NQH][ND], uyg[NQH][ND];
#pragma unroll
   for (int q2 = 0; q2 < NQH; q2++)
   {
#pragma unroll
      for (int i1 = 0; i1 < ND; i1++)
      {
         double sb = T.B[q2][0] * u[0][i1], sg = T.G[q2][0] * u[0][i1];
#pragma unroll
         for (int i2 = 1; i2 < ND; i2++)
         {
            sb = fma(T.B[q2][i2], u[i2][i1], sb);
            sg = fma(T.G[q2][i2], u[i2][i1], sg);
         }
         uyb[q2][i1] = sb;
         uyg[q2][i1] = sg;
      }
   }

   double r[ND * ND];
#pragma unroll
   for (int i = 0; i < ND * ND; i++) { r[i] = 0.0; }
   ZD H00[NQH][NQ], H01[NQH][NQ], H11[NQH][NQ]; // [q2][q1]

#pragma unroll
   for (int q2 = 0; q2 < NQH; q2++)
   {
      const double J00 = fma(dd[0], T.xq[q2], a0[0]), J10 = fma(dd[1], T.xq[q2], a0[1]);
      ZD gh0[NQ], gh1[NQ];
#pragma unroll
      for (int q1 = 0; q1 < NQ; q1++)
      {
         // x-step: reference gradient at (q2, q1)
         double rg0 = T.G[q1][0] * uyb[q2][0], rg1 = T.B[q1][0] * uyg[q2][0];
#pragma unroll
         for (int i1 = 1; i1 < ND; i1++)
         {
            rg0 = fma(T.G[q1][i1], uyb[q2][i1], rg0);
            rg1 = fma(T.B[q1][i1], uyg[q2][i1], rg1);
         }
         const double J01 = Jc1[q1][0], J11 = Jc1[q1][1];
         const double det = J00 * J11 - J01 * J10;
         const double rdet = frcp(det);
         double xin[2];
         xin[0] = (J11 * rg0 - J10 * rg1) * rdet;
         xin[1] = (J00 * rg1 - J01 * rg0) * rdet;
         TA xs[2];
         xs[0] = ad_seed<2, 2>(xin[0], 0);
         xs[1] = ad_seed<2, 2>(xin[1], 1);
         const TA res = f(xs, nullptr);
         const double wq = (T.wq[q2] * T.wq[q1]) * wsgn;
         // H^ = (w/det) adj H adj^T ; g^ = w adj g
         const double sc = wq * rdet;
         const ZD g00 = res.H(hidx<2>(0, 0)), g01 = res.H(hidx<2>(0, 1)), g11 = res.H(hidx<2>(1, 1));
         const ZD M00 = zfmac(g01, -J01, zmulc(g00, J11)), M01 = zfmac(g11, -J01, zmulc(g01, J11));
         const ZD M10 = zfmac(g01, J00, zmulc(g00, -J10)), M11 = zfmac(g11, J00, zmulc(g01, -J10));
         H00[q2][q1] = zmulc(zfmac(M01, -J01, zmulc(M00, J11)), sc);
         H01[q2][q1] = zmulc(zfmac(M01, J00, zmulc(M00, -J10)), sc);
         H11[q2][q1] = zmulc(zfmac(M11, J00, zmulc(M10, -J10)), sc);
         gh0[q1] = zmulc(zfmac(res.G(1), -J01, zmulc(res.G(0), J11)), wq);
         gh1[q1] = zmulc(zfmac(res.G(1), J00, zmulc(res.G(0), -J10)), wq);
      }
      // element vector, own rows: r[i2][i1] += B[q2][i2] t0[i1] + G[q2][i2] t1[i1]
      ZD t0[ND], t1[ND];
#pragma unroll
      for (int i1 = 0; i1 < ND; i1++)
      {
         ZD s0 {0.0, true}, s1 {0.0, true};
#pragma unroll
         for (int q1 = 0; q1 < NQ; q1++)
         {
            s0 = zfmac(gh0[q1], T.G[q1][i1], s0);
            s1 = zfmac(gh1[q1], T.B[q1][i1], s1);
         }
         t0[i1] = s0;
         t1[i1] = s1;
      }
#pragma unroll
      for (int i2 = 0; i2 < ND; i2++)
      {
#pragma unroll
         for (int i1 = 0; i1 < ND; i1++)
         {
            zacc(r[i2 * ND + i1], t0[i1], T.B[q2][i2]);
            zacc(r[i2 * ND + i1], t1[i1], T.G[q2][i2]);
         }
      }
   }

   // ---- element vector: keep row i2, send row ND-1-i2 ------------------------------------------------------
   {
      int e = 0;
#pragma unroll
      for (int i2 = 0; 2 * i2 <= ND - 1; i2++)
      {
#pragma unroll
         for (int i1 = 0; i1 < ND; i1++)
         {
            const double got = shfl_xor1(r[(ND - 1 - i2) * ND + i1]);
            ysink(e, r[i2 * ND + i1] + got);
            e++;
         }
      }
   }
   pre_matrix();

   // ---- element matrix ---------------------------------------------------------------------------------------
   int e = 0;
#pragma unroll
   for (int j1 = 0; j1 < ND; j1++)
   {
#pragma unroll
      for (int i1 = 0; i1 <= j1; i1++)
      {
         ZD T00[NQH], T01[NQH], T10[NQH], T11[NQH];
#pragma unroll
         for (int q2 = 0; q2 < NQH; q2++)
         {
            ZD s00 {0.0, true}, s01 {0.0, true}, s10 {0.0, true}, s11 {0.0, true};
#pragma unroll
            for (int q1 = 0; q1 < NQ; q1++)
            {
               s00 = zfmac(H00[q2][q1], T.GG[q1][i1][j1], s00);
               s01 = zfmac(H01[q2][q1], T.BG[q1][j1][i1], s01);
               if (i1 != j1) { s10 = zfmac(H01[q2][q1], T.BG[q1][i1][j1], s10); }
               s11 = zfmac(H11[q2][q1], T.BB[q1][i1][j1], s11);
            }
            T00[q2] = s00;
            T01[q2] = s01;
            T10[q2] = (i1 != j1) ? s10 : s01;
            T11[q2] = s11;
         }
         // the (i2, j2) entries of the block: own half's partial sums
         double v[ND][ND];
#pragma unroll
         for (int i2 = 0; i2 < ND; i2++)
         {
#pragma unroll
            for (int j2 = 0; j2 < ND; j2++)
            {
               if (i1 == j1 && i2 > j2) { continue; }
               double s = 0.0;
#pragma unroll
               for (int q2 = 0; q2 < NQH; q2++)
               {
                  zacc(s, T00[q2], T.BB[q2][i2][j2]);
                  zacc(s, T01[q2], T.BG[q2][i2][j2]);
                  zacc(s, T10[q2], T.BG[q2][j2][i2]);
                  zacc(s, T11[q2], T.GG[q2][i2][j2]);
               }
               v[i2][j2] = s;
            }
         }
         // exchange: keep (i2, j2), send its mirror image (same walk as sf2d_pair_keep_v)
#pragma unroll
         for (int i2 = 0; i2 < ND; i2++)
         {
#pragma unroll
            for (int j2 = 0; j2 < ND; j2++)
            {
               if (i1 == j1 && i2 > j2) { continue; }
               int mi = ND - 1 - i2, mj = ND - 1 - j2;
               if (i1 == j1 && mi > mj) { const int tt = mi; mi = mj; mj = tt; }
               const bool keep = (i2 < mi) || (i2 == mi && j2 <= mj);
               if (!keep) { continue; }
               const double got = shfl_xor1(v[mi][mj]);
               vsink(e, v[i2][j2] + got);
               e++;
            }
         }
      }
   }
}
#endif // __CUDACC__

} // namespace madb

// madb_sf2d.cuh -- sum-factorised element computation for 2-D tensor-product elements,
// one scalar H1 field with ADEval::GRAD (ex1, ex2, config 2).
//
// Same mathematics as the generic qpoint() of madb_kernels.cuh -- the reference's
// AssembleElementVector / AssembleElementGrad (src/ad_intg.hpp:202-257, :260-334) with the
// gradient and Hessian pulled back to reference coordinates -- but the dof <-> quadrature
// contractions are done one direction at a time on the 1-D tables B, G:
//   interpolation   u[i2][i1] -> (ub, ug)[i2][q1] -> reference gradient at (q2,q1)
//   residual        r[i2][i1] += B[q2][i2] (sum_q1 G[q1][i1] g^0) + G[q2][i2] (sum_q1 B[q1][i1] g^1)
//   Jacobian        A[(i2,i1),(j2,j1)] += BB[q2][i2][j2] T00[i1][j1] + BG[q2][i2][j2] T01[i1][j1]
//                                       + BG[q2][j2][i2] T01[j1][i1] + GG[q2][i2][j2] T11[i1][j1]
//                   with Tab[i1][j1] = sum_q1 (1-D product table)[q1][i1][j1] H^ab(q2,q1)
// For order 2 / 4x4 points this needs ~1.1 k FP64 operations for the element matrix instead of
// ~2.0 k (direct B^T D B on the upper triangle), and ~340 instead of ~580 for interpolation + residual.
// The bilinear geometry is evaluated in closed form: J(:,0) = a0 + d eta, J(:,1) = c0 + d xi; the
// pull-back uses adj(J) so that one reciprocal per point suffices:
//   grad u = adj^T (ref grad) / det ;  g^ = w adj g ;  H^ = (w / det) adj H adj^T   (w = reference weight).
#pragma once
#include "madb_config.cuh"
#include <algorithm>
#include <cmath>

namespace madb
{

template <int ND, int NQ> struct Sf2dTab
{
   double B[NQ][ND], G[NQ][ND];
   double BB[NQ][ND][ND], BG[NQ][ND][ND], GG[NQ][ND][ND]; // products at one 1-D point: BG[q][i][j] = B[q][i] G[q][j]
   // Mirror halves of the product tables (q < NQH; q' = NQ-1-q, i' = ND-1-i): Xs = (X[q][i][j] + X[q][i'][j']) / 2,
   // Xd = (X[q][i][j] - X[q][i'][j']) / 2.  1-D nodes and points symmetric about 1/2 give B[q'][i'] = B[q][i],
   // G[q'][i'] = -G[q][i], so the entries (i2,j2) and (i2',j2') of one (i1,j1) block share their even and odd sums
   // over q2: 18 instead of 32 FP64 operations per pair of entries (element matrix phase of element_compute_sf2d).
   static constexpr int NQH = (NQ + 1) / 2;
   double BBs[NQH][ND][ND], BBd[NQH][ND][ND], BGs[NQH][ND][ND], BGd[NQH][ND][ND], GGs[NQH][ND][ND], GGd[NQH][ND][ND];
   double xq[NQ], wq[NQ];
   int mirror_ok; // the 1-D tables have the mirror symmetry the even/odd contraction relies on (checked by fill_sf2d)
};
struct Sf2dNone
{
};

/// configurations the sum-factorised 2-D path covers
template <class Cfg> constexpr bool sf2d_cfg()
{
   if constexpr (Cfg::TENSOR && Cfg::DIM == 2 && Cfg::NF == 1)
   {
      using F = typename Cfg::template field<0>;
      return F::VDIM == 1 && F::HAS_GRAD && !F::HAS_VALUE && F::ROLE == ROLE_INPUT;
   }
   else { return false; }
}
template <class Cfg, bool OK = sf2d_cfg<Cfg>()> struct Sf2dTabFor
{
   using type = Sf2dNone;
};
template <class Cfg> struct Sf2dTabFor<Cfg, true>
{
   using type = Sf2dTab<Cfg::template field<0>::ND1D, Cfg::NQ1D>;
};

template <int ND, int NQ> void fill_sf2d(Sf2dTab<ND, NQ> &T, const double *b1d, const double *g1d, const double *xq, const double *wq)
{
   for (int q = 0; q < NQ; q++)
   {
      T.xq[q] = xq[q];
      T.wq[q] = wq[q];
      for (int i = 0; i < ND; i++)
      {
         T.B[q][i] = b1d[q * ND + i];
         T.G[q][i] = g1d[q * ND + i];
      }
      for (int i = 0; i < ND; i++)
      {
         for (int j = 0; j < ND; j++)
         {
            T.BB[q][i][j] = T.B[q][i] * T.B[q][j];
            T.BG[q][i][j] = T.B[q][i] * T.G[q][j];
            T.GG[q][i][j] = T.G[q][i] * T.G[q][j];
         }
      }
   }
   double dev = 0.0, scale = 0.0;
   for (int q = 0; q < NQ; q++)
   {
      for (int i = 0; i < ND; i++)
      {
         const double b0 = T.B[q][i], b1 = T.B[NQ - 1 - q][ND - 1 - i], g0 = T.G[q][i], g1 = T.G[NQ - 1 - q][ND - 1 - i];
         dev = std::max(dev, std::max(std::fabs(b0 - b1), std::fabs(g0 + g1)));
         scale = std::max(scale, std::max(std::fabs(b0), std::fabs(g0)));
      }
   }
   T.mirror_ok = dev <= 1e-13 * scale;
   for (int q = 0; q < Sf2dTab<ND, NQ>::NQH; q++)
   {
      for (int i = 0; i < ND; i++)
      {
         for (int j = 0; j < ND; j++)
         {
            const int im = ND - 1 - i, jm = ND - 1 - j;
            T.BBs[q][i][j] = 0.5 * (T.BB[q][i][j] + T.BB[q][im][jm]);
            T.BBd[q][i][j] = 0.5 * (T.BB[q][i][j] - T.BB[q][im][jm]);
            T.BGs[q][i][j] = 0.5 * (T.BG[q][i][j] + T.BG[q][im][jm]);
            T.BGd[q][i][j] = 0.5 * (T.BG[q][i][j] - T.BG[q][im][jm]);
            T.GGs[q][i][j] = 0.5 * (T.GG[q][i][j] + T.GG[q][im][jm]);
            T.GGd[q][i][j] = 0.5 * (T.GG[q][i][j] - T.GG[q][im][jm]);
         }
      }
   }
}
template <int ND, int NQ> inline bool sf2d_mirror_ok(const Sf2dTab<ND, NQ> &T) { return T.mirror_ok != 0; }
inline bool sf2d_mirror_ok(const Sf2dNone &) { return true; }
inline void fill_sf2d(Sf2dNone &, const double *, const double *, const double *, const double *) {}

} // namespace madb

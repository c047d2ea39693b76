// madb_kernels.cuh -- whole-mesh batched element kernels (one thread per element).
//
// One launch per element colour replaces the reference's per-element virtual
// calls from MFEM's element loop:
//   gather  (GetElementVDofs + GetSubVector)               -> vmap loads
//   geometry (Tr.SetIntPoint / Weight / InverseJacobian)   -> from 2^DIM vertices
//   CalcInputShapes + x = allshapes^T elfun                -> src/ad_intg.hpp:118-154,:242
//   f.Gradient / f.Hessian                                 -> one AD<N,1>/AD<N,2> pass
//   elvect += allshapes*jac, elmat += B H B^T              -> src/ad_intg.hpp:245-255,:307-331,:698-727
//   AddElementVector / AddSubMatrix                        -> direct scatter through vmap / e2csr
// Physical shapes are never formed: the gradient and Hessian are pulled back to
// reference coordinates (g^ = w T^T g, H^ = w T^T H T with T = blockdiag(1, J^-T)),
// so the test/trial contraction uses the constant reference tables only.
//
// Determinism: elements of one colour share no dof, so the scatter uses plain
// loads/stores; colours run in stream order, hence every y[i] / vals[p] is
// summed in a fixed (colour) order.  Bit 31 of a map entry marks the first
// contribution to that slot: it stores instead of accumulating, so neither y
// nor vals needs a memset.
#pragma once
#include "madb_config.cuh"
#include "madb_host.hpp"
#include "madb_sf2d.cuh"
#include <cuda_runtime.h>

namespace madb
{


template <class Func, class Cfg> struct AsmArgs
{
   static constexpr int NQF = Func::N_QPRM - Cfg::N_FIELD_QPRM;
   static_assert(Func::N_INPUT == Cfg::N_INPUT, "functional n_input must match shapedim*vdim summed over the input spaces");
   static_assert(NQF >= 0, "functional expects fewer per-point parameters than the parameter fields supply");
   int begin, end; // sorted-element range of this launch (one colour)
   int stride;     // padded element count = SoA stride of all maps
   int write_y, write_vals;
   const int *e2n;       // [NGN][stride] vertex ids, lexicographic
   const double *coords; // [nnodes][DIM]
   const double *xe;     // 2-D: vertex coordinates per element [4][stride][2] (sum-factorised path)
   const int *vmap;      // [NVD][stride] index into x / y (bit 31: first touch)
   const int *pmap;      // [NDOF_ALL-NVD][stride] indices into pdata
   const double *pdata[Cfg::NF]; // parameter-field vectors (null for input fields)
   const double *qf;     // [NQF][NQ][stride] per-point parameters given as quadrature functions
   const int *e2csr;     // [NVD*NVD][stride] CSR positions (bit 31: first touch)
   const double *x;
   const double *v; // MODE_ACT: direction
   double *y;
   double *vals;
   double *energy; // [stride] per-element energies (sorted order)
   const int *perm; // sorted position -> element
   double *cvalue, *cgrad; // MODE_COEF: value [e][q] and gradient [e][q][N] at the points
   double *chess;          // MODE_COEF: Hessian [e][q][N][N] (HessianCoefficient, src/ad_native.hpp:300-323)
   int coef_variant;       // MODE_COEF: 1 = cgrad is ParamGradient::Eval as written (src/mmto.cpp:25-37)
   double fparams[Func::N_PARAM > 0 ? Func::N_PARAM : 1];
   Tables<Cfg> tab;
   typename Sf2dTabFor<Cfg>::type sf; // 1-D tables of the sum-factorised 2-D path (empty otherwise)
};

/// Basic-block boundary ptxas cannot remove (the stride is never negative, which it cannot know).  ptxas schedules
/// inside basic blocks and sinks loads whose results are needed late to the end of theirs: without the boundaries the
/// map loads and the value loads of the prefetch end up next to each other at the end of the matrix phase.
__device__ __forceinline__ void bb_break(const int never_negative)
{
   for (int k = never_negative; k < 0; k++) { __nanosleep(1); }
}
#ifndef MADB_QLOOP_UNROLL
#define MADB_QLOOP_UNROLL 1 // unroll factor of the quadrature loop of configurations registered with UNROLLQ = false
#endif
#define MADB_PRAGMA_(x) _Pragma(#x)
#define MADB_UNROLL(n) MADB_PRAGMA_(unroll n)
#ifndef MADB_QPOINT_BB
#define MADB_QPOINT_BB 1 // generic qpoint(): bit 0: boundary between the per-point phase (inputs, AD, pull-back) and the contraction for
                         // element matrices of more than 12 dofs (config 5: 4.51 -> 4.18 ms; config 4's 8-dof block loses 3 %, so
                         // small matrices are left alone), bit 1: between the test-function dofs' fields (no further gain)
#endif
/// functionals that implement ParamGradient::Eval as written provide param_gradient_as_written(x, qp, J)
template <class F, class = void> struct has_param_gradient : std::false_type
{
};
template <class F> struct has_param_gradient<F, std::void_t<decltype(F::HAS_PARAM_GRADIENT)>> : std::true_type
{
};

template <int DIM> MADB_HD void invert(const double (&J)[DIM][DIM], double (&Ji)[DIM][DIM], double &det)
{
   if constexpr (DIM == 1)
   {
      det = J[0][0];
      Ji[0][0] = frcp(det);
   }
   else if constexpr (DIM == 2)
   {
      det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      const double t = frcp(det);
      Ji[0][0] = J[1][1] * t;
      Ji[0][1] = -J[0][1] * t;
      Ji[1][0] = -J[1][0] * t;
      Ji[1][1] = J[0][0] * t;
   }
   else
   {
      const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
      const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
      const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
      const double t = frcp(det);
      Ji[0][0] = c00 * t;
      Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * t;
      Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * t;
      Ji[1][0] = c01 * t;
      Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * t;
      Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * t;
      Ji[2][0] = c02 * t;
      Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * t;
      Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * t;
   }
}

MADB_HD constexpr int symidx(int a, int b) { return a <= b ? b * (b + 1) / 2 + a : a * (a + 1) / 2 + b; }

/// Everything the functional needs at one quadrature point, then the
/// contribution of that point to the element vector / matrix.
template <class Func, class Cfg>
__device__ __forceinline__ void qpoint_inputs(const AsmArgs<Func, Cfg> &a, const Tables<Cfg> &tab, const int q, const int t,
                                              const double (&X)[Cfg::NGN][Cfg::DIM],
                                              const double (&u)[Cfg::NDOF_ALL],
                                              double (&xin)[Cfg::N_INPUT],
                                              double (&qp)[Func::N_QPRM > 0 ? Func::N_QPRM : 1],
                                              double (&Ji)[Cfg::DIM][Cfg::DIM], double &w)
{
   constexpr int DIM = Cfg::DIM, NF = Cfg::NF;
   using Args = AsmArgs<Func, Cfg>;

   // ---- geometry: J = sum_k X_k (x) dN_k/dxi ; Weight = det J ; J^-1 ------------
   double J[DIM][DIM];
#pragma unroll
   for (int i = 0; i < DIM; i++)
   {
#pragma unroll
      for (int j = 0; j < DIM; j++) { J[i][j] = 0.0; }
   }
#pragma unroll
   for (int k = 0; k < Cfg::NGN; k++)
   {
#pragma unroll
      for (int j = 0; j < DIM; j++)
      {
         const double g = tab.gdphi[q][k][j];
#pragma unroll
         for (int i = 0; i < DIM; i++) { J[i][j] = fma(X[k][i], g, J[i][j]); }
      }
   }
   double detJ;
   invert<DIM>(J, Ji, detJ);
   w = tab.w[q] * detJ; // ip.weight * Tr.Weight()

   // ---- inputs x (physical) and field parameters ---------------------------------
   static_for<NF>([&](auto F)
   {
      constexpr int fi = decltype(F)::value;
      using Fd = typename Cfg::template field<fi>;
      constexpr int nd = Cfg::template nd<fi>(), sd = Cfg::template sd<fi>();
      constexpr int toff = Cfg::template toff<fi>(), doff = Cfg::template doff<fi>();
#pragma unroll
      for (int c = 0; c < Fd::VDIM; c++)
      {
         double val = 0.0, rg[DIM];
#pragma unroll
         for (int k = 0; k < DIM; k++) { rg[k] = 0.0; }
#pragma unroll
         for (int i = 0; i < nd; i++)
         {
            const double ui = u[doff + c * nd + i];
            if constexpr (Fd::HAS_VALUE) { val = fma(tab.phi[q][toff + i], ui, val); }
            if constexpr (Fd::HAS_GRAD)
            {
#pragma unroll
               for (int k = 0; k < DIM; k++) { rg[k] = fma(tab.dphi[q][toff + i][k], ui, rg[k]); }
            }
         }
         double *dst = Cfg::template is_input<fi>() ? (xin + Cfg::template xoff<fi>()) : (qp + Cfg::template poff<fi>());
         int slot = c * sd;
         if constexpr (Fd::HAS_VALUE) { dst[slot++] = val; }
         if constexpr (Fd::HAS_GRAD)
         {
            // physical gradient: gshape = dshape * J^-1  (CalcPhysDShape)
#pragma unroll
            for (int j = 0; j < DIM; j++)
            {
               double s = 0.0;
#pragma unroll
               for (int k = 0; k < DIM; k++) { s = fma(Ji[k][j], rg[k], s); }
               dst[slot + j] = s;
            }
         }
      }
   });
#pragma unroll
   for (int k = 0; k < Args::NQF; k++) { qp[Cfg::N_FIELD_QPRM + k] = a.qf[((size_t)k * Cfg::NQ + q) * a.stride + t]; }
}

/// Contribution of one quadrature point to the element vector / matrix / energy.
/// B0, B1: only the matrix entries (a, b), a <= b, with b in [B0, B1) are accumulated (a thread may own a
/// slice of the upper triangle: large element matrices are split over several threads, see k_element).
template <class Func, class Cfg, int MODE, int B0 = 0, int B1 = Cfg::NVD>
__device__ __forceinline__ void qpoint(const AsmArgs<Func, Cfg> &a, const Tables<Cfg> &tab, const int q, const int t,
                                       const double (&X)[Cfg::NGN][Cfg::DIM],
                                       const double (&u)[Cfg::NDOF_ALL],
                                       const double (&vdir)[(MODE & MODE_ACT) ? Cfg::NVD : 1],
                                       const Func &f,
                                       double (&r)[(MODE & (MODE_RES | MODE_ACT)) ? Cfg::NVD : 1],
                                       double (&A)[(MODE & MODE_JAC) ? Cfg::NSYM : 1],
                                       double &energy)
{
   constexpr int DIM = Cfg::DIM, NF = Cfg::NF, N = Cfg::N_INPUT;
   double xin[N];
   double qp[Func::N_QPRM > 0 ? Func::N_QPRM : 1];
   double Ji[DIM][DIM], w;
   qpoint_inputs<Func, Cfg>(a, tab, q, t, X, u, xin, qp, Ji, w);

   if constexpr (MODE == MODE_COEF)
   {
      // DifferentiableCoefficient::Eval / Gradient().Eval at the point (src/ad_native.hpp:272-283,315-317):
      // value and gradient of f w.r.t. its inputs, stored as quadrature functions [e][q][.]
      const size_t pt = (size_t)a.perm[t] * Cfg::NQ + q;
      if constexpr (has_param_gradient<Func>::value)
      {
         if (a.coef_variant == 1)
         {
            // ParametrizedFunctional::ParamGradient::Eval as written (src/mmto.cpp:25-37)
            double J[N];
            f.param_gradient_as_written(xin, qp, J);
            if (a.cvalue) { a.cvalue[pt] = f(xin, qp); }
            if (a.cgrad)
            {
#pragma unroll
               for (int m = 0; m < N; m++) { a.cgrad[pt * N + m] = J[m]; }
            }
            return;
         }
      }
      if (a.chess)
      {
         using T2 = AD<N, 2>;
         T2 xs[N];
#pragma unroll
         for (int m = 0; m < N; m++) { xs[m] = ad_seed<N, 2>(xin[m], m); }
         const T2 res = f(xs, qp);
         if (a.cvalue) { a.cvalue[pt] = res.v; }
         if (a.cgrad)
         {
#pragma unroll
            for (int m = 0; m < N; m++) { a.cgrad[pt * N + m] = res.g[m]; }
         }
#pragma unroll
         for (int m = 0; m < N; m++)
         {
#pragma unroll
            for (int n = 0; n < N; n++) { a.chess[(pt * N + m) * N + n] = res.hess(m, n); }
         }
         return;
      }
      using T = AD<N, 1>;
      T xs[N];
#pragma unroll
      for (int m = 0; m < N; m++) { xs[m] = ad_seed<N, 1>(xin[m], m); }
      const T res = f(xs, qp);
      if (a.cvalue) { a.cvalue[pt] = res.v; }
      if (a.cgrad)
      {
#pragma unroll
         for (int m = 0; m < N; m++) { a.cgrad[pt * N + m] = res.g[m]; }
      }
      return;
   }
   else if constexpr (MODE == MODE_ENERGY)
   {
      energy += f(xin, qp) * w; // src/ad_intg.hpp:196
      return;
   }
   else
   {
      constexpr int ORDER = (MODE & (MODE_JAC | MODE_ACT)) ? 2 : 1;
      using T = AD<N, ORDER>;
      T xs[N];
#pragma unroll
      for (int m = 0; m < N; m++) { xs[m] = ad_seed<N, ORDER>(xin[m], m); }
      const T res = f(xs, qp);

      // ---- pull back to reference coordinates: g^ = w T^T g ; H^ = w T^T H T ------
      ZD gh[N];
      ZD Hh[(ORDER >= 2) ? N : 1][(ORDER >= 2) ? N : 1];
      static_for<NF>([&](auto F)
      {
         constexpr int fi = decltype(F)::value;
         using Fd = typename Cfg::template field<fi>;
         if constexpr (Cfg::template is_input<fi>())
         {
            constexpr int sd = Cfg::template sd<fi>(), xo = Cfg::template xoff<fi>();
            constexpr int gofs = Fd::HAS_VALUE ? 1 : 0;
#pragma unroll
            for (int c = 0; c < Fd::VDIM; c++)
            {
               const int base = xo + c * sd;
               if constexpr (Fd::HAS_VALUE) { gh[base] = zmulc(res.G(base), w); }
               if constexpr (Fd::HAS_GRAD)
               {
#pragma unroll
                  for (int k = 0; k < DIM; k++)
                  {
                     ZD s {0.0, true};
#pragma unroll
                     for (int j = 0; j < DIM; j++) { s = zfmac(res.G(base + gofs + j), Ji[k][j], s); }
                     gh[base + gofs + k] = zmulc(s, w);
                  }
               }
            }
         }
      });
      if constexpr (ORDER >= 2)
      {
         // M = H T (columns), then H^ = w T^T M (rows)
         ZD M[N][N];
         static_for<NF>([&](auto F)
         {
            constexpr int fi = decltype(F)::value;
            using Fd = typename Cfg::template field<fi>;
            if constexpr (Cfg::template is_input<fi>())
            {
               constexpr int sd = Cfg::template sd<fi>(), xo = Cfg::template xoff<fi>();
               constexpr int gofs = Fd::HAS_VALUE ? 1 : 0;
#pragma unroll
               for (int c = 0; c < Fd::VDIM; c++)
               {
                  const int base = xo + c * sd;
#pragma unroll
                  for (int m = 0; m < N; m++)
                  {
                     if constexpr (Fd::HAS_VALUE) { M[m][base] = res.H(symidx_h<N>(m, base)); }
                     if constexpr (Fd::HAS_GRAD)
                     {
#pragma unroll
                        for (int k = 0; k < DIM; k++)
                        {
                           ZD s {0.0, true};
#pragma unroll
                           for (int j = 0; j < DIM; j++) { s = zfmac(res.H(symidx_h<N>(m, base + gofs + j)), Ji[k][j], s); }
                           M[m][base + gofs + k] = s;
                        }
                     }
                  }
               }
            }
         });
         static_for<NF>([&](auto F)
         {
            constexpr int fi = decltype(F)::value;
            using Fd = typename Cfg::template field<fi>;
            if constexpr (Cfg::template is_input<fi>())
            {
               constexpr int sd = Cfg::template sd<fi>(), xo = Cfg::template xoff<fi>();
               constexpr int gofs = Fd::HAS_VALUE ? 1 : 0;
#pragma unroll
               for (int c = 0; c < Fd::VDIM; c++)
               {
                  const int base = xo + c * sd;
#pragma unroll
                  for (int n = 0; n < N; n++)
                  {
                     if constexpr (Fd::HAS_VALUE) { Hh[base][n] = zmulc(M[base][n], w); }
                     if constexpr (Fd::HAS_GRAD)
                     {
#pragma unroll
                        for (int k = 0; k < DIM; k++)
                        {
                           ZD s {0.0, true};
#pragma unroll
                           for (int j = 0; j < DIM; j++) { s = zfmac(M[base + gofs + j][n], Ji[k][j], s); }
                           Hh[base + gofs + k][n] = zmulc(s, w);
                        }
                     }
                  }
               }
            }
         });
      }

      // ---- MODE_ACT: y^ = H^ v^ with v^ the reference-space inputs of the direction --
      ZD yh[(MODE & MODE_ACT) ? N : 1];
      if constexpr ((MODE & MODE_ACT) != 0)
      {
         double vh[N];
         static_for<NF>([&](auto F)
         {
            constexpr int fi = decltype(F)::value;
            using Fd = typename Cfg::template field<fi>;
            if constexpr (Cfg::template is_input<fi>())
            {
               constexpr int nd = Cfg::template nd<fi>(), sd = Cfg::template sd<fi>();
               constexpr int toff = Cfg::template toff<fi>(), vo = Cfg::template voff<fi>(), xo = Cfg::template xoff<fi>();
#pragma unroll
               for (int c = 0; c < Fd::VDIM; c++)
               {
                  double val = 0.0, rg[DIM];
#pragma unroll
                  for (int k = 0; k < DIM; k++) { rg[k] = 0.0; }
#pragma unroll
                  for (int i = 0; i < nd; i++)
                  {
                     const double vi = vdir[vo + c * nd + i];
                     if constexpr (Fd::HAS_VALUE) { val = fma(tab.phi[q][toff + i], vi, val); }
                     if constexpr (Fd::HAS_GRAD)
                     {
#pragma unroll
                        for (int k = 0; k < DIM; k++) { rg[k] = fma(tab.dphi[q][toff + i][k], vi, rg[k]); }
                     }
                  }
                  int slot = xo + c * sd;
                  if constexpr (Fd::HAS_VALUE) { vh[slot++] = val; }
                  if constexpr (Fd::HAS_GRAD)
                  {
#pragma unroll
                     for (int k = 0; k < DIM; k++) { vh[slot + k] = rg[k]; }
                  }
               }
            }
         });
#pragma unroll
         for (int m = 0; m < N; m++)
         {
            ZD s {0.0, true};
#pragma unroll
            for (int n = 0; n < N; n++) { s = zfmac(Hh[m][n], vh[n], s); }
            yh[m] = s;
         }
      }

      // ---- test-function contraction: elvect += B g^ ; elmat += B H^ B^T ----------
      if constexpr ((MADB_QPOINT_BB & 1) != 0 && (MODE & MODE_JAC) != 0 && Cfg::NVD > 12) { bb_break(a.stride); }
      static_for<NF>([&](auto FB)
      {
         constexpr int fb = decltype(FB)::value;
         using Fb = typename Cfg::template field<fb>;
         if constexpr (Cfg::template is_input<fb>())
         {
            if constexpr ((MADB_QPOINT_BB & 2) != 0 && (MODE & MODE_JAC) != 0 && fb > 0) { bb_break(a.stride); }
            constexpr int ndb = Cfg::template nd<fb>(), sdb = Cfg::template sd<fb>();
            constexpr int tob = Cfg::template toff<fb>(), vob = Cfg::template voff<fb>(), xob = Cfg::template xoff<fb>();
#pragma unroll
            for (int jb = 0; jb < ndb; jb++)
            {
               // reference basis vector of trial/test dof jb: [phi?, dphi/dxi_k?]
               double bv[sdb];
               {
                  int s = 0;
                  if constexpr (Fb::HAS_VALUE) { bv[s++] = tab.phi[q][tob + jb]; }
                  if constexpr (Fb::HAS_GRAD)
                  {
#pragma unroll
                     for (int k = 0; k < DIM; k++) { bv[s + k] = tab.dphi[q][tob + jb][k]; }
                  }
               }
#pragma unroll
               for (int cb = 0; cb < Fb::VDIM; cb++)
               {
                  const int b = vob + cb * ndb + jb;
                  const int sb = xob + cb * sdb;
                  if constexpr ((MODE & MODE_RES) != 0)
                  {
                     ZD s {0.0, true};
#pragma unroll
                     for (int k = 0; k < sdb; k++) { s = zfmac(gh[sb + k], bv[k], s); }
                     if (!s.z) { r[b] += s.v; }
                  }
                  if constexpr ((MODE & MODE_ACT) != 0)
                  {
                     ZD s {0.0, true};
#pragma unroll
                     for (int k = 0; k < sdb; k++) { s = zfmac(yh[sb + k], bv[k], s); }
                     if (!s.z) { r[b] += s.v; }
                  }
                  if constexpr ((MODE & MODE_JAC) != 0)
                  {
                     if (b < B0 || b >= B1) { continue; }
                     // tcol[m] = sum_k H^[m][sb+k] bv[k]
                     ZD tcol[N];
#pragma unroll
                     for (int m = 0; m < N; m++)
                     {
                        ZD s {0.0, true};
#pragma unroll
                        for (int k = 0; k < sdb; k++) { s = zfmac(Hh[m][sb + k], bv[k], s); }
                        tcol[m] = s;
                     }
                     static_for<NF>([&](auto FA)
                     {
                        constexpr int fa = decltype(FA)::value;
                        using Fa = typename Cfg::template field<fa>;
                        if constexpr (Cfg::template is_input<fa>() && fa <= fb)
                        {
                           constexpr int nda = Cfg::template nd<fa>(), sda = Cfg::template sd<fa>();
                           constexpr int toa = Cfg::template toff<fa>(), voa = Cfg::template voff<fa>(), xoa = Cfg::template xoff<fa>();
#pragma unroll
                           for (int ca = 0; ca < Fa::VDIM; ca++)
                           {
#pragma unroll
                              for (int ia = 0; ia < nda; ia++)
                              {
                                 const int aa = voa + ca * nda + ia;
                                 if (aa <= b)
                                 {
                                    const int sa = xoa + ca * sda;
                                    ZD s {0.0, true};
                                    int k0 = 0;
                                    if constexpr (Fa::HAS_VALUE) { s = zfmac(tcol[sa], tab.phi[q][toa + ia], s); k0 = 1; }
                                    if constexpr (Fa::HAS_GRAD)
                                    {
#pragma unroll
                                       for (int k = 0; k < DIM; k++) { s = zfmac(tcol[sa + k0 + k], tab.dphi[q][toa + ia][k], s); }
                                    }
                                    if (!s.z) { A[symidx(aa, b)] += s.v; }
                                 }
                              }
                           }
                        }
                     });
                  }
               }
            }
         }
      });
   }
}

#ifndef MADB_SF2D_FENCE
#define MADB_SF2D_FENCE 1
#endif
/// Sum-factorised gather + quadrature loop (madb_sf2d.cuh) for 2-D scalar H1 fields with ADEval::GRAD.
/// The element vector is returned in r; the entries of the upper triangle of the element matrix are handed
/// to sink(k, value), k = symidx(I, J), one by one at the end (they never all live in registers: the
/// pulled-back Hessians of the NQ x NQ points are kept instead, 3 doubles per point).
/// pre_matrix() (k_patch_ws) runs between the quadrature loop (r is final there) and the emission of the matrix entries.
struct NoHook
{
   __device__ __forceinline__ void operator()() const {}
   __device__ __forceinline__ void operator()(int) const {}
};
/// Inputs of one element of the sum-factorised 2-D path: vertex coordinates, dof values (and the direction of ACT)
template <class Cfg, bool ACT> struct Sf2dIn
{
   static constexpr int ND = Cfg::template field<0>::ND1D;
   double X[4][2];
   double u[ND][ND], vd[ACT ? ND : 1][ACT ? ND : 1];
};
#ifndef MADB_SF2D_XE
#define MADB_SF2D_XE 0 // 1: vertex coordinates from a per-element array (one dependent load less, +32 MB of traffic per assembly of
                       // config 2): no measurable gain (0.266 vs 0.264 ms), off
#endif
/// vertex coordinates of sorted element t (MADB_SF2D_XE: one coalesced 16-byte load per vertex from the per-element array)
template <class Func, class Cfg, bool ACT>
__device__ __forceinline__ void sf2d_gather_x(const AsmArgs<Func, Cfg> &a, const int t, Sf2dIn<Cfg, ACT> &in)
{
#if MADB_SF2D_XE
   const double2 *xe = reinterpret_cast<const double2 *>(a.xe);
#pragma unroll
   for (int k = 0; k < 4; k++)
   {
      const double2 v = __ldg(xe + (size_t)k * a.stride + t);
      in.X[k][0] = v.x;
      in.X[k][1] = v.y;
   }
#else
#pragma unroll
   for (int k = 0; k < 4; k++)
   {
      const int n = a.e2n[(size_t)k * a.stride + t];
      in.X[k][0] = a.coords[(size_t)n * 2];
      in.X[k][1] = a.coords[(size_t)n * 2 + 1];
   }
#endif
}
template <class Func, class Cfg, bool ACT>
__device__ __forceinline__ void sf2d_gather(const AsmArgs<Func, Cfg> &a, const int t, Sf2dIn<Cfg, ACT> &in)
{
   constexpr int ND = Cfg::template field<0>::ND1D, NVD = Cfg::NVD;
   sf2d_gather_x<Func, Cfg, ACT>(a, t, in);
#pragma unroll
   for (int i = 0; i < NVD; i++)
   {
      const int idx = a.vmap[(size_t)i * a.stride + t] & 0x7fffffff;
      in.u[i / ND][i % ND] = a.x[idx];
      if constexpr (ACT) { in.vd[i / ND][i % ND] = a.v[idx]; }
   }
}
template <class Func, class Cfg, int MODE, class Sink, class HookPre = NoHook, class HookMid = NoHook>
__device__ __forceinline__ void element_compute_sf2d_core(const AsmArgs<Func, Cfg> &a, const Sf2dIn<Cfg, (MODE & MODE_ACT) != 0> &in,
                                                          double (&r)[(MODE & (MODE_RES | MODE_ACT)) ? Cfg::NVD : 1], Sink &&sink,
                                                          HookPre &&pre_matrix = HookPre(), HookMid &&mid_matrix = HookMid())
{
   constexpr int ND = Cfg::template field<0>::ND1D, NQ = Cfg::NQ1D, NVD = Cfg::NVD;
   constexpr bool RES = (MODE & MODE_RES) != 0, JAC = (MODE & MODE_JAC) != 0, ACT = (MODE & MODE_ACT) != 0;
   constexpr int ORDER = (JAC || ACT) ? 2 : 1;
   static_assert(Func::N_INPUT == 2 && Func::N_QPRM == 0, "sum-factorised 2-D path: scalar field, GRAD, no per-point parameters");
   const auto &T = a.sf;
   const auto &X = in.X;
   const auto &u = in.u;
   const auto &vd = in.vd;
   Func f;
   f.load(a.fparams);

   // ---- bilinear geometry: dx/dxi = a0 + d eta, dx/deta = c0 + d xi ---------------------------
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

   // ---- x-step of the interpolation ----------------------------------------------------------
   double ub[ND][NQ], ug[ND][NQ], vb[ACT ? ND : 1][ACT ? NQ : 1], vg[ACT ? ND : 1][ACT ? NQ : 1];
#pragma unroll
   for (int i2 = 0; i2 < ND; i2++)
   {
#pragma unroll
      for (int q = 0; q < NQ; q++)
      {
         double sb = T.B[q][0] * u[i2][0], sg = T.G[q][0] * u[i2][0];
#pragma unroll
         for (int i1 = 1; i1 < ND; i1++)
         {
            sb = fma(T.B[q][i1], u[i2][i1], sb);
            sg = fma(T.G[q][i1], u[i2][i1], sg);
         }
         ub[i2][q] = sb;
         ug[i2][q] = sg;
         if constexpr (ACT)
         {
            double tb = T.B[q][0] * vd[i2][0], tg = T.G[q][0] * vd[i2][0];
#pragma unroll
            for (int i1 = 1; i1 < ND; i1++)
            {
               tb = fma(T.B[q][i1], vd[i2][i1], tb);
               tg = fma(T.G[q][i1], vd[i2][i1], tg);
            }
            vb[i2][q] = tb;
            vg[i2][q] = tg;
         }
      }
   }

   if constexpr (RES || ACT)
   {
#pragma unroll
      for (int i = 0; i < NVD; i++) { r[i] = 0.0; }
   }
   ZD H00[JAC ? NQ : 1][JAC ? NQ : 1], H01[JAC ? NQ : 1][JAC ? NQ : 1], H11[JAC ? NQ : 1][JAC ? NQ : 1]; // [q2][q1]

#pragma unroll
   for (int q2 = 0; q2 < NQ; q2++)
   {
      const double J00 = fma(dd[0], T.xq[q2], a0[0]), J10 = fma(dd[1], T.xq[q2], a0[1]);
      ZD gh0[NQ], gh1[NQ];                    // reference-space gradient (RES) or H^ v^ (ACT), weighted
#pragma unroll
      for (int q1 = 0; q1 < NQ; q1++)
      {
         // y-step of the interpolation: reference gradient at (q2, q1)
         double rg0 = T.B[q2][0] * ug[0][q1], rg1 = T.G[q2][0] * ub[0][q1];
#pragma unroll
         for (int i2 = 1; i2 < ND; i2++)
         {
            rg0 = fma(T.B[q2][i2], ug[i2][q1], rg0);
            rg1 = fma(T.G[q2][i2], ub[i2][q1], rg1);
         }
         const double J01 = Jc1[q1][0], J11 = Jc1[q1][1];
         const double det = J00 * J11 - J01 * J10;
         const double rdet = frcp(det);
         // physical gradient = J^-T (ref grad) = adj^T (ref grad) / det,  adj = [[J11,-J01],[-J10,J00]]
         double xin[2];
         xin[0] = (J11 * rg0 - J10 * rg1) * rdet;
         xin[1] = (J00 * rg1 - J01 * rg0) * rdet;
         using TA = AD<2, ORDER>;
         TA xs[2];
         xs[0] = ad_seed<2, ORDER>(xin[0], 0);
         xs[1] = ad_seed<2, ORDER>(xin[1], 1);
         const TA res = f(xs, nullptr);
         const double wq = T.wq[q2] * T.wq[q1];
         ZD h00 {0.0, true}, h01 {0.0, true}, h11 {0.0, true};
         if constexpr (ORDER >= 2)
         {
            // H^ = (w/det) adj H adj^T
            const double sc = wq * rdet;
            const ZD g00 = res.H(hidx<2>(0, 0)), g01 = res.H(hidx<2>(0, 1)), g11 = res.H(hidx<2>(1, 1));
            const ZD M00 = zfmac(g01, -J01, zmulc(g00, J11)), M01 = zfmac(g11, -J01, zmulc(g01, J11));
            const ZD M10 = zfmac(g01, J00, zmulc(g00, -J10)), M11 = zfmac(g11, J00, zmulc(g01, -J10));
            h00 = zmulc(zfmac(M01, -J01, zmulc(M00, J11)), sc);
            h01 = zmulc(zfmac(M01, J00, zmulc(M00, -J10)), sc);
            h11 = zmulc(zfmac(M11, J00, zmulc(M10, -J10)), sc);
         }
         if constexpr (JAC)
         {
            H00[q2][q1] = h00;
            H01[q2][q1] = h01;
            H11[q2][q1] = h11;
         }
         if constexpr (RES)
         {
            // g^ = w adj g
            gh0[q1] = zmulc(zfmac(res.G(1), -J01, zmulc(res.G(0), J11)), wq);
            gh1[q1] = zmulc(zfmac(res.G(1), J00, zmulc(res.G(0), -J10)), wq);
         }
         if constexpr (ACT)
         {
            // y^ = H^ v^ with v^ the reference gradient of the direction
            double vr0 = T.B[q2][0] * vg[0][q1], vr1 = T.G[q2][0] * vb[0][q1];
#pragma unroll
            for (int i2 = 1; i2 < ND; i2++)
            {
               vr0 = fma(T.B[q2][i2], vg[i2][q1], vr0);
               vr1 = fma(T.G[q2][i2], vb[i2][q1], vr1);
            }
            gh0[q1] = zfmac(h01, vr1, zmulc(h00, vr0));
            gh1[q1] = zfmac(h11, vr1, zmulc(h01, vr0));
         }
      }

      // ---- element vector: r[i2][i1] += B[q2][i2] t0[i1] + G[q2][i2] t1[i1] ------------------------
      if constexpr (RES || ACT)
      {
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
      mid_matrix(100 + q2); // end of a row of points (hook codes: 0.. block of the matrix phase, 100 + q2, 200 + block: between its two stages)
   }

   pre_matrix();
   // ---- element matrix: for every pair (i1 <= j1) of 1-D x-indices contract over q1, then emit the (i2, j2) block ----
   if constexpr (JAC)
   {
#pragma unroll
      for (int j1 = 0; j1 < ND; j1++)
      {
#pragma unroll
         for (int i1 = 0; i1 <= j1; i1++)
         {
            // Scheduling fence between two (i1, j1) blocks: left alone, ptxas hoists the q1-contractions of the later
            // blocks above the emission of the earlier ones and spills about 50 registers per element.  The stores of
            // the previous block stay above the fence ("memory"), the pulled-back Hessians are "redefined" by it.
            mid_matrix(j1 * (j1 + 1) / 2 + i1); // k_patch_ws: loads of the next element are issued between two blocks
            if (MADB_SF2D_FENCE && (i1 + j1) > 0)
            {
#pragma unroll
               for (int q2 = 0; q2 < NQ; q2++)
               {
#pragma unroll
                  for (int q1 = 0; q1 < NQ; q1++)
                  {
                     if (!H00[q2][q1].z) { asm volatile("" : "+d"(H00[q2][q1].v)::"memory"); }
                     if (!H01[q2][q1].z) { asm volatile("" : "+d"(H01[q2][q1].v)::"memory"); }
                     if (!H11[q2][q1].z) { asm volatile("" : "+d"(H11[q2][q1].v)::"memory"); }
                  }
               }
            }
            // per q2: Tab = sum_q1 (1-D products)[q1][i1][j1] H^ab(q2,q1); kept as even / odd parts over q2 (mirror
            // q2' = NQ-1-q2): P multiplies the symmetrised tables, M the antisymmetrised ones; T01 / T10 multiply
            // B G products, which change sign under the mirror (madb_sf2d.cuh)
            constexpr int NQH = (NQ + 1) / 2;
            ZD P00[NQH], P01[NQH], P10[NQH], P11[NQH], M00[NQH], M01[NQH], M10[NQH], M11[NQH];
#pragma unroll
            for (int q = 0; q < NQH; q++)
            {
               ZD t00[2], t01[2], t10[2], t11[2];
#pragma unroll
               for (int h = 0; h < 2; h++)
               {
                  const int q2 = h ? NQ - 1 - q : q;
                  ZD s00 {0.0, true}, s01 {0.0, true}, s10 {0.0, true}, s11 {0.0, true};
                  if (h == 0 || q2 != q)
                  {
#pragma unroll
                     for (int q1 = 0; q1 < NQ; q1++)
                     {
                        s00 = zfmac(H00[q2][q1], T.GG[q1][i1][j1], s00);
                        s01 = zfmac(H01[q2][q1], T.BG[q1][j1][i1], s01); // G[q1][i1] B[q1][j1]
                        if (i1 != j1) { s10 = zfmac(H01[q2][q1], T.BG[q1][i1][j1], s10); } // G[q1][j1] B[q1][i1]
                        s11 = zfmac(H11[q2][q1], T.BB[q1][i1][j1], s11);
                     }
                  }
                  t00[h] = s00;
                  t01[h] = s01;
                  t10[h] = (i1 != j1) ? s10 : s01;
                  t11[h] = s11;
               }
               if (q == NQ - 1 - q)
               {
                  P00[q] = M00[q] = t00[0];
                  P01[q] = M01[q] = t01[0];
                  P10[q] = M10[q] = t10[0];
                  P11[q] = M11[q] = t11[0];
               }
               else
               {
                  P00[q] = zadd(t00[0], t00[1]);
                  M00[q] = zsub(t00[0], t00[1]);
                  P11[q] = zadd(t11[0], t11[1]);
                  M11[q] = zsub(t11[0], t11[1]);
                  P01[q] = zsub(t01[0], t01[1]);
                  M01[q] = zadd(t01[0], t01[1]);
                  if (i1 != j1)
                  {
                     P10[q] = zsub(t10[0], t10[1]);
                     M10[q] = zadd(t10[0], t10[1]);
                  }
                  else
                  {
                     P10[q] = P01[q];
                     M10[q] = M01[q];
                  }
               }
            }
            mid_matrix(200 + j1 * (j1 + 1) / 2 + i1);
#pragma unroll
            for (int i2 = 0; i2 < ND; i2++)
            {
#pragma unroll
               for (int j2 = 0; j2 < ND; j2++)
               {
                  if (i1 == j1 && i2 > j2) { continue; }
                  // entry (I, J), I = (i2, i1), J = (j2, j1), and its mirror partner (i2', j2') (for i1 == j1 the block
                  // is symmetric and the partner is stored as its transpose when i2' > j2')
                  int pi = ND - 1 - i2, pj = ND - 1 - j2;
                  if (i1 == j1 && pi > pj)
                  {
                     const int tmp = pi;
                     pi = pj;
                     pj = tmp;
                  }
                  const int lin = i2 * ND + j2, plin = pi * ND + pj;
                  if (plin < lin) { continue; } // emitted together with its partner
                  double se = 0.0;
#pragma unroll
                  for (int q = 0; q < NQH; q++)
                  {
                     zacc(se, P00[q], T.BBs[q][i2][j2]);
                     zacc(se, P01[q], T.BGs[q][i2][j2]);
                     zacc(se, P10[q], T.BGs[q][j2][i2]);
                     zacc(se, P11[q], T.GGs[q][i2][j2]);
                  }
                  const int I = i2 * ND + i1, J = j2 * ND + j1;
                  if (plin == lin) { sink(symidx(I, J), se); }
                  else
                  {
                     double so = 0.0;
#pragma unroll
                     for (int q = 0; q < NQH; q++)
                     {
                        zacc(so, M00[q], T.BBd[q][i2][j2]);
                        zacc(so, M01[q], T.BGd[q][i2][j2]);
                        zacc(so, M10[q], T.BGd[q][j2][i2]);
                        zacc(so, M11[q], T.GGd[q][i2][j2]);
                     }
                     sink(symidx(I, J), se + so);
                     sink(symidx(pi * ND + i1, pj * ND + j1), se - so);
                  }
               }
            }
         }
      }
   }
}

/// gather + computation of sorted element t
template <class Func, class Cfg, int MODE, class Sink, class HookPre = NoHook>
__device__ __forceinline__ void element_compute_sf2d(const AsmArgs<Func, Cfg> &a, const int t,
                                                     double (&r)[(MODE & (MODE_RES | MODE_ACT)) ? Cfg::NVD : 1], Sink &&sink,
                                                     HookPre &&pre_matrix = HookPre())
{
   Sf2dIn<Cfg, (MODE & MODE_ACT) != 0> in;
   sf2d_gather<Func, Cfg, (MODE & MODE_ACT) != 0>(a, t, in);
   element_compute_sf2d_core<Func, Cfg, MODE>(a, in, r, sink, pre_matrix);
}

/// does <functional, configuration, mode> take the sum-factorised 2-D path?
template <class Func, class Cfg, int MODE> constexpr bool use_sf2d()
{
#ifdef MADB_NO_SF2D
   return false;
#else
   return sf2d_cfg<Cfg>() && Func::N_QPRM == 0 && (MODE == MODE_RES || MODE == (MODE_RES | MODE_JAC) || MODE == MODE_ACT);
#endif
}

/// Gather + quadrature loop of sorted element t: element vector r, upper triangle of the element matrix A
/// (columns b in [B0, B1) only).
template <class Func, class Cfg, int MODE, bool UNROLLQ, int B0 = 0, int B1 = Cfg::NVD>
__device__ __forceinline__ void element_compute(const AsmArgs<Func, Cfg> &a, const Tables<Cfg> &tab, const int t,
                                                double (&r)[(MODE & (MODE_RES | MODE_ACT)) ? Cfg::NVD : 1],
                                                double (&A)[(MODE & MODE_JAC) ? Cfg::NSYM : 1], double &energy)
{
   constexpr int DIM = Cfg::DIM, NVD = Cfg::NVD;
   if constexpr (use_sf2d<Func, Cfg, MODE>())
   {
      energy = 0.0;
      element_compute_sf2d<Func, Cfg, MODE>(a, t, r, [&](int k, double v) { if constexpr ((MODE & MODE_JAC) != 0) { A[k] = v; } });
      return;
   }

   // ---- gather: vertices, dofs of all fields ---------------------------------------
   double X[Cfg::NGN][DIM];
#pragma unroll
   for (int k = 0; k < Cfg::NGN; k++)
   {
      const int n = a.e2n[(size_t)k * a.stride + t];
#pragma unroll
      for (int d = 0; d < DIM; d++) { X[k][d] = a.coords[(size_t)n * DIM + d]; }
   }
   double u[Cfg::NDOF_ALL];
   double vdir[(MODE & MODE_ACT) ? NVD : 1];
   static_for<Cfg::NF>([&](auto F)
   {
      constexpr int fi = decltype(F)::value;
      using Fd = typename Cfg::template field<fi>;
      constexpr int n = Cfg::template nd<fi>() * Fd::VDIM, doff = Cfg::template doff<fi>();
      if constexpr (Cfg::template is_input<fi>())
      {
         constexpr int vo = Cfg::template voff<fi>();
#pragma unroll
         for (int i = 0; i < n; i++)
         {
            const int idx = a.vmap[(size_t)(vo + i) * a.stride + t] & 0x7fffffff;
            u[doff + i] = a.x[idx];
            if constexpr ((MODE & MODE_ACT) != 0) { vdir[vo + i] = a.v[idx]; }
         }
      }
      else
      {
         constexpr int po = doff - Cfg::template voff<fi>();
#pragma unroll
         for (int i = 0; i < n; i++) { u[doff + i] = a.pdata[fi][a.pmap[(size_t)(po + i) * a.stride + t]]; }
      }
   });

   Func f;
   f.load(a.fparams);

   energy = 0.0;
   if constexpr ((MODE & (MODE_RES | MODE_ACT)) != 0)
   {
#pragma unroll
      for (int i = 0; i < NVD; i++) { r[i] = 0.0; }
   }
   if constexpr ((MODE & MODE_JAC) != 0)
   {
#pragma unroll
      for (int i = 0; i < Cfg::NSYM; i++) { A[i] = 0.0; }
   }

   if constexpr (UNROLLQ)
   {
#pragma unroll
      for (int q = 0; q < Cfg::NQ; q++) { qpoint<Func, Cfg, MODE, B0, B1>(a, tab, q, t, X, u, vdir, f, r, A, energy); }
   }
   else
   {
      MADB_UNROLL(MADB_QLOOP_UNROLL)
      for (int q = 0; q < Cfg::NQ; q++) { qpoint<Func, Cfg, MODE, B0, B1>(a, tab, q, t, X, u, vdir, f, r, A, energy); }
   }
}

/// shapedim of the field that element-vector index b belongs to
template <class Cfg, int F = 0> MADB_HD constexpr int sd_of_vdof(int b)
{
   if constexpr (F >= Cfg::NF) { return 1; }
   else
   {
      if constexpr (Cfg::template is_input<F>())
      {
         constexpr int n = Cfg::template nd<F>() * Cfg::template field<F>::VDIM, vo = Cfg::template voff<F>();
         if (b >= vo && b < vo + n) { return Cfg::template sd<F>(); }
      }
      return sd_of_vdof<Cfg, F + 1>(b);
   }
}
/// FP64 operations per point spent on column b of the upper triangle: tcol (N_INPUT * sd_b) + entries (sum_{a<=b} sd_a)
template <class Cfg> MADB_HD constexpr int column_cost(int b)
{
   int c = Cfg::N_INPUT * sd_of_vdof<Cfg>(b);
   for (int a = 0; a <= b; a++) { c += sd_of_vdof<Cfg>(a); }
   return c;
}
/// First column of slice k when the upper triangle of the element matrix is cut into nparts slices of about
/// equal cost; slice 0 also carries the element vector (about 2 sum_b sd_b operations per point).
template <class Cfg> MADB_HD constexpr int tri_split(int nparts, int k)
{
   constexpr int n = Cfg::NVD;
   if (k >= nparts) { return n; }
   int res = 0, total = 0;
   for (int b = 0; b < n; b++) { res += 2 * sd_of_vdof<Cfg>(b); }
   total = res;
   for (int b = 0; b < n; b++) { total += column_cost<Cfg>(b); }
   int cum = res, b = 0;
   while (b < n && cum * nparts < k * total) { cum += column_cost<Cfg>(b); b++; }
   return k == 0 ? 0 : b;
}
/// threads per element in the colour-scatter kernel: large element matrices (ex4 / ex5 blocks, order-2
/// elasticity) are split so that a thread's slice of the upper triangle stays in registers
template <class Cfg, int MODE> constexpr int element_parts() { return ((MODE & MODE_JAC) != 0 && Cfg::NVD > 12) ? 4 : 1; }

/// compute + colour scatter of slice PART of NPART of sorted element t
template <class Func, class Cfg, int MODE, bool UNROLLQ, int PART, int NPART>
__device__ __forceinline__ void element_body(const AsmArgs<Func, Cfg> &a, const int t)
{
   constexpr int NVD = Cfg::NVD;
   constexpr int B0 = tri_split<Cfg>(NPART, PART), B1 = tri_split<Cfg>(NPART, PART + 1);
   // slices other than the first do not touch the element vector
   constexpr int PMODE = (PART == 0) ? MODE : (MODE & ~(MODE_RES | MODE_ACT));
   double r[(PMODE & (MODE_RES | MODE_ACT)) ? NVD : 1];
   double A[(PMODE & MODE_JAC) ? Cfg::NSYM : 1];
   double energy;
   element_compute<Func, Cfg, PMODE, UNROLLQ, B0, B1>(a, a.tab, t, r, A, energy);

   // ---- scatter --------------------------------------------------------------------
   // Map entries are read through the read-only path in row batches, ahead of the
   // dependent loads/stores, so that a warp keeps NVD independent requests in flight
   // (the first version serialised 81 map-load -> store round trips: 75 % of all
   // stall samples, profiles/r01_v1_k_element.md).
   if constexpr (PMODE == MODE_ENERGY) { a.energy[t] = energy; }
   if constexpr ((PMODE & (MODE_RES | MODE_ACT)) != 0)
   {
      if (a.write_y)
      {
         double *__restrict__ y = a.y;
         int m[NVD];
         double old[NVD];
#pragma unroll
         for (int i = 0; i < NVD; i++) { m[i] = __ldg(a.vmap + (size_t)i * a.stride + t); }
#pragma unroll
         for (int i = 0; i < NVD; i++) { old[i] = (m[i] < 0) ? 0.0 : y[m[i] & 0x7fffffff]; }
#pragma unroll
         for (int i = 0; i < NVD; i++) { y[m[i] & 0x7fffffff] = old[i] + r[i]; }
      }
   }
   if constexpr ((PMODE & MODE_JAC) != 0)
   {
      if (a.write_vals)
      {
         double *__restrict__ vals = a.vals;
         const int *__restrict__ e2csr = a.e2csr + t;
         // entries (i, j) of this slice: max(i, j) in [B0, B1)
#pragma unroll
         for (int i = 0; i < NVD; i++)
         {
            const int j0 = (i >= B0) ? 0 : B0, j1 = (i >= B1) ? 0 : B1;
            int m[NVD];
            double old[NVD];
#pragma unroll
            for (int j = 0; j < NVD; j++) { if (j >= j0 && j < j1) { m[j] = __ldg(e2csr + (size_t)(i * NVD + j) * a.stride); } }
#pragma unroll
            for (int j = 0; j < NVD; j++) { if (j >= j0 && j < j1) { old[j] = (m[j] < 0) ? 0.0 : vals[m[j] & 0x7fffffff]; } }
#pragma unroll
            for (int j = 0; j < NVD; j++) { if (j >= j0 && j < j1) { vals[m[j] & 0x7fffffff] = old[j] + A[symidx(i, j)]; } }
         }
      }
   }
}

/// Colour-scatter kernel: one launch per colour, blockIdx.y = slice of the element matrix (element_parts).
template <class Func, class Cfg, int MODE, bool UNROLLQ>
__global__ void __launch_bounds__(128) k_element(const __grid_constant__ AsmArgs<Func, Cfg> a)
{
   constexpr int NPART = element_parts<Cfg, MODE>();
   const int t = a.begin + blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= a.end) { return; }
   if constexpr (NPART == 1) { element_body<Func, Cfg, MODE, UNROLLQ, 0, 1>(a, t); }
   else
   {
      static_assert(NPART == 4, "slices are dispatched on blockIdx.y");
      switch (blockIdx.y)
      {
         case 0: element_body<Func, Cfg, MODE, UNROLLQ, 0, 4>(a, t); break;
         case 1: element_body<Func, Cfg, MODE, UNROLLQ, 1, 4>(a, t); break;
         case 2: element_body<Func, Cfg, MODE, UNROLLQ, 2, 4>(a, t); break;
         default: element_body<Func, Cfg, MODE, UNROLLQ, 3, 4>(a, t); break;
      }
   }
}

} // namespace madb

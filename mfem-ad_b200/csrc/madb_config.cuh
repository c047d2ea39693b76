// madb_config.cuh -- compile-time description of an element configuration.
//
// A Config<DIM, NQ1D, Field...> fixes what the reference decides at run time
// inside InitInputShapes / CalcInputShapes (src/ad_intg.hpp:68-154, :363-466):
// which finite-element quantities (ADEval flags, src/_ad_intg.hpp:24-36) of
// which spaces feed the functional, and the layout of the input vector
//    x = [ space0: comp0 [value? grad?], comp1 ... | space1: ... ]
// i.e. x[xoff_s + k + shapedim_s * c]   (src/ad_intg.hpp:228,242,560-566).
// PARAM fields are interpolated like inputs but land in the per-point
// parameter vector (Evaluator GridFunction sources, src/ad_native.cpp:166-171)
// and carry no test functions.
#pragma once
#include "madb_ad.cuh"
#include <type_traits>
#include <utility>

namespace madb
{

enum : unsigned // ADEval, src/_ad_intg.hpp:24-36
{
   EV_QVALUE = 1u << 0, EV_VALUE = 1u << 1, EV_GRAD = 1u << 2, EV_DIV = 1u << 3,
   EV_CURL = 1u << 4, EV_HESSIAN = 1u << 5, EV_VECTOR = 1u << 6, EV_VECFE = 1u << 7
};
enum { ROLE_INPUT = 0, ROLE_PARAM = 1 };

MADB_HD constexpr int ipow(int b, int e) { return e == 0 ? 1 : b * ipow(b, e - 1); }

template <int ND1D_, int VDIM_, unsigned MODE_, int ROLE_ = ROLE_INPUT> struct Field
{
   static constexpr int ND1D = ND1D_, VDIM = VDIM_, ROLE = ROLE_;
   static constexpr unsigned MODE = MODE_;
   static constexpr bool HAS_VALUE = (MODE_ & EV_VALUE) != 0, HAS_GRAD = (MODE_ & EV_GRAD) != 0;
};

template <int I> using IC = std::integral_constant<int, I>;
template <class F, int... Is> MADB_HD void static_for_impl(F &&f, std::integer_sequence<int, Is...>) { (f(IC<Is> {}), ...); }
template <int N, class F> MADB_HD void static_for(F &&f) { static_for_impl(f, std::make_integer_sequence<int, N> {}); }

template <int I, class... Ts> struct type_at;
template <class T, class... Ts> struct type_at<0, T, Ts...> { using type = T; };
template <int I, class T, class... Ts> struct type_at<I, T, Ts...> { using type = typename type_at<I - 1, Ts...>::type; };

template <int DIM_, int NQ1D_, class... Fs> struct Config
{
   static constexpr bool TENSOR = true;
   static constexpr int DIM = DIM_, NQ1D = NQ1D_, NF = sizeof...(Fs);
   static constexpr int NQ = ipow(NQ1D_, DIM_);
   static constexpr int NGN = ipow(2, DIM_); // geometry nodes (order-1 isoparametric map)
   template <int F> using field = typename type_at<F, Fs...>::type;

   template <int F> static constexpr int nd() { return ipow(field<F>::ND1D, DIM); }
   template <int F> static constexpr int sd() { return (field<F>::HAS_VALUE ? 1 : 0) + (field<F>::HAS_GRAD ? DIM : 0); }
   template <int F> static constexpr int nslots() { return sd<F>() * field<F>::VDIM; }
   template <int F> static constexpr bool is_input() { return field<F>::ROLE == ROLE_INPUT; }

   // offsets: inputs -> x, params -> qprm
   template <int F> static constexpr int xoff()
   {
      if constexpr (F == 0) { return 0; }
      else { return xoff<F - 1>() + (is_input<F - 1>() ? nslots<F - 1>() : 0); }
   }
   template <int F> static constexpr int poff()
   {
      if constexpr (F == 0) { return 0; }
      else { return poff<F - 1>() + (is_input<F - 1>() ? 0 : nslots<F - 1>()); }
   }
   // element-vector offset (inputs only): [field][comp][dof]
   template <int F> static constexpr int voff()
   {
      if constexpr (F == 0) { return 0; }
      else { return voff<F - 1>() + (is_input<F - 1>() ? nd<F - 1>() * field<F - 1>::VDIM : 0); }
   }
   // offset of the field's dofs in the gathered-dof array (all fields)
   template <int F> static constexpr int doff()
   {
      if constexpr (F == 0) { return 0; }
      else { return doff<F - 1>() + nd<F - 1>() * field<F - 1>::VDIM; }
   }
   // offset of the field's basis tables
   template <int F> static constexpr int toff()
   {
      if constexpr (F == 0) { return 0; }
      else { return toff<F - 1>() + nd<F - 1>(); }
   }
   static constexpr int N_INPUT = xoff<NF>();
   static constexpr int N_FIELD_QPRM = poff<NF>();
   static constexpr int NVD = voff<NF>();    // element vector size
   static constexpr int NDOF_ALL = doff<NF>(); // gathered dofs incl. parameter fields
   static constexpr int NTAB = toff<NF>();   // sum of nd over fields
   static constexpr int NSYM = NVD * (NVD + 1) / 2;
};

// ---- simplex elements (triangles: ex5.cpp:72-73) ------------------------------------------------
// Non-tensor bases: a field is described by its number of scalar dofs per element (3 for P1, 6 for P2, 1 for the
// constant), the rule by its number of points (MFEM's triangle rules: 6 points for order 4, 12 for order 6).  Only the
// table-driven generic element computation applies (no sum factorisation); the geometry is the affine map of the
// 3 vertices.  Same static interface as Config.
template <int ND_, int VDIM_, unsigned MODE_, int ROLE_ = ROLE_INPUT> struct SField
{
   static constexpr int ND = ND_, ND1D = 0, VDIM = VDIM_, ROLE = ROLE_;
   static constexpr unsigned MODE = MODE_;
   static constexpr bool HAS_VALUE = (MODE_ & EV_VALUE) != 0, HAS_GRAD = (MODE_ & EV_GRAD) != 0;
};
template <int NQ_, class... Fs> struct SConfig
{
   static constexpr bool TENSOR = false;
   static constexpr int DIM = 2, NQ1D = 0, NF = sizeof...(Fs);
   static constexpr int NQ = NQ_;
   static constexpr int NGN = 3;
   template <int F> using field = typename type_at<F, Fs...>::type;
   template <int F> static constexpr int nd() { return field<F>::ND; }
   template <int F> static constexpr int sd() { return (field<F>::HAS_VALUE ? 1 : 0) + (field<F>::HAS_GRAD ? DIM : 0); }
   template <int F> static constexpr int nslots() { return sd<F>() * field<F>::VDIM; }
   template <int F> static constexpr bool is_input() { return field<F>::ROLE == ROLE_INPUT; }
   template <int F> static constexpr int xoff()
   {
      if constexpr (F == 0) { return 0; }
      else { return xoff<F - 1>() + (is_input<F - 1>() ? nslots<F - 1>() : 0); }
   }
   template <int F> static constexpr int poff()
   {
      if constexpr (F == 0) { return 0; }
      else { return poff<F - 1>() + (is_input<F - 1>() ? 0 : nslots<F - 1>()); }
   }
   template <int F> static constexpr int voff()
   {
      if constexpr (F == 0) { return 0; }
      else { return voff<F - 1>() + (is_input<F - 1>() ? nd<F - 1>() * field<F - 1>::VDIM : 0); }
   }
   template <int F> static constexpr int doff()
   {
      if constexpr (F == 0) { return 0; }
      else { return doff<F - 1>() + nd<F - 1>() * field<F - 1>::VDIM; }
   }
   template <int F> static constexpr int toff()
   {
      if constexpr (F == 0) { return 0; }
      else { return toff<F - 1>() + nd<F - 1>(); }
   }
   static constexpr int N_INPUT = xoff<NF>();
   static constexpr int N_FIELD_QPRM = poff<NF>();
   static constexpr int NVD = voff<NF>();
   static constexpr int NDOF_ALL = doff<NF>();
   static constexpr int NTAB = toff<NF>();
   static constexpr int NSYM = NVD * (NVD + 1) / 2;
};

/// Basis tables at the quadrature points, passed by value in the kernel
/// parameter space so that fully unrolled code reads them as constant-bank
/// operands.  phi[q][i] = value of scalar basis i of the field at point q
/// (lexicographic, x fastest -- the reference's tensor rule order),
/// dphi[q][i][k] = reference-space derivative d/dxi_k.
template <class Cfg> struct Tables
{
   double phi[Cfg::NQ][Cfg::NTAB];
   double dphi[Cfg::NQ][Cfg::NTAB][Cfg::DIM];
   double gdphi[Cfg::NQ][Cfg::NGN][Cfg::DIM]; // geometry (order-1) basis derivatives
   double w[Cfg::NQ];                         // reference quadrature weights
};

} // namespace madb

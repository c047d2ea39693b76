// madb_ad.cuh -- forward-mode AD scalars evaluated per quadrature point in registers.
//
// Replaces the reference's seed-by-seed nested duals
//   ADReal_t  = future::dual<real_t,real_t>       (src/ad_native.hpp:42)
//   AD2Real_t = future::dual<ADReal_t,ADReal_t>   (src/ad_native.hpp:47)
// and the n / n(n+1)/2 re-evaluation loops of ADFunction::Gradient / Hessian
// (src/ad_native.cpp:188-201, :211-230) by ONE pass with a vector-tangent
// dual (value + N gradients) or hyper-dual (value + N gradients + packed
// symmetric N x N second derivatives).  Results agree with the reference to
// rounding (addition order only).
//
// nvcc does not fold 0.0*x or x+0.0 in FP64 (no nnan/nsz switch), so every
// derivative slot carries a "structurally zero" flag.  Inputs are seeded with
// compile-time flags; after full inlining/unrolling the flags are constants,
// the branches fold away and only structurally non-zero arithmetic is emitted
// (checked in SASS: minimal-surface Hessian = 62 FP64 ops vs 42 hand-written).
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define MADB_HD __host__ __device__ __forceinline__
#else
#define MADB_HD inline
#endif

namespace madb
{

// Branch-free reciprocal / reciprocal square root for NORMAL positive-or-negative arguments away from the
// over/underflow range: MUFU.RCP64H / MUFU.RSQ64H seed (about 20 bits) + two Newton steps (quadratic: 2^-40, 2^-80 ->
// rounding error only).  CUDA's 1.0/x and rsqrt(x) carry a slow-path test and CALL per use: every use splits the
// unrolled quadrature loop into basic blocks and ptxas cannot interleave the FP64 chains of neighbouring points
// (k_patch_ws: 64 such blocks per element).  0, inf, NaN and denormal arguments give inf/NaN, not IEEE results:
// use them only where the argument is known to be regular (Jacobian determinants, arguments tested by the caller).
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double frcp(double x)
{
   double r;
   asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
   double e = fma(-x, r, 1.0);
   r = fma(r, e, r);
   e = fma(-x, r, 1.0);
   return fma(r, e, r);
}
__device__ __forceinline__ double frsqrt(double x)
{
   double r;
   asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
   const double hx = 0.5 * x;
   double e = fma(-hx * r, r, 0.5); // (1 - x r^2) / 2
   r = fma(r, e, r);
   e = fma(-hx * r, r, 0.5);
   return fma(r, e, r);
}
#else
inline double frcp(double x) { return 1.0 / x; }
inline double frsqrt(double x) { return 1.0 / std::sqrt(x); }
#endif

// (value, structurally-zero) pair used inside the operators
struct ZD
{
   double v;
   bool z;
};
MADB_HD ZD zmul(ZD a, ZD b) { if (a.z || b.z) { return {0.0, true}; } return {a.v * b.v, false}; }
MADB_HD ZD zmulc(ZD a, double b) { if (a.z) { return {0.0, true}; } return {a.v * b, false}; }
MADB_HD ZD zadd(ZD a, ZD b) { if (a.z) { return b; } if (b.z) { return a; } return {a.v + b.v, false}; }
MADB_HD ZD zsub(ZD a, ZD b) { if (b.z) { return a; } if (a.z) { return {-b.v, false}; } return {a.v - b.v, false}; }
MADB_HD ZD zneg(ZD a) { if (a.z) { return a; } return {-a.v, false}; }
// a*b + c with fused multiply-add when all are live
MADB_HD ZD zfma(ZD a, ZD b, ZD c)
{
   if (a.z || b.z) { return c; }
   if (c.z) { return {a.v * b.v, false}; }
   return {fma(a.v, b.v, c.v), false};
}
MADB_HD ZD zfmac(ZD a, double b, ZD c)
{
   if (a.z) { return c; }
   if (c.z) { return {a.v * b, false}; }
   return {fma(a.v, b, c.v), false};
}

/// acc += a*b, skipped when a is structurally zero
MADB_HD void zacc(double &acc, ZD a, double b) { if (!a.z) { acc = fma(a.v, b, acc); } }

// packed upper-triangular index of (i,j), i<=j, row-major
template <int N> MADB_HD constexpr int hidx(int i, int j) { return i * N - (i * (i - 1)) / 2 + (j - i); }

template <int N> MADB_HD constexpr int symidx_h(int i, int j) { return i <= j ? hidx<N>(i, j) : hidx<N>(j, i); }

/// AD scalar of differentiation order ORDER (1: dual, 2: hyper-dual) in N variables
template <int N, int ORDER> struct AD
{
   static constexpr int NV = N;
   static constexpr int NS = (ORDER >= 2) ? N * (N + 1) / 2 : 1;
   double v;
   double g[N];
   double h[NS];
   bool gz[N];
   bool hz[NS];

   MADB_HD AD() : v(0.0)
   {
#pragma unroll
      for (int i = 0; i < N; i++) { g[i] = 0.0; gz[i] = true; }
#pragma unroll
      for (int k = 0; k < NS; k++) { h[k] = 0.0; hz[k] = true; }
   }
   MADB_HD AD(double a) : AD() { v = a; }
   MADB_HD AD(int a) : AD() { v = a; }
   MADB_HD ZD G(int i) const { return {g[i], gz[i]}; }
   MADB_HD ZD H(int k) const { return {h[k], hz[k]}; }
   MADB_HD void setG(int i, ZD t) { g[i] = t.v; gz[i] = t.z; }
   MADB_HD void setH(int k, ZD t) { h[k] = t.v; hz[k] = t.z; }
   /// second derivative d2/dxi dxj (either order)
   MADB_HD double hess(int i, int j) const { return (i <= j) ? h[hidx<N>(i, j)] : h[hidx<N>(j, i)]; }
};

/// independent variable k with value x (the reference's seeding, src/ad_native.cpp:196,219-222)
template <int N, int ORDER> MADB_HD AD<N, ORDER> ad_seed(double x, int k)
{
   AD<N, ORDER> r(x);
   r.g[k] = 1.0;
   r.gz[k] = false;
   return r;
}

template <class T> struct ad_traits
{
   static constexpr int order = 0;
   static constexpr int n = 0;
};
template <int N, int O> struct ad_traits<AD<N, O>>
{
   static constexpr int order = O;
   static constexpr int n = N;
};

MADB_HD double ad_value(double a) { return a; }
template <int N, int O> MADB_HD double ad_value(const AD<N, O> &a) { return a.v; }

// ---- addition / subtraction ------------------------------------------------
template <int N, int O> MADB_HD AD<N, O> operator+(const AD<N, O> &a, const AD<N, O> &b)
{
   AD<N, O> r;
   r.v = a.v + b.v;
#pragma unroll
   for (int i = 0; i < N; i++) { r.setG(i, zadd(a.G(i), b.G(i))); }
   if constexpr (O >= 2)
   {
#pragma unroll
      for (int k = 0; k < AD<N, O>::NS; k++) { r.setH(k, zadd(a.H(k), b.H(k))); }
   }
   return r;
}
template <int N, int O> MADB_HD AD<N, O> operator-(const AD<N, O> &a, const AD<N, O> &b)
{
   AD<N, O> r;
   r.v = a.v - b.v;
#pragma unroll
   for (int i = 0; i < N; i++) { r.setG(i, zsub(a.G(i), b.G(i))); }
   if constexpr (O >= 2)
   {
#pragma unroll
      for (int k = 0; k < AD<N, O>::NS; k++) { r.setH(k, zsub(a.H(k), b.H(k))); }
   }
   return r;
}
template <int N, int O> MADB_HD AD<N, O> operator-(const AD<N, O> &a)
{
   AD<N, O> r;
   r.v = -a.v;
#pragma unroll
   for (int i = 0; i < N; i++) { r.setG(i, zneg(a.G(i))); }
   if constexpr (O >= 2)
   {
#pragma unroll
      for (int k = 0; k < AD<N, O>::NS; k++) { r.setH(k, zneg(a.H(k))); }
   }
   return r;
}
template <int N, int O> MADB_HD AD<N, O> operator+(const AD<N, O> &a, double b) { AD<N, O> r = a; r.v = a.v + b; return r; }
template <int N, int O> MADB_HD AD<N, O> operator+(double a, const AD<N, O> &b) { AD<N, O> r = b; r.v = a + b.v; return r; }
template <int N, int O> MADB_HD AD<N, O> operator-(const AD<N, O> &a, double b) { AD<N, O> r = a; r.v = a.v - b; return r; }
template <int N, int O> MADB_HD AD<N, O> operator-(double a, const AD<N, O> &b) { AD<N, O> r = -b; r.v = a - b.v; return r; }

// ---- multiplication --------------------------------------------------------
template <int N, int O> MADB_HD AD<N, O> operator*(const AD<N, O> &a, const AD<N, O> &b)
{
   AD<N, O> r;
   r.v = a.v * b.v;
#pragma unroll
   for (int i = 0; i < N; i++) { r.setG(i, zfmac(a.G(i), b.v, zmulc(b.G(i), a.v))); }
   if constexpr (O >= 2)
   {
#pragma unroll
      for (int i = 0; i < N; i++)
      {
#pragma unroll
         for (int j = i; j < N; j++)
         {
            const int k = hidx<N>(i, j);
            ZD t = zfmac(a.H(k), b.v, zmulc(b.H(k), a.v));
            t = zfma(a.G(i), b.G(j), t);
            t = zfma(a.G(j), b.G(i), t);
            r.setH(k, t);
         }
      }
   }
   return r;
}
template <int N, int O> MADB_HD AD<N, O> operator*(const AD<N, O> &a, double b)
{
   AD<N, O> r;
   r.v = a.v * b;
#pragma unroll
   for (int i = 0; i < N; i++) { r.setG(i, zmulc(a.G(i), b)); }
   if constexpr (O >= 2)
   {
#pragma unroll
      for (int k = 0; k < AD<N, O>::NS; k++) { r.setH(k, zmulc(a.H(k), b)); }
   }
   return r;
}
template <int N, int O> MADB_HD AD<N, O> operator*(double a, const AD<N, O> &b) { return b * a; }

// ---- unary chain rule: r = f(a) given f, f', f'' at a.v ---------------------
template <int N, int O> MADB_HD AD<N, O> ad_chain(const AD<N, O> &a, double f, double f1, double f2)
{
   AD<N, O> r;
   r.v = f;
#pragma unroll
   for (int i = 0; i < N; i++) { r.setG(i, zmulc(a.G(i), f1)); }
   if constexpr (O >= 2)
   {
#pragma unroll
      for (int i = 0; i < N; i++)
      {
#pragma unroll
         for (int j = i; j < N; j++)
         {
            const int k = hidx<N>(i, j);
            r.setH(k, zfmac(zmul(a.G(i), a.G(j)), f2, zmulc(a.H(k), f1)));
         }
      }
   }
   return r;
}

// ---- division --------------------------------------------------------------
template <int N, int O> MADB_HD AD<N, O> ad_inv(const AD<N, O> &b)
{
   const double i1 = 1.0 / b.v;
   const double i2 = i1 * i1;
   return ad_chain(b, i1, -i2, 2.0 * i2 * i1);
}
template <int N, int O> MADB_HD AD<N, O> operator/(const AD<N, O> &a, const AD<N, O> &b) { return a * ad_inv(b); }
template <int N, int O> MADB_HD AD<N, O> operator/(const AD<N, O> &a, double b) { return a * (1.0 / b); }
template <int N, int O> MADB_HD AD<N, O> operator/(double a, const AD<N, O> &b) { return ad_inv(b) * a; }

// ---- compound assignment ---------------------------------------------------
template <int N, int O, class B> MADB_HD AD<N, O> &operator+=(AD<N, O> &a, const B &b) { a = a + b; return a; }
template <int N, int O, class B> MADB_HD AD<N, O> &operator-=(AD<N, O> &a, const B &b) { a = a - b; return a; }
template <int N, int O, class B> MADB_HD AD<N, O> &operator*=(AD<N, O> &a, const B &b) { a = a * b; return a; }
template <int N, int O, class B> MADB_HD AD<N, O> &operator/=(AD<N, O> &a, const B &b) { a = a / b; return a; }

// ---- comparisons act on values (mfem dual semantics) ------------------------
#define MADB_CMP(op)                                                                                      \
   template <int N, int O> MADB_HD bool operator op(const AD<N, O> &a, const AD<N, O> &b) { return a.v op b.v; } \
   template <int N, int O> MADB_HD bool operator op(const AD<N, O> &a, double b) { return a.v op b; }        \
   template <int N, int O> MADB_HD bool operator op(double a, const AD<N, O> &b) { return a op b.v; }
MADB_CMP(<)
MADB_CMP(>)
MADB_CMP(<=)
MADB_CMP(>=)
MADB_CMP(==)
MADB_CMP(!=)
#undef MADB_CMP

// ---- elementary functions ---------------------------------------------------
template <int N, int O> MADB_HD AD<N, O> sqrt(const AD<N, O> &a)
{
#if defined(__CUDA_ARCH__)
   // one reciprocal square root instead of a square root and two divisions:
   //   r = a^-1/2,  f = a r (one Newton correction),  f' = r/2,  f'' = -r^3/4
   // branch-free (see frsqrt): denormal / huge arguments are rescaled by selects, 0 keeps the IEEE results of the
   // reference (0, inf, -inf), negative arguments give NaN; only +inf differs (NaN instead of inf)
   {
      const bool tiny = a.v < 1e-290, huge = a.v > 1e290;
      const double sc = tiny ? 0x1p600 : (huge ? 0x1p-600 : 1.0), rs = tiny ? 0x1p300 : (huge ? 0x1p-300 : 1.0);
      double r = frsqrt(a.v * sc) * rs;
      r = (a.v == 0.0) ? (double)INFINITY : r;
      double s = a.v * r;
      s = fma(0.5 * r, fma(-s, s, a.v), s);
      s = (a.v == 0.0) ? 0.0 : s;
      const double f1 = 0.5 * r;
      return ad_chain(a, s, f1, -0.5 * f1 * (r * r));
   }
#else
   const double s = ::sqrt(a.v);
   const double f1 = 0.5 / s;
   return ad_chain(a, s, f1, -0.5 * f1 / a.v);
#endif
}
template <int N, int O> MADB_HD AD<N, O> exp(const AD<N, O> &a)
{
   const double e = ::exp(a.v);
   return ad_chain(a, e, e, e);
}
template <int N, int O> MADB_HD AD<N, O> log(const AD<N, O> &a)
{
   const double i1 = 1.0 / a.v;
   return ad_chain(a, ::log(a.v), i1, -i1 * i1);
}
template <int N, int O> MADB_HD AD<N, O> sin(const AD<N, O> &a)
{
   double s, c;
   ::sincos(a.v, &s, &c);
   return ad_chain(a, s, c, -s);
}
template <int N, int O> MADB_HD AD<N, O> cos(const AD<N, O> &a)
{
   double s, c;
   ::sincos(a.v, &s, &c);
   return ad_chain(a, c, -s, -c);
}
/// a^p for a real exponent; small integer exponents (the SIMP penalisation p = 3, src/mmto.hpp:24) by multiplication
MADB_HD double rpow(double a, double p)
{
   if (p == 1.0) { return a; }
   if (p == 2.0) { return a * a; }
   if (p == 3.0) { return a * a * a; }
   if (p == 0.0) { return 1.0; }
   return ::pow(a, p);
}
template <int N, int O> MADB_HD AD<N, O> pow(const AD<N, O> &a, double p);
template <int N, int O> MADB_HD AD<N, O> rpow(const AD<N, O> &a, double p) { return pow(a, p); }

template <int N, int O> MADB_HD AD<N, O> pow(const AD<N, O> &a, double p)
{
   // one transcendental instead of three: a^(p-2) (or a^(p-1) for first order), the rest by multiplication.
   // a == 0 keeps the reference's values pow(0, p), p pow(0, p-1), p (p-1) pow(0, p-2) (src/mmto.hpp:24 uses p = 3).
   if (a.v == 0.0)
   {
      const double f2 = (O >= 2) ? p * (p - 1.0) * ::pow(a.v, p - 2.0) : 0.0;
      return ad_chain(a, ::pow(a.v, p), p * ::pow(a.v, p - 1.0), f2);
   }
   if constexpr (O >= 2)
   {
      const double pm2 = rpow(a.v, p - 2.0), pm1 = pm2 * a.v;
      return ad_chain(a, pm1 * a.v, p * pm1, p * (p - 1.0) * pm2);
   }
   else
   {
      const double pm1 = rpow(a.v, p - 1.0);
      return ad_chain(a, pm1 * a.v, p * pm1, 0.0);
   }
}
using ::cos;
using ::exp;
using ::log;
using ::pow;
using ::sin;
using ::sqrt;

// dual-aware max/min with sub-gradient average on ties: src/ad_native.hpp:695-749
MADB_HD double max(double a, double b) { return a > b ? a : b; }
MADB_HD double min(double a, double b) { return a < b ? a : b; }
template <int N, int O> MADB_HD AD<N, O> max(const AD<N, O> &a, const AD<N, O> &b)
{
   if (a.v > b.v) { return a; }
   else if (a.v < b.v) { return b; }
   else { return 0.5 * (a + b); }
}
template <int N, int O> MADB_HD AD<N, O> min(const AD<N, O> &a, const AD<N, O> &b)
{
   if (a.v < b.v) { return a; }
   else if (a.v > b.v) { return b; }
   else { return 0.5 * (a + b); }
}
template <int N, int O> MADB_HD AD<N, O> max(const AD<N, O> &a, double b)
{
   if (a.v > b) { return a; }
   else if (a.v < b) { return AD<N, O>(b); }
   else { return 0.5 * (a + b); }
}
template <int N, int O> MADB_HD AD<N, O> min(const AD<N, O> &a, double b)
{
   if (a.v < b) { return a; }
   else if (a.v > b) { return AD<N, O>(b); }
   else { return 0.5 * (a + b); }
}

// ---- fixed-size vector with the TAutoDiffVector / mfem::Vector subset that
//      AD_IMPL bodies use (Size, [], (), dot product) -------------------------
template <class T, int N> struct SVec
{
   T d[N];
   MADB_HD static constexpr int Size() { return N; }
   MADB_HD T &operator[](int i) { return d[i]; }
   MADB_HD const T &operator[](int i) const { return d[i]; }
   MADB_HD T &operator()(int i) { return d[i]; }
   MADB_HD const T &operator()(int i) const { return d[i]; }
   MADB_HD T *GetData() { return d; }
   MADB_HD const T *GetData() const { return d; }
};
template <class T, int N> MADB_HD T operator*(const SVec<T, N> &a, const SVec<T, N> &b)
{
   T s = a[0] * b[0];
#pragma unroll
   for (int i = 1; i < N; i++) { s += a[i] * b[i]; }
   return s;
}

} // namespace madb

// madb_functionals.cuh -- device-side pointwise functionals (the AD_IMPL bodies).
//
// Each functional is a trivially-copyable struct
//    static constexpr int N_INPUT   number of differentiated inputs
//    static constexpr int N_PARAM   run-time constants, re-read at every call
//                                   (the reference mutates them between calls:
//                                   ex2.cpp:98 eps, ex4.cpp:187 alpha)
//    static constexpr int N_QPRM    per-quadrature-point parameters
//                                   (= Evaluator::val, src/ad_native.cpp:120-179)
//    void load(const double *p)     consume N_PARAM constants (own first, then children)
//    template <class T> T operator()(const T *x, const double *qp) const
// evaluated with T = double (energy), AD<N,1> (residual), AD<N,2> (Jacobian).
// Composition (ADPGFunctional, DiffEnergy, ...) is static (templates) instead
// of the reference's virtual nested calls (src/pg.hpp:210, src/ad_native.hpp:523).
#pragma once
#include "madb_ad.cuh"

namespace madb
{

// ex0.cpp:20  f = sin(x0) e^{x1} + x2^3
struct Ex0Function
{
   static constexpr int N_INPUT = 3, N_PARAM = 0, N_QPRM = 0;
   MADB_HD void load(const double *) {}
   template <class T> MADB_HD T operator()(const T *x, const double *) const
   {
      return sin(x[0]) * exp(x[1]) + pow(x[2], 3.0);
   }
};

// Probe of the AD square root (a user-body building block; mfem's dual sqrt: value sqrt(a), derivative 0.5/sqrt(a)):
// f = sqrt(x0) x1.  The device sqrt is branch-free (madb_ad.cuh); the probe pins its results at 0, denormal, tiny,
// huge and negative arguments against the IEEE results of the reference's formula (tests/test_gpu_latent.py).
struct SqrtProbe
{
   static constexpr int N_INPUT = 2, N_PARAM = 0, N_QPRM = 0;
   MADB_HD void load(const double *) {}
   template <class T> MADB_HD T operator()(const T *x, const double *) const { return sqrt(x[0]) * x[1]; }
};

// ex0.cpp:23-35  F = (sin(x0 x1), cos(x0 x1 x2))   (ADVectorFunction, AD_VEC_IMPL)
struct Ex0VectorFunction
{
   static constexpr int N_INPUT = 3, N_OUTPUT = 2, N_PARAM = 0;
   MADB_HD void load(const double *) {}
   template <class T> MADB_HD void operator()(const T *x, T *result) const
   {
      result[0] = sin(x[0] * x[1]);
      result[1] = cos(x[0] * x[1] * x[2]);
   }
};

// src/ad_native.hpp:413-420
template <int N> struct MassEnergy
{
   static constexpr int N_INPUT = N, N_PARAM = 0, N_QPRM = 0;
   MADB_HD void load(const double *) {}
   template <class T> MADB_HD T operator()(const T *x, const double *) const
   {
      T s = x[0] * x[0];
#pragma unroll
      for (int i = 1; i < N; i++) { s += x[i] * x[i]; }
      return 0.5 * s;
   }
};

// Linear form (f, v): energy f(x) u with f a per-point parameter (a Coefficient sampled at the points).  Its gradient
// is the load vector MFEM's DomainLFIntegrator assembles (ex4.cpp:145-148); its Hessian is zero.
struct LoadFunctional
{
   static constexpr int N_INPUT = 1, N_PARAM = 0, N_QPRM = 1;
   MADB_HD void load(const double *) {}
   template <class T> MADB_HD T operator()(const T *x, const double *qp) const { return x[0] * qp[0]; }
};

// Vector load (f, v) = sum_c f_c(x) u_c on a vector space with ADEval::VALUE | VECTOR: VectorDomainLFIntegrator (ex3.cpp:64-67)
template <int N> struct VectorLoadFunctional
{
   static constexpr int N_INPUT = N, N_PARAM = 0, N_QPRM = N;
   MADB_HD void load(const double *) {}
   template <class T> MADB_HD T operator()(const T *x, const double *qp) const
   {
      T s = x[0] * qp[0];
#pragma unroll
      for (int c = 1; c < N; c++) { s += x[c] * qp[c]; }
      return s;
   }
};

// src/ad_native.hpp:421-481.  KDIM selects the K kind at compile time
// (0 none, 1 scalar, DIM diagonal, DIM*DIM full, column-major) instead of the
// per-point size dispatch of the reference.  K is a constant here;
// DiffusionEnergyQ reads it from the per-point parameters.
template <int DIM, int KDIM, bool QP = false> struct DiffusionEnergy
{
   static constexpr int N_INPUT = DIM, N_PARAM = QP ? 0 : KDIM, N_QPRM = QP ? KDIM : 0;
   double K[KDIM > 0 ? KDIM : 1];
   MADB_HD void load(const double *p)
   {
      if constexpr (!QP) { for (int i = 0; i < KDIM; i++) { K[i] = p[i]; } }
   }
   template <class T> MADB_HD T operator()(const T *gradu, const double *qp) const
   {
      const double *Kp = QP ? qp : K;
      if constexpr (KDIM == 0 || KDIM == 1)
      {
         T s = gradu[0] * gradu[0];
#pragma unroll
         for (int i = 1; i < DIM; i++) { s += gradu[i] * gradu[i]; }
         if constexpr (KDIM == 0) { return 0.5 * s; }
         else { return (0.5 * Kp[0]) * s; }
      }
      else if constexpr (KDIM == DIM)
      {
         T result = Kp[0] * gradu[0] * gradu[0];
#pragma unroll
         for (int i = 1; i < DIM; i++) { result += Kp[i] * gradu[i] * gradu[i]; }
         return 0.5 * result;
      }
      else
      {
         T result = T();
#pragma unroll
         for (int j = 0; j < DIM; j++)
         {
#pragma unroll
            for (int i = 0; i < DIM; i++) { result += Kp[i + DIM * j] * gradu[i] * gradu[j]; }
         }
         return 0.5 * result;
      }
   }
};

// src/ad_native.hpp:527-566 ; gradu[i*dim + j]
template <int DIM> struct LinearElasticityEnergy
{
   static constexpr int N_INPUT = DIM * DIM, N_PARAM = 2, N_QPRM = 0;
   double lambda, mu;
   MADB_HD void load(const double *p) { lambda = p[0]; mu = p[1]; }
   template <class T> MADB_HD static T body(const T *gradu, double lambda, double mu)
   {
      T divnorm = gradu[0];
#pragma unroll
      for (int i = 1; i < DIM; i++) { divnorm += gradu[i * DIM + i]; }
      divnorm = divnorm * divnorm;
      T h1_norm = T();
#pragma unroll
      for (int i = 0; i < DIM; i++)
      {
#pragma unroll
         for (int j = 0; j < DIM; j++)
         {
            T symm = 0.5 * (gradu[i * DIM + j] + gradu[j * DIM + i]);
            h1_norm += symm * symm;
         }
      }
      return 0.5 * lambda * divnorm + mu * h1_norm;
   }
   template <class T> MADB_HD T operator()(const T *gradu, const double *) const { return body(gradu, lambda, mu); }
};

// ex2.cpp:12-24
template <int DIM> struct MinimalSurfaceEnergy
{
   static constexpr int N_INPUT = DIM, N_PARAM = 1, N_QPRM = 0;
   double eps;
   MADB_HD void load(const double *p) { eps = p[0]; }
   template <class T> MADB_HD T operator()(const T *gradu, const double *) const
   {
      T h1_norm = gradu[0] * gradu[0];
#pragma unroll
      for (int i = 1; i < DIM; i++) { h1_norm += gradu[i] * gradu[i]; }
      return sqrt(h1_norm + 1.0) + eps * h1_norm;
   }
};

// ex4.cpp:15-28 ; x = [u, grad u]
template <int DIM> struct ObstacleEnergy
{
   static constexpr int N_INPUT = DIM + 1, N_PARAM = 0, N_QPRM = 0;
   MADB_HD void load(const double *) {}
   template <class T> MADB_HD T operator()(const T *x, const double *) const
   {
      T result = x[1] * x[1];
#pragma unroll
      for (int i = 2; i < DIM + 1; i++) { result += x[i] * x[i]; }
      return result * 0.5;
   }
};

// ex5.cpp:15-22
template <int DIM> struct GradientObstacleEnergy
{
   static constexpr int N_INPUT = DIM, N_PARAM = 0, N_QPRM = 0;
   MADB_HD void load(const double *) {}
   template <class T> MADB_HD T operator()(const T *x, const double *) const
   {
      T s = x[0] * x[0];
#pragma unroll
      for (int i = 1; i < DIM; i++) { s += x[i] * x[i]; }
      return s * 0.5;
   }
};

// src/_dof_pg.hpp:9-15
template <int N> struct EmptyEnergy
{
   static constexpr int N_INPUT = N, N_PARAM = 0, N_QPRM = 0;
   MADB_HD void load(const double *) {}
   template <class T> MADB_HD T operator()(const T *, const double *) const { return T(); }
};

// src/ad_native.hpp:483-525: energy(x - target); target is a per-point parameter
template <class E> struct DiffEnergy
{
   static constexpr int N_INPUT = E::N_INPUT, N_PARAM = E::N_PARAM, N_QPRM = E::N_INPUT + E::N_QPRM;
   E energy;
   MADB_HD void load(const double *p) { energy.load(p); }
   template <class T> MADB_HD T operator()(const T *x, const double *qp) const
   {
      T diff[N_INPUT];
#pragma unroll
      for (int i = 0; i < N_INPUT; i++) { diff[i] = x[i] - qp[i]; }
      return energy(diff, qp + N_INPUT);
   }
};

// Equality constraints c_0 .. c_{k-1} of Lagrangian / ALFunctional (std::vector<ADFunction*> eq_con in the reference), composed
// statically; parameters of the constraints follow each other.
template <class... Cs> struct ConList
{
   static constexpr int NC = 0, N_PARAM = 0;
   MADB_HD void load(const double *) {}
};
template <class C0, class... Cs> struct ConList<C0, Cs...>
{
   static constexpr int NC = 1 + (int)sizeof...(Cs), N_PARAM = C0::N_PARAM + ConList<Cs...>::N_PARAM;
   C0 c;
   ConList<Cs...> rest;
   MADB_HD void load(const double *p)
   {
      c.load(p);
      rest.load(p + C0::N_PARAM);
   }
   /// c_K(x)
   template <int K, class T> MADB_HD T con(const T *x, const double *qp) const
   {
      if constexpr (K == 0) { return c(x, qp); }
      else { return rest.template con<K - 1>(x, qp); }
   }
   /// sum_i c_i(x) lambda_i  (src/ad_native.hpp:617)
   template <class T> MADB_HD T lagr(const T *x, const T *lambda, const double *qp) const
   {
      T r = c(x, qp) * lambda[0];
      if constexpr (sizeof...(Cs) > 0) { r += rest.lagr(x, lambda + 1, qp); }
      return r;
   }
   /// sum_i c~_i (lambda_i + mu/2 c~_i), c~_i = c_i(x) - rhs_i  (src/ad_native.hpp:684-690)
   template <class T> MADB_HD T al(const T *x, const double *rhs, const double *lambda, const double mu, const double *qp) const
   {
      T cx = c(x, qp) - rhs[0];
      T r = cx * (lambda[0] + (mu * 0.5) * cx);
      if constexpr (sizeof...(Cs) > 0) { r += rest.al(x, rhs + 1, lambda + 1, mu, qp); }
      return r;
   }
   template <int K, class T> MADB_HD T al_con(const T *x, const double *rhs, const double *qp) const
   {
      if constexpr (K == 0) { return c(x, qp) - rhs[0]; }
      else { return rest.template al_con<K - 1>(x, rhs + 1, qp); }
   }
};

// src/ad_native.hpp:570-621: Lagrangian f(x) + sum_i lambda_i c_i(x); inputs [x, lambda_0 .. lambda_{k-1}].
// MODE is the reference's eval_mode (-1 full, -2 objective only, i >= 0 constraint i), a structural integer here.
template <class F, int MODE, class... Cs> struct LagrangianN
{
   static_assert(((F::N_INPUT == Cs::N_INPUT) && ...) && F::N_QPRM == 0 && ((Cs::N_QPRM == 0) && ...), "objective and constraints act on the same x");
   static constexpr int NF = F::N_INPUT, NC = (int)sizeof...(Cs);
   static_assert(MODE < NC, "eval_mode: constraint index out of range");
   static constexpr int N_INPUT = NF + NC, N_PARAM = F::N_PARAM + ConList<Cs...>::N_PARAM, N_QPRM = 0;
   F f;
   ConList<Cs...> cons;
   MADB_HD void load(const double *p)
   {
      f.load(p);
      cons.load(p + F::N_PARAM);
   }
   template <class T> MADB_HD T operator()(const T *x_and_lambda, const double *qp) const
   {
      if constexpr (MODE >= 0) { return cons.template con<MODE>(x_and_lambda, qp); }
      else
      {
         T result = f(x_and_lambda, qp);
         if constexpr (MODE == -2) { return result; }
         else { return result + cons.lagr(x_and_lambda, x_and_lambda + NF, qp); }
      }
   }
};
template <class F, class C, int MODE> using LagrangianOf = LagrangianN<F, MODE, C>;

// src/ad_native.hpp:624-691: augmented Lagrangian f(x) + sum_i c~_i (lambda_i + mu/2 c~_i), c~_i = c_i(x) - rhs_i;
// parameters [mu, rhs_0 .., lambda_0 ..] (SetPenalty / SetEqRHS / SetLambda), then those of f and of the constraints.
template <class F, int MODE, class... Cs> struct ALFunctionalN
{
   static_assert(((F::N_INPUT == Cs::N_INPUT) && ...) && F::N_QPRM == 0 && ((Cs::N_QPRM == 0) && ...), "objective and constraints act on the same x");
   static constexpr int NC = (int)sizeof...(Cs);
   static_assert(MODE < NC, "eval_mode: constraint index out of range");
   static constexpr int N_INPUT = F::N_INPUT, N_PARAM = 1 + 2 * NC + F::N_PARAM + ConList<Cs...>::N_PARAM, N_QPRM = 0;
   double mu, rhs[NC], lambda[NC];
   F f;
   ConList<Cs...> cons;
   MADB_HD void load(const double *p)
   {
      mu = p[0];
      for (int i = 0; i < NC; i++) { rhs[i] = p[1 + i]; lambda[i] = p[1 + NC + i]; }
      f.load(p + 1 + 2 * NC);
      cons.load(p + 1 + 2 * NC + F::N_PARAM);
   }
   template <class T> MADB_HD T operator()(const T *x, const double *qp) const
   {
      if constexpr (MODE >= 0) { return cons.template al_con<MODE>(x, rhs, qp); }
      else
      {
         T result = f(x, qp);
         if constexpr (MODE == -2) { return result; }
         else { return result + cons.al(x, rhs, lambda, mu, qp); }
      }
   }
};
template <class F, class C, int MODE> using ALFunctionalOf = ALFunctionalN<F, MODE, C>;

// ---------------------------------------------------------------------------
// Dual entropies (src/pg.hpp:253-376): gradients are the latent->primal maps
// ---------------------------------------------------------------------------
// src/pg.hpp:259-278
struct ShannonEntropy
{
   static constexpr int N_INPUT = 1, N_PARAM = 2, N_QPRM = 0;
   double bound, sign;
   MADB_HD void load(const double *p) { bound = p[0]; sign = p[1]; }
   template <class T> MADB_HD T operator()(const T *x, const double *) const
   {
      return sign * (exp(x[0] * sign)) + bound * x[0];
   }
};

// src/pg.hpp:281-322.  Parameters are the constructor arguments (lower, upper);
// the reference binds member upper_bound to evaluator slot 0 (= lower argument)
// and lower_bound to slot 1 (:291-295), so effectively shift = upper argument,
// scale = lower - upper.  Replicated here.
struct FermiDiracEntropy
{
   static constexpr int N_INPUT = 1, N_PARAM = 2, N_QPRM = 0;
   double shift, scale;
   MADB_HD void load(const double *p)
   {
      const double upper_bound = p[0], lower_bound = p[1];
      shift = lower_bound;         // :305
      scale = upper_bound - shift; // :306
   }
   template <class T> MADB_HD T operator()(const T *x, const double *) const
   {
      T z = x[0] * scale;
      if (z > 0) { return z + log(1.0 + exp(-z)) + shift * x[0]; }
      else { return log(1.0 + exp(z)) + shift * x[0]; }
   }
};

// src/pg.hpp:324-342 ; QP: the bound is a spatial coefficient (ex5.cpp:114-117)
template <int DIM, bool QP = false> struct HellingerEntropy
{
   static constexpr int N_INPUT = DIM, N_PARAM = QP ? 0 : 1, N_QPRM = QP ? 1 : 0;
   double scale;
   MADB_HD void load(const double *p) { if constexpr (!QP) { scale = p[0]; } }
   template <class T> MADB_HD T operator()(const T *x, const double *qp) const
   {
      const double s = QP ? qp[0] : scale;
      T xx = x[0] * x[0];
#pragma unroll
      for (int i = 1; i < DIM; i++) { xx += x[i] * x[i]; }
      return sqrt(1 + xx * (s * s));
   }
};

// src/pg.hpp:347-376 (log-sum-exp; max chain with tie averaging)
template <int N> struct SimplexEntropy
{
   static constexpr int N_INPUT = N, N_PARAM = 1, N_QPRM = 0;
   double scale;
   MADB_HD void load(const double *p) { scale = p[0]; }
   template <class T> MADB_HD T operator()(const T *x, const double *) const
   {
      T maxval = x[0];
#pragma unroll
      for (int i = 1; i < N; i++) { maxval = max(maxval, x[i]); }
      T sum_exp = exp(x[0] - maxval);
#pragma unroll
      for (int i = 1; i < N; i++) { sum_exp += exp(x[i] - maxval); }
      return scale * (maxval + log(sum_exp));
   }
};

// ---------------------------------------------------------------------------
// Proximal-Galerkin functionals (src/pg.hpp:58-243)
//   L(u,psi) = f(x) + (sum_j x[primal_idx+j]*(psi_j - psi_k,j) - E*(psi)) / alpha
// x = [f inputs, psi]; psi_k is a per-point parameter (first N_E slots of qp).
// ---------------------------------------------------------------------------
template <class F, class E, int PRIMAL_IDX> struct PGFunctional
{
   static constexpr int NF = F::N_INPUT, NE = E::N_INPUT;
   static constexpr int N_INPUT = NF + NE, N_PARAM = 1 + F::N_PARAM + E::N_PARAM;
   static constexpr int N_QPRM = NE + F::N_QPRM + E::N_QPRM;
   double alpha;
   F f;
   E entropy;
   MADB_HD void load(const double *p)
   {
      alpha = p[0];
      f.load(p + 1);
      entropy.load(p + 1 + F::N_PARAM);
   }
   template <class T> MADB_HD T operator()(const T *x_psi, const double *qp) const
   {
      const T *psi = x_psi + NF;
      const double *psi_k = qp;
      T cross_entropy = x_psi[PRIMAL_IDX] * (psi[0] - psi_k[0]);
#pragma unroll
      for (int j = 1; j < NE; j++) { cross_entropy += x_psi[PRIMAL_IDX + j] * (psi[j] - psi_k[j]); }
      T dual_entropy_sum = entropy(psi, qp + NE + F::N_QPRM);
      return f(x_psi, qp + NE) + (cross_entropy - dual_entropy_sum) / alpha;
   }
};

// ADPGFunctional with TWO entropies (the multi-entropy constructors, src/pg.hpp:105-127, and the loop of :193-213):
//   L = f(x) + sum_i [ sum_j x[idx_i + j] (psi_i,j - psi_k,i,j) - E*_i(psi_i) ] / alpha,   inputs [x, psi_1, psi_2],
// per-point parameters [psi_k,1, psi_k,2, those of f, E_1, E_2].  The latent blocks follow each other after the inputs
// of f (dual_idx[i] = f.n_input + sizes of the earlier entropies).  NOTE: the reference's multi-entropy constructor
// never fills dual_idx (src/pg.hpp:110: value-initialised to 0), so as written every psi_i aliases the first inputs
// of f; no driver uses that constructor.  The device (and the oracle) implement the evident intent.
template <class F, class E1, int IDX1, class E2, int IDX2> struct PGFunctional2
{
   static constexpr int NF = F::N_INPUT, NE1 = E1::N_INPUT, NE2 = E2::N_INPUT;
   static constexpr int N_INPUT = NF + NE1 + NE2, N_PARAM = 1 + F::N_PARAM + E1::N_PARAM + E2::N_PARAM;
   static constexpr int N_QPRM = NE1 + NE2 + F::N_QPRM + E1::N_QPRM + E2::N_QPRM;
   static_assert(NF >= IDX1 + NE1 && NF >= IDX2 + NE2, "ADPGFunctional: f.n_input must be larger than primal_begin[i] + dual_entropy.n_input[i]");
   double alpha;
   F f;
   E1 e1;
   E2 e2;
   MADB_HD void load(const double *p)
   {
      alpha = p[0];
      f.load(p + 1);
      e1.load(p + 1 + F::N_PARAM);
      e2.load(p + 1 + F::N_PARAM + E1::N_PARAM);
   }
   template <class T> MADB_HD T operator()(const T *x_psi, const double *qp) const
   {
      const T *psi1 = x_psi + NF, *psi2 = x_psi + NF + NE1;
      const double *pk1 = qp, *pk2 = qp + NE1;
      T cross_entropy = x_psi[IDX1] * (psi1[0] - pk1[0]);
#pragma unroll
      for (int j = 1; j < NE1; j++) { cross_entropy += x_psi[IDX1 + j] * (psi1[j] - pk1[j]); }
      T dual_entropy_sum = e1(psi1, qp + NE1 + NE2 + F::N_QPRM);
#pragma unroll
      for (int j = 0; j < NE2; j++) { cross_entropy += x_psi[IDX2 + j] * (psi2[j] - pk2[j]); }
      dual_entropy_sum += e2(psi2, qp + NE1 + NE2 + F::N_QPRM + E1::N_QPRM);
      return f(x_psi, qp + NE1 + NE2) + (cross_entropy - dual_entropy_sum) / alpha;
   }
};

// src/pg.hpp:216-243: lambda-form  f(x) + x.lambda - E*(psi_k + alpha*lambda)/alpha
template <class F, class E, int PRIMAL_IDX> struct LambdaPGFunctional
{
   static constexpr int NF = F::N_INPUT, NE = E::N_INPUT;
   static constexpr int N_INPUT = NF + NE, N_PARAM = 1 + F::N_PARAM + E::N_PARAM;
   static constexpr int N_QPRM = NE + F::N_QPRM + E::N_QPRM;
   double alpha;
   F f;
   E entropy;
   MADB_HD void load(const double *p)
   {
      alpha = p[0];
      f.load(p + 1);
      entropy.load(p + 1 + F::N_PARAM);
   }
   template <class T> MADB_HD T operator()(const T *x_lambda, const double *qp) const
   {
      const T *lambda = x_lambda + NF;
      T psi[NE];
      T cross_entropy = x_lambda[PRIMAL_IDX] * lambda[0];
#pragma unroll
      for (int j = 0; j < NE; j++)
      {
         psi[j] = qp[j] + alpha * lambda[j];
         if (j > 0) { cross_entropy += x_lambda[PRIMAL_IDX + j] * lambda[j]; }
      }
      T dual_entropy_sum = entropy(psi, qp + NE + F::N_QPRM);
      return f(x_lambda, qp + NE) + cross_entropy - dual_entropy_sum / alpha;
   }
};

// ---------------------------------------------------------------------------
// mmto pieces (src/mmto.hpp)
// ---------------------------------------------------------------------------
// src/mmto.hpp:9-28  sum_i E_i x_i^p
template <int N> struct SIMPFunction
{
   static constexpr int N_INPUT = N, N_PARAM = N + 1, N_QPRM = 0;
   double E[N];
   double p;
   MADB_HD void load(const double *q)
   {
      for (int i = 0; i < N; i++) { E[i] = q[i]; }
      p = q[N];
   }
   template <class T> MADB_HD T operator()(const T *x, const double *) const
   {
      T result = E[0] * rpow(x[0], p);
#pragma unroll
      for (int i = 1; i < N; i++) { result += E[i] * rpow(x[i], p); }
      return result;
   }
};

// src/mmto.hpp:154-189: elastic energy whose lambda, mu are per-point parameters
// (evaluator slots dim*dim, dim*dim+1 -- :169-170; here the first two qp slots)
template <int DIM> struct ParametrizedCompliance
{
   static constexpr int N_INPUT = DIM * DIM, N_PARAM = 0, N_QPRM = 2;
   MADB_HD void load(const double *) {}
   template <class T> MADB_HD T operator()(const T *gradu, const double *qp) const
   {
      return LinearElasticityEnergy<DIM>::body(gradu, qp[0], qp[1]);
   }
};


// ParametrizedCompliance whose lambda(rho), mu(rho) are evaluated in-kernel from the design
// field at the point (src/mmto.hpp:43-109 "normal mode": only the f_i(design) are evaluated,
// :103-108).  FL, FM: the parameter functions (e.g. SIMPFunction); qp = rho at the point.
template <int DIM, class FL, class FM> struct ParametrizedComplianceOf
{
   static_assert(FL::N_INPUT == FM::N_INPUT, "lambda and mu take the same design vector");
   static constexpr int N_INPUT = DIM * DIM, N_PARAM = FL::N_PARAM + FM::N_PARAM, N_QPRM = FL::N_INPUT;
   FL fl;
   FM fm;
   MADB_HD void load(const double *p) { fl.load(p); fm.load(p + FL::N_PARAM); }
   template <class T> MADB_HD T operator()(const T *gradu, const double *rho) const
   {
      const double lambda = fl(rho, (const double *)nullptr), mu = fm(rho, (const double *)nullptr);
      return LinearElasticityEnergy<DIM>::body(gradu, lambda, mu);
   }
};

// The same energy as a function of the DESIGN with the state gradient as per-point
// parameter: its gradient w.r.t. rho is the design sensitivity dF/drho_j; the reference's
// ParamGradient (src/mmto.cpp:25-37) returns dF/drho_j + (m-1) F with m = 2 parameter
// functions -- both are formed from value and gradient of this functional.
template <int DIM, class FL, class FM> struct DesignComplianceOf
{
   static constexpr int N_INPUT = FL::N_INPUT, N_PARAM = FL::N_PARAM + FM::N_PARAM, N_QPRM = DIM * DIM;
   FL fl;
   FM fm;
   MADB_HD void load(const double *p) { fl.load(p); fm.load(p + FL::N_PARAM); }
   static constexpr bool HAS_PARAM_GRADIENT = true;
   /// ParametrizedFunctional::ParamGradient::Eval exactly as written (src/mmto.cpp:25-37): for every parameter
   /// function f_i and design component j, slot i of the evaluator is overwritten with df_i/drho_j while the other
   /// slots keep their VALUES, and the parent is re-evaluated: J[j] = sum_i parent(state; ..., df_i/drho_j, ...).
   /// For this parent (linear in lambda, mu) that is dF/drho_j + (m-1) F, m = 2.
   MADB_HD void param_gradient_as_written(const double *rho, const double *gradu, double *J) const
   {
      using T1 = AD<N_INPUT, 1>;
      T1 r[N_INPUT];
#pragma unroll
      for (int m = 0; m < N_INPUT; m++) { r[m] = ad_seed<N_INPUT, 1>(rho[m], m); }
      const T1 lam = fl(r, (const double *)nullptr), mu = fm(r, (const double *)nullptr);
#pragma unroll
      for (int j = 0; j < N_INPUT; j++)
      {
         double acc = 0.0;                                                       // J = 0.0 (:9)
         acc += LinearElasticityEnergy<DIM>::body(gradu, lam.g[j], mu.v);       // slot 0 <- dlambda/drho_j (:32-33)
         acc += LinearElasticityEnergy<DIM>::body(gradu, lam.v, mu.g[j]);       // slot 1 <- dmu/drho_j
         J[j] = acc;
      }
   }
   template <class T> MADB_HD T operator()(const T *rho, const double *gradu) const
   {
      const T lambda = fl(rho, (const double *)nullptr), mu = fm(rho, (const double *)nullptr);
      double div = 0.0;
#pragma unroll
      for (int i = 0; i < DIM; i++) { div += gradu[i * DIM + i]; }
      double h1 = 0.0;
#pragma unroll
      for (int i = 0; i < DIM; i++)
      {
#pragma unroll
         for (int j = 0; j < DIM; j++)
         {
            const double s = 0.5 * (gradu[i * DIM + j] + gradu[j * DIM + i]);
            h1 += s * s;
         }
      }
      return (0.5 * div * div) * lambda + h1 * mu;
   }
};

// ---------------------------------------------------------------------------
// Single-space VECTOR integrator, reference arithmetic.
// ADNonlinearFormIntegrator<...|VECTOR>::AssembleElementGrad (src/ad_intg.hpp:283-326) views the n x n Hessian
// (n = SD*VD, x index s + SD*c) as Hs[SD x VD*SD*VD] (:292), forms Hx = allshapes*Hs (:312) and takes CONTIGUOUS
// windows Hx + (c*VD + r)*SD*dof as if they held the (c, r) block (:318); the (c, r) part is added untransposed at
// (r, c) as well (:320-324).  In index terms: block (c, r), r <= c, of the element matrix is B G B^T with
//     G(t, s1) = H[s1 + SD*c1][s2 + SD*c2],   k = (c*VD + r)*SD + t,  c1 = k % VD,  s2 = (k / VD) % SD,  c2 = k / (VD*SD),
// which is the intended H[(t,c)][(s1,r)] only for special H (LinearElasticityEnergy with lambda == mu).  The wrapper
// applies exactly this permutation to the second derivatives of F's result, so that the index-consistent contraction
// of the kernels reproduces the reference's element matrix.  The packed symmetric storage keeps the (a <= b) half:
// for the energies the reference uses with this integrator (LinearElasticityEnergy, ParametrizedCompliance) the
// permuted matrix is symmetric for any lambda, mu (checked against the oracle's literal emulation in the tests).
// Value and gradient (energy, residual) are untouched.
// ---------------------------------------------------------------------------
template <class F, int SD, int VD> struct RefVectorOf
{
   static_assert(F::N_INPUT == SD * VD, "single vector space: n_input = shapedim * vdim");
   static constexpr int N_INPUT = F::N_INPUT, N_PARAM = F::N_PARAM, N_QPRM = F::N_QPRM;
   F f;
   MADB_HD void load(const double *p) { f.load(p); }
   template <class T> MADB_HD T operator()(const T *x, const double *qp) const
   {
      const T r = f(x, qp);
      if constexpr (ad_traits<T>::order >= 2)
      {
         constexpr int N = N_INPUT;
         T out = r;
#pragma unroll
         for (int a = 0; a < N; a++)
         {
#pragma unroll
            for (int b = a; b < N; b++)
            {
               const int ca = a / SD, t = a % SD, cb = b / SD, s1 = b % SD; // ca <= cb
               // block (ca, cb) is the untransposed copy of block (c, r) = (cb, ca) (:320-324); ca == cb: the block itself
               const int k = (cb * VD + ca) * SD + t;
               const int c1 = k % VD, s2 = (k / VD) % SD, c2 = k / (VD * SD);
               out.setH(hidx<N>(a, b), r.H(symidx_h<N>(s1 + SD * c1, s2 + SD * c2)));
            }
         }
         return out;
      }
      else { return r; }
   }
};

} // namespace madb

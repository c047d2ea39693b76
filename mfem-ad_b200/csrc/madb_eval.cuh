// madb_eval.cuh -- pointwise AD evaluation on the device (value, gradient,
// Hessian of a functional at many points): the device counterpart of
// ADFunction::operator() / Gradient / Hessian (src/ad_native.cpp:181-230),
// used for the ex0 known answers (config 1), for DifferentiableCoefficient-style
// latent->primal maps (src/ad_native.hpp:267-323) and by the nodal LVPP update.
#pragma once
#include "madb_ad.cuh"
#include <cuda_runtime.h>

#include <map>
#include <string>

namespace madb
{

struct EvalOps
{
   int (*launch)(cudaStream_t, int npts, const double *fparams_host, const double *x, const double *qprm,
                 double *value, double *grad, double *hess);
   int n_input, n_fparam, n_qprm;
   // nodal proximal-Galerkin terms (scalar entropies only, else null)
   int (*dofpg)(cudaStream_t, int n, const double *fparams_host, double alpha, const double *u, const double *psi,
                const double *psik, const double *w, double *r_u, double *r_psi, double *d_pp, double *d_up);
};
std::map<std::string, EvalOps> &eval_registry();
struct EvalRegistrar
{
   EvalRegistrar(const std::string &key, const EvalOps &ops);
};

template <class Func> struct EvalArgs
{
   int npts;
   const double *x, *qprm;
   double *value, *grad, *hess;
   double fparams[Func::N_PARAM > 0 ? Func::N_PARAM : 1];
};

template <class Func> __global__ void __launch_bounds__(128) k_eval(const EvalArgs<Func> a)
{
   constexpr int N = Func::N_INPUT;
   const int p = blockIdx.x * blockDim.x + threadIdx.x;
   if (p >= a.npts) { return; }
   Func f;
   f.load(a.fparams);
   double qp[Func::N_QPRM > 0 ? Func::N_QPRM : 1];
#pragma unroll
   for (int k = 0; k < Func::N_QPRM; k++) { qp[k] = a.qprm[(size_t)p * Func::N_QPRM + k]; }
   using T = AD<N, 2>;
   T xs[N];
#pragma unroll
   for (int m = 0; m < N; m++) { xs[m] = ad_seed<N, 2>(a.x[(size_t)p * N + m], m); }
   const T r = f(xs, qp);
   if (a.value) { a.value[p] = r.v; }
   if (a.grad)
   {
#pragma unroll
      for (int i = 0; i < N; i++) { a.grad[(size_t)p * N + i] = r.g[i]; }
   }
   if (a.hess)
   {
#pragma unroll
      for (int i = 0; i < N; i++)
      {
#pragma unroll
         for (int j = 0; j < N; j++) { a.hess[((size_t)p * N + i) * N + j] = r.hess(i, j); }
      }
   }
}

// DOF-collocated proximal-Galerkin terms (ADDofPGNonlinearFormIntegrator, src/dof_pg.hpp:113-125,
// :210-228): one thread per dof j with nodal weight w_j (an explicit input):
//   r_u[j]  += (psi_j - psi_k,j) w_j/alpha        r_psi[j]  = (u_j - E*'(psi_j)) w_j/alpha
//   d_pp[j]  = -E*''(psi_j) w_j/alpha             d_up[j]   = w_j/alpha
template <class Func> struct DofPGArgs
{
   int n;
   double alpha;
   const double *u, *psi, *psik, *w;
   double *r_u, *r_psi, *d_pp, *d_up;
   double fparams[Func::N_PARAM > 0 ? Func::N_PARAM : 1];
};
template <class Func> __global__ void __launch_bounds__(256) k_dofpg(const DofPGArgs<Func> a)
{
   const int j = blockIdx.x * blockDim.x + threadIdx.x;
   if (j >= a.n) { return; }
   Func f;
   f.load(a.fparams);
   const double ps = a.psi[j];
   AD<1, 2> xs[1] = {ad_seed<1, 2>(ps, 0)};
   const AD<1, 2> r = f(xs, (const double *)nullptr);
   const double ww = a.w[j] / a.alpha;
   if (a.r_u) { a.r_u[j] += (ps - a.psik[j]) * ww; }
   if (a.r_psi) { a.r_psi[j] = (a.u[j] - r.g[0]) * ww; }
   if (a.d_pp) { a.d_pp[j] = -r.h[0] * ww; }
   if (a.d_up) { a.d_up[j] = ww; }
}
template <class Func> int dofpg_launch(cudaStream_t s, int n, const double *fp, double alpha, const double *u,
                                       const double *psi, const double *psik, const double *w, double *r_u,
                                       double *r_psi, double *d_pp, double *d_up)
{
   DofPGArgs<Func> a;
   a.n = n; a.alpha = alpha; a.u = u; a.psi = psi; a.psik = psik; a.w = w;
   a.r_u = r_u; a.r_psi = r_psi; a.d_pp = d_pp; a.d_up = d_up;
   for (int i = 0; i < Func::N_PARAM; i++) { a.fparams[i] = fp[i]; }
   k_dofpg<Func><<<(n + 255) / 256, 256, 0, s>>>(a);
   return (int)cudaGetLastError();
}
template <class Func> constexpr auto dofpg_ptr()
{
   if constexpr (Func::N_INPUT == 1 && Func::N_QPRM == 0) { return &dofpg_launch<Func>; }
   else { return (decltype(&dofpg_launch<Func>))nullptr; }
}

template <class Func> int eval_launch(cudaStream_t s, int npts, const double *fp, const double *x, const double *qprm,
                                      double *value, double *grad, double *hess)
{
   EvalArgs<Func> a;
   a.npts = npts; a.x = x; a.qprm = qprm; a.value = value; a.grad = grad; a.hess = hess;
   for (int i = 0; i < Func::N_PARAM; i++) { a.fparams[i] = fp[i]; }
   k_eval<Func><<<(npts + 127) / 128, 128, 0, s>>>(a);
   return (int)cudaGetLastError();
}

// ---- ADVectorFunction (src/ad_native.hpp:198-265, src/ad_native.cpp:232-276): F: R^n -> R^m -------------------
// value F[m], Jacobian J[m][n] (row = output) and Hessians H[m][n][n], one hyper-dual pass for all outputs.
// A vector functional provides N_INPUT, N_OUTPUT, N_PARAM, load() and
//    template <class T> void operator()(const T *x, T *result) const     (the AD_VEC_IMPL body)
struct VecEvalOps
{
   int (*launch)(cudaStream_t, int npts, const double *fparams_host, const double *x, double *value, double *jac, double *hess);
   int n_input, n_output, n_fparam;
};
std::map<std::string, VecEvalOps> &vec_eval_registry();
struct VecEvalRegistrar
{
   VecEvalRegistrar(const std::string &key, const VecEvalOps &ops);
};
template <class Func> struct VecEvalArgs
{
   int npts;
   const double *x;
   double *value, *jac, *hess;
   double fparams[Func::N_PARAM > 0 ? Func::N_PARAM : 1];
};
template <class Func> __global__ void __launch_bounds__(128) k_eval_vec(const VecEvalArgs<Func> a)
{
   constexpr int N = Func::N_INPUT, M = Func::N_OUTPUT;
   const int p = blockIdx.x * blockDim.x + threadIdx.x;
   if (p >= a.npts) { return; }
   Func f;
   f.load(a.fparams);
   using T = AD<N, 2>;
   T xs[N], res[M];
#pragma unroll
   for (int m = 0; m < N; m++) { xs[m] = ad_seed<N, 2>(a.x[(size_t)p * N + m], m); }
   f(xs, res);
#pragma unroll
   for (int o = 0; o < M; o++)
   {
      if (a.value) { a.value[(size_t)p * M + o] = res[o].v; }
      if (a.jac)
      {
#pragma unroll
         for (int i = 0; i < N; i++) { a.jac[((size_t)p * M + o) * N + i] = res[o].g[i]; }
      }
      if (a.hess)
      {
#pragma unroll
         for (int i = 0; i < N; i++)
         {
#pragma unroll
            for (int j = 0; j < N; j++) { a.hess[(((size_t)p * M + o) * N + i) * N + j] = res[o].hess(i, j); }
         }
      }
   }
}
template <class Func> int vec_eval_launch(cudaStream_t s, int npts, const double *fp, const double *x, double *value, double *jac,
                                          double *hess)
{
   VecEvalArgs<Func> a;
   a.npts = npts; a.x = x; a.value = value; a.jac = jac; a.hess = hess;
   for (int i = 0; i < Func::N_PARAM; i++) { a.fparams[i] = fp[i]; }
   k_eval_vec<Func><<<(npts + 127) / 128, 128, 0, s>>>(a);
   return (int)cudaGetLastError();
}

#define MADB_EVAL_CAT2(a, b) a##b
#define MADB_EVAL_CAT(a, b) MADB_EVAL_CAT2(a, b)
#define MADB_EVAL_INSTANCE(KIND, FUNC)                                                                              \
   static ::madb::EvalRegistrar MADB_EVAL_CAT(madb_evreg_, __COUNTER__)(                                            \
      std::string(KIND) + "|n" + std::to_string(FUNC::N_INPUT),                                                     \
      ::madb::EvalOps {&::madb::eval_launch<FUNC>, FUNC::N_INPUT, FUNC::N_PARAM, FUNC::N_QPRM, ::madb::dofpg_ptr<FUNC>()});

#define MADB_VEC_EVAL_INSTANCE(KIND, FUNC)                                                                          \
   static ::madb::VecEvalRegistrar MADB_EVAL_CAT(madb_vevreg_, __COUNTER__)(                                        \
      std::string(KIND) + "|n" + std::to_string(FUNC::N_INPUT) + "m" + std::to_string(FUNC::N_OUTPUT),              \
      ::madb::VecEvalOps {&::madb::vec_eval_launch<FUNC>, FUNC::N_INPUT, FUNC::N_OUTPUT, FUNC::N_PARAM});

} // namespace madb

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
   a.value[p] = r.v;
#pragma unroll
   for (int i = 0; i < N; i++) { a.grad[(size_t)p * N + i] = r.g[i]; }
#pragma unroll
   for (int i = 0; i < N; i++)
   {
#pragma unroll
      for (int j = 0; j < N; j++) { a.hess[((size_t)p * N + i) * N + j] = r.hess(i, j); }
   }
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

#define MADB_EVAL_CAT2(a, b) a##b
#define MADB_EVAL_CAT(a, b) MADB_EVAL_CAT2(a, b)
#define MADB_EVAL_INSTANCE(KIND, FUNC)                                                                              \
   static ::madb::EvalRegistrar MADB_EVAL_CAT(madb_evreg_, __COUNTER__)(                                            \
      std::string(KIND) + "|n" + std::to_string(FUNC::N_INPUT),                                                     \
      ::madb::EvalOps {&::madb::eval_launch<FUNC>, FUNC::N_INPUT, FUNC::N_PARAM, FUNC::N_QPRM});

} // namespace madb

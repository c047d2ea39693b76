// Out-of-tree user functional, written exactly as INTEGRATION.md section 4 describes: the body a user of the reference
// would put into AD_IMPL(T, V, M, x, ...) (src/ad_native.hpp:332-365), as a template over the scalar type, plus one
// MADB_INSTANCE line per <functional, element configuration>.  tests/test_plugin.py compiles this file with nvcc into
// tests/plugin/libmyenergy.so against the installed headers and libmadb.so, loads it next to the library and assembles
// with it.
#include "madb_functionals.cuh"
#include "madb_registry.cuh"

template <int DIM> struct MyEnergy
{
   static constexpr int N_INPUT = DIM, N_PARAM = 1, N_QPRM = 0;
   double kappa;
   MADB_HD void load(const double *p) { kappa = p[0]; }
   template <class T> MADB_HD T operator()(const T *g, const double *) const
   {
      T s = g[0] * g[0];
      for (int i = 1; i < DIM; i++) { s += g[i] * g[i]; }
      return exp(kappa * s) + 0.5 * s;
   }
};
using E2 = MyEnergy<2>;
using Q2 = madb::Config<2, 4, madb::Field<3, 1, madb::EV_GRAD>>; // H1 order 2, 4x4 points
using Q1 = madb::Config<2, 3, madb::Field<2, 1, madb::EV_GRAD>>; // H1 order 1, 3x3 points
MADB_INSTANCE("myenergy", E2, Q2, true)
MADB_INSTANCE("myenergy", E2, Q1, true)

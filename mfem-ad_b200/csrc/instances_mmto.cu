// Fused-kernel instances for the multi-material topology-optimisation pieces (config 4):
//   state block:      displacement Q1 (vdim 2, GRAD|VECTOR) with lambda(rho), mu(rho) from the
//                     5-material design field (SIMP) -- ParametrizedCompliance, src/mmto.hpp:154-189
//   design gradient:  ParamGradient at the quadrature points, src/mmto.cpp:4-38
//   latent map:       rho = grad E*(psi) = softmax(psi) at the quadrature points (SimplexEntropy)
#include "madb_functionals.cuh"
#include "madb_registry.cuh"
using namespace madb;

using SIMP5 = SIMPFunction<5>;
using State = ParametrizedComplianceOf<2, SIMP5, SIMP5>;
using Design = DesignComplianceOf<2, SIMP5, SIMP5>;
using Simplex5 = SimplexEntropy<5>;

using CfgState = Config<2, 3, Field<2, 2, EV_GRAD>, Field<2, 5, EV_VALUE, ROLE_PARAM>>;
using CfgDesign = Config<2, 3, Field<2, 5, EV_VALUE>, Field<2, 2, EV_GRAD, ROLE_PARAM>>;
using CfgLatent = Config<2, 3, Field<2, 5, EV_VALUE>>;
// quadrature loop NOT unrolled: unrolled, the kernel is 13.8 k instructions (220 KB) and instruction-cache bound (2x slower)
MADB_INSTANCE("paramcompliance[simp,simp]", State, CfgState, false)
MADB_INSTANCE_REFVEC("paramcompliance[simp,simp]", State, CfgState, false) // single-space arithmetic as written
MADB_INSTANCE("designcompliance[simp,simp]", Design, CfgDesign, false)
MADB_INSTANCE("simplex", Simplex5, CfgLatent, false)

// ex4.cpp:124-128,200: latent->primal map U(psi) of the obstacle problem on the L2 latent space
using CfgMapO1 = Config<2, 4, Field<1, 1, EV_VALUE>>;
using CfgMapO2 = Config<2, 5, Field<2, 1, EV_VALUE>>;
MADB_INSTANCE("fermidirac", FermiDiracEntropy, CfgMapO1, false)
MADB_INSTANCE("fermidirac", FermiDiracEntropy, CfgMapO2, false)

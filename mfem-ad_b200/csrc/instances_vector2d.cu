// Fused-kernel instances: vector H1 spaces, ADEval::GRAD | ADEval::VECTOR (ex3, config-4 state block).
#include "madb_functionals.cuh"
#include "madb_registry.cuh"
using namespace madb;

using Elast2 = LinearElasticityEnergy<2>;
using V1 = Config<2, 3, Field<2, 2, EV_GRAD>>; // order 1
using V2 = Config<2, 4, Field<3, 2, EV_GRAD>>; // order 2
MADB_INSTANCE("elasticity", Elast2, V1, true)
MADB_INSTANCE("elasticity", Elast2, V2, false)
// reference arithmetic of the single-space VECTOR integrator: the default for one vector space
MADB_INSTANCE_REFVEC("elasticity", Elast2, V1, true)
MADB_INSTANCE_REFVEC("elasticity", Elast2, V2, false)

// vector load vectors (VectorDomainLFIntegrator, ex3.cpp:64-67) on (H1)^2 of order 1 and 2, MFEM's rule of order 2p
using VL1 = Config<2, 2, Field<2, 2, EV_VALUE>>;
using VL2 = Config<2, 3, Field<3, 2, EV_VALUE>>;
MADB_INSTANCE("vload:2", VectorLoadFunctional<2>, VL1, true)
MADB_INSTANCE("vload:2", VectorLoadFunctional<2>, VL2, true)

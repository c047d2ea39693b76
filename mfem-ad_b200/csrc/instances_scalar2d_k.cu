// Fused-kernel instances: 2-D scalar H1 space, DiffusionEnergy with K (split from instances_scalar2d.cu for build time).
#include "madb_functionals.cuh"
#include "madb_registry.cuh"
using namespace madb;

using Q1 = Config<2, 3, Field<2, 1, EV_GRAD>>;
using Q2 = Config<2, 4, Field<3, 1, EV_GRAD>>;

// DiffusionEnergy with a constant K: scalar / diagonal / full (src/ad_native.hpp:421-481; the K kind is a structural
// integer of the functional: "diffusion:1|2|4") and with a spatially varying scalar K given as a quadrature function
// ("diffusionq:1": Coefficient-type evaluator source sampled at the points, madb_integrator_qpoint_coords)
using DiffK1 = DiffusionEnergy<2, 1>;
using DiffK2 = DiffusionEnergy<2, 2>;
using DiffK4 = DiffusionEnergy<2, 4>;
MADB_INSTANCE("diffusion:1", DiffK1, Q1, true)
MADB_INSTANCE("diffusion:2", DiffK2, Q1, true)
MADB_INSTANCE("diffusion:4", DiffK4, Q1, true)
MADB_INSTANCE("diffusion:1", DiffK1, Q2, true)
MADB_INSTANCE("diffusion:2", DiffK2, Q2, true)
MADB_INSTANCE("diffusion:4", DiffK4, Q2, true)
using DiffQ1 = DiffusionEnergy<2, 1, true>;
MADB_INSTANCE("diffusionq:1", DiffQ1, Q2, true)

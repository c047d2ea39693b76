// Fused-kernel instances: 2-D scalar H1 order 3 (split from instances_scalar2d.cu for build time).
#include "madb_functionals.cuh"
#include "madb_registry.cuh"
using namespace madb;

using Diff2 = DiffusionEnergy<2, 0>;
using MinS2 = MinimalSurfaceEnergy<2>;

// scalar H1 order 3 (4x4 dofs, 5x5 points): 64-element patches, 4 threads per element
using Q3 = Config<2, 5, Field<4, 1, EV_GRAD>>;
MADB_INSTANCE("diffusion:0", Diff2, Q3, false)
MADB_INSTANCE("minsurf", MinS2, Q3, false)

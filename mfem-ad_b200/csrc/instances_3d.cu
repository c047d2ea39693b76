// Fused-kernel instances in 3-D: direct (assembled) kernels for order 1 and the
// sum-factorised matrix-free kernels for orders 2 and 3 (config 3).
#include "madb_functionals.cuh"
#include "madb_kernels3d.cuh"
#include "madb_registry.cuh"
using namespace madb;

using Diff3 = DiffusionEnergy<3, 0>;
using MinS3 = MinimalSurfaceEnergy<3>;
using H1 = Config<3, 3, Field<2, 1, EV_GRAD>>; // trilinear hexes, 3^3 points, assembled Jacobian
MADB_INSTANCE("diffusion:0", Diff3, H1, false)
MADB_INSTANCE("minsurf", MinS3, H1, false)

MADB_INSTANCE_SUMFAC3D("diffusion:0", Diff3, 3, 4)
MADB_INSTANCE_SUMFAC3D("minsurf", MinS3, 3, 4)
MADB_INSTANCE_SUMFAC3D("diffusion:0", Diff3, 4, 5)
MADB_INSTANCE_SUMFAC3D("minsurf", MinS3, 4, 5)

// Fused-kernel instances: 2-D scalar H1 space, ADEval::GRAD (ex1, ex2, config 2).
#include "madb_functionals.cuh"
#include "madb_registry.cuh"
using namespace madb;

using Q1 = Config<2, 3, Field<2, 1, EV_GRAD>>; // order 1, default rule 2p+2 -> 3x3 points
using Q2 = Config<2, 4, Field<3, 1, EV_GRAD>>; // order 2 -> 4x4 points (config 2)
using Diff2 = DiffusionEnergy<2, 0>;
using MinS2 = MinimalSurfaceEnergy<2>;

MADB_INSTANCE("diffusion:0", Diff2, Q1, true)
MADB_INSTANCE("diffusion:0", Diff2, Q2, true)
MADB_INSTANCE("minsurf", MinS2, Q1, true)
MADB_INSTANCE("minsurf", MinS2, Q2, true)

// Lagrangian f(grad u) + lambda c(grad u): H1 order 1 x L2 order 0 (block integrator, GRAD / VALUE);
// augmented Lagrangian on the order-2 space (sum-factorised path)
using LagDM = LagrangianOf<Diff2, MinS2, -1>;
using ALDM = ALFunctionalOf<Diff2, MinS2, -1>;
using Q1L0 = Config<2, 3, Field<2, 1, EV_GRAD>, Field<1, 1, EV_VALUE>>;
MADB_INSTANCE("lagrangian:-1[diffusion:0,minsurf]", LagDM, Q1L0, true)
MADB_INSTANCE("al:-1[diffusion:0,minsurf]", ALDM, Q2, true)

// two equality constraints: Lagrangian on H1 order 1 x (L2 order 0)^2 (multipliers as a vector space), augmented
// Lagrangian on the order-2 space
using Lag2 = LagrangianN<Diff2, -1, MinS2, Diff2>;
using AL2 = ALFunctionalN<Diff2, -1, MinS2, Diff2>;
using Q1L0x2 = Config<2, 3, Field<2, 1, EV_GRAD>, Field<1, 2, EV_VALUE>>;
MADB_INSTANCE("lagrangian:-1[diffusion:0,minsurf,diffusion:0]", Lag2, Q1L0x2, true)
MADB_INSTANCE("al:-1[diffusion:0,minsurf,diffusion:0]", AL2, Q2, true)

// Fused-kernel instances on triangles (ex5.cpp:72-73: Mesh::MakeCartesian2D(10, 10, Element::TRIANGLE)).
// Table-driven generic element computation (no sum factorisation on simplices), patch assembly as everywhere else.
//   scalar H1 P1 / P2 with ADEval::GRAD, default rule 2p+2: order 4 -> 6 points, order 6 -> 12 points
//   ex5: H1 P2 (GRAD) x (H1 P1)^2 (VALUE | VECTOR), psi_k parameter on the latent space, 12 points
//   load vectors: order 2p -> 3 / 6 points
#include "madb_functionals.cuh"
#include "madb_registry.cuh"
using namespace madb;

using T1 = SConfig<6, SField<3, 1, EV_GRAD>>;
using T2 = SConfig<12, SField<6, 1, EV_GRAD>>;
using Diff2 = DiffusionEnergy<2, 0>;
using MinS2 = MinimalSurfaceEnergy<2>;
MADB_INSTANCE("diffusion:0", Diff2, T1, true)
MADB_INSTANCE("diffusion:0", Diff2, T2, true)
MADB_INSTANCE("minsurf", MinS2, T1, true)
MADB_INSTANCE("minsurf", MinS2, T2, true)

using PGGrad = PGFunctional<GradientObstacleEnergy<2>, HellingerEntropy<2>, 0>;
using PGGradQ = PGFunctional<GradientObstacleEnergy<2>, HellingerEntropy<2, true>, 0>;
using Ex5T = SConfig<12, SField<6, 1, EV_GRAD>, SField<3, 2, EV_VALUE>, SField<3, 2, EV_VALUE, ROLE_PARAM>>;
MADB_INSTANCE("pg:0[gradobstacle,hellinger]", PGGrad, Ex5T, false)
MADB_INSTANCE("pg:0[gradobstacle,hellingerq]", PGGradQ, Ex5T, false)

using L1 = SConfig<3, SField<3, 1, EV_VALUE>>;
using L2 = SConfig<6, SField<6, 1, EV_VALUE>>;
MADB_INSTANCE("load", LoadFunctional, L1, true)
MADB_INSTANCE("load", LoadFunctional, L2, true)

// Fused-kernel instances: proximal-Galerkin (LVPP) block systems in 2-D.
//   ex4 (obstacle):  H1(p+1) x L2(p-1), modes VALUE|GRAD / VALUE, psi_k parameter on the L2 space,
//                    rule order 3p+3 (ex4.cpp:99-104,137-142)
//   ex5 (gradient constraint, here on quads): H1(p) scalar x H1(p-1) vector latent,
//                    modes GRAD / VALUE|VECTOR, default rule (ex5.cpp:88-91,135-140)
#include "madb_functionals.cuh"
#include "madb_registry.cuh"
using namespace madb;

using PGObs = PGFunctional<ObstacleEnergy<2>, FermiDiracEntropy, 0>;
using PGGrad = PGFunctional<GradientObstacleEnergy<2>, HellingerEntropy<2>, 0>;

// ex4 -o 1: H1 p2 x L2 p0, rule order 6 -> 4x4 ; ex4 -o 2 (default, config 5): H1 p3 x L2 p1, order 9 -> 5x5
using Ex4o1 = Config<2, 4, Field<3, 1, EV_VALUE | EV_GRAD>, Field<1, 1, EV_VALUE>, Field<1, 1, EV_VALUE, ROLE_PARAM>>;
using Ex4o2 = Config<2, 5, Field<4, 1, EV_VALUE | EV_GRAD>, Field<2, 1, EV_VALUE>, Field<2, 1, EV_VALUE, ROLE_PARAM>>;
// ADLambdaPGFunctional (src/pg.hpp:216-243; unused by the reference's drivers): the lambda form on the ex4 -o 1 spaces
using LamPGObs = LambdaPGFunctional<ObstacleEnergy<2>, FermiDiracEntropy, 0>;
MADB_INSTANCE("lambdapg:0[obstacle,fermidirac]", LamPGObs, Ex4o1, false)
MADB_INSTANCE("pg:0[obstacle,fermidirac]", PGObs, Ex4o1, false)
MADB_INSTANCE("pg:0[obstacle,fermidirac]", PGObs, Ex4o2, false)

// ex5 -o 2: H1 p2 x (H1 p1)^2, rule order 6 -> 4x4
using Ex5o2 = Config<2, 4, Field<3, 1, EV_GRAD>, Field<2, 2, EV_VALUE>, Field<2, 2, EV_VALUE, ROLE_PARAM>>;
MADB_INSTANCE("pg:0[gradobstacle,hellinger]", PGGrad, Ex5o2, false)
// ex5.cpp:114-117: the Hellinger bound as a spatial coefficient -> a quadrature-function parameter of the entropy
using PGGradQ = PGFunctional<GradientObstacleEnergy<2>, HellingerEntropy<2, true>, 0>;
MADB_INSTANCE("pg:0[gradobstacle,hellingerq]", PGGradQ, Ex5o2, false)

// ADEval::QVALUE latent variable (src/ad_intg.hpp:127, quadrature-space unknowns of src/tools.hpp:156-177): H1 p1 primal,
// one latent value per point of the default 3x3 rule = VALUE on the nodal L2 space of order 2 on the rule's points
using Ex4q = Config<2, 3, Field<2, 1, EV_VALUE | EV_GRAD>, Field<3, 1, EV_VALUE>, Field<3, 1, EV_VALUE, ROLE_PARAM>>;
MADB_INSTANCE("pg:0[obstacle,fermidirac]", PGObs, Ex4q, false)

// ADPGFunctional with two entropies (src/pg.hpp:105-127): bound constraint on u (FermiDirac, primal index 0) and gradient
// bound (Hellinger, primal index 1) on the obstacle energy; H1 p2 x L2 p0 x (L2 p0)^2, previous latents as parameters
using PG2 = PGFunctional2<ObstacleEnergy<2>, FermiDiracEntropy, 0, HellingerEntropy<2>, 1>;
using PG2cfg = Config<2, 4, Field<3, 1, EV_VALUE | EV_GRAD>, Field<1, 1, EV_VALUE>, Field<1, 2, EV_VALUE>,
                      Field<1, 1, EV_VALUE, ROLE_PARAM>, Field<1, 2, EV_VALUE, ROLE_PARAM>>;
MADB_INSTANCE("pg:0,1[obstacle,fermidirac,hellinger]", PG2, PG2cfg, false)

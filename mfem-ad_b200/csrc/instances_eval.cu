// Pointwise AD instances (ex0 known answers, entropy maps).
#include "madb_eval.cuh"
#include "madb_functionals.cuh"
using namespace madb;

using MinS2 = MinimalSurfaceEnergy<2>;
using Simplex5 = SimplexEntropy<5>;
using Simplex3 = SimplexEntropy<3>;
using Hell2 = HellingerEntropy<2>;
using SIMP5 = SIMPFunction<5>;
using Elast2 = LinearElasticityEnergy<2>;
using Elast3 = LinearElasticityEnergy<3>;
using PGObsFD = PGFunctional<ObstacleEnergy<2>, FermiDiracEntropy, 0>;
using LamPGObsFD = LambdaPGFunctional<ObstacleEnergy<2>, FermiDiracEntropy, 0>;

MADB_EVAL_INSTANCE("ex0", Ex0Function)
MADB_EVAL_INSTANCE("sqrtprobe", SqrtProbe)
MADB_EVAL_INSTANCE("minsurf", MinS2)
MADB_EVAL_INSTANCE("shannon", ShannonEntropy)
MADB_EVAL_INSTANCE("fermidirac", FermiDiracEntropy)
MADB_EVAL_INSTANCE("hellinger", Hell2)
MADB_EVAL_INSTANCE("simplex", Simplex5)
MADB_EVAL_INSTANCE("simplex", Simplex3)
MADB_EVAL_INSTANCE("simp", SIMP5)
MADB_EVAL_INSTANCE("elasticity", Elast2)
MADB_EVAL_INSTANCE("elasticity", Elast3)
MADB_EVAL_INSTANCE("pg:0[obstacle,fermidirac]", PGObsFD)
MADB_EVAL_INSTANCE("lambdapg:0[obstacle,fermidirac]", LamPGObsFD)

// ADVectorFunction of ex0 (ex0.cpp:23-35): value, Jacobian and Hessians on the device
MADB_VEC_EVAL_INSTANCE("ex0vec", Ex0VectorFunction)

// Lagrangian / ALFunctional (src/ad_native.hpp:570-691; unused by the reference's drivers)
using LagDM = LagrangianOf<DiffusionEnergy<2, 0>, MinS2, -1>;
using ALDM = ALFunctionalOf<DiffusionEnergy<2, 0>, MinS2, -1>;
MADB_EVAL_INSTANCE("lagrangian:-1[diffusion:0,minsurf]", LagDM)
MADB_EVAL_INSTANCE("al:-1[diffusion:0,minsurf]", ALDM)

using HellQ2 = HellingerEntropy<2, true>;
MADB_EVAL_INSTANCE("hellingerq", HellQ2)
MADB_EVAL_INSTANCE("mass", MassEnergy<1>)
using DiffK4e = DiffusionEnergy<2, 4>;
MADB_EVAL_INSTANCE("diffusion:4", DiffK4e)
using DiffMass1e = DiffEnergy<MassEnergy<1>>;
MADB_EVAL_INSTANCE("diff[mass]", DiffMass1e)

// several equality constraints (std::vector<ADFunction*> eq_con, src/ad_native.hpp:583,648) and two entropies
// (src/pg.hpp:105-127): statically composed lists
using Lag2 = LagrangianN<DiffusionEnergy<2, 0>, -1, MinS2, DiffusionEnergy<2, 0>>;
using AL2 = ALFunctionalN<DiffusionEnergy<2, 0>, -1, MinS2, DiffusionEnergy<2, 0>>;
using Lag2c1 = LagrangianN<DiffusionEnergy<2, 0>, 1, MinS2, DiffusionEnergy<2, 0>>;
MADB_EVAL_INSTANCE("lagrangian:-1[diffusion:0,minsurf,diffusion:0]", Lag2)
MADB_EVAL_INSTANCE("lagrangian:1[diffusion:0,minsurf,diffusion:0]", Lag2c1)
MADB_EVAL_INSTANCE("al:-1[diffusion:0,minsurf,diffusion:0]", AL2)
using PG2ObsFDHell = PGFunctional2<ObstacleEnergy<2>, FermiDiracEntropy, 0, Hell2, 1>;
MADB_EVAL_INSTANCE("pg:0,1[obstacle,fermidirac,hellinger]", PG2ObsFDHell)

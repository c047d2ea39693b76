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

// MassEnergy (src/ad_native.hpp:413-420) and DiffEnergy (:483-525: energy(x - target), the target a per-point
// parameter) on scalar spaces with ADEval::VALUE: L2 projection-type forms
using Q1V = Config<2, 3, Field<2, 1, EV_VALUE>>;
using Q2V = Config<2, 4, Field<3, 1, EV_VALUE>>;
using Mass1 = MassEnergy<1>;
using DiffMass1 = DiffEnergy<Mass1>;
MADB_INSTANCE("mass", Mass1, Q1V, true)
MADB_INSTANCE("mass", Mass1, Q2V, true)
MADB_INSTANCE("diff[mass]", DiffMass1, Q1V, true)
MADB_INSTANCE("diff[mass]", DiffMass1, Q2V, true)
// the target as a GridFunction parameter of the same space (Evaluator GridFunction source)
using Q2VP = Config<2, 4, Field<3, 1, EV_VALUE>, Field<3, 1, EV_VALUE, ROLE_PARAM>>;
MADB_INSTANCE("diff[mass]", DiffMass1, Q2VP, true)

// scalar H1 order 3 (4x4 dofs, 5x5 points): 64-element patches, 4 threads per element
using Q3 = Config<2, 5, Field<4, 1, EV_GRAD>>;
MADB_INSTANCE("diffusion:0", Diff2, Q3, false)
MADB_INSTANCE("minsurf", MinS2, Q3, false)

// Fused-kernel instances: 2-D scalar spaces with ADEval::VALUE (split from instances_scalar2d.cu for build time).
#include "madb_functionals.cuh"
#include "madb_registry.cuh"
using namespace madb;

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

// Load vectors b_i = (f, phi_i) (DomainLFIntegrator, ex4.cpp:145-148): "load" on scalar H1 spaces of order 1 - 3 with
// MFEM's linear-form rule (order 2p -> p+1 points per direction) and with the AD rules of the drivers (p+2 points; 5x5
// for ex4's order-3 space)
using L1a = Config<2, 2, Field<2, 1, EV_VALUE>>;
using L2a = Config<2, 3, Field<3, 1, EV_VALUE>>;
using L3a = Config<2, 4, Field<4, 1, EV_VALUE>>;
using L3b = Config<2, 5, Field<4, 1, EV_VALUE>>;
MADB_INSTANCE("load", LoadFunctional, L1a, true)
MADB_INSTANCE("load", LoadFunctional, Q1V, true)
MADB_INSTANCE("load", LoadFunctional, L2a, true)
MADB_INSTANCE("load", LoadFunctional, Q2V, true)
MADB_INSTANCE("load", LoadFunctional, L3a, false)
MADB_INSTANCE("load", LoadFunctional, L3b, false)

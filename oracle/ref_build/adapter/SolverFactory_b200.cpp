// Replaces ONE translation unit of the reference when the adapter is linked in: src/shared/SolverFactory.cpp
// (/root/reference/src/shared/SolverFactory.cpp:19-52), whose constructor is where solvers are registered. The only
// difference is the creator bound to "MHD_Muscl_3D" (the key has to stay: HydroParams::setup derives nbvar = 8,
// ghostWidth = 3 and mhdEnabled from that exact string, HydroParams.cpp:64-71).
#include "shared/SolverFactory.h"

#include "shared/SolverBase.h"

#include "muscl/SolverHydroMuscl.h"
#include "muscl/SolverMHDMuscl.h"
#include "SolverMHDMusclB200.h"

namespace ppkMHD
{

SolverFactory::SolverFactory()
{
  registerSolver("Hydro_Muscl_2D", &muscl::SolverHydroMuscl<2>::create);
  registerSolver("Hydro_Muscl_3D", &muscl::SolverHydroMuscl<3>::create);
  registerSolver("MHD_Muscl_2D", &muscl::SolverMHDMuscl<2>::create);
  registerSolver("MHD_Muscl_3D", &muscl::SolverMHDMusclB200::create); // was &muscl::SolverMHDMuscl<3>::create
}

} // namespace ppkMHD

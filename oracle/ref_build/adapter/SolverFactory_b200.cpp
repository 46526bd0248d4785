// Replaces ONE translation unit of the reference when the adapter is linked in: the one that defines
// ppkMHD::SolverFactory::SolverFactory() (/root/reference/src/shared/SolverFactory.cpp:19-52), i.e. the place where solver
// names are bound to creation functions. Every key keeps its creator except "MHD_Muscl_3D", which now creates the adapter
// (the key has to stay: HydroParams::setup derives nbvar = 8, ghostWidth = 3 and mhdEnabled from that exact string,
// HydroParams.cpp:64-71).
#include "shared/SolverFactory.h"

#include "muscl/SolverHydroMuscl.h"
#include "muscl/SolverMHDMuscl.h"
#include "SolverMHDMusclB200.h"

namespace ppkMHD
{

namespace
{
struct Binding
{
  const char * key;
  SolverBase * (*create)(HydroParams &, ConfigMap &);
};
const Binding bindings[] = {
  { "MHD_Muscl_3D", &muscl::SolverMHDMusclB200::create }, // was &muscl::SolverMHDMuscl<3>::create
  { "MHD_Muscl_2D", &muscl::SolverMHDMuscl<2>::create },
  { "Hydro_Muscl_3D", &muscl::SolverHydroMuscl<3>::create },
  { "Hydro_Muscl_2D", &muscl::SolverHydroMuscl<2>::create },
};
} // namespace

SolverFactory::SolverFactory()
{
  for (const Binding & b : bindings)
    registerSolver(b.key, b.create);
}

} // namespace ppkMHD

// Run parameters filled from a ConfigMap: the settings contract of the reference
// (src/shared/HydroParams.h:28-65,70-234; HydroParams.cpp:28-217, 223-414, 421-453).
// Only what the MHD_Muscl_3D path reads is kept; the MPI Cartesian topology is replaced by the
// multi-GPU slab topology (mx,my,mz + this process's position), see setup_multi_gpu().
#pragma once
#include <string>

#include "../../include/ppkmhd_b200.h"
#include "ConfigMap.h"

namespace ppkMHD {

using real_t = double;  // the reference is built with PPKMHD_USE_DOUBLE (CMakeLists.txt:28,210-212)

enum DimensionType { TWO_D = 2, THREE_D = 3 };
enum VarIndex { ID = 0, IP = 1, IE = 1, IU = 2, IV = 3, IW = 4, IA = 5, IB = 6, IC = 7, IBX = 5, IBY = 6, IBZ = 7 };
enum BoundaryConditionType { BC_UNDEFINED = 0, BC_DIRICHLET = 1, BC_NEUMANN = 2, BC_PERIODIC = 3, BC_COPY = 4 };
enum RiemannSolverType { RIEMANN_APPROX = 0, RIEMANN_LLF = 1, RIEMANN_HLL = 2, RIEMANN_HLLC = 3, RIEMANN_HLLD = 4 };
enum NeighborLocation { X_MIN = 0, X_MAX = 1, Y_MIN = 2, Y_MAX = 3, Z_MIN = 4, Z_MAX = 5 };

struct HydroSettings {  // HydroParams.h:28-65
  real_t gamma0 = 1.4, gamma6 = 1.0, cfl = 1.0, slope_type = 2.0;
  int iorder = 1;
  real_t smallr = 1e-8, smallc = 1e-8, smallp = 1e-6, smallpp = 1e-6;
  real_t cIso = 0, Omega0 = 0.0, cp = 0.0, mu = 0.0, kappa = 0.0;
};

struct HydroParams {
  int nStepmax = 0;
  real_t tEnd = 0.0;
  int nOutput = 0;
  bool enableOutput = true;
  bool mhdEnabled = false;
  int nlog = 10;
  int nx = 0, ny = 0, nz = 0;
  int ghostWidth = 2;
  int nbvar = 4;
  DimensionType dimType = TWO_D;
  int imin = 0, imax = 0, jmin = 0, jmax = 0, kmin = 0, kmax = 0;
  int isize = 0, jsize = 0, ksize = 0;
  real_t xmin = 0.0, xmax = 1.0, ymin = 0.0, ymax = 1.0, zmin = 0.0, zmax = 1.0;
  real_t dx = 0.0, dy = 0.0, dz = 0.0;
  BoundaryConditionType boundary_type_xmin = BC_UNDEFINED, boundary_type_xmax = BC_UNDEFINED,
                        boundary_type_ymin = BC_UNDEFINED, boundary_type_ymax = BC_UNDEFINED,
                        boundary_type_zmin = BC_UNDEFINED, boundary_type_zmax = BC_UNDEFINED;
  bool ioVTK = true, ioHDF5 = false;
  HydroSettings settings;
  int niter_riemann = 10;
  int riemannSolverType = RIEMANN_APPROX;
  int implementationVersion = 0;

  // multi-GPU slab topology (replaces the USE_MPI block of the reference, HydroParams.h:190-212)
  int mx = 1, my = 1, mz = 1;
  int myRank = 0, nProcs = 1;
  int myMpiPos[3] = {0, 0, 0};
  int neighborsRank[6] = {0, 0, 0, 0, 0, 0};
  BoundaryConditionType neighborsBC[6] = {BC_UNDEFINED, BC_UNDEFINED, BC_UNDEFINED, BC_UNDEFINED, BC_UNDEFINED, BC_UNDEFINED};
  int device = 0;             // CUDA device of this process
  int forcedRank = -1, forcedNranks = 1;  // set before setup() to bypass the RANK/WORLD_SIZE env
  bool exactArithmetic = true;  // [cuda] exact_arithmetic (default on: bit-identical to the reference)

  virtual ~HydroParams() = default;
  // `rank`/`nranks` = position in the process group (env RANK/WORLD_SIZE or explicit); must equal mx*my*mz
  virtual void setup(ConfigMap &configMap);
  void setup_multi_gpu(ConfigMap &configMap, int rank, int nranks);
  void init();
  void print();

  // POD handed to the C ABI
  ppk_mhd3d_params to_c_params() const;
};

}  // namespace ppkMHD

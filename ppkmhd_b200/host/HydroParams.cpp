#include "HydroParams.h"

#include <cstdio>
#include <cstdlib>
#include <iostream>

namespace ppkMHD {

void HydroParams::setup(ConfigMap &cfg) {
  // [run]
  nStepmax = (int)cfg.getInteger("run", "nstepmax", 1000);
  tEnd = cfg.getFloat("run", "tend", 0.0);
  nOutput = (int)cfg.getInteger("run", "noutput", 100);
  if (nOutput == 0) enableOutput = false;
  nlog = (int)cfg.getInteger("run", "nlog", 10);

  // the solver key decides nbvar / ghostWidth (HydroParams.cpp:42-130): any other key keeps the
  // hydro 2-D defaults and only prints a warning, exactly like the reference
  const std::string solver_name = cfg.getString("run", "solver_name", "unknown");
  if (solver_name == "MHD_Muscl_3D") {
    dimType = THREE_D; nbvar = 8; ghostWidth = 3; mhdEnabled = true;
  } else if (solver_name == "MHD_Muscl_2D") {
    dimType = TWO_D; nbvar = 8; ghostWidth = 3; mhdEnabled = true;
  } else if (solver_name == "Hydro_Muscl_3D") {
    dimType = THREE_D; nbvar = 5; ghostWidth = 2;
  } else if (solver_name == "Hydro_Muscl_2D") {
    dimType = TWO_D; nbvar = 4; ghostWidth = 2;
  } else {
    std::cerr << "Solver name not valid : " << solver_name << "\n";
  }

  // [mesh]
  nx = (int)cfg.getInteger("mesh", "nx", 1);
  ny = (int)cfg.getInteger("mesh", "ny", 1);
  nz = (int)cfg.getInteger("mesh", "nz", 1);
  xmin = cfg.getFloat("mesh", "xmin", 0.0);
  ymin = cfg.getFloat("mesh", "ymin", 0.0);
  zmin = cfg.getFloat("mesh", "zmin", 0.0);
  xmax = cfg.getFloat("mesh", "xmax", 1.0);
  ymax = cfg.getFloat("mesh", "ymax", 1.0);
  zmax = cfg.getFloat("mesh", "zmax", 1.0);
  auto bc = [&](const char *key) {
    return static_cast<BoundaryConditionType>(cfg.getInteger("mesh", key, BC_DIRICHLET));
  };
  boundary_type_xmin = bc("boundary_type_xmin");
  boundary_type_xmax = bc("boundary_type_xmax");
  boundary_type_ymin = bc("boundary_type_ymin");
  boundary_type_ymax = bc("boundary_type_ymax");
  boundary_type_zmin = bc("boundary_type_zmin");
  boundary_type_zmax = bc("boundary_type_zmax");

  // [hydro] -- every number goes through float (ConfigMap::getFloat)
  settings.gamma0 = cfg.getFloat("hydro", "gamma0", 1.4);
  settings.cfl = cfg.getFloat("hydro", "cfl", 0.5);
  settings.iorder = (int)cfg.getInteger("hydro", "iorder", 2);
  settings.slope_type = cfg.getFloat("hydro", "slope_type", 1.0);
  settings.smallc = cfg.getFloat("hydro", "smallc", 1e-10);
  settings.smallr = cfg.getFloat("hydro", "smallr", 1e-10);
  settings.cp = cfg.getFloat("hydro", "cp", 0.0);
  settings.mu = cfg.getFloat("hydro", "mu", 0.0);
  settings.kappa = cfg.getFloat("hydro", "kappa", 0.0);
  niter_riemann = (int)cfg.getInteger("hydro", "niter_riemann", 10);

  const std::string riemann = cfg.getString("hydro", "riemann", "approx");
  if (riemann == "approx") riemannSolverType = RIEMANN_APPROX;
  else if (riemann == "llf") riemannSolverType = RIEMANN_LLF;
  else if (riemann == "hll") riemannSolverType = RIEMANN_HLL;
  else if (riemann == "hllc") riemannSolverType = RIEMANN_HLLC;
  else if (riemann == "hlld") riemannSolverType = RIEMANN_HLLD;
  else {
    std::cout << "Riemann Solver specified in parameter file is invalid\n";
    std::cout << "Use the default one : approx\n";
    riemannSolverType = RIEMANN_APPROX;
  }

  implementationVersion = (int)cfg.getFloat("OTHER", "implementationVersion", 0);
  if (implementationVersion != 0 && implementationVersion != 1 && implementationVersion != 2) {
    std::cout << "Implementation version is invalid (must be 0, 1 or 2)\n";
    std::cout << "Use the default : 0\n";
    implementationVersion = 0;
  }

  // [cuda] (new section, ignored by the reference)
  exactArithmetic = cfg.getBool("cuda", "exact_arithmetic", true);
  device = (int)cfg.getInteger("cuda", "device", -1);

  init();

  // process-group shape: explicit, else the env of torchrun-like launchers, default single process
  int rank = forcedRank, nranks = forcedNranks;
  if (rank < 0) {
    rank = 0;
    nranks = 1;
    if (const char *e = getenv("RANK")) rank = atoi(e);
    if (const char *e = getenv("WORLD_SIZE")) nranks = atoi(e);
  }
  setup_multi_gpu(cfg, rank, nranks);
}

void HydroParams::setup_multi_gpu(ConfigMap &cfg, int rank, int nranks) {
  // [mpi] mx,my,mz (HydroParams.cpp:231-233) keep their meaning: a periodic Cartesian grid of sub-domains (slabs,
  // pencils or blocks).
  mx = (int)cfg.getInteger("mpi", "mx", 1);
  my = (int)cfg.getInteger("mpi", "my", 1);
  mz = (int)cfg.getInteger("mpi", "mz", 1);
  if (mx < 1) mx = 1;
  if (my < 1) my = 1;
  if (mz < 1) mz = 1;
  nProcs = mx * my * mz;
  if (nranks != nProcs) {
    if (nProcs != 1 && nranks == 1) {
      // a multi-slab ini driven by a single process: take the slab given by `rank`
    } else if (nProcs != nranks) {
      std::cerr << "Inconsistent Cartesian topology: mx*my*mz must match the number of processes !!!\n";
    }
  }
  myRank = rank;
  // MPI_Cart_create(dims = {mx, my, mz}) numbers the processes with z fastest (HydroParams.cpp:249-277)
  const int r = nProcs > 1 ? rank % nProcs : 0;
  myMpiPos[2] = r % mz;
  myMpiPos[1] = (r / mz) % my;
  myMpiPos[0] = r / (mz * my);
  auto rank_of = [&](int cx, int cy, int cz) { return (((cx + mx) % mx) * my + (cy + my) % my) * mz + (cz + mz) % mz; };
  neighborsRank[X_MIN] = rank_of(myMpiPos[0] - 1, myMpiPos[1], myMpiPos[2]);
  neighborsRank[X_MAX] = rank_of(myMpiPos[0] + 1, myMpiPos[1], myMpiPos[2]);
  neighborsRank[Y_MIN] = rank_of(myMpiPos[0], myMpiPos[1] - 1, myMpiPos[2]);
  neighborsRank[Y_MAX] = rank_of(myMpiPos[0], myMpiPos[1] + 1, myMpiPos[2]);
  neighborsRank[Z_MIN] = rank_of(myMpiPos[0], myMpiPos[1], myMpiPos[2] - 1);
  neighborsRank[Z_MAX] = rank_of(myMpiPos[0], myMpiPos[1], myMpiPos[2] + 1);
  for (int f = 0; f < 6; ++f) neighborsBC[f] = BC_COPY;  // HydroParams.cpp:300-351
  if (myMpiPos[0] == 0) neighborsBC[X_MIN] = boundary_type_xmin;
  if (myMpiPos[0] == mx - 1) neighborsBC[X_MAX] = boundary_type_xmax;
  if (myMpiPos[1] == 0) neighborsBC[Y_MIN] = boundary_type_ymin;
  if (myMpiPos[1] == my - 1) neighborsBC[Y_MAX] = boundary_type_ymax;
  if (myMpiPos[2] == 0) neighborsBC[Z_MIN] = boundary_type_zmin;
  if (myMpiPos[2] == mz - 1) neighborsBC[Z_MAX] = boundary_type_zmax;
  // resolution of the GLOBAL grid (HydroParams.cpp:400-402)
  dx = (xmax - xmin) / (nx * mx);
  dy = (ymax - ymin) / (ny * my);
  dz = (zmax - zmin) / (nz * mz);
  if (device < 0) {
    device = 0;
    if (const char *e = getenv("LOCAL_RANK")) device = atoi(e);
  }
}

void HydroParams::init() {  // HydroParams.cpp:421-453
  imin = jmin = kmin = 0;
  imax = nx - 1 + 2 * ghostWidth;
  jmax = ny - 1 + 2 * ghostWidth;
  kmax = nz - 1 + 2 * ghostWidth;
  isize = imax - imin + 1;
  jsize = jmax - jmin + 1;
  ksize = kmax - kmin + 1;
  dx = (xmax - xmin) / nx;
  dy = (ymax - ymin) / ny;
  dz = (zmax - zmin) / nz;
  settings.smallp = settings.smallc * settings.smallc / settings.gamma0;
  settings.smallpp = settings.smallr * settings.smallp;
  settings.gamma6 = (settings.gamma0 + 1.0) / (2.0 * settings.gamma0);
  if (implementationVersion != 0 && implementationVersion != 1 && implementationVersion != 2) {
    fprintf(stderr, "The implementation version parameter should 0,1 or 2 !!!");
    fprintf(stderr, "Check your parameter file, section OTHER");
    exit(EXIT_FAILURE);
  }
}

void HydroParams::print() {  // same table as HydroParams.cpp:459-506
  printf("##########################\n");
  printf("Simulation run parameters:\n");
  printf("##########################\n");
  printf("nx         : %d\n", nx);
  printf("ny         : %d\n", ny);
  printf("nz         : %d\n", nz);
  printf("dx         : %f\n", dx);
  printf("dy         : %f\n", dy);
  printf("dz         : %f\n", dz);
  printf("imin       : %d\n", imin);
  printf("imax       : %d\n", imax);
  printf("jmin       : %d\n", jmin);
  printf("jmax       : %d\n", jmax);
  printf("kmin       : %d\n", kmin);
  printf("kmax       : %d\n", kmax);
  printf("ghostWidth : %d\n", ghostWidth);
  printf("nbvar      : %d\n", nbvar);
  printf("nStepmax   : %d\n", nStepmax);
  printf("tEnd       : %f\n", tEnd);
  printf("nOutput    : %d\n", nOutput);
  printf("gamma0     : %f\n", settings.gamma0);
  printf("gamma6     : %f\n", settings.gamma6);
  printf("cfl        : %f\n", settings.cfl);
  printf("smallr     : %12.10f\n", settings.smallr);
  printf("smallc     : %12.10f\n", settings.smallc);
  printf("smallp     : %12.10f\n", settings.smallp);
  printf("smallpp    : %g\n", settings.smallpp);
  printf("cp (specific heat)          : %g\n", settings.cp);
  printf("mu (dynamic visosity)       : %g\n", settings.mu);
  printf("kappa (thermal diffusivity) : %g\n", settings.kappa);
  printf("iorder     : %d\n", settings.iorder);
  printf("slope_type : %f\n", settings.slope_type);
  printf("riemann    : %d\n", riemannSolverType);
  printf("implementation version : %d\n", implementationVersion);
  printf("multi-GPU topology     : %dx%dx%d, this sub-domain at (%d,%d,%d), device %d, %s arithmetic\n", mx, my, mz, myMpiPos[0],
         myMpiPos[1], myMpiPos[2], device, exactArithmetic ? "exact (no FMA)" : "fast (FMA)");
  printf("##########################\n");
}

ppk_mhd3d_params HydroParams::to_c_params() const {
  ppk_mhd3d_params p{};
  p.nx = nx; p.ny = ny; p.nz = nz; p.ghost_width = ghostWidth;
  p.xmin = xmin; p.xmax = xmax; p.ymin = ymin; p.ymax = ymax; p.zmin = zmin; p.zmax = zmax;
  p.dx = dx; p.dy = dy; p.dz = dz;
  p.boundary_type[0] = boundary_type_xmin; p.boundary_type[1] = boundary_type_xmax;
  p.boundary_type[2] = boundary_type_ymin; p.boundary_type[3] = boundary_type_ymax;
  p.boundary_type[4] = boundary_type_zmin; p.boundary_type[5] = boundary_type_zmax;
  p.gamma0 = settings.gamma0; p.cfl = settings.cfl; p.slope_type = settings.slope_type;
  p.smallr = settings.smallr; p.smallc = settings.smallc; p.smallp = settings.smallp;
  p.riemann_solver = riemannSolverType;
  p.implementation_version = implementationVersion;
  p.mx = mx; p.my = my; p.mz = mz;
  p.rank_x = myMpiPos[0]; p.rank_y = myMpiPos[1]; p.rank_z = myMpiPos[2];
  p.device = device;
  p.exact_arithmetic = exactArithmetic ? 1 : 0;
  return p;
}

}  // namespace ppkMHD

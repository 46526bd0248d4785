// C entry points of the host layer (include/ppkmhd_b200_host.h) and the program of src/main.cpp.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <thread>

#include "../../include/ppkmhd_b200_host.h"
#include "SolverMHDMusclCuda3D.h"

using namespace ppkMHD;

namespace {
HydroParams params_for(ConfigMap &cfg, int rank_z) {
  HydroParams p;
  const int mz = (int)cfg.getInteger("mpi", "mz", 1);
  p.forcedRank = rank_z;
  p.forcedNranks = mz < 1 ? 1 : mz;
  p.setup(cfg);
  return p;
}

// NCCL bootstrap without MPI: rank 0 publishes the 128-byte unique id in a file next to the output,
// the other processes poll for it (the reference gets its communicator from MPI_Cart_create instead).
int exchange_unique_id(const std::string &path, int rank, unsigned char id[128]) {
  if (rank == 0) {
    if (int rc = ppk_nccl_get_unique_id(id)) return rc;
    const std::string tmp = path + ".tmp";
    std::ofstream(tmp, std::ios::binary).write((const char *)id, 128);
    std::rename(tmp.c_str(), path.c_str());
    return 0;
  }
  for (int tries = 0; tries < 6000; ++tries) {
    std::ifstream in(path, std::ios::binary);
    if (in && in.read((char *)id, 128) && in.gcount() == 128) return 0;
    std::this_thread::sleep_for(std::chrono::milliseconds(10));
  }
  return PPK_ERR_NCCL;
}
}  // namespace

extern "C" {

int ppk_params_from_ini(const char *ini_text, int rank_z, ppk_mhd3d_params *out, double *t_end, int *nstepmax) {
  if (!ini_text || !out) return PPK_ERR_INVALID_ARGUMENT;
  ConfigMap cfg(ini_text, (int)strlen(ini_text));
  HydroParams p = params_for(cfg, rank_z);
  *out = p.to_c_params();
  if (t_end) *t_end = p.tEnd;
  if (nstepmax) *nstepmax = p.nStepmax;
  return 0;
}

int ppk_init_condition_from_ini(const char *ini_text, int rank_z, double *u_host) {
  if (!ini_text || !u_host) return PPK_ERR_INVALID_ARGUMENT;
  ConfigMap cfg(ini_text, (int)strlen(ini_text));
  HydroParams p = params_for(cfg, rank_z);
  DataArray3dHost U(p.isize, p.jsize, p.ksize, p.nbvar);
  init_problem(p, cfg, cfg.getString("hydro", "problem", "unknown"), U);
  memcpy(u_host, U.data(), U.size() * sizeof(double));
  return 0;
}

int ppk_init_condition_2d_from_ini(const char *ini_text, double *u_host) {
  if (!ini_text || !u_host) return PPK_ERR_INVALID_ARGUMENT;
  ConfigMap cfg(ini_text, (int)strlen(ini_text));
  HydroParams p = params_for(cfg, 0);
  DataArray3dHost U(p.isize, p.jsize, 1, p.nbvar);
  init_problem_2d(p, cfg, cfg.getString("hydro", "problem", "unknown"), U);
  memcpy(u_host, U.data(), U.size() * sizeof(double));
  return 0;
}

int ppk_run_ini(const char *ini_path, int rank, int nranks) {
  if (!ini_path) return PPK_ERR_INVALID_ARGUMENT;
  ConfigMap configMap = broadcast_parameters(ini_path);
  if (configMap.ParseError() < 0) {
    fprintf(stderr, "cannot read parameter file %s\n", ini_path);
    return PPK_ERR_INVALID_ARGUMENT;
  }
  HydroParams params;
  if (rank >= 0) {
    params.forcedRank = rank;
    params.forcedNranks = nranks;
  }
  params.setup(configMap);
  const std::string solver_name = configMap.getString("run", "solver_name", "Unknown");
  SolverBase *solver = SolverFactory::Instance().create(solver_name, params, configMap);

  if (params.nProcs > 1) {
    unsigned char id[128];
    const char *port = getenv("MASTER_PORT");
    const std::string path = configMap.getString("output", "outputDir", "./") + "/.ppk_nccl_id_" + (port ? port : "0");
    if (int rc = exchange_unique_id(path, params.myRank, id)) {
      fprintf(stderr, "NCCL unique-id exchange failed: %s\n", ppk_last_error_string());
      return rc;
    }
    static_cast<SolverMHDMusclCuda3D *>(solver)->comm_init(id);
    if (params.myRank == 0) std::remove(path.c_str());
  }

  if (params.nOutput != 0) solver->save_solution();
  if (params.myRank == 0) std::cout << "Start computation....\n";
  solver->timers[TIMER_TOTAL]->start();
  while (!solver->finished()) solver->next_iteration();
  solver->timers[TIMER_TOTAL]->stop();
  if (params.nOutput != 0) solver->save_solution();
  if (params.myRank == 0) printf("final time is %f\n", solver->m_t);
  print_solver_monitoring_info(solver);
  delete solver;
  return 0;
}

}  // extern "C"

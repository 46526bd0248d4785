// C entry points of the host layer (include/ppkmhd_b200_host.h) and the program of src/main.cpp.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <thread>
#include <sys/stat.h>
#include <unistd.h>
#include <ctime>

#include "../../include/ppkmhd_b200_host.h"
#include "SolverMHDMusclCuda3D.h"

using namespace ppkMHD;

namespace {
HydroParams params_for(ConfigMap &cfg, int rank_z) {
  HydroParams p;
  // `rank_z` is the process rank of the sub-domain: the slab index for z-slabs, (x*my + y)*mz + z in general
  const long n = cfg.getInteger("mpi", "mx", 1) * cfg.getInteger("mpi", "my", 1) * cfg.getInteger("mpi", "mz", 1);
  p.forcedRank = rank_z;
  p.forcedNranks = n < 1 ? 1 : (int)n;
  p.setup(cfg);
  return p;
}

// NCCL bootstrap without MPI: rank 0 publishes the 128-byte unique id in a file next to the output,
// the other processes poll for it (the reference gets its communicator from MPI_Cart_create instead).
int exchange_unique_id(const std::string &path, int rank, unsigned char id[128]) {
  if (rank == 0 && id) {
    if (int rc = ppk_nccl_get_unique_id(id)) return rc;
    const std::string tmp = path + ".tmp";
    std::ofstream(tmp, std::ios::binary).write((const char *)id, 128);
    std::rename(tmp.c_str(), path.c_str());
    return 0;
  }
  // a file left behind by a crashed run must not be mistaken for this run's id: the name carries the launcher's pid
  // (see ppk_run_ini) and the file has to be younger than this process
  static const time_t loaded = time(nullptr);  // (first call; ppk_run_ini touches it before it creates the solver)
  if (rank < 0) return 0;
  for (int tries = 0; tries < 6000; ++tries) {
    struct stat sb;
    if (stat(path.c_str(), &sb) == 0 && sb.st_mtime >= loaded - 120) {
      std::ifstream in(path, std::ios::binary);
      if (in && in.read((char *)id, 128) && in.gcount() == 128) return 0;
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(10));
  }
  fprintf(stderr, "rank %d: no NCCL unique id appeared at %s within 60 s\n", rank, path.c_str());
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

int ppk_save_data_from_ini(const char *ini_text, int rank_z, const double *u_host, int i_step) {
  if (!ini_text || !u_host) return PPK_ERR_INVALID_ARGUMENT;
  ConfigMap cfg(ini_text, (int)strlen(ini_text));
  HydroParams p = params_for(cfg, rank_z);
  DataArray3dHost U(p.isize, p.jsize, p.dimType == TWO_D ? 1 : p.ksize, p.nbvar);
  memcpy(U.data(), u_host, U.size() * sizeof(double));
  std::map<int, std::string> names = {{ID, "rho"}, {IP, "energy"}, {IU, "rho_vx"}, {IV, "rho_vy"}, {IW, "rho_vz"},
                                      {IA, "bx"},  {IB, "by"},     {IC, "bz"}};  // SolverBase.cpp:49-56
  io::IO_ReadWrite writer(p, cfg, names);
  writer.save_data(U, i_step, 0.0, "");
  return writer.hdf5_failed ? PPK_ERR_UNSUPPORTED : 0;  // [output] hdf5_enabled without a usable libhdf5
}

int ppk_load_data_from_ini(const char *ini_text, int rank_z, double *u_host, int *i_step, double *time) {
  if (!ini_text || !u_host) return PPK_ERR_INVALID_ARGUMENT;
  ConfigMap cfg(ini_text, (int)strlen(ini_text));
  HydroParams p = params_for(cfg, rank_z);
  DataArray3dHost U(p.isize, p.jsize, p.dimType == TWO_D ? 1 : p.ksize, p.nbvar);
  memcpy(U.data(), u_host, U.size() * sizeof(double));  // (cells the file does not cover keep the caller's values)
  std::map<int, std::string> names = {{ID, "rho"}, {IP, "energy"}, {IU, "rho_vx"}, {IV, "rho_vy"}, {IW, "rho_vz"},
                                      {IA, "bx"},  {IB, "by"},     {IC, "bz"}};
  io::IO_ReadWrite reader(p, cfg, names);
  std::string why;
  int step = 0;
  real_t t = 0;
  if (!reader.load_data(U, step, t, &why)) {
    fprintf(stderr, "ppk_load_data_from_ini: %s\n", why.c_str());
    return PPK_ERR_UNSUPPORTED;
  }
  memcpy(u_host, U.data(), U.size() * sizeof(double));
  if (i_step) *i_step = step;
  if (time) *time = t;
  return 0;
}

int ppk_hdf5_available(void) { return io::hdf5_available(nullptr) ? 1 : 0; }

int ppk_write_xdmf_from_ini(const char *ini_text, int total_number_of_steps, int single_step) {
  if (!ini_text) return PPK_ERR_INVALID_ARGUMENT;
  ConfigMap cfg(ini_text, (int)strlen(ini_text));
  HydroParams p = params_for(cfg, 0);
  std::map<int, std::string> names = {{ID, "rho"}, {IP, "energy"}, {IU, "rho_vx"}, {IV, "rho_vy"}, {IW, "rho_vz"},
                                      {IA, "bx"},  {IB, "by"},     {IC, "bz"}};
  io::writeXdmfForHdf5Wrapper(p, cfg, names, total_number_of_steps, single_step != 0);
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
  exchange_unique_id("", -1, nullptr);  // records the start time of this run (see the staleness check there)
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
    // one file per launch: the ranks of a launch share MASTER_PORT (torchrun) and their parent process (the launcher)
    const char *nonce = getenv("PPK_RUN_NONCE");
    const std::string path = configMap.getString("output", "outputDir", "./") + "/.ppk_nccl_id_" + (port ? port : "0") + "_" +
                             (nonce ? std::string(nonce) : std::to_string((long)getppid()));
    SolverMHDMusclCuda3D *solver3d = dynamic_cast<SolverMHDMusclCuda3D *>(solver);
    if (!solver3d) {
      fprintf(stderr, "[mpi] mx*my*mz > 1 is only supported by MHD_Muscl_3D; %s is single-GPU\n", solver_name.c_str());
      delete solver;
      return PPK_ERR_UNSUPPORTED;
    }
    if (params.myRank == 0) std::remove(path.c_str());  // (rank 0 writes it through a rename: never half-visible)
    if (int rc = exchange_unique_id(path, params.myRank, id)) {
      fprintf(stderr, "NCCL unique-id exchange failed: %s\n", ppk_last_error_string());
      return rc;
    }
    solver3d->comm_init(id);
    if (params.myRank == 0) std::remove(path.c_str());
  }

  if (params.nOutput != 0) solver->save_solution();
  if (params.myRank == 0) std::cout << "Start computation....\n";
  solver->timers[TIMER_TOTAL]->start();
  while (!solver->finished()) solver->next_iteration();
  solver->timers[TIMER_TOTAL]->stop();
  if (params.nOutput != 0) solver->save_solution();
  // the Xdmf wrapper of the HDF5 series (src/main.cpp:163-170)
  if (configMap.getBool("output", "hdf5_enabled", false) && params.myRank == 0)
    io::writeXdmfForHdf5Wrapper(params, configMap, solver->m_variables_names, solver->m_times_saved - 1, false);
  if (params.myRank == 0) printf("final time is %f\n", solver->m_t);
  print_solver_monitoring_info(solver);
  delete solver;
  return 0;
}

}  // extern "C"

#include "SolverBase.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "SolverMHDMusclCuda3D.h"

namespace ppkMHD {

SolverBase::SolverBase(HydroParams &params_, ConfigMap &configMap_) : params(params_), configMap(configMap_) {
  read_config();
  for (int id : {TIMER_TOTAL, TIMER_IO, TIMER_DT, TIMER_BOUNDARIES, TIMER_NUM_SCHEME}) timers[id] = std::make_shared<Timer>();
  // dataset names of the reference (SolverBase.cpp:48-55)
  m_variables_names[ID] = "rho";
  m_variables_names[IP] = "energy";
  m_variables_names[IU] = "rho_vx";
  m_variables_names[IV] = "rho_vy";
  m_variables_names[IW] = "rho_vz";
  m_variables_names[IA] = "bx";
  m_variables_names[IB] = "by";
  m_variables_names[IC] = "bz";
}

SolverBase::~SolverBase() = default;

void SolverBase::read_config() {  // SolverBase.cpp:116-145 (all numbers through float)
  m_t = configMap.getFloat("run", "tCurrent", 0.0);
  m_tEnd = configMap.getFloat("run", "tEnd", 0.0);
  m_dt = m_tEnd;
  m_cfl = configMap.getFloat("hydro", "cfl", 1.0);
  m_nlog = (int)configMap.getFloat("run", "nlog", 10);
  m_iteration = 0;
  m_problem_name = configMap.getString("hydro", "problem", "unknown");
  m_solver_name = configMap.getString("run", "solver_name", "unknown");
  m_restart_run_enabled = configMap.getInteger("run", "restart_enabled", 0) != 0;
  m_restart_run_filename = configMap.getString("run", "restart_filename", "");
}

void SolverBase::compute_dt() {
  m_dt = compute_dt_local();  // a multi-GPU solver returns the already-reduced value
  if (m_t + m_dt > m_tEnd) m_dt = m_tEnd - m_t;
}
double SolverBase::compute_dt_local() { return m_tEnd; }
int SolverBase::finished() { return m_t >= (m_tEnd - 1e-14) || m_iteration >= params.nStepmax; }
void SolverBase::next_iteration() {
  next_iteration_impl();
  ++m_iteration;
  m_t += m_dt;
}
void SolverBase::next_iteration_impl() {}
void SolverBase::save_solution() {
  save_solution_impl();
  ++m_times_saved;
}
void SolverBase::save_solution_impl() {}

int SolverBase::should_save_solution() {
  const double interval = m_tEnd / params.nOutput;
  if (params.nOutput < 0) return 1;
  if ((m_t - (m_times_saved - 1) * interval) > interval) return 1;
  if (std::fabs(m_t - m_tEnd) < 1e-12) return 1;  // ISFUZZYNULL
  return 0;
}

void SolverBase::init_io() { m_io_reader_writer = std::make_shared<io::IO_ReadWrite>(params, configMap, m_variables_names); }

void SolverBase::save_data(DataArray3dHost &Uhost, int iStep, real_t time) {
  m_io_reader_writer->save_data(Uhost, iStep, time, "");
}
void SolverBase::save_data_debug(DataArray3dHost &Uhost, int iStep, real_t time, const std::string &debug_name) {
  m_io_reader_writer->save_data(Uhost, iStep, time, debug_name);
}
void SolverBase::load_data(DataArray3dHost &Uhost, int &iStep, real_t &time) {
  std::string why;
  if (!m_io_reader_writer->load_data(Uhost, iStep, time, &why)) {  // a restart that cannot be honoured must not run something else
    fprintf(stderr, "load_data: %s\n", why.c_str());
    exit(EXIT_FAILURE);
  }
}

// ---------------------------------------------------------------------------------------------
SolverFactory &SolverFactory::Instance() {
  static SolverFactory instance;
  return instance;
}

SolverFactory::SolverFactory() {
  // the key is the reference's: it is what makes HydroParams::setup choose nbvar=8, ghostWidth=3
  registerSolver("MHD_Muscl_3D", &SolverMHDMusclCuda3D::create);
  registerSolver("MHD_Muscl_2D", &SolverMHDMusclCuda2D::create);
}

SolverBase *SolverFactory::create(const std::string &solver_name, HydroParams &params, ConfigMap &configMap) {
  auto it = m_solverCreateMap.find(solver_name);
  if (it != m_solverCreateMap.end()) {
    SolverBase *solver = it->second(params, configMap);
    solver->init_io();
    return solver;
  }
  printf("############ WARNING: ############\n");
  printf("%s: is not recognized as a valid application name key.\n", solver_name.c_str());
  printf("Valid solver names are:\n");
  for (auto &kv : m_solverCreateMap) printf("%s\n", kv.first.c_str());
  printf("############ WARNING: ############\n");
  printf("Solver application name not found\n");
  std::abort();
  return nullptr;
}

void print_solver_monitoring_info(SolverBase *solver) {
  const double t_tot = solver->timers[TIMER_TOTAL]->elapsed();
  const double t_comp = solver->timers[TIMER_NUM_SCHEME]->elapsed();
  const double t_dt = solver->timers[TIMER_DT]->elapsed();
  const double t_bound = solver->timers[TIMER_BOUNDARIES]->elapsed();
  const double t_io = solver->timers[TIMER_IO]->elapsed();
  if (solver->params.myRank != 0) return;
  printf("total       time : %5.3f secondes\n", t_tot);
  printf("godunov     time : %5.3f secondes %5.2f%%\n", t_comp, 100 * t_comp / t_tot);
  printf("compute dt  time : %5.3f secondes %5.2f%%\n", t_dt, 100 * t_dt / t_tot);
  printf("boundaries  time : %5.3f secondes %5.2f%%\n", t_bound, 100 * t_bound / t_tot);
  printf("io          time : %5.3f secondes %5.2f%%\n", t_io, 100 * t_io / t_tot);
  // the reference counts ghost cells (m_nCells = isize*jsize*ksize, SolverMHDMuscl.h:256)
  printf("Perf             : %5.3f number of Mcell-updates/s\n",
         (double)solver->m_iteration * solver->m_nCells * solver->params.nProcs / t_tot * 1e-6);
}

}  // namespace ppkMHD

// The plug-in boundary of the reference, kept as is: abstract solver with the time-loop state that
// main() and the monitoring report read directly (src/shared/SolverBase.h:50-261), and the
// name -> create-function factory (src/shared/SolverFactory.h:43-116).
#pragma once
#include <chrono>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "ConfigMap.h"
#include "HydroParams.h"

namespace ppkMHD {

// host mirror of a DataArray3d in the GPU layout of the reference (LayoutLeft: i fastest, variable
// slowest, src/shared/kokkos_shared.h:25,44-53) so that IO code keeps its Uhost(i,j,k,iVar) idiom
class DataArray3dHost {
public:
  DataArray3dHost() = default;
  DataArray3dHost(int isize, int jsize, int ksize, int nbvar)
    : n_{isize, jsize, ksize, nbvar}, data_((size_t)isize * jsize * ksize * nbvar, 0.0) {}
  real_t &operator()(int i, int j, int k, int v) {
    return data_[(size_t)i + (size_t)n_[0] * ((size_t)j + (size_t)n_[1] * ((size_t)k + (size_t)n_[2] * (size_t)v))];
  }
  const real_t &operator()(int i, int j, int k, int v) const { return const_cast<DataArray3dHost &>(*this)(i, j, k, v); }
  int extent(int d) const { return n_[d]; }
  real_t *data() { return data_.data(); }
  const real_t *data() const { return data_.data(); }
  size_t size() const { return data_.size(); }

private:
  int n_[4] = {0, 0, 0, 0};
  std::vector<real_t> data_;
};

enum TimerIds { TIMER_TOTAL = 0, TIMER_IO = 1, TIMER_DT = 2, TIMER_BOUNDARIES = 3, TIMER_NUM_SCHEME = 4 };  // SolverBase.h:35-42
enum SolverType { SOLVER_UNDEFINED = 0, SOLVER_MUSCL_HANCOCK = 1 };

// wall-clock accumulator with the start/stop/elapsed interface of the reference's timers
// (src/utils/monitoring/OpenMPTimer.h, CudaTimer.h:17-67)
class Timer {
public:
  void start() { t0_ = clock::now(); }
  void stop() { total_ += std::chrono::duration<double>(clock::now() - t0_).count(); }
  double elapsed() const { return total_; }

private:
  using clock = std::chrono::steady_clock;
  clock::time_point t0_{};
  double total_ = 0.0;
};

namespace io { class IO_ReadWrite; }

class SolverBase {
public:
  SolverBase(HydroParams &params, ConfigMap &configMap);
  virtual ~SolverBase();

  HydroParams &params;   // owned by main(), must outlive the solver (SolverBase.h:58-59)
  ConfigMap &configMap;
  int solver_type = SOLVER_UNDEFINED;
  std::map<int, std::shared_ptr<Timer>> timers;

  // time-loop state read by main() / print_solver_monitoring_info
  real_t m_t = 0, m_dt = 0, m_tEnd = 0, m_cfl = 1;
  int m_nlog = 10;
  int m_iteration = 0;
  long m_nCells = -1, m_nDofsPerCell = -1;
  int m_times_saved = 0;
  std::string m_problem_name, m_solver_name;
  bool m_restart_run_enabled = false;
  std::string m_restart_run_filename;
  std::map<int, std::string> m_variables_names;

  virtual void read_config();
  virtual void compute_dt();            // local dt -> global MIN -> clamp to tEnd (SolverBase.cpp:149-179)
  virtual double compute_dt_local();
  virtual int finished();               // SolverBase.cpp:196-201
  virtual void next_iteration();        // SolverBase.cpp:206-220
  virtual void next_iteration_impl();
  virtual void save_solution();         // SolverBase.cpp:235-244
  virtual void save_solution_impl();
  virtual int should_save_solution();   // SolverBase.cpp:265-290
  virtual void init_io();               // called by SolverFactory::create after construction
  virtual void read_restart_file() {}   // an empty TODO in the reference too (SolverBase.cpp:255-260); restart = load_data below
  // ghost fill of the solver's current array, all faces in the reference's order X -> Y -> Z. The GPU path fills a direction
  // per launch, so the per-face make_boundary(Udata, faceId, mhd) of the reference (SolverBase.h:191-193) has no counterpart;
  // the two whole-array entry points map onto the same call (the engine exchanges the faces a neighbour owns, fills the others)
  virtual void make_boundaries() {}
  virtual void make_boundaries_serial() { make_boundaries(); }  // SolverBase.cpp:527-537
  virtual void make_boundaries_mpi() { make_boundaries(); }     // SolverBase.cpp:610-693

  void save_data(DataArray3dHost &Uhost, int iStep, real_t time);                                       // SolverBase.cpp:286-310
  void save_data_debug(DataArray3dHost &Uhost, int iStep, real_t time, const std::string &debug_name);  // SolverBase.h:166-179
  // restart: fills Uhost, iStep and time from [run] restart_filename (SolverBase.h:184-187 -> IO_ReadWrite::load_data);
  // the reference has no error path here (HDF5 is a build-time option): without a usable libhdf5 this prints why and exits
  void load_data(DataArray3dHost &Uhost, int &iStep, real_t &time);

protected:
  std::shared_ptr<io::IO_ReadWrite> m_io_reader_writer;
};

using SolverCreateFn = SolverBase *(*)(HydroParams &params, ConfigMap &configMap);

class SolverFactory {
public:
  static SolverFactory &Instance();
  void registerSolver(const std::string &key, SolverCreateFn cfn) { m_solverCreateMap[key] = cfn; }
  // unknown key: prints the valid names and std::abort()s, like the reference (SolverFactory.h:105-113)
  SolverBase *create(const std::string &solver_name, HydroParams &params, ConfigMap &configMap);

private:
  SolverFactory();
  std::map<std::string, SolverCreateFn> m_solverCreateMap;
};

void print_solver_monitoring_info(SolverBase *solver);  // src/shared/solver_utils.h:16-60

namespace io {
// VTK ImageData writer, byte-compatible with save_VTK_3D (src/utils/io/IO_VTK.cpp:211-408) and, for a
// decomposed run, save_VTK_3D_mpi + write_pvti_header (:630-1008).
class IO_ReadWrite {
public:
  IO_ReadWrite(HydroParams &params, ConfigMap &configMap, std::map<int, std::string> &variables_names);
  void save_data(DataArray3dHost &Uhost, int iStep, real_t time, const std::string &debug_name);
  // restart (IO_ReadWrite::load_data_impl, src/utils/io/IO_ReadWrite.cpp:245-287): [run] restart_filename (.h5) -> Uhost,
  // output counter and time of the file; false (with the reason) when HDF5 is unavailable or the file does not fit
  bool load_data(DataArray3dHost &Uhost, int &iStep, real_t &time, std::string *why);
  bool vtk_enabled = true, hdf5_enabled = false;
  bool hdf5_failed = false;  // an HDF5 output was requested and could not be written (reported once)

private:
  HydroParams &params;
  ConfigMap &configMap;
  std::map<int, std::string> &variables_names;
};
void save_VTK_3D(const DataArray3dHost &Uhost, HydroParams &params, ConfigMap &configMap, int nbvar,
                 const std::map<int, std::string> &variables_names, int iStep, const std::string &debug_name);
// 2-D files of save_VTK_2D (src/utils/io/IO_VTK.cpp:24-206): Uhost is (isize, jsize, 1, nbvar)
void save_VTK_2D(const DataArray3dHost &Uhost, HydroParams &params, ConfigMap &configMap, int nbvar,
                 const std::map<int, std::string> &variables_names, int iStep, const std::string &debug_name);
// HDF5 + XDMF (IO_HDF5.cpp here; src/utils/io/IO_HDF5.h:73-526, :1537-2153 and IO_HDF5.cpp:16-378 in the reference).
// libhdf5 is resolved at run time (dlopen): hdf5_available() says whether it was found.
bool hdf5_available(std::string *why);
void writeXdmfForHdf5Wrapper(HydroParams &params, ConfigMap &configMap, const std::map<int, std::string> &variables_names,
                             int totalNumberOfSteps, bool singleStep);
bool save_HDF5(const DataArray3dHost &Uhost, HydroParams &params, ConfigMap &configMap, const std::map<int, std::string> &variables_names,
               int iStep, real_t totalTime, std::string *why);
bool load_HDF5(DataArray3dHost &Uhost, HydroParams &params, ConfigMap &configMap, const std::map<int, std::string> &variables_names,
               const std::string &filename, int &iStep, real_t &totalTime, std::string *why);
void save_VTK_3D_slab(const DataArray3dHost &Uhost, HydroParams &params, ConfigMap &configMap, int nbvar,
                      const std::map<int, std::string> &variables_names, int iStep, const std::string &debug_name);
}  // namespace io

}  // namespace ppkMHD

// CPU harness for the SolverBase members a reference-side caller may use besides the time loop (src/shared/SolverBase.h:137,
// 166-198): save_data_debug, load_data, read_restart_file, make_boundaries_serial / _mpi. A dummy solver stands in for the GPU one
// (constructing SolverMHDMusclCuda3D needs a device); built and run by tests/test_host_layer.py.
#include <cstdio>
#include <cstring>

#include "SolverBase.h"

using namespace ppkMHD;

struct DummySolver : SolverBase {
  DummySolver(HydroParams &p, ConfigMap &c) : SolverBase(p, c) {}
  int fills = 0;
  void make_boundaries() override { ++fills; }
};

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  ConfigMap configMap(argv[1]);
  if (configMap.ParseError() != 0) return 3;
  HydroParams params;
  params.forcedRank = 0;
  params.forcedNranks = 1;
  params.setup(configMap);
  DummySolver s(params, configMap);
  s.init_io();
  DataArray3dHost U(params.isize, params.jsize, params.ksize, params.nbvar);
  for (size_t i = 0; i < U.size(); ++i) U.data()[i] = (double)(i % 97) * 0.125;
  if (!strcmp(argv[2], "debug")) {
    s.save_data(U, 3, 0.5);
    s.save_data_debug(U, 3, 0.5, "afterBC");
    s.read_restart_file();
    s.make_boundaries_serial();
    s.make_boundaries_mpi();
    printf("fills=%d\n", s.fills);
    return s.fills == 2 ? 0 : 1;
  }
  if (!strcmp(argv[2], "load")) {  // no libhdf5 in this image: must print the reason and exit(EXIT_FAILURE), never return
    int step = -1;
    real_t t = -1;
    s.load_data(U, step, t);
    printf("returned step=%d t=%g\n", step, (double)t);
    return 0;
  }
  return 2;
}

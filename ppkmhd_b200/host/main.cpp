// ppkMHD_b200 <parameter file> : same command line and behaviour as the reference's ppkMHD
// (src/main.cpp:51-189) for the MHD_Muscl_3D solver, without Kokkos.
#include <cstdio>
#include <cstdlib>

#include "../../include/ppkmhd_b200_host.h"

int main(int argc, char *argv[]) {
  if (argc != 2) {
    fprintf(stderr, "Error: wrong number of argument; input filename must be the only parameter on the command line\n");
    return EXIT_FAILURE;
  }
  return ppk_run_ini(argv[1], -1, 1) == 0 ? EXIT_SUCCESS : EXIT_FAILURE;
}

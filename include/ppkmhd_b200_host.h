/*
 * ppkmhd_b200_host.h -- C entry points of the C++ host layer (ppkmhd_b200/host/), i.e. of the part that
 * mirrors the reference's settings / driver code above the solver:
 *   ConfigMap + HydroParams::setup   (src/utils/config/ConfigMap.cpp, src/shared/HydroParams.cpp:28-217)
 *   SolverMHDMuscl<3>::init          (src/muscl/SolverMHDMuscl.h:653-713, MHDInitFunctors3D.h)
 *   main()                           (src/main.cpp:51-189)
 * They exist so that non-C++ callers (the Python tests, bench.py) go through exactly the same parsing and
 * initial-condition code as the ppkMHD_b200 executable.
 */
#ifndef PPKMHD_B200_HOST_H
#define PPKMHD_B200_HOST_H
#include "ppkmhd_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Parse `ini_text` (NUL-terminated) like the reference's main does for slab `rank_z` of the [mpi] mz slabs
 * and fill the POD handed to ppk_mhd3d_create (device and exact_arithmetic come from the optional [cuda]
 * section: device=-1 -> LOCAL_RANK or 0, exact_arithmetic=true). t_end / nstepmax receive [run] tEnd,
 * nStepmax. */
int ppk_params_from_ini(const char *ini_text, int rank_z, ppk_mhd3d_params *params, double *t_end, int *nstepmax);

/* Evaluate the initial condition selected by [hydro] problem on the host (orszag_tang, blast, field_loop, implode,
 * kelvin_helmholtz, rotor, wave: SolverMHDMuscl<3>::init, src/muscl/SolverMHDMuscl.h:653-713; anything else falls back
 * to orszag_tang like the reference) into u_host (8*isize*jsize*ksize doubles). */
int ppk_init_condition_from_ini(const char *ini_text, int rank_z, double *u_host);

/* 2-D path ([run] solver_name=MHD_Muscl_2D): the initial condition selected by [hydro] problem (orszag_tang, blast, rotor,
 * field_loop, kelvin_helmholtz: src/muscl/MHDInitFunctors2D.h; anything else falls back to orszag_tang with a message) into
 * u_host (8*isize*jsize doubles). */
int ppk_init_condition_2d_from_ini(const char *ini_text, double *u_host);

/* SolverBase::save_data (src/shared/SolverBase.cpp:286-310 -> IO_ReadWrite::save_data, src/utils/io/IO_ReadWrite.cpp:60-130)
 * for the state `u_host` (8*isize*jsize*ksize doubles of slab `rank_z`, ghosts included) as output number `i_step`:
 * a single .vti (IO_VTK.cpp:211-408) for an undecomposed run; with [mpi] mz > 1 the piece `_time%07d_mpi%05d.vti` of this
 * slab (IO_VTK.cpp:630-853) and, from slab 0, the `.pvti` header naming every piece (IO_VTK.cpp:860-1008). Host only:
 * needs no GPU (the solvers call the same writer after ppk_mhd3d_download). */
int ppk_save_data_from_ini(const char *ini_text, int rank_z, const double *u_host, int i_step);

/* IO_ReadWrite::load_data (src/utils/io/IO_ReadWrite.cpp:245-287 -> Load_HDF5<d>::load, src/utils/io/IO_HDF5.h:1537-2153): read the
 * restart file named by [run] restart_filename (.h5) into `u_host` (8*isize*jsize*ksize doubles: the interior, or the whole array
 * when the file carries its ghost zones) and return the file's "time step" and "total time" attributes. What
 * SolverMHDMuscl<dim>::init_restart (SolverMHDMuscl.h:615-643) calls; PPK_ERR_UNSUPPORTED without a usable libhdf5 or when the
 * file does not fit the [mesh] sizes. */
int ppk_load_data_from_ini(const char *ini_text, int rank_z, double *u_host, int *i_step, double *time);

/* 1 when a libhdf5 (>= 1.10) was found at run time (dlopen; PPK_HDF5_LIB overrides the search), else 0. The reference
 * decides this at build time (USE_HDF5); with 0, [output] hdf5_enabled=true makes ppk_save_data_from_ini return
 * PPK_ERR_UNSUPPORTED after writing the VTK files, and [run] restart_enabled=true stops the program with a message. */
int ppk_hdf5_available(void);

/* writeXdmfForHdf5Wrapper (src/utils/io/IO_HDF5.cpp:16-378): the Xdmf light-data file <outputPrefix>.xmf (or
 * <outputPrefix>_%07d.xmf when single_step) in the current directory, naming the datasets of <outputPrefix>_%07d.h5 for
 * the outputs 0 .. total_number_of_steps. Plain text: needs no HDF5 library. */
int ppk_write_xdmf_from_ini(const char *ini_text, int total_number_of_steps, int single_step);

/* The whole program of src/main.cpp: read the ini file, create the solver through SolverFactory, run the
 * time loop, write VTK output, print the monitoring table. rank < 0: take RANK / WORLD_SIZE from the env. */
int ppk_run_ini(const char *ini_path, int rank, int nranks);

#ifdef __cplusplus
}
#endif
#endif

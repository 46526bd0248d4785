// Reference-side adapter: what a ppkMHD maintainer adds to the reference tree (as src/muscl/SolverMHDMusclB200.h) to run
// the "MHD_Muscl_3D" hot path on libppkmhd_b200.so.  It is compiled HERE against the reference's own, unmodified
// headers and objects (SolverBase, HydroParams, ConfigMap, kokkos_shared.h, MHDInitFunctors3D.h, IO_ReadWrite) by
// oracle/ref_build/Makefile.adapter -> oracle/_ref/ppkMHD_b200adapter, and tests/test_reference_adapter.py requires that
// binary's .vti output to equal the unmodified reference's byte for byte.
//
// Interface it implements: ppkMHD::SolverBase (/root/reference/src/shared/SolverBase.h:50-261); sequence it mirrors:
// the constructor and next_iteration_impl / save_solution_impl of SolverMHDMuscl<3>
// (/root/reference/src/muscl/SolverMHDMuscl.h:256-420, 747-784, 896-907).  Nothing of the reference is modified.
#ifndef SOLVER_MHD_MUSCL_B200_H_
#define SOLVER_MHD_MUSCL_B200_H_

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "shared/SolverBase.h"
#include "shared/HydroParams.h"
#include "shared/kokkos_shared.h"
#include "muscl/MHDInitFunctors3D.h"
#include "shared/problems/BlastParams.h"
#include "shared/problems/ImplodeParams.h"
#include "shared/problems/RotorParams.h"
#include "shared/problems/WaveParams.h"

#include <ppkmhd_b200.h>  // the C ABI (this repository: include/ppkmhd_b200.h)

namespace ppkMHD
{
namespace muscl
{

class SolverMHDMusclB200 : public SolverBase
{
public:
  static SolverBase *
  create(HydroParams & params, ConfigMap & configMap)
  {
    return new SolverMHDMusclB200(params, configMap);
  }

  SolverMHDMusclB200(HydroParams & params, ConfigMap & configMap)
    : SolverBase(params, configMap)
    , U("U", params.isize, params.jsize, params.ksize, params.nbvar)
    , Uhost(Kokkos::create_mirror(U))
    , m_soa((size_t)params.isize * params.jsize * params.ksize * params.nbvar)
  {
    solver_type = SOLVER_MUSCL_HANCOCK;
    m_nCells = (long long)params.isize * params.jsize * params.ksize; // SolverMHDMuscl.h:256
    m_nDofsPerCell = 1;

    ppk_mhd3d_params q{};
    q.nx = params.nx; q.ny = params.ny; q.nz = params.nz;
    q.ghost_width = params.ghostWidth;
    q.xmin = params.xmin; q.xmax = params.xmax;
    q.ymin = params.ymin; q.ymax = params.ymax;
    q.zmin = params.zmin; q.zmax = params.zmax;
    q.dx = params.dx; q.dy = params.dy; q.dz = params.dz;
    const int bc[6] = { params.boundary_type_xmin, params.boundary_type_xmax, params.boundary_type_ymin,
                        params.boundary_type_ymax, params.boundary_type_zmin, params.boundary_type_zmax };
    for (int f = 0; f < 6; ++f)
      q.boundary_type[f] = bc[f]; // enum BoundaryConditionType == enum ppk_bc (shared/enums.h)
    q.gamma0 = params.settings.gamma0;
    q.cfl = params.settings.cfl;
    q.slope_type = params.settings.slope_type;
    q.smallr = params.settings.smallr;
    q.smallc = params.settings.smallc;
    q.smallp = params.settings.smallp;
    q.riemann_solver = params.riemannSolverType; // enum RiemannSolverType == enum ppk_riemann
    q.implementation_version = params.implementationVersion;
    q.mx = q.my = q.mz = 1;
#ifdef USE_MPI // (members of HydroParams that only exist in an MPI build, HydroParams.h:150-180)
    q.mx = params.mx; q.my = params.my; q.mz = params.mz;
    q.rank_x = params.myMpiPos[0]; q.rank_y = params.myMpiPos[1]; q.rank_z = params.myMpiPos[2];
    m_rank = params.myRank;
#endif
    q.device = 0;
    q.exact_arithmetic = configMap.getInteger("b200", "exact_arithmetic", 1);
    check(ppk_mhd3d_create(&q, &m_handle));
#ifdef USE_MPI // NCCL bootstrap over the reference's own MPI communicator
    if (params.nProcs > 1)
    {
      unsigned char id[128];
      if (m_rank == 0)
        check(ppk_nccl_get_unique_id(id));
      params.communicator->bcast(id, 128, hydroSimu::MpiComm::CHAR, 0);
      check(ppk_mhd3d_comm_init(m_handle, id, params.nProcs, params.myRank));
    }
#endif

    // the reference's own problem set-up (SolverMHDMuscl<3>::init, SolverMHDMuscl.h:649-713), run by its own functors
    // on the default execution space of this build (OpenMP: the "device" array lives in host memory)
    init(U);
    upload();
    check(ppk_mhd3d_set_time(m_handle, m_t, m_tEnd, m_iteration));
    check(ppk_mhd3d_make_boundaries(m_handle)); // SolverMHDMuscl.h:393
    compute_dt();                               // SolverMHDMuscl.h:402 (convertToPrimitives + compute_dt)

    if (m_rank == 0)
    {
      std::cout << "##########################" << "\n";
      std::cout << "Solver is " << m_solver_name << " (libppkmhd_b200: " << ppk_version_string() << ")\n";
      std::cout << "Problem (init condition) is " << m_problem_name << "\n";
      std::cout << "##########################" << "\n";
      params.print();
      std::cout << "##########################" << "\n";
      std::cout << "Memory requested : " << (ppk_mhd3d_device_bytes(m_handle) / 1e6) << " MBytes\n";
      std::cout << "##########################" << "\n";
    }
  }

  ~SolverMHDMusclB200() override { ppk_mhd3d_destroy(m_handle); }

  //! already the global value: on a decomposed run the library all-reduces 1/dt over NCCL
  double
  compute_dt_local() override
  {
    double dt = 0.0;
    check(ppk_mhd3d_compute_dt(m_handle, &dt));
    return dt;
  }

  void
  next_iteration_impl() override // SolverMHDMuscl.h:747-784
  {
    if (m_iteration % m_nlog == 0 && m_rank == 0)
      printf("time step=%7d (dt=% 10.8f t=% 10.8f)\n", m_iteration, m_dt, m_t);
    if (params.enableOutput && should_save_solution())
    {
      if (m_rank == 0)
        std::cout << "Output results at time t=" << m_t << " step " << m_iteration << " dt=" << m_dt << std::endl;
      save_solution();
    }
    timers[TIMER_NUM_SCHEME]->start();
    check(ppk_mhd3d_step(m_handle)); // godunov_unsplit_impl, v0 branch (SolverMHDMuscl.cpp:465-517), dt included
    double dt = 0.0;
    check(ppk_mhd3d_get_time(m_handle, nullptr, &dt, nullptr));
    timers[TIMER_NUM_SCHEME]->stop();
    m_dt = dt; // SolverBase::next_iteration then does ++m_iteration; m_t += m_dt (the addition the device did too)
  }

  void
  save_solution_impl() override // SolverMHDMuscl.h:896-907
  {
    timers[TIMER_IO]->start();
    check(ppk_mhd3d_download(m_handle, m_soa.data()));
    const int isize = params.isize, jsize = params.jsize, ksize = params.ksize, nbvar = params.nbvar;
    for (int v = 0; v < nbvar; ++v)
      for (int k = 0; k < ksize; ++k)
        for (int j = 0; j < jsize; ++j)
          for (int i = 0; i < isize; ++i)
            U(i, j, k, v) = m_soa[i + (size_t)isize * (j + (size_t)jsize * (k + (size_t)ksize * v))];
    save_data(U, Uhost, m_times_saved, m_t);
    timers[TIMER_IO]->stop();
  }

private:
  ppk_mhd3d *              m_handle = nullptr;
  int                      m_rank = 0;
  DataArray3d              U;     //!< the reference's array type: what its init functors fill and its writers read
  DataArray3d::HostMirror  Uhost;
  std::vector<double>      m_soa; //!< the C ABI's layout: i fastest, then j, k, variable (LayoutLeft, any Kokkos backend)

  static void
  check(int rc)
  {
    if (rc)
    {
      fprintf(stderr, "ppkmhd_b200: %s\n", ppk_last_error_string());
      std::abort();
    }
  }

  //! through the layout-independent accessor: an OpenMP build of the reference stores DataArray3d LayoutRight
  void
  upload()
  {
    const int isize = params.isize, jsize = params.jsize, ksize = params.ksize, nbvar = params.nbvar;
    for (int v = 0; v < nbvar; ++v)
      for (int k = 0; k < ksize; ++k)
        for (int j = 0; j < jsize; ++j)
          for (int i = 0; i < isize; ++i)
            m_soa[i + (size_t)isize * (j + (size_t)jsize * (k + (size_t)ksize * v))] = U(i, j, k, v);
    check(ppk_mhd3d_upload(m_handle, m_soa.data()));
  }

  //! SolverMHDMuscl<3>::init (SolverMHDMuscl.h:649-713): same dispatch, same functors (restart needs HDF5)
  void
  init(DataArray3d Udata)
  {
    if (!m_problem_name.compare("blast"))
    {
      BlastParams blastParams = BlastParams(configMap);
      InitBlastFunctor3D_MHD::apply(params, blastParams, Udata);
    }
    else if (!m_problem_name.compare("implode"))
    {
      ImplodeParams implodeParams = ImplodeParams(configMap);
      InitImplodeFunctor3D_MHD::apply(params, implodeParams, Udata);
    }
    else if (!m_problem_name.compare("kelvin_helmholtz"))
    {
      KHParams khParams = KHParams(configMap);
      InitKelvinHelmholtzFunctor3D_MHD::apply(params, khParams, Udata);
    }
    else if (!m_problem_name.compare("rotor"))
    {
      RotorParams rotorParams = RotorParams(configMap);
      InitRotorFunctor3D_MHD::apply(params, rotorParams, Udata);
    }
    else if (!m_problem_name.compare("field_loop") || !m_problem_name.compare("field loop"))
    {
      FieldLoopParams flParams = FieldLoopParams(configMap);
      InitFieldLoopFunctor3D_MHD::apply(params, flParams, Udata);
    }
    else if (!m_problem_name.compare("wave"))
    {
      WaveParams wParams = WaveParams(configMap);
      InitWaveFunctor3D_MHD::apply(params, wParams, Udata);
    }
    else
    {
      if (m_problem_name.compare("orszag_tang"))
      {
        std::cout << "Problem : " << m_problem_name << " is not recognized / implemented." << std::endl;
        std::cout << "Use default - Orszag-Tang vortex" << std::endl;
        m_problem_name = "orszag_tang";
      }
      OrszagTangParams otParams = OrszagTangParams(configMap);
      InitOrszagTangFunctor3D::apply(params, otParams, Udata);
    }
  }
}; // class SolverMHDMusclB200

} // namespace muscl
} // namespace ppkMHD

#endif // SOLVER_MHD_MUSCL_B200_H_

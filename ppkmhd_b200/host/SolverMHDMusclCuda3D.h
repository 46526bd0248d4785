// Drop-in for SolverMHDMuscl<3> (src/muscl/SolverMHDMuscl.h:48-200) registered under "MHD_Muscl_3D":
// same SolverBase interface and time-loop semantics, the Kokkos functors of godunov_unsplit_impl
// (src/muscl/SolverMHDMuscl.cpp:465-517) replaced by the C ABI of include/ppkmhd_b200.h.
#pragma once
#include "SolverBase.h"

namespace ppkMHD {

// problem parameter blocks of the reference (src/shared/problems/*.h), float-precision parsing included
struct BlastParams {
  real_t blast_radius, blast_center_x, blast_center_y, blast_center_z;
  real_t blast_density_in, blast_density_out, blast_pressure_in, blast_pressure_out;
  explicit BlastParams(ConfigMap &configMap);
};
struct OrszagTangParams {
  real_t kt;
  explicit OrszagTangParams(ConfigMap &configMap) { kt = configMap.getFloat("OrszagTang", "kt", 0.0); }
};
struct FieldLoopParams {
  real_t radius, density_in, amplitude, vflow;
  explicit FieldLoopParams(ConfigMap &configMap);
};

// host-side initial conditions (libm sin/cos/sqrt like the reference's OpenMP build); they fill the
// whole array, ghosts included, exactly where the reference's init functors do
struct ImplodeParams {  // src/shared/problems/ImplodeParams.h:14-62
  double rho_out, p_out, u_out, v_out, w_out, Bx_out, By_out, Bz_out;
  double rho_in, p_in, u_in, v_in, w_in, Bx_in, By_in, Bz_in;
  int shape;
  explicit ImplodeParams(ConfigMap &configMap);
};
struct KHParams {  // src/shared/problems/KHParams.h:17-92
  double d_in, d_out, pressure, vflow_in, vflow_out, amplitude, inner_size = 0.0, w0 = 0.0, delta = 0.0;
  bool p_sine, p_sine_rob, p_rand;
  int seed = 0, mode = 0;
  explicit KHParams(ConfigMap &configMap);
};
struct RotorParams {  // src/shared/problems/RotorParams.h:13-26
  double r0, r1, u0, p0, b0;
  explicit RotorParams(ConfigMap &configMap);
};
struct WaveParams {  // src/shared/problems/WaveParams.h:13-150 (linear wave on an axis rotated by (a2, a3))
  double wave_amplitude, wave_V0 = 0.0, rev[7], d0, p0, sin_a2, cos_a2, sin_a3, cos_a3, dby, dbz, bx0, by0, bz0, k_par;
  int wave_type;
  explicit WaveParams(ConfigMap &configMap);
};
void init_wave(const HydroParams &params, const WaveParams &wp, DataArray3dHost &U);                  // MHDInitFunctors3D.h:1034-1308
void init_implode(const HydroParams &params, const ImplodeParams &ip, DataArray3dHost &U);            // MHDInitFunctors3D.h:34-150
void init_kelvin_helmholtz(const HydroParams &params, const KHParams &kh, DataArray3dHost &U);        // MHDInitFunctors3D.h:420-622
void init_rotor(const HydroParams &params, const RotorParams &rp, DataArray3dHost &U);                // MHDInitFunctors3D.h:627-757
void init_orszag_tang(const HydroParams &params, const OrszagTangParams &ot, DataArray3dHost &U);  // MHDInitFunctors3D.h:264-415
void init_blast(const HydroParams &params, const BlastParams &b, DataArray3dHost &U);              // MHDInitFunctors3D.h:155-259
void init_field_loop(const HydroParams &params, const FieldLoopParams &fl, DataArray3dHost &U);    // MHDInitFunctors3D.h:759-1023
// SolverMHDMuscl<dim>::init dispatch (SolverMHDMuscl.h:653-713); returns the problem name actually used
std::string init_problem(const HydroParams &params, ConfigMap &configMap, const std::string &problem, DataArray3dHost &U);
// 2-D path (MHD_Muscl_2D): U is (isize, jsize, 1, 8). InitOrszagTangFunctor2D, src/muscl/MHDInitFunctors2D.h:241-385
void init_orszag_tang_2d(const HydroParams &params, DataArray3dHost &U);
void init_blast_2d(const HydroParams &params, const BlastParams &b, DataArray3dHost &U);               // MHDInitFunctors2D.h:144-236
void init_rotor_2d(const HydroParams &params, const RotorParams &rp, DataArray3dHost &U);              // MHDInitFunctors2D.h:588-668
void init_field_loop_2d(const HydroParams &params, const FieldLoopParams &fl, DataArray3dHost &U);     // MHDInitFunctors2D.h:715-945
void init_kelvin_helmholtz_2d(const HydroParams &params, const KHParams &kh, DataArray3dHost &U);      // MHDInitFunctors2D.h:394-583
void init_implode_2d(const HydroParams &params, const ImplodeParams &ip, DataArray3dHost &U);           // MHDInitFunctors2D.h:36-139
// SolverMHDMuscl<2>::init dispatch; the reference's 2-D wave functor is empty: message + Orszag-Tang, like an unknown name
std::string init_problem_2d(const HydroParams &params, ConfigMap &configMap, const std::string &problem, DataArray3dHost &U);

class SolverMHDMusclCuda3D : public SolverBase {
public:
  SolverMHDMusclCuda3D(HydroParams &params, ConfigMap &configMap);
  ~SolverMHDMusclCuda3D() override;
  static SolverBase *create(HydroParams &params, ConfigMap &configMap) { return new SolverMHDMusclCuda3D(params, configMap); }

  double compute_dt_local() override;
  void next_iteration_impl() override;
  void save_solution_impl() override;
  void make_boundaries() override;

  // attach the NCCL communicator of a decomposed run (unique id produced on rank 0)
  void comm_init(const void *unique_id_128);
  ppk_mhd3d *handle() { return m_handle; }
  DataArray3dHost Uhost;

private:
  void godunov_unsplit();
  ppk_mhd3d *m_handle = nullptr;
};

// Drop-in for SolverMHDMuscl<2> (src/muscl/SolverMHDMuscl.h, dim = 2) registered under "MHD_Muscl_2D": the 2-D time
// loop on one GPU through the ppk_mhd2d_* C ABI. Uhost is (isize, jsize, 1, 8).
class SolverMHDMusclCuda2D : public SolverBase {
public:
  SolverMHDMusclCuda2D(HydroParams &params, ConfigMap &configMap);
  ~SolverMHDMusclCuda2D() override;
  static SolverBase *create(HydroParams &params, ConfigMap &configMap) { return new SolverMHDMusclCuda2D(params, configMap); }

  double compute_dt_local() override;
  void next_iteration_impl() override;
  void save_solution_impl() override;
  void make_boundaries() override;
  DataArray3dHost Uhost;

private:
  ppk_mhd2d *m_handle = nullptr;
};

}  // namespace ppkMHD

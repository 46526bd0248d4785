/*
 * oracle/mhd3d_oracle.h  --  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, CPU restatement of ppkMHD's 3-D MUSCL-Hancock + constrained-transport MHD step
 * ("implementationVersion = 0", the deterministic store-everything variant), written so that the
 * CUDA product path in ppkmhd_b200/ can be checked against an independent implementation.
 *
 *   * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 *     include, link, load or execute anything under oracle/.  The product (ppkmhd_b200/, include/)
 *     never does, and fails loudly when its CUDA library is missing.
 *   * PARITY PIN: this restatement is pinned bit-for-bit against outputs of the unmodified
 *     reference (oracle/_ref/ppkMHD, built from /root/reference by oracle/ref_build/Makefile):
 *     tests/golden/ holds the reference's .vti states and tests/test_oracle_vs_golden.py
 *     requires exact equality (see tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line (relative to /root/reference/) it follows.
 * Arithmetic keeps the reference's operation order; compile with -ffp-contract=off (no FMA), as the
 * reference's own x86-64 build contains none.
 *
 * Data layout: one array U[var][k][j][i] ("LayoutLeft": i fastest, variable slowest),
 *   index = i + isize*(j + jsize*(k + ksize*var)), ghost cells included (ghost width 3),
 *   var order ID=0 rho, IP=1 E (or p), IU,IV,IW = 2,3,4, IA,IB,IC = 5,6,7 (B on the LOWER faces).
 */
#ifndef PPK_MHD3D_ORACLE_H
#define PPK_MHD3D_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_ID = 0, ORC_IP = 1, ORC_IU = 2, ORC_IV = 3, ORC_IW = 4, ORC_IA = 5, ORC_IB = 6, ORC_IC = 7, ORC_NVAR = 8 };
/* src/shared/enums.h BoundaryConditionType */
enum { ORC_BC_UNDEFINED = 0, ORC_BC_DIRICHLET = 1, ORC_BC_NEUMANN = 2, ORC_BC_PERIODIC = 3, ORC_BC_COPY = 4 };

typedef struct orc_params {
  /* local (per-rank) interior sizes and ghost width: HydroParams.cpp:133-135, 64-71 */
  int nx, ny, nz, gw;
  int isize, jsize, ksize;
  /* global domain bounds (already float-rounded by the caller, as ConfigMap::getFloat does) */
  double xmin, xmax, ymin, ymax, zmin, zmax;
  double dx, dy, dz;
  /* BC per face, order xmin,xmax,ymin,ymax,zmin,zmax; ORC_BC_COPY = interior (rank-to-rank) face */
  int bc[6];
  /* HydroSettings: HydroParams.cpp:158-163, :441 */
  double gamma0, cfl, slope_type, smallr, smallc, smallp;
  /* Cartesian decomposition (HydroParams.cpp:231-233) and this rank's position in it */
  int mx, my, mz;
  int px, py, pz;
  /* face Riemann solver, enum RiemannSolverType of src/shared/enums.h: 1 llf, 2 hll, 4 hlld (HydroParams.cpp:175-198);
   * the edge EMFs always use the 2-D HLLD solver (MHDRunFunctors3D.h:2100-2238) */
  int riemann;
} orc_params;

/* float-precision parse of an ini value: ConfigMap.cpp:37-46 (strtof) */
double orc_parse_float(const char *text, double default_value);

/* HydroParams::init (HydroParams.cpp:421-441) + dx rule of setup_mpi (:400-402). */
void orc_params_finalize(orc_params *p);

long orc_ncells(const orc_params *p); /* isize*jsize*ksize */

/* ---- initial conditions (write the whole array incl. ghosts exactly like the reference) ---- */
void orc_init_orszag_tang(const orc_params *p, double kt, double *U);             /* MHDInitFunctors3D.h:264-415 */
void orc_init_blast(const orc_params *p, double radius, double cx, double cy, double cz,
                    double density_in, double density_out, double pressure_in, double pressure_out,
                    double *U);                                                     /* MHDInitFunctors3D.h:155-259 */
void orc_init_field_loop(const orc_params *p, double radius, double density_in, double amplitude,
                         double vflow, double *U);                                  /* MHDInitFunctors3D.h:759-1023 */
void orc_init_implode(const orc_params *p, const double outer[8], const double inner[8], int shape,
                      double *U);                                                   /* MHDInitFunctors3D.h:34-150 */
void orc_init_kelvin_helmholtz(const orc_params *p, double d_in, double d_out, double pressure, double vflow_in,
                               double vflow_out, int mode, double w0, double delta, int sine_robertson,
                               double *U);                                          /* MHDInitFunctors3D.h:420-622 */
void orc_init_rotor(const orc_params *p, double r0, double r1, double u0, double p0, double b0,
                    double *U);                                                     /* MHDInitFunctors3D.h:627-757 */
int orc_init_wave(const orc_params *p, double wave_amplitude, int wave_type, double *U); /* MHDInitFunctors3D.h:1034-1308, WaveParams.h */

/* ---- the step, one function per reference functor ---- */
void orc_make_boundary(const orc_params *p, double *U, int face);                   /* BoundariesFunctors.h:749-1053 */
void orc_make_boundaries(const orc_params *p, double *U);                           /* SolverBase.cpp:527-537 */
void orc_convert_to_primitives(const orc_params *p, const double *U, double *Q);    /* MHDRunFunctors3D.h:88-163 */
double orc_compute_inv_dt(const orc_params *p, const double *Q);                    /* MHDRunFunctors3D.h:16-83 */
double orc_compute_dt_local(const orc_params *p, const double *Q);                  /* SolverMHDMuscl.h:724-741 */

/* scratch for the v0 step: 24 arrays of 8 + E,dA,dB,dC,Emf of 3 => (18+3)*8+15 = 183 doubles/cell */
typedef struct orc_scratch orc_scratch;
orc_scratch *orc_scratch_create(const orc_params *p);
void orc_scratch_destroy(orc_scratch *s);

/* godunov_unsplit_impl v0 AFTER make_boundaries, convertToPrimitives and compute_dt:
 * U_out = U_in; E; dB; trace; fluxes; emf; update; CT update   (SolverMHDMuscl.cpp:477, 490-517) */
void orc_godunov_v0(const orc_params *p, const double *U_in, const double *Q, double *U_out,
                    orc_scratch *s, double dt);

/* Whole single-rank step as SolverMHDMuscl<3>::godunov_unsplit_impl (SolverMHDMuscl.cpp:465-517)
 * + SolverBase::compute_dt clamp (SolverBase.cpp:174-177). Returns the dt used. */
double orc_step(const orc_params *p, double *U_in, double *U_out, double *Q, orc_scratch *s,
                double t, double t_end);

/* debug access to intermediates (pointers into scratch; 8 or 3 components, LayoutLeft) */
const double *orc_scratch_array(const orc_scratch *s, const char *name);

/* diagnostics over interior cells: sums of the 8 conserved variables and max |div B| (first
 * differences of face B; needs valid upper ghost faces, i.e. call after orc_make_boundaries). */
void orc_diagnostics(const orc_params *p, const double *U, double sums[8], double *max_divb);

#ifdef __cplusplus
}
#endif
#endif

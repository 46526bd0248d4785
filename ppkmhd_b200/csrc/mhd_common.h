// Shared between the engine (host) and the two kernel builds (exact / fast arithmetic).
#pragma once
#include <cuda_runtime.h>

namespace ppk {

// variable order of the reference (src/shared/enums.h:17-34)
enum { ID = 0, IP = 1, IU = 2, IV = 3, IW = 4, IA = 5, IB = 6, IC = 7, NBVAR = 8 };
// RiemannSolverType of the reference (src/shared/enums.h); the 3-D MHD path implements llf, hll and hlld
enum { RIEMANN_APPROX = 0, RIEMANN_LLF = 1, RIEMANN_HLL = 2, RIEMANN_HLLC = 3, RIEMANN_HLLD = 4 };
enum { BC_UNDEFINED = 0, BC_DIRICHLET = 1, BC_NEUMANN = 2, BC_PERIODIC = 3, BC_COPY = 4 };

// "Trace basis": what ComputeTraceFunctor3D_MHD (MHDRunFunctors3D.h:543-856) would expand into
// 18 stored states (144 doubles per cell). Every one of those states is an exact 1- or 2-addition
// combination of these 32 numbers (+ the 3 lower-face values of the +1 neighbours) (MHDBaseFunctor3D.h:898-1112), so we store the 32 and rebuild a
// state in registers where a Riemann problem needs it.
enum {
  BQ = 0,       // 8: half-step-updated cell-centred primitives r,p,u,v,w,A,B,C (order ID..IC)
  BSX = 8,      // 7: halved x-slopes  r,p,u,v,w,B,C
  BSY = 15,     // 7: halved y-slopes  r,p,u,v,w,A,C
  BSZ = 22,     // 7: halved z-slopes  r,p,u,v,w,A,B
  BFACE = 29,   // 3: half-step-updated LOWER-face fields AL,BL,CL (a cell's upper-face value is bitwise its +1
                //    neighbour's lower-face value, so it is not stored)
  NBASIS = 32
};
// limited transverse slopes of the face-centred field (DeltaA/B/C of the reference, only the
// 6 non-trivial ones): dA/dy, dA/dz, dB/dx, dB/dz, dC/dx, dC/dy
enum { NDBF = 6, NFLUX = 5, NEMF = 3, NELEC = 3 };

struct GridParams {
  int nx, ny, nz, gw;
  int isize, jsize, ksize;
  long long ncell;  // isize*jsize*ksize
  double dx, dy, dz;
  double idx, idy, idz;  // 1/dx, 1/dy, 1/dz (fast-arithmetic build only)
  double gamma0, cfl, slope_type, smallr, smallc, smallp;
  int wrap_x;   // x periodic on both faces: values at i = nx+gw equal those at i = gw (see launch_flux / k_update)
  int riemann;  // face Riemann solver (the edge EMFs always use the 2-D HLLD solver, like the reference)
  int bc[6];  // effective BC of this slab's faces (BC_COPY on faces owned by the halo exchange)
};

// Device-resident time-loop state (replaces SolverBase::m_t, m_dt, m_iteration living on the host).
struct StepState {
  double t;
  double t_end;
  double dt;
  double dtdx, dtdy, dtdz;         // dt/dx, dt/dy, dt/dz of the current step (SolverMHDMuscl.cpp:490-492)
  unsigned long long inv_dt_bits;  // max over cells of sum_d (c_f,d + |v_d|)/dx_d as ordered bits
  long long iteration;
};

// kernel kinds, for the per-kernel timers
enum KernelKind {
  KK_BOUNDARY = 0, KK_PRIM_DT, KK_FINALIZE_DT, KK_ELEC_DBF, KK_TRACE,
  KK_FLUX_X, KK_FLUX_Y, KK_FLUX_Z, KK_EMF_Z, KK_EMF_Y, KK_EMF_X, KK_UPDATE, KK_DIAG, KK_HALO, KK_CONSUME, KK_HYDRO, KK_UPDATE_CT, KK_DT_ONLY, KK_PRODUCER, KK_RIEMANN_ALL, KK_PLANE_GROUP, KK_XZ_GROUP, KK_COUNT
};

// Launchers exported by each arithmetic build (mhd_kernels.cu compiled twice).
struct KernelTable {
  const char *mode;
  // x / y ghost fill of the planes k in [k0, k1); dir 2 (z) always fills whole planes
  void (*boundary)(const GridParams &g, double *U, int dir, int k0, int k1, cudaStream_t s);
  void (*prim_dt)(const GridParams &g, const double *U, double *Q, StepState *st, int k0, int k1, cudaStream_t s);
  void (*finalize_dt)(const GridParams &g, StepState *st, cudaStream_t s);
  void (*advance_time)(StepState *st, cudaStream_t s);
  void (*elec_dbf)(const GridParams &g, const double *U, const double *Q, double *E, double *DBF, cudaStream_t s);
  void (*trace)(const GridParams &g, const StepState *st, const double *U, const double *Q, const double *E,
                double *BASIS, void *tma, cudaStream_t s);
  // `tma`: context from tma_create (TMA-staged tiles) or nullptr (plain loads)
  void (*flux)(const GridParams &g, int dir, const double *BASIS, double *F, const void *tma, cudaStream_t s);
  void (*emf)(const GridParams &g, int edir, const double *BASIS, const double *DBF, double *EMF, const void *tma, cudaStream_t s);
  void (*update)(const GridParams &g, const StepState *st, const double *Uin, double *Uout, const double *Fx,
                 const double *Fy, const double *Fz, const double *EMF, int k0, int k1, cudaStream_t s);
  void (*diagnostics)(const GridParams &g, const double *U, double *out9, cudaStream_t s);
  void (*fastmath_selftest)(int n, const double *x, double *rcp, double *sq, double *rsq, cudaStream_t s);
  // fused fluxes + EMFs + update; split != 0: two launches (hydro part, CT part)
  void (*consume)(const GridParams &g, const StepState *st, const double *BASIS, const double *DBF, const double *Uin,
                  double *Uout, int split, cudaStream_t s);
  // tensor maps of the basis / face-slope arrays for the TMA-staged kernels (nullptr: not applicable)
  void *(*tma_create)(const GridParams &g, const double *BASIS, const double *DBF);
  void (*tma_destroy)(void *ctx);
  // streamed pipeline: z-marching HLLD fluxes + hydro update (no flux array), then the CT update alone
  void (*hydro)(const GridParams &g, const StepState *st, const double *BASIS, const double *Uin, double *Uout, cudaStream_t s);
  void (*update_ct)(const GridParams &g, const StepState *st, const double *Uin, double *Uout, const double *EMF, cudaStream_t s);
  // A[comp][.., i = nx+gw] <- A[comp][.., i = gw] when x is periodic (debug arrays only)
  void (*wrap_x_column)(const GridParams &g, double *A, int ncomp, cudaStream_t s);
  // 2-D path (MHD_Muscl_2D, v0): mhd2d_kernels.inc. S = the 8 state arrays of the trace (64 numbers per cell),
  // FX / FY = 6 flux components per face, EMF = 1 number per corner
  void (*boundary2d)(const GridParams &g, double *U, int dir, cudaStream_t s);
  void (*prim_dt2d)(const GridParams &g, const double *U, double *Q, StepState *st, cudaStream_t s);
  void (*trace2d)(const GridParams &g, const StepState *st, const double *U, const double *Q, double *S, cudaStream_t s);
  void (*flux_emf2d)(const GridParams &g, const double *S, double *FX, double *FY, double *EMF, cudaStream_t s);
  void (*update2d)(const GridParams &g, const StepState *st, const double *Uin, double *Uout, const double *FX,
                   const double *FY, const double *EMF, cudaStream_t s);
  // tiled pipeline (round 2): CFL reduction alone over the planes [k0, k1); fused producer U -> basis + face-field
  // slopes of the planes [k0, k1) (context: tensor maps of the two conservative arrays; returns -1 if unusable);
  // the six flux / EMF tasks as one L2-ordered launch (returns -1 if unusable)
  void (*dt_only)(const GridParams &g, const double *U, StepState *st, int k0, int k1, cudaStream_t s);
  void *(*prod_create)(const GridParams &g, const double *U0, const double *U1);
  void (*prod_destroy)(void *ctx);
  int (*producer)(const GridParams &g, const StepState *st, const void *ctx, const double *U, double *BASIS, double *DBF,
                  int k0, int k1, cudaStream_t s);
  int (*riemann_all)(const GridParams &g, const double *BASIS, const double *DBF, double *F0, double *F1, double *F2,
                     double *EMF, const void *tma, cudaStream_t s);
  // block decomposition: pack (pack != 0) the gw layers starting at index c0 along dir (0: x, 1: y) into `buf`, or unpack
  // `buf` into them; buffer shape = the reference's border buffers (see k_face_copy)
  void (*face_copy)(const GridParams &g, double *U, double *buf, int dir, int c0, int pack, cudaStream_t s);
  // x-faces + y-faces + z-edges (the three Riemann tasks that read plane k only) in one launch on shared TMA tiles;
  // returns -1 when unavailable (the caller launches flux(0), flux(1), emf(2) instead)
  int (*plane_group)(const GridParams &g, const double *BASIS, const double *DBF, double *F0, double *F1, double *EMF,
                     const void *tma, cudaStream_t s);
  // z-faces + y-edges (the two Riemann tasks that read row j only) in one launch on shared x-z TMA tiles; returns -1 when
  // unavailable (the caller launches flux(2), emf(1) instead)
  int (*xz_group)(const GridParams &g, const double *BASIS, const double *DBF, double *F2, double *EMF, const void *tma,
                  cudaStream_t s);
};

const KernelTable *kernel_table_exact();
const KernelTable *kernel_table_fast();

}  // namespace ppk

/*
 * ppkmhd_b200.h -- C ABI of the B200-native 3-D MUSCL-Hancock + constrained-transport MHD step.
 *
 * The reference (pkestene/ppkMHD) has no FFI: its hot path is reached through the C++ plug-in
 * interface SolverBase / SolverFactory (src/shared/SolverBase.h:50-261, SolverFactory.h:43-116) and
 * executes Kokkos functors.  This header is the boundary the replacement solver
 * (ppkmhd_b200/host/SolverMHDMusclCuda3D.*, registered as "MHD_Muscl_3D") calls instead of those
 * functors.  Each entry point names the reference member / functor sequence it replaces.
 *
 * Conventions: plain pointers and sizes, no C++ / torch types.  Every function returns 0 on
 * success, otherwise a non-zero status (a cudaError_t, or PPK_ERR_* below) and
 * ppk_last_error_string() describes it.  A handle drives ONE GPU (one slab of the domain) and is
 * not re-entrant.  There is no CPU fallback: without a CUDA device ppk_mhd3d_create fails.
 *
 * Array layout (the reference's DataArray3d on a GPU, src/shared/kokkos_shared.h:25,44-53):
 *   U[i + isize*(j + jsize*(k + ksize*var))], isize = nx + 2*3 ..., ghost cells included,
 *   var: 0 rho, 1 E, 2..4 momentum x,y,z, 5..7 Bx,By,Bz on the LOWER x,y,z faces (enums.h:17-34).
 */
#ifndef PPKMHD_B200_H
#define PPKMHD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PPK_NBVAR 8
#define PPK_GHOST_WIDTH 3

enum ppk_status {
  PPK_OK = 0,
  PPK_ERR_INVALID_ARGUMENT = 10001,
  PPK_ERR_UNSUPPORTED = 10002, /* e.g. riemann = approx|hllc, mx*my != 1, ghost width != 3 */
  PPK_ERR_NO_DEVICE = 10003,
  PPK_ERR_NCCL = 10004,
  PPK_ERR_STATE = 10005
};

/* src/shared/enums.h BoundaryConditionType */
enum ppk_bc { PPK_BC_UNDEFINED = 0, PPK_BC_DIRICHLET = 1, PPK_BC_NEUMANN = 2, PPK_BC_PERIODIC = 3, PPK_BC_COPY = 4 };
/* src/shared/enums.h RiemannSolverType */
enum ppk_riemann { PPK_RIEMANN_APPROX = 0, PPK_RIEMANN_LLF = 1, PPK_RIEMANN_HLL = 2, PPK_RIEMANN_HLLC = 3, PPK_RIEMANN_HLLD = 4 };

/* POD mirror of what the functors read from HydroParams / HydroSettings
 * (src/shared/HydroParams.h:28-65,70-234; passed by value into every functor, MHDBaseFunctor3D.h:24-28).
 * Floating-point members must already carry the reference's float-precision parsing
 * (ConfigMap::getFloat, src/utils/config/ConfigMap.cpp:37-46). */
typedef struct ppk_mhd3d_params {
  int nx, ny, nz;      /* LOCAL interior sizes ([mesh] nx,ny,nz are per rank, HydroParams.cpp:400-410) */
  int ghost_width;     /* must be 3 (HydroParams.cpp:64-71) */
  double xmin, xmax, ymin, ymax, zmin, zmax; /* global domain bounds */
  double dx, dy, dz;   /* (xmax-xmin)/(nx*mx) ... (HydroParams.cpp:400-402) */
  int boundary_type[6];/* xmin,xmax,ymin,ymax,zmin,zmax of the GLOBAL domain (enum ppk_bc) */
  double gamma0, cfl, slope_type, smallr, smallc, smallp; /* smallp = smallc*smallc/gamma0 (HydroParams.cpp:441) */
  int riemann_solver;  /* enum ppk_riemann: hlld, hll or llf for the face fluxes (the edge EMFs always use 2-D HLLD) */
  int implementation_version; /* 0 (the deterministic variant, SolverMHDMuscl.cpp:494-517) or 1 (the reference's atomic-scatter
                                 variant of the same arithmetic: runs v0's deterministic kernels, so results equal v0's, which
                                 the reference's own v1 matches to ~1e-15). 3-D rejects 2 (PPK_ERR_UNSUPPORTED); the 2-D entry
                                 points accept 0, 1 and 2 and always run the v0 formulation (v2 agrees to ~1e-15, not bitwise) */
  int mx, my, mz;      /* Cartesian decomposition ([mpi] mx,my,mz): slabs, pencils or blocks; z-slabs (mx = my = 1) are the
                          fast path (no pack kernels, exchange overlapped with the update) */
  int rank_x, rank_y, rank_z; /* position of this sub-domain (replaces myMpiPos, HydroParams.cpp:269-277) */
  int device;          /* CUDA device ordinal to run on */
  int exact_arithmetic;/* 1: kernels compiled --fmad=false, reference operation order => bit-identical
                          to the reference's OpenMP build; 0: same order, FMA contraction allowed */
} ppk_mhd3d_params;

typedef struct ppk_mhd3d ppk_mhd3d; /* opaque per-GPU solver state (U, U2, Q and all scratch arrays) */

/* Replaces the allocation part of SolverMHDMuscl<3>::SolverMHDMuscl (src/muscl/SolverMHDMuscl.h:321-386). */
int ppk_mhd3d_create(const ppk_mhd3d_params *params, ppk_mhd3d **handle);
/* Replaces the View destructors run by `delete solver` (src/main.cpp:183). */
int ppk_mhd3d_destroy(ppk_mhd3d *handle);

/* Host <-> device copy of the CURRENT conservative array (U or U2 by iteration parity,
 * SolverMHDMuscl.h:793-805, :896-907). `u_host` holds 8*isize*jsize*ksize doubles, ghosts included.
 * Replaces Kokkos::deep_copy(Uhost, Udata) (src/utils/io/IO_VTK.cpp:253) and its inverse after init(). */
int ppk_mhd3d_upload(ppk_mhd3d *handle, const double *u_host);
int ppk_mhd3d_download(ppk_mhd3d *handle, double *u_host);
/* The same copy enqueued on the handle's stream WITHOUT waiting for it (pinned `u_host`): with one handle per batch in
 * flight, the download of one batch overlaps the upload and the step of the next ones (bench.py's e2e leg).
 * ppk_mhd3d_synchronize(handle) makes `u_host` valid. Plays the role of the asynchronous deep_copy(exec_space, ...)
 * overloads of Kokkos the reference never uses (its IO_VTK.cpp:253 copy is blocking). */
int ppk_mhd3d_download_async(ppk_mhd3d *handle, double *u_host);

/* Split-phase host transfers: batches pipelined on ONE handle (bench.py's end-to-end leg at sizes where several handles do not
 * fit in HBM). A third conservative array rotates with U and U2 (and a fourth one, allocated the first time an upload
 * finds the third still being downloaded):
 *   ppk_mhd3d_stage_upload  : asynchronous H2D of a full state (pinned `u_host`) into the staging array, on its own copy stream;
 *                             overlaps a running step and a running stage_download of another array;
 *   ppk_mhd3d_stage_swap    : the staged array becomes the current array (the previous current array becomes the staging
 *                             array); the compute stream waits for the copy, nothing else does;
 *   ppk_mhd3d_stage_download: asynchronous D2H of the current array on a second copy stream, ordered after the work already
 *                             enqueued on the compute stream; whatever overwrites that array later waits for the copy.
 * ppk_mhd3d_synchronize waits for all of it. Same role as ppk_mhd3d_upload / ppk_mhd3d_download_async (the reference's only
 * counterpart is the blocking Kokkos::deep_copy of src/utils/io/IO_VTK.cpp:253). */
int ppk_mhd3d_stage_upload(ppk_mhd3d *handle, const double *u_host);
int ppk_mhd3d_stage_swap(ppk_mhd3d *handle);
int ppk_mhd3d_stage_download(ppk_mhd3d *handle, double *u_host);

/* Time bookkeeping of SolverBase (m_t, m_tEnd, m_iteration; SolverBase.cpp:119-124, 206-220). */
int ppk_mhd3d_set_time(ppk_mhd3d *handle, double t, double t_end, long iteration);
/* Synchronises the stream; returns m_t, the last m_dt and m_iteration. Any pointer may be NULL. */
int ppk_mhd3d_get_time(ppk_mhd3d *handle, double *t, double *dt, long *iteration);

/* SolverMHDMuscl<3>::make_boundaries (src/muscl/SolverMHDMuscl.cpp:49-64 ->
 * SolverBase::make_boundaries_serial / _mpi, SolverBase.cpp:527-537, 610-693) on the current array. */
int ppk_mhd3d_make_boundaries(ppk_mhd3d *handle);

/* convertToPrimitives + SolverBase::compute_dt (SolverBase.cpp:149-179) on the current array,
 * as the constructor does (SolverMHDMuscl.h:399-402). Synchronous; *dt receives the global dt. */
int ppk_mhd3d_compute_dt(ppk_mhd3d *handle, double *dt);

/* One SolverBase::next_iteration (SolverBase.cpp:206-220) = godunov_unsplit_impl
 * (src/muscl/SolverMHDMuscl.cpp:465-517, v0) + ++m_iteration, m_t += m_dt.
 * Asynchronous: enqueues the kernels (ghost fill/halo exchange, primitives + CFL reduction, edge
 * electric field + face-B slopes, Hancock trace, HLLD face fluxes, edge EMFs, conservative + CT update)
 * and returns; dt and t live in device memory. */
int ppk_mhd3d_step(ppk_mhd3d *handle);
/* `nsteps` steps back to back without host synchronisation (the loop of src/main.cpp:154-159 for a
 * run that is nStepmax-limited). dt is clamped on the device so that t never passes t_end. */
int ppk_mhd3d_run(ppk_mhd3d *handle, int nsteps);
int ppk_mhd3d_synchronize(ppk_mhd3d *handle);

/* Sums of the 8 conserved variables over the local interior and max |div B| (first differences of
 * face B). Synchronous. No reference counterpart (the reference has no diagnostics); used by the
 * parity tests of SURVEY 8(d). */
int ppk_mhd3d_diagnostics(ppk_mhd3d *handle, double sums[8], double *max_divb);

/* ---- multi-GPU: z-slab decomposition, one handle (one process) per GPU -------------------- */
/* Replaces MpiCommCart / MPI_Sendrecv / MPI_Allreduce (SolverBase.cpp:842-925, :149-179;
 * HydroParams.cpp:249-259) by NCCL send/recv of ghost k-planes over NVLink and an NCCL max-allreduce
 * of 1/dt. `unique_id` is the 128-byte ncclUniqueId produced on rank 0 by ppk_nccl_get_unique_id
 * and distributed by the caller (e.g. torch.distributed broadcast). */
int ppk_nccl_get_unique_id(void *unique_id_128_bytes);
int ppk_mhd3d_comm_init(ppk_mhd3d *handle, const void *unique_id_128_bytes, int nranks, int rank);

/* The message list of one z-halo exchange for the slab described by `params` (pure host function, no GPU,
 * no handle): offsets are in doubles from the start of the conservative array; `count` doubles each.
 * These are the offsets of CopyDataArray_To_BorderBuf<ZMIN/ZMAX> / CopyBorderBuf_To_DataArray
 * (mpiBorderUtils.h:117-123, 286-292) expressed on the array itself: send k in [gw,2gw) down and
 * k in [nz,nz+gw) up, receive into k in [nz+gw,nz+2gw) and [0,gw). Returns the number of messages
 * (0 for mz == 1; faces with a physical, non-periodic BC are not exchanged) or -1 on bad arguments;
 * at most `capacity` entries are written. The engine posts exactly this list inside one NCCL group. */
typedef struct ppk_halo_msg {
  int peer;           /* global rank of the neighbour: (rank_x*my + rank_y)*mz + rank_z' (= rank_z' for z-slabs) */
  int is_send;        /* 1: send, 0: receive */
  int var;            /* variable plane 0..7 */
  long long offset;   /* in doubles from U[0] */
  long long count;    /* doubles */
} ppk_halo_msg;
int ppk_mhd3d_halo_plan(const ppk_mhd3d_params *params, int capacity, ppk_halo_msg *msgs);

/* Block decomposition ([mpi] mx, my > 1; HydroParams.cpp:231-351): the gw ghost layers of an x (dir 0) or y (dir 1) face are
 * strided in memory, so they travel as ONE packed message per face with the shape of the reference's border buffers
 * (borderBufSend_xmin_3d(gw, jsize, ksize, nbvar), SolverBase.cpp:77-95; pack / unpack = CopyDataArray_To_BorderBuf /
 * CopyBorderBuf_To_DataArray, mpiBorderUtils.h:184-330), ghosts of the other directions included:
 *   dir 0: buf[g + gw*(j + jsize*(k + ksize*v))] = U[first_layer+g, j, k, v]
 *   dir 1: buf[i + isize*(g + gw*(k + ksize*v))] = U[i, first_layer+g, k, v]
 * Returns the messages of this sub-domain for direction `dir` (0 when that direction is not decomposed, at most 4) in
 * the order every rank posts them inside one NCCL group. `peer` is a global rank: rank = (rank_x*my + rank_y)*mz + rank_z,
 * the layout MPI_Cart_create gives the reference. make_boundaries order stays X, then Y, then Z (SolverBase.cpp:618-691). */
typedef struct ppk_face_msg {
  int peer;           /* global rank of the neighbour */
  int is_send;        /* 1: send, 0: receive */
  int hi_face;        /* 0: the lower face of this sub-domain along dir, 1: the upper one */
  int first_layer;    /* index along dir of the first of the gw layers packed (send) or overwritten (receive) */
  long long count;    /* doubles in the message: gw * jsize (or isize) * ksize * 8 */
} ppk_face_msg;
int ppk_mhd3d_face_plan(const ppk_mhd3d_params *params, int dir, ppk_face_msg msgs[4]);

/* ---- plumbing ----------------------------------------------------------------------------- */
/* Run on a caller-owned stream (a cudaStream_t, e.g. torch's current stream) instead of the
 * handle's own non-blocking stream. */
int ppk_mhd3d_set_stream(ppk_mhd3d *handle, void *cuda_stream);
/* Kernel schedule of one step (results are identical, bit for bit in exact mode). The default is the one that measured
 * fastest at 256^3 and at 512^3: UNFUSED; every schedule stays selectable and tested.
 *   PPK_PIPELINE_UNFUSED (default): ghost fill | primitives + CFL | edge E + face-B slopes | Hancock trace (TMA-staged) | x-faces +
 *       y-faces + z-edges in ONE launch on shared TMA tiles (the three Riemann tasks that read plane k only) | z-faces | y-edges |
 *       x-edges (one TMA-staged kernel each) | update;
 *       stores Fluxes_x|y|z and Emf like the reference's v0 (what ppk_mhd3d_debug_array exposes);
 *   PPK_PIPELINE_FUSED: after the trace, ONE z-marching consumer kernel = HLLD fluxes x,y,z + edge EMFs z,y,x +
 *       conservative and CT update; fluxes and EMFs never reach HBM (2.2x less DRAM traffic for that part, but
 *       slower today: see DESIGN.md);
 *   PPK_PIPELINE_FUSED_SPLIT: that consumer as two kernels (fluxes + hydro update, EMFs + CT update);
 *   PPK_PIPELINE_STREAMED: after the trace, one z-marching kernel whose threads exchange one-sided states (warp
 *       shuffles / shared memory) solves the three HLLD fluxes of every cell and applies the hydro update (no flux
 *       array), then the TMA-staged EMF kernels and the CT update of the field.
 *   PPK_PIPELINE_ORDERED: UNFUSED with the six flux / EMF kernels replaced by ONE launch whose CTAs are dispatched in the
 *       order (y-slab, plane, task, tile): the basis numbers of a plane come from HBM for the first task that touches them
 *       and from the L2 for the others (the plane-k tasks are one merged work item here too). Works on decomposed runs like UNFUSED;
 *       falls back to UNFUSED where the TMA tiles do not exist (odd nx, nx < 32).
 *   PPK_PIPELINE_TILED (even nx >= 32, mz = 1): ghost fill | CFL reduction (reads U only) |
 *       ONE fused producer kernel = primitives + edge electric field + face-field slopes + hydro slopes + Hancock trace
 *       on TMA-staged U tiles marching in z (Q and E never reach HBM) | ONE launch with the six flux / EMF tasks ordered
 *       (y-slab, plane, task, tile) so that the basis is read from HBM once and from the L2 five times | update. */
enum ppk_pipeline { PPK_PIPELINE_UNFUSED = 0, PPK_PIPELINE_FUSED = 1, PPK_PIPELINE_FUSED_SPLIT = 2, PPK_PIPELINE_STREAMED = 3, PPK_PIPELINE_TILED = 4, PPK_PIPELINE_ORDERED = 5 };
int ppk_mhd3d_set_pipeline(ppk_mhd3d *handle, int pipeline);
int ppk_mhd3d_get_pipeline(ppk_mhd3d *handle); /* the schedule in use (enum ppk_pipeline), -1 for a null handle */
/* Per-kernel CUDA-event timing (replaces the coarse timers of SolverBase.h:35-42 for profiling).
 * While enabled every kernel launch is bracketed by events on the launch stream. */
int ppk_mhd3d_profile(ppk_mhd3d *handle, int enable);
/* The launches timed since ppk_mhd3d_profile(handle, 1), in launch order: kernel name, whether it ran on the handle's
 * communication stream (NCCL exchange, the deferred dt all-reduce) or on the compute stream, start and end in ms after that
 * call. A CUDA-event timeline of the decomposed step (what runs under what); returns the number of entries written. */
int ppk_mhd3d_kernel_timeline(ppk_mhd3d *handle, int capacity, const char **names, int *on_comm_stream, double *start_ms, double *end_ms);
/* Accumulated milliseconds and launch counts per kernel since the last reset; returns the number of
 * kernel kinds (<= capacity). `names[i]` points to static strings. */
int ppk_mhd3d_kernel_times(ppk_mhd3d *handle, int capacity, const char **names, double *ms, long *launches, int reset);
/* Total number of kernels launched by this handle since creation (bench.py's gpu_launches). */
long ppk_mhd3d_launch_count(ppk_mhd3d *handle);
/* Copy an internal device array to the host for tests: "U","U2","Q" (8 comps), "ElecField" (3),
 * "dbf" (6), "basis" (32), "Fluxes_x|y|z" (5), "Emf" (3). Returns the number of components. */
int ppk_mhd3d_debug_array(ppk_mhd3d *handle, const char *name, double *host_out, int *ncomp);
/* Bytes of device memory held by the handle. */
long long ppk_mhd3d_device_bytes(ppk_mhd3d *handle);

/* Evaluate the fast build's reciprocal / sqrt / rsqrt primitives (MUFU seed + one cubic Newton step) on n
 * host values (device 0 / current device); tests compare them with IEEE results (<= 2 ulp). */
int ppk_selftest_fastmath(int n, const double *x_host, double *rcp_out, double *sqrt_out, double *rsqrt_out);

const char *ppk_last_error_string(void);
const char *ppk_version_string(void);

/* ---- 2-D path: SolverMHDMuscl<2> registered as "MHD_Muscl_2D" (src/shared/SolverFactory.cpp:26-30) ------------------------
 * One step = SolverMHDMuscl<2>::godunov_unsplit_impl, implementationVersion 0 (src/muscl/SolverMHDMuscl.cpp:373-417):
 * make_boundaries, deep_copy, ConvertToPrimitivesFunctor2D_MHD, ComputeDtFunctor2D_MHD + SolverBase::compute_dt,
 * ComputeTraceFunctor2D_MHD, ComputeFluxesAndStoreFunctor2D_MHD, ComputeEmfAndStoreFunctor2D, UpdateFunctor2D_MHD,
 * UpdateEmfFunctor2D (src/muscl/MHDRunFunctors2D.h), then ++m_iteration, m_t += m_dt. Single GPU. The parameter block is
 * the 3-D one; nz, dz, the z bounds and the z faces are ignored. Arrays are u[var][j][i] with ghosts:
 * 8 * (ny+6) * (nx+6) doubles. implementationVersion 0, 1 and 2 all run the v0 formulation (the reference's own three variants
 * agree to ~1e-15 in 2-D). Same status codes and error string as the 3-D entry points. */
typedef struct ppk_mhd2d ppk_mhd2d;
int ppk_mhd2d_create(const ppk_mhd3d_params *params, ppk_mhd2d **handle);
int ppk_mhd2d_destroy(ppk_mhd2d *handle);
int ppk_mhd2d_upload(ppk_mhd2d *handle, const double *u_host);
int ppk_mhd2d_download(ppk_mhd2d *handle, double *u_host);
int ppk_mhd2d_set_time(ppk_mhd2d *handle, double t, double t_end, long iteration);
int ppk_mhd2d_get_time(ppk_mhd2d *handle, double *t, double *dt, long *iteration);
int ppk_mhd2d_make_boundaries(ppk_mhd2d *handle);
/* convertToPrimitives + SolverBase::compute_dt on the current array, as the constructor does (SolverMHDMuscl.h:399-402) */
int ppk_mhd2d_compute_dt(ppk_mhd2d *handle, double *dt);
int ppk_mhd2d_step(ppk_mhd2d *handle);
int ppk_mhd2d_run(ppk_mhd2d *handle, int nsteps);
int ppk_mhd2d_synchronize(ppk_mhd2d *handle);
long ppk_mhd2d_launch_count(ppk_mhd2d *handle);

#ifdef __cplusplus
}
#endif
#endif /* PPKMHD_B200_H */

// Engine behind the C ABI of include/ppkmhd_b200.h: owns the device arrays of one z-slab, sequences
// the kernels of one time step on CUDA streams, exchanges ghost planes with NCCL.
//
// Step schedule (replaces SolverMHDMuscl<3>::godunov_unsplit_impl, src/muscl/SolverMHDMuscl.cpp:465-517):
//   main stream : BC x | BC y | ------- prim+CFL (planes needing no z ghost) ---- | wait | BC z(phys) |
//                 prim+CFL (edge planes) | [allreduce max 1/dt] | dt | E+dB | trace | flux x,y,z | emf z,y,x |
//                 update(+CT, +copy) | t += dt
//   comm stream :              wait(x,y) | ncclSend/Recv ghost k-planes (8 vars, both z neighbours) | signal
// dt, t and the iteration counter live in device memory (StepState), so a run of many steps needs no
// host synchronisation at all.
#include "../../include/ppkmhd_b200.h"
#include "mhd_common.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace ppk;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
  g_last_error = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess)                                                                          \
      return fail((int)e_, std::string(#expr) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                             std::to_string(__LINE__) + ")");                                       \
  } while (0)

// ---- NCCL, resolved at run time so that single-GPU use has no NCCL dependency -------------------
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.ok) return 0;
  // prefer a libnccl already mapped into the process (e.g. the one bundled with torch)
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return fail(PPK_ERR_NCCL, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
  g_nccl.lib = lib;
#define SYM(field, name)                                                               \
  *(void **)(&g_nccl.field) = dlsym(lib, name);                                        \
  if (!g_nccl.field) return fail(PPK_ERR_NCCL, std::string("missing NCCL symbol ") + name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(AllReduce, "ncclAllReduce");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.ok = true;
  return 0;
}
#define NCCL_TRY(expr)                                                                         \
  do {                                                                                         \
    ncclResult_t r_ = (expr);                                                                  \
    if (r_ != ncclSuccess)                                                                     \
      return fail(PPK_ERR_NCCL, std::string(#expr) + ": " + g_nccl.GetErrorString(r_));       \
  } while (0)

const char *const kKernelNames[KK_COUNT] = {"boundary", "prim_dt", "finalize_dt", "elec_dbf", "trace",
                                            "flux_x", "flux_y", "flux_z", "emf_z", "emf_y", "emf_x",
                                            "update", "diagnostics", "halo_exchange", "consume", "hydro", "update_ct",
                                            "dt_only", "producer", "riemann_all", "flux_xy_emf_z", "flux_z_emf_y"};

}  // namespace

struct ppk_mhd3d {
  ppk_mhd3d_params params{};
  GridParams g{};
  const KernelTable *kt = nullptr;
  int device = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr, comm_stream = nullptr;
  cudaEvent_t ev_xy = nullptr, ev_halo = nullptr, ev_dt = nullptr;
  double *U[2] = {nullptr, nullptr};
  double *Q = nullptr, *E = nullptr, *DBF = nullptr, *BASIS = nullptr, *F[3] = {nullptr, nullptr, nullptr}, *EMF = nullptr;
  StepState *st = nullptr;
  double *diag = nullptr;
  void *tma = nullptr;  // tensor maps for the TMA-staged flux / EMF kernels
  void *prod = nullptr; // tensor maps of U / U2 for the fused producer (tiled pipeline)
  // split-phase host transfers (ppk_mhd3d_stage_*): a third conservative array and two copy streams
  double *Ustage = nullptr, *Uspare = nullptr;  // upload target / a second one while the first is still being downloaded
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t ev_h2d = nullptr, ev_main = nullptr;
  struct Pending { double *buf; cudaEvent_t done; };
  std::vector<Pending> d2h_pending;  // arrays a staged download is still reading
  long long bytes = 0;
  long launches = 0;
  long host_iteration = 0;  // parity selects U / U2 like SolverMHDMuscl::godunov_unsplit (SolverMHDMuscl.h:793-805)
  // profiling
  bool profile = false;
  struct Timed { int kind; cudaEvent_t a, b; int on_comm; };
  std::vector<Timed> pending;
  // optional timeline of the launches timed since ppk_mhd3d_profile(h, 1): start / end relative to that call
  cudaEvent_t t0 = nullptr;
  struct Span { int kind; int on_comm; float start_ms, end_ms; };
  std::vector<Span> timeline;
  std::vector<cudaEvent_t> pool;
  double acc_ms[KK_COUNT] = {0};
  long acc_n[KK_COUNT] = {0};
  // NCCL
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0, zlo = 0, zhi = 0;
  bool exch_lo = false, exch_hi = false;  // z faces filled by the halo exchange
  bool exch_xy[4] = {false, false, false, false};  // x-lo, x-hi, y-lo, y-hi faces filled by the (packed) halo exchange
  double *fbuf[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};  // [dir][send lo, send hi, recv lo, recv hi]
  int pipeline = PPK_PIPELINE_UNFUSED;
  double *halo_posted = nullptr;  // array whose z exchange was started at the end of the previous step
  bool early_halo = true, defer_dt = true;

  double *cur() { return U[host_iteration & 1]; }
  double *nxt() { return U[(host_iteration + 1) & 1]; }
};

namespace {

struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int now;
    cudaGetDevice(&now);
    if (now != prev) cudaSetDevice(prev);
  }
};

cudaEvent_t get_event(ppk_mhd3d *h) {
  if (!h->pool.empty()) {
    cudaEvent_t e = h->pool.back();
    h->pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

struct Scope {  // brackets one kernel launch with events when profiling
  ppk_mhd3d *h;
  int kind;
  cudaStream_t s;
  cudaEvent_t a = nullptr, b = nullptr;
  Scope(ppk_mhd3d *h_, int kind_, cudaStream_t s_) : h(h_), kind(kind_), s(s_) {
    if (h->profile) {
      a = get_event(h);
      b = get_event(h);
      cudaEventRecord(a, s);
    }
  }
  ~Scope() {
    h->launches += 1;
    if (h->profile) {
      cudaEventRecord(b, s);
      h->pending.push_back({kind, a, b, s == h->comm_stream ? 1 : 0});
    }
  }
};

int collect_timings(ppk_mhd3d *h) {
  for (auto &p : h->pending) {
    CUDA_TRY(cudaEventSynchronize(p.b));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, p.a, p.b));
    h->acc_ms[p.kind] += ms;
    h->acc_n[p.kind] += 1;
    if (h->t0 && h->timeline.size() < 4096) {
      float t_a = 0.f;
      if (cudaEventElapsedTime(&t_a, h->t0, p.a) == cudaSuccess) h->timeline.push_back({p.kind, p.on_comm, t_a, t_a + ms});
      else (void)cudaGetLastError();
    }
    h->pool.push_back(p.a);
    h->pool.push_back(p.b);
  }
  h->pending.clear();
  return 0;
}

int alloc_doubles(ppk_mhd3d *h, double **p, long long n) {
  CUDA_TRY(cudaMalloc((void **)p, (size_t)n * sizeof(double)));
  CUDA_TRY(cudaMemsetAsync(*p, 0, (size_t)n * sizeof(double), h->stream));  // Kokkos::View zero-initialises
  h->bytes += n * (long long)sizeof(double);
  return 0;
}

// z ghost planes through NCCL: replaces CopyDataArray_To_BorderBuf + MPI_Sendrecv + CopyBorderBuf_To_DataArray
// (SolverBase.cpp:842-925, mpiBorderUtils.h:36-330). In the (i fastest, variable slowest) layout the 3 ghost
// planes of one variable are contiguous, so no pack/unpack kernel exists: 8 sends + 8 receives per face.
// The message list is ppk_mhd3d_halo_plan (a pure host function, also driven over gloo by the CPU tests).
int halo_exchange_z(ppk_mhd3d *h, double *U, cudaStream_t s) {
  ppk_halo_msg msgs[4 * NBVAR];
  const int n = ppk_mhd3d_halo_plan(&h->params, 4 * NBVAR, msgs);
  if (n < 0) return fail(PPK_ERR_STATE, "halo plan failed");
  Scope sc(h, KK_HALO, s);
  NCCL_TRY(g_nccl.GroupStart());
  for (int m = 0; m < n; ++m) {
    if (msgs[m].is_send) NCCL_TRY(g_nccl.Send(U + msgs[m].offset, (size_t)msgs[m].count, ncclDouble, msgs[m].peer, h->comm, s));
    else NCCL_TRY(g_nccl.Recv(U + msgs[m].offset, (size_t)msgs[m].count, ncclDouble, msgs[m].peer, h->comm, s));
  }
  NCCL_TRY(g_nccl.GroupEnd());
  return 0;
}

// x or y ghost layers of a block-decomposed run: pack kernel -> one grouped ncclSend / ncclRecv per face -> unpack kernel
// (copy_boundaries + transfert_boundaries_3d + copy_boundaries_back of the reference, SolverBase.cpp:700-925). Runs on
// stream s, in order: the layers that are sent already carry the ghosts of the directions exchanged before.
int halo_exchange_face(ppk_mhd3d *h, double *U, int dir, cudaStream_t s) {
  ppk_face_msg msgs[4];
  const int n = ppk_mhd3d_face_plan(&h->params, dir, msgs);
  if (n <= 0) return n < 0 ? fail(PPK_ERR_STATE, "face plan failed") : 0;
  for (int b = 0; b < 4; ++b)
    if (!h->fbuf[dir][b]) { if (int rc = alloc_doubles(h, &h->fbuf[dir][b], msgs[0].count)) return rc; }
  Scope sc(h, KK_HALO, s);
  for (int m = 0; m < n; ++m)
    if (msgs[m].is_send) h->kt->face_copy(h->g, U, h->fbuf[dir][msgs[m].hi_face], dir, msgs[m].first_layer, 1, s);
  NCCL_TRY(g_nccl.GroupStart());
  for (int m = 0; m < n; ++m) {
    double *buf = h->fbuf[dir][(msgs[m].is_send ? 0 : 2) + msgs[m].hi_face];
    if (msgs[m].is_send) NCCL_TRY(g_nccl.Send(buf, (size_t)msgs[m].count, ncclDouble, msgs[m].peer, h->comm, s));
    else NCCL_TRY(g_nccl.Recv(buf, (size_t)msgs[m].count, ncclDouble, msgs[m].peer, h->comm, s));
  }
  NCCL_TRY(g_nccl.GroupEnd());
  for (int m = 0; m < n; ++m)
    if (!msgs[m].is_send) h->kt->face_copy(h->g, U, h->fbuf[dir][2 + msgs[m].hi_face], dir, msgs[m].first_layer, 0, s);
  h->launches += 2 * n - 1;
  return 0;
}

// an array that a staged download (ppk_mhd3d_stage_download) is still reading must not be overwritten on stream s
int wait_staged_reads(ppk_mhd3d *h, const double *buf, cudaStream_t s) {
  for (auto &pd : h->d2h_pending)
    if (pd.buf == buf) CUDA_TRY(cudaStreamWaitEvent(s, pd.done, 0));
  return 0;
}

// Q and the edge electric field only exist for the schedules that store them (the tiled pipeline keeps both on chip)
int ensure_prim_arrays(ppk_mhd3d *h) {
  const long long n = h->g.ncell;
  if (!h->Q) { if (int rc = alloc_doubles(h, &h->Q, NBVAR * n)) return rc; }
  if (!h->E) { if (int rc = alloc_doubles(h, &h->E, NELEC * n)) return rc; }
  return 0;
}

// make_boundaries on array U, optionally overlapped with the part of prim+CFL that needs no z ghost.
// `defer_dt`: the caller finishes the time step size itself (enqueue_step overlaps the all-reduce with E + dB).
int boundaries_and_primitives(ppk_mhd3d *h, double *U, bool with_prim, bool defer_dt = false) {
  const GridParams &g = h->g;
  cudaStream_t s = h->stream;
  const bool exch = h->exch_lo || h->exch_hi;
  const bool posted = exch && h->halo_posted == U;  // the z exchange of this array started at the end of the last step
  const bool block = h->exch_xy[0] || h->exch_xy[1] || h->exch_xy[2] || h->exch_xy[3];
  if ((exch || block) && !h->comm) return fail(PPK_ERR_STATE, "mx*my*mz > 1 but ppk_mhd3d_comm_init was not called");
  if (with_prim) { if (int rc = ensure_prim_arrays(h)) return rc; }
  if (block) {
    // block decomposition: X -> Y (-> Z below) in the reference's order (SolverBase.cpp:618-691), every direction =
    // local fill of the physical faces, then the exchange of the others; stream-ordered, no overlap
    { Scope sc(h, KK_BOUNDARY, s); h->kt->boundary(g, U, 0, 0, g.ksize, s); }
    if (int rc = halo_exchange_face(h, U, 0, s)) return rc;
    { Scope sc(h, KK_BOUNDARY, s); h->kt->boundary(g, U, 1, 0, g.ksize, s); }
    if (int rc = halo_exchange_face(h, U, 1, s)) return rc;
  } else if (posted) {
    // the edge planes already carry their x / y ghosts (they are being sent), the z ghost planes are being received
    { Scope sc(h, KK_BOUNDARY, s); h->kt->boundary(g, U, 0, 2 * g.gw, g.nz, s); }
    { Scope sc(h, KK_BOUNDARY, s); h->kt->boundary(g, U, 1, 2 * g.gw, g.nz, s); }
  } else {
    { Scope sc(h, KK_BOUNDARY, s); h->kt->boundary(g, U, 0, 0, g.ksize, s); }
    { Scope sc(h, KK_BOUNDARY, s); h->kt->boundary(g, U, 1, 0, g.ksize, s); }
  }
  if (exch) {
    if (!posted) {
      CUDA_TRY(cudaEventRecord(h->ev_xy, s));
      CUDA_TRY(cudaStreamWaitEvent(h->comm_stream, h->ev_xy, 0));
      if (int rc = halo_exchange_z(h, U, h->comm_stream)) return rc;
      CUDA_TRY(cudaEventRecord(h->ev_halo, h->comm_stream));
    }
    h->halo_posted = nullptr;
    if (with_prim) {  // interior planes: Q(k) reads U(k) and U(k+1), both inside [gw, nz+gw)
      Scope sc(h, KK_PRIM_DT, s);
      h->kt->prim_dt(g, U, h->Q, h->st, g.gw, g.nz + g.gw - 1, s);
    }
    CUDA_TRY(cudaStreamWaitEvent(s, h->ev_halo, 0));
  }
  { Scope sc(h, KK_BOUNDARY, s); h->kt->boundary(g, U, 2, 0, g.ksize, s); }  // physical / locally periodic z faces
  if (with_prim) {
    if (exch) {
      { Scope sc(h, KK_PRIM_DT, s); h->kt->prim_dt(g, U, h->Q, h->st, 0, g.gw, s); }
      { Scope sc(h, KK_PRIM_DT, s); h->kt->prim_dt(g, U, h->Q, h->st, g.nz + g.gw - 1, g.ksize - 1, s); }
    } else {
      Scope sc(h, KK_PRIM_DT, s);
      h->kt->prim_dt(g, U, h->Q, h->st, 0, g.ksize - 1, s);
    }
    if (defer_dt) return 0;
    if (h->comm && h->nranks > 1) {
      // MPI_Allreduce(MIN) of dt (SolverBase.cpp:152-165) == max-allreduce of 1/dt: the division
      // cfl/invDt is monotonic, so min_r(cfl/invDt_r) and cfl/max_r(invDt_r) are the same double.
      NCCL_TRY(g_nccl.AllReduce(&h->st->inv_dt_bits, &h->st->inv_dt_bits, 1, ncclDouble, ncclMax, h->comm, s));
    }
    { Scope sc(h, KK_FINALIZE_DT, s); h->kt->finalize_dt(g, h->st, s); }
  }
  return 0;
}

// The tiled pipeline (round 2): ghost fill | CFL reduction (reads U only) | dt | fused producer U -> basis + face-field
// slopes (TMA-staged, z-marching; Q and the edge electric field stay on chip) | the six flux / EMF tasks as one
// L2-ordered launch | update. Same arithmetic, same results as the unfused pipeline, bit for bit in exact mode.
int enqueue_step_tiled(ppk_mhd3d *h) {
  const GridParams &g = h->g;
  cudaStream_t s = h->stream;
  double *Uin = h->cur(), *Uout = h->nxt();
  if (h->exch_lo || h->exch_hi) return fail(PPK_ERR_STATE, "the tiled pipeline is single-slab for now");
  if (int rc = wait_staged_reads(h, Uout, s)) return rc;
  if (!h->F[0] || !h->EMF) {
    if (int rc = ppk_mhd3d_set_pipeline(h, PPK_PIPELINE_TILED)) return rc;
  }
  if (int rc = boundaries_and_primitives(h, Uin, false)) return rc;
  { Scope sc(h, KK_DT_ONLY, s); h->kt->dt_only(g, Uin, h->st, g.gw, g.nz + g.gw, s); }
  { Scope sc(h, KK_FINALIZE_DT, s); h->kt->finalize_dt(g, h->st, s); }
  {
    Scope sc(h, KK_PRODUCER, s);
    if (h->kt->producer(g, h->st, h->prod, Uin, h->BASIS, h->DBF, 2, g.ksize - 2, s) != 0)
      return fail(PPK_ERR_STATE, "fused producer unavailable for this handle");
  }
  {
    Scope sc(h, KK_RIEMANN_ALL, s);
    if (h->kt->riemann_all(g, h->BASIS, h->DBF, h->F[0], h->F[1], h->F[2], h->EMF, h->tma, s) != 0) {
      // PPK_RALL=0: the six tasks as six launches (A/B of the L2-ordered launch)
      h->kt->flux(g, 0, h->BASIS, h->F[0], h->tma, s);
      h->kt->flux(g, 1, h->BASIS, h->F[1], h->tma, s);
      h->kt->flux(g, 2, h->BASIS, h->F[2], h->tma, s);
      h->kt->emf(g, 2, h->BASIS, h->DBF, h->EMF, h->tma, s);
      h->kt->emf(g, 1, h->BASIS, h->DBF, h->EMF, h->tma, s);
      h->kt->emf(g, 0, h->BASIS, h->DBF, h->EMF, h->tma, s);
      h->launches += 5;
    }
  }
  { Scope sc(h, KK_UPDATE, s); h->kt->update(g, h->st, Uin, Uout, h->F[0], h->F[1], h->F[2], h->EMF, 0, g.ksize, s); }
  h->kt->advance_time(h->st, s);
  h->launches += 1;
  h->host_iteration += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int enqueue_step(ppk_mhd3d *h) {
  const GridParams &g = h->g;
  cudaStream_t s = h->stream;
  double *Uin = h->cur(), *Uout = h->nxt();
  if (h->pipeline == PPK_PIPELINE_TILED) return enqueue_step_tiled(h);
  if (int rc = wait_staged_reads(h, Uout, s)) return rc;
  const bool multi = h->comm && h->nranks > 1 && h->defer_dt;
  if (int rc = boundaries_and_primitives(h, Uin, true, multi)) return rc;
  if (multi) {
    // the dt all-reduce (latency + rank skew) runs on the comm stream under E + dB, which do not need dt
    CUDA_TRY(cudaEventRecord(h->ev_xy, s));
    CUDA_TRY(cudaStreamWaitEvent(h->comm_stream, h->ev_xy, 0));
    NCCL_TRY(g_nccl.AllReduce(&h->st->inv_dt_bits, &h->st->inv_dt_bits, 1, ncclDouble, ncclMax, h->comm, h->comm_stream));
    { Scope sc(h, KK_FINALIZE_DT, h->comm_stream); h->kt->finalize_dt(g, h->st, h->comm_stream); }
    CUDA_TRY(cudaEventRecord(h->ev_dt, h->comm_stream));
  }
  { Scope sc(h, KK_ELEC_DBF, s); h->kt->elec_dbf(g, Uin, h->Q, h->E, h->DBF, s); }
  if (multi) CUDA_TRY(cudaStreamWaitEvent(s, h->ev_dt, 0));
  { Scope sc(h, KK_TRACE, s); h->kt->trace(g, h->st, Uin, h->Q, h->E, h->BASIS, h->tma, s); }
  if (h->pipeline == PPK_PIPELINE_UNFUSED || h->pipeline == PPK_PIPELINE_ORDERED) {
    if (!h->F[0]) {  // flux / EMF arrays are allocated on first use of this pipeline
      if (int rc = ppk_mhd3d_set_pipeline(h, h->pipeline)) return rc;
    }
    bool one_launch = false;
    if (h->pipeline == PPK_PIPELINE_ORDERED) {  // the six tasks as ONE launch, ordered (y-slab, plane, task, tile) for the L2
      Scope sc(h, KK_RIEMANN_ALL, s);
      one_launch = h->kt->riemann_all(g, h->BASIS, h->DBF, h->F[0], h->F[1], h->F[2], h->EMF, h->tma, s) == 0;
    }
    if (!one_launch) {
      bool grouped = false;  // x-faces, y-faces and z-edges read plane k only: one launch on shared tiles where available
      { Scope sc(h, KK_PLANE_GROUP, s); grouped = h->kt->plane_group(g, h->BASIS, h->DBF, h->F[0], h->F[1], h->EMF, h->tma, s) == 0; }
      if (!grouped) {
        h->launches -= 1;
        { Scope sc(h, KK_FLUX_X, s); h->kt->flux(g, 0, h->BASIS, h->F[0], h->tma, s); }
        { Scope sc(h, KK_FLUX_Y, s); h->kt->flux(g, 1, h->BASIS, h->F[1], h->tma, s); }
        { Scope sc(h, KK_EMF_Z, s); h->kt->emf(g, 2, h->BASIS, h->DBF, h->EMF, h->tma, s); }
      }
      bool grouped_xz = false;  // z-faces and y-edges read row j only: one launch on shared x-z tiles where available
      { Scope sc(h, KK_XZ_GROUP, s); grouped_xz = h->kt->xz_group(g, h->BASIS, h->DBF, h->F[2], h->EMF, h->tma, s) == 0; }
      if (!grouped_xz) {
        h->launches -= 1;
        { Scope sc(h, KK_FLUX_Z, s); h->kt->flux(g, 2, h->BASIS, h->F[2], h->tma, s); }
        { Scope sc(h, KK_EMF_Y, s); h->kt->emf(g, 1, h->BASIS, h->DBF, h->EMF, h->tma, s); }
      }
      { Scope sc(h, KK_EMF_X, s); h->kt->emf(g, 0, h->BASIS, h->DBF, h->EMF, h->tma, s); }
    }
    const bool exch = h->exch_lo || h->exch_hi;
    if (exch && h->comm && h->early_halo && g.nz >= 2 * g.gw) {
      // Update the planes the neighbours need first, fill their x / y ghosts and start the z exchange of the NEXT
      // step; it then runs under the update of the remaining planes and the next step's primitives. The z ghost
      // planes of Uout are not written here: they are being received.
      const int gw = g.gw, nz = g.nz;
      { Scope sc(h, KK_UPDATE, s); h->kt->update(g, h->st, Uin, Uout, h->F[0], h->F[1], h->F[2], h->EMF, gw, 2 * gw, s); }
      { Scope sc(h, KK_UPDATE, s); h->kt->update(g, h->st, Uin, Uout, h->F[0], h->F[1], h->F[2], h->EMF, nz, nz + gw, s); }
      for (int dir = 0; dir < 2; ++dir) {
        { Scope sc(h, KK_BOUNDARY, s); h->kt->boundary(g, Uout, dir, gw, 2 * gw, s); }
        { Scope sc(h, KK_BOUNDARY, s); h->kt->boundary(g, Uout, dir, nz, nz + gw, s); }
      }
      CUDA_TRY(cudaEventRecord(h->ev_xy, s));
      CUDA_TRY(cudaStreamWaitEvent(h->comm_stream, h->ev_xy, 0));
      if (int rc = halo_exchange_z(h, Uout, h->comm_stream)) return rc;
      CUDA_TRY(cudaEventRecord(h->ev_halo, h->comm_stream));
      h->halo_posted = Uout;
      { Scope sc(h, KK_UPDATE, s); h->kt->update(g, h->st, Uin, Uout, h->F[0], h->F[1], h->F[2], h->EMF, 2 * gw, nz, s); }
    } else {
      Scope sc(h, KK_UPDATE, s);
      h->kt->update(g, h->st, Uin, Uout, h->F[0], h->F[1], h->F[2], h->EMF, 0, g.ksize, s);
    }
  } else if (h->pipeline == PPK_PIPELINE_STREAMED) {
    if (!h->EMF) {
      if (int rc = ppk_mhd3d_set_pipeline(h, PPK_PIPELINE_STREAMED)) return rc;
    }
    { Scope sc(h, KK_HYDRO, s); h->kt->hydro(g, h->st, h->BASIS, Uin, Uout, s); }
    { Scope sc(h, KK_EMF_Z, s); h->kt->emf(g, 2, h->BASIS, h->DBF, h->EMF, h->tma, s); }
    { Scope sc(h, KK_EMF_Y, s); h->kt->emf(g, 1, h->BASIS, h->DBF, h->EMF, h->tma, s); }
    { Scope sc(h, KK_EMF_X, s); h->kt->emf(g, 0, h->BASIS, h->DBF, h->EMF, h->tma, s); }
    { Scope sc(h, KK_UPDATE_CT, s); h->kt->update_ct(g, h->st, Uin, Uout, h->EMF, s); }
  } else {
    Scope sc(h, KK_CONSUME, s);
    h->kt->consume(g, h->st, h->BASIS, h->DBF, Uin, Uout, h->pipeline == PPK_PIPELINE_FUSED_SPLIT, s);
    if (h->pipeline == PPK_PIPELINE_FUSED_SPLIT) h->launches += 1;
  }
  h->kt->advance_time(h->st, s);
  h->launches += 1;
  h->host_iteration += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int create_impl(const ppk_mhd3d_params *p, ppk_mhd3d *h);
int create2d_impl(const ppk_mhd3d_params *p, ppk_mhd2d *h);

// a posted z exchange writes ghost planes of the current array: finish it before the host reads or replaces that array
int settle_halo(ppk_mhd3d *h) {
  if (h->halo_posted) CUDA_TRY(cudaStreamSynchronize(h->comm_stream));
  return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

const char *ppk_last_error_string(void) { return g_last_error.c_str(); }
const char *ppk_version_string(void) { return "ppkmhd_b200 0.1 (sm_100a; MHD_Muscl_3D v0; exact+fast fp64)"; }

int ppk_mhd3d_create(const ppk_mhd3d_params *p, ppk_mhd3d **out) {
  if (!p || !out) return fail(PPK_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (p->ghost_width != PPK_GHOST_WIDTH) return fail(PPK_ERR_UNSUPPORTED, "ghost_width must be 3 (MHD_Muscl_3D)");
  if (p->riemann_solver != PPK_RIEMANN_HLLD && p->riemann_solver != PPK_RIEMANN_HLL && p->riemann_solver != PPK_RIEMANN_LLF)
    return fail(PPK_ERR_UNSUPPORTED, "riemann must be hlld, hll or llf (the reference's 'approx' and 'hllc' are silent no-op fluxes for MHD, RiemannSolvers_MHD.h:372-392)");
  // v1 of the reference is v0's arithmetic with the update scattered through atomic_add (SolverMHDMuscl.cpp:518-534):
  // its result differs from v0 only by the (run-to-run varying) order of those additions, ~1e-15. Both versions map to
  // the same deterministic kernels here. v2 is a different formulation (and faulty in 3-D).
  if (p->implementation_version != 0 && p->implementation_version != 1)
    return fail(PPK_ERR_UNSUPPORTED, "implementationVersion must be 0 or 1 (v1 runs v0's deterministic kernels; v2 is not implemented)");
  if (p->mx < 1 || p->my < 1 || p->mz < 1) return fail(PPK_ERR_INVALID_ARGUMENT, "mx, my, mz must be >= 1");
  if (p->rank_x < 0 || p->rank_x >= p->mx || p->rank_y < 0 || p->rank_y >= p->my || p->rank_z < 0 || p->rank_z >= p->mz)
    return fail(PPK_ERR_INVALID_ARGUMENT, "rank_x / rank_y / rank_z out of range");
  if (p->nx < 3 || p->ny < 3 || p->nz < 3) return fail(PPK_ERR_INVALID_ARGUMENT, "nx, ny, nz must be >= 3 (ghost width)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(PPK_ERR_NO_DEVICE, "no CUDA device: ppkmhd_b200 has no CPU fallback");
  if (p->device < 0 || p->device >= ndev) return fail(PPK_ERR_INVALID_ARGUMENT, "device ordinal out of range");

  ppk_mhd3d *h = new ppk_mhd3d();
  if (int rc = create_impl(p, h)) {  // any failure (stream / event creation, an out-of-memory cudaMalloc at 512^3, ...)
    const std::string msg = g_last_error;
    ppk_mhd3d_destroy(h);              // releases whatever was already allocated
    g_last_error = msg;
    return rc;
  }
  *out = h;
  return 0;
}

} // extern "C"
namespace {
int create_impl(const ppk_mhd3d_params *p, ppk_mhd3d *h) {
  h->params = *p;
  h->device = p->device;
  h->kt = p->exact_arithmetic ? kernel_table_exact() : kernel_table_fast();
  GridParams &g = h->g;
  g.nx = p->nx; g.ny = p->ny; g.nz = p->nz; g.gw = p->ghost_width;
  g.isize = p->nx + 2 * g.gw; g.jsize = p->ny + 2 * g.gw; g.ksize = p->nz + 2 * g.gw;
  g.ncell = (long long)g.isize * g.jsize * g.ksize;
  g.dx = p->dx; g.dy = p->dy; g.dz = p->dz;
  g.idx = 1.0 / p->dx; g.idy = 1.0 / p->dy; g.idz = 1.0 / p->dz;
  g.gamma0 = p->gamma0; g.cfl = p->cfl; g.slope_type = p->slope_type;
  g.smallr = p->smallr; g.smallc = p->smallc; g.smallp = p->smallp;
  g.riemann = p->riemann_solver;
  // (the periodic-copy shortcut only holds where the x ghosts are this sub-domain's own cells)
  g.wrap_x = (p->mx == 1 && p->boundary_type[0] == PPK_BC_PERIODIC && p->boundary_type[1] == PPK_BC_PERIODIC) ? 1 : 0;
  if (const char *e = getenv("PPK_WRAP_X")) g.wrap_x = g.wrap_x && atoi(e) != 0;
  for (int f = 0; f < 6; ++f) g.bc[f] = p->boundary_type[f];
  // x / y faces of a block decomposition ([mpi] mx, my > 1): same rule as for z below
  {
    const int m[2] = {p->mx, p->my}, pos[2] = {p->rank_x, p->rank_y};
    for (int d = 0; d < 2; ++d) {
      if (m[d] <= 1) continue;
      h->exch_xy[2 * d] = pos[d] != 0 || p->boundary_type[2 * d] == PPK_BC_PERIODIC;
      h->exch_xy[2 * d + 1] = pos[d] != m[d] - 1 || p->boundary_type[2 * d + 1] == PPK_BC_PERIODIC;
      for (int side = 0; side < 2; ++side)
        if (h->exch_xy[2 * d + side]) g.bc[2 * d + side] = BC_COPY;
    }
    if (p->mx > 1 || p->my > 1) h->early_halo = false;  // the early z exchange assumes x / y ghosts are local fills
  }
  // z faces of a decomposed run: inner faces, and outer faces of a periodic domain, belong to the halo
  // exchange (HydroParams.cpp:300-351: neighborsBC = BC_COPY unless on the outer boundary).
  if (p->mz > 1) {
    const bool lo_outer = p->rank_z == 0, hi_outer = p->rank_z == p->mz - 1;
    h->exch_lo = !lo_outer || p->boundary_type[4] == PPK_BC_PERIODIC;
    h->exch_hi = !hi_outer || p->boundary_type[5] == PPK_BC_PERIODIC;
    if (h->exch_lo) g.bc[4] = BC_COPY;
    if (h->exch_hi) g.bc[5] = BC_COPY;
    h->zlo = (p->rank_z - 1 + p->mz) % p->mz;
    h->zhi = (p->rank_z + 1) % p->mz;
  }
  DeviceGuard guard(h->device);
  CUDA_TRY(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  {
    // the comm stream carries short kernels (NCCL send/recv, the dt all-reduce) that must not queue behind the
    // thousands of CTAs of a compute kernel launched a moment earlier on the main stream: highest priority
    int lo = 0, hi = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    const char *e = getenv("PPK_COMM_PRIORITY");
    const bool high = !e || atoi(e) != 0;
    CUDA_TRY(cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, high ? hi : lo));
  }
  h->stream = h->own_stream;
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_xy, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_dt, cudaEventDisableTiming));
  if (const char *e = getenv("PPK_EARLY_HALO")) h->early_halo = h->early_halo && atoi(e) != 0;
  if (const char *e = getenv("PPK_DEFER_DT")) h->defer_dt = atoi(e) != 0;
  int rc = 0;
  const long long n = g.ncell;
  if ((rc = alloc_doubles(h, &h->U[0], NBVAR * n)) || (rc = alloc_doubles(h, &h->U[1], NBVAR * n)) ||
      (rc = alloc_doubles(h, &h->DBF, NDBF * n)) || (rc = alloc_doubles(h, &h->BASIS, NBASIS * n)) ||
      (rc = alloc_doubles(h, &h->diag, 16)))
    return rc;
  h->tma = h->kt->tma_create(g, h->BASIS, h->DBF);
  h->prod = h->kt->prod_create(g, h->U[0], h->U[1]);
  // Default schedule, from the A/B measurements of profiles/r2 (B200, Orszag-Tang kt=1): one kernel per functor, with the three
  // Riemann tasks that read plane k only (x-faces, y-faces, z-edges) merged into one launch on shared tiles (mhd_pgroup.inc):
  // 6.84 ms at 256^3 (tiled 7.25, ordered 7.10), 56.0 ms at 512^3 (ordered 56.9, tiled 61.5).
  h->pipeline = PPK_PIPELINE_UNFUSED;
  if (const char *e = getenv("PPK_PIPELINE")) {
    const int want = atoi(e);
    if (want == PPK_PIPELINE_ORDERED && !h->tma) h->pipeline = PPK_PIPELINE_UNFUSED;
    else if (want != PPK_PIPELINE_TILED || (h->tma && h->prod && p->mz == 1)) h->pipeline = want;
  }
  CUDA_TRY(cudaMalloc((void **)&h->st, sizeof(StepState)));
  StepState st0{};
  st0.t = 0.0; st0.t_end = 1e300; st0.dt = 0.0; st0.inv_dt_bits = 0ull; st0.iteration = 0;
  CUDA_TRY(cudaMemcpyAsync(h->st, &st0, sizeof(st0), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}
}  // namespace
extern "C" {

int ppk_mhd3d_destroy(ppk_mhd3d *h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  if (h->comm && g_nccl.ok) g_nccl.CommDestroy(h->comm);
  for (double *p : {h->U[0], h->U[1], h->Ustage, h->Uspare, h->Q, h->E, h->DBF, h->BASIS, h->F[0], h->F[1], h->F[2], h->EMF, h->diag})
    if (p) cudaFree(p);
  for (int d = 0; d < 2; ++d)
    for (int b = 0; b < 4; ++b)
      if (h->fbuf[d][b]) cudaFree(h->fbuf[d][b]);
  for (auto &pd : h->d2h_pending) cudaEventDestroy(pd.done);
  if (h->t0) cudaEventDestroy(h->t0);
  if (h->ev_h2d) cudaEventDestroy(h->ev_h2d);
  if (h->ev_main) cudaEventDestroy(h->ev_main);
  if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
  if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
  if (h->st) cudaFree(h->st);
  if (h->tma && h->kt) h->kt->tma_destroy(h->tma);
  if (h->prod && h->kt) h->kt->prod_destroy(h->prod);
  for (auto &p : h->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto e : h->pool) cudaEventDestroy(e);
  if (h->ev_xy) cudaEventDestroy(h->ev_xy);
  if (h->ev_halo) cudaEventDestroy(h->ev_halo);
  if (h->ev_dt) cudaEventDestroy(h->ev_dt);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  delete h;
  return 0;
}

int ppk_mhd3d_upload(ppk_mhd3d *h, const double *u_host) {
  if (!h || !u_host) return fail(PPK_ERR_INVALID_ARGUMENT, "null argument");
  DeviceGuard guard(h->device);
  if (int rc = settle_halo(h)) return rc;
  h->halo_posted = nullptr;  // the array is replaced: its ghosts are exchanged again at the next step
  if (int rc = wait_staged_reads(h, h->cur(), h->stream)) return rc;
  CUDA_TRY(cudaMemcpyAsync(h->cur(), u_host, (size_t)NBVAR * h->g.ncell * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return 0;
}

int ppk_mhd3d_download(ppk_mhd3d *h, double *u_host) {
  if (!h || !u_host) return fail(PPK_ERR_INVALID_ARGUMENT, "null argument");
  DeviceGuard guard(h->device);
  if (int rc = settle_halo(h)) return rc;
  CUDA_TRY(cudaMemcpyAsync(u_host, h->cur(), (size_t)NBVAR * h->g.ncell * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

int ppk_mhd3d_download_async(ppk_mhd3d *h, double *u_host) {
  if (!h || !u_host) return fail(PPK_ERR_INVALID_ARGUMENT, "null argument");
  DeviceGuard guard(h->device);
  if (int rc = settle_halo(h)) return rc;
  CUDA_TRY(cudaMemcpyAsync(u_host, h->cur(), (size_t)NBVAR * h->g.ncell * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  return 0;
}

int ppk_mhd3d_set_time(ppk_mhd3d *h, double t, double t_end, long iteration) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  DeviceGuard guard(h->device);
  if (int rc = settle_halo(h)) return rc;
  // keep the array that is current now current after the parity change
  if (((iteration ^ h->host_iteration) & 1) != 0) std::swap(h->U[0], h->U[1]);
  StepState st{};
  st.t = t; st.t_end = t_end; st.dt = 0.0; st.inv_dt_bits = 0ull; st.iteration = iteration;
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaMemcpy(h->st, &st, sizeof(st), cudaMemcpyHostToDevice));
  h->host_iteration = iteration;
  return 0;
}

int ppk_mhd3d_get_time(ppk_mhd3d *h, double *t, double *dt, long *iteration) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  DeviceGuard guard(h->device);
  StepState st{};
  CUDA_TRY(cudaMemcpyAsync(&st, h->st, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (t) *t = st.t;
  if (dt) *dt = st.dt;
  if (iteration) *iteration = (long)st.iteration;
  return 0;
}

int ppk_mhd3d_make_boundaries(ppk_mhd3d *h) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  DeviceGuard guard(h->device);
  if (int rc = boundaries_and_primitives(h, h->cur(), false)) return rc;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int ppk_mhd3d_compute_dt(ppk_mhd3d *h, double *dt) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  DeviceGuard guard(h->device);
  if (h->pipeline == PPK_PIPELINE_TILED) {  // the tiled pipeline never stores Q: CFL reduction straight from U
    if (int rc = boundaries_and_primitives(h, h->cur(), false)) return rc;
    { Scope sc(h, KK_DT_ONLY, h->stream); h->kt->dt_only(h->g, h->cur(), h->st, h->g.gw, h->g.nz + h->g.gw, h->stream); }
    { Scope sc(h, KK_FINALIZE_DT, h->stream); h->kt->finalize_dt(h->g, h->st, h->stream); }
  } else if (int rc = boundaries_and_primitives(h, h->cur(), true)) return rc;
  CUDA_TRY(cudaGetLastError());
  return ppk_mhd3d_get_time(h, nullptr, dt, nullptr);
}

int ppk_mhd3d_step(ppk_mhd3d *h) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  DeviceGuard guard(h->device);
  return enqueue_step(h);
}

int ppk_mhd3d_run(ppk_mhd3d *h, int nsteps) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  DeviceGuard guard(h->device);
  for (int s = 0; s < nsteps; ++s)
    if (int rc = enqueue_step(h)) return rc;
  return 0;
}

int ppk_mhd3d_synchronize(ppk_mhd3d *h) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  DeviceGuard guard(h->device);
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->comm_stream));
  if (h->h2d_stream) CUDA_TRY(cudaStreamSynchronize(h->h2d_stream));
  if (h->d2h_stream) CUDA_TRY(cudaStreamSynchronize(h->d2h_stream));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---- split-phase host transfers: batches pipelined on ONE handle ---------------------------------------------
// Three conservative arrays rotate: the one a step reads (current), the one it writes, and a staging array that the
// next batch is uploaded into while the step runs and the previous result is still being downloaded.
static int stage_init(ppk_mhd3d *h) {
  if (h->Ustage) return 0;
  if (int rc = alloc_doubles(h, &h->Ustage, NBVAR * h->g.ncell)) return rc;
  CUDA_TRY(cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_h2d, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming));
  CUDA_TRY(cudaStreamSynchronize(h->stream));  // (the zero fill of the new array)
  return 0;
}

int ppk_mhd3d_stage_upload(ppk_mhd3d *h, const double *u_host) {
  if (!h || !u_host) return fail(PPK_ERR_INVALID_ARGUMENT, "null argument");
  DeviceGuard guard(h->device);
  if (int rc = stage_init(h)) return rc;
  // Four arrays are busy at once when batches are pipelined (being uploaded | step input | step output | being downloaded):
  // if the staging array is still the source of a running download, upload into the spare one instead
  auto busy = [&](const double *buf) {
    for (auto &pd : h->d2h_pending)
      if (pd.buf == buf && cudaEventQuery(pd.done) == cudaErrorNotReady) return true;
    return false;
  };
  if (busy(h->Ustage)) {
    if (!h->Uspare) {
      if (int rc = alloc_doubles(h, &h->Uspare, NBVAR * h->g.ncell)) return rc;
      CUDA_TRY(cudaStreamSynchronize(h->stream));  // (its zero fill)
    }
    if (!busy(h->Uspare)) std::swap(h->Ustage, h->Uspare);
  }
  // whatever still reads the chosen array finishes first (and nothing else is waited for)
  for (auto &pd : h->d2h_pending)
    if (pd.buf == h->Ustage) CUDA_TRY(cudaStreamWaitEvent(h->h2d_stream, pd.done, 0));
  CUDA_TRY(cudaMemcpyAsync(h->Ustage, u_host, (size_t)NBVAR * h->g.ncell * sizeof(double), cudaMemcpyHostToDevice, h->h2d_stream));
  CUDA_TRY(cudaEventRecord(h->ev_h2d, h->h2d_stream));
  return 0;
}

int ppk_mhd3d_stage_swap(ppk_mhd3d *h) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  if (!h->Ustage) return fail(PPK_ERR_STATE, "ppk_mhd3d_stage_swap before ppk_mhd3d_stage_upload");
  DeviceGuard guard(h->device);
  if (int rc = settle_halo(h)) return rc;
  h->halo_posted = nullptr;
  CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_h2d, 0));  // the step reads the staged array after its copy
  std::swap(h->U[h->host_iteration & 1], h->Ustage);
  return 0;
}

int ppk_mhd3d_stage_download(ppk_mhd3d *h, double *u_host) {
  if (!h || !u_host) return fail(PPK_ERR_INVALID_ARGUMENT, "null argument");
  DeviceGuard guard(h->device);
  if (int rc = stage_init(h)) return rc;
  if (int rc = settle_halo(h)) return rc;
  double *src = h->cur();
  CUDA_TRY(cudaEventRecord(h->ev_main, h->stream));
  CUDA_TRY(cudaStreamWaitEvent(h->d2h_stream, h->ev_main, 0));
  CUDA_TRY(cudaMemcpyAsync(u_host, src, (size_t)NBVAR * h->g.ncell * sizeof(double), cudaMemcpyDeviceToHost, h->d2h_stream));
  ppk_mhd3d::Pending *slot = nullptr;
  for (auto &pd : h->d2h_pending)
    if (pd.buf == src) slot = &pd;
  if (!slot) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->d2h_pending.push_back({src, e});
    slot = &h->d2h_pending.back();
  }
  CUDA_TRY(cudaEventRecord(slot->done, h->d2h_stream));
  return 0;  // (whoever overwrites `src` later waits for slot->done: wait_staged_reads)
}

int ppk_mhd3d_diagnostics(ppk_mhd3d *h, double sums[8], double *max_divb) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  DeviceGuard guard(h->device);
  if (int rc = boundaries_and_primitives(h, h->cur(), false)) return rc;  // upper ghost faces feed div B
  CUDA_TRY(cudaMemsetAsync(h->diag, 0, 16 * sizeof(double), h->stream));
  { Scope sc(h, KK_DIAG, h->stream); h->kt->diagnostics(h->g, h->cur(), h->diag, h->stream); }
  double out[9];
  CUDA_TRY(cudaMemcpyAsync(out, h->diag, sizeof(out), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (sums) memcpy(sums, out, 8 * sizeof(double));
  if (max_divb) *max_divb = out[8];
  return 0;
}

// =================================================================================================
// 2-D path (MHD_Muscl_2D, implementationVersion 0): SolverMHDMuscl<2>::godunov_unsplit_impl
// (src/muscl/SolverMHDMuscl.cpp:373-417) as boundary x, boundary y, primitives + CFL, dt, trace, fluxes + EMF, update
// =================================================================================================
struct ppk_mhd2d {
  GridParams g{};
  const KernelTable *kt = nullptr;
  int device = 0;
  cudaStream_t stream = nullptr;
  double *U[2] = {nullptr, nullptr}, *Q = nullptr, *S = nullptr, *FX = nullptr, *FY = nullptr, *EMF = nullptr;
  StepState *st = nullptr;
  long launches = 0, host_iteration = 0;
  double *cur() { return U[host_iteration & 1]; }
  double *nxt() { return U[(host_iteration + 1) & 1]; }
};

int ppk_mhd2d_create(const ppk_mhd3d_params *p, ppk_mhd2d **out) {
  if (!p || !out) return fail(PPK_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (p->ghost_width != PPK_GHOST_WIDTH) return fail(PPK_ERR_UNSUPPORTED, "ghost_width must be 3 (MHD_Muscl_2D)");
  if (p->riemann_solver != PPK_RIEMANN_HLLD && p->riemann_solver != PPK_RIEMANN_HLL && p->riemann_solver != PPK_RIEMANN_LLF)
    return fail(PPK_ERR_UNSUPPORTED, "riemann must be hlld, hll or llf");
  // v0, v1 (atomic scatter) and v2 (slopes + updated primitives) of the reference are three formulations of the same 2-D
  // scheme: its own outputs differ by ~1e-15 between them (tests/golden2d_v2/README.md). All three run the v0 kernels.
  if (p->implementation_version < 0 || p->implementation_version > 2)
    return fail(PPK_ERR_UNSUPPORTED, "implementationVersion must be 0, 1 or 2");
  if (p->mx != 1 || p->my != 1) return fail(PPK_ERR_UNSUPPORTED, "the 2-D path is single-GPU (mx = my = 1)");
  if (p->nx < 3 || p->ny < 3) return fail(PPK_ERR_INVALID_ARGUMENT, "nx, ny must be >= 3 (ghost width)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(PPK_ERR_NO_DEVICE, "no CUDA device: ppkmhd_b200 has no CPU fallback");
  if (p->device < 0 || p->device >= ndev) return fail(PPK_ERR_INVALID_ARGUMENT, "device ordinal out of range");
  ppk_mhd2d *h = new ppk_mhd2d();
  if (int rc = create2d_impl(p, h)) {
    const std::string msg = g_last_error;
    ppk_mhd2d_destroy(h);
    g_last_error = msg;
    return rc;
  }
  *out = h;
  return 0;
}

} // extern "C"
namespace {
int create2d_impl(const ppk_mhd3d_params *p, ppk_mhd2d *h) {
  h->device = p->device;
  h->kt = p->exact_arithmetic ? kernel_table_exact() : kernel_table_fast();
  GridParams &g = h->g;
  g.nx = p->nx; g.ny = p->ny; g.nz = 1; g.gw = p->ghost_width;
  g.isize = p->nx + 2 * g.gw; g.jsize = p->ny + 2 * g.gw; g.ksize = 1;
  g.ncell = (long long)g.isize * g.jsize;
  g.dx = p->dx; g.dy = p->dy; g.dz = 1.0;
  g.idx = 1.0 / p->dx; g.idy = 1.0 / p->dy; g.idz = 1.0;
  g.gamma0 = p->gamma0; g.cfl = p->cfl; g.slope_type = p->slope_type;
  g.smallr = p->smallr; g.smallc = p->smallc; g.smallp = p->smallp;
  g.riemann = p->riemann_solver;
  g.wrap_x = 0;
  for (int f = 0; f < 6; ++f) g.bc[f] = p->boundary_type[f];
  DeviceGuard guard(h->device);
  CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  const size_t n = (size_t)g.ncell * sizeof(double);
  for (double **a : {&h->U[0], &h->U[1], &h->Q}) {
    CUDA_TRY(cudaMalloc((void **)a, NBVAR * n));
    CUDA_TRY(cudaMemsetAsync(*a, 0, NBVAR * n, h->stream));
  }
  CUDA_TRY(cudaMalloc((void **)&h->S, 8 * NBVAR * n));
  CUDA_TRY(cudaMemsetAsync(h->S, 0, 8 * NBVAR * n, h->stream));
  CUDA_TRY(cudaMalloc((void **)&h->FX, 6 * n));
  CUDA_TRY(cudaMalloc((void **)&h->FY, 6 * n));
  CUDA_TRY(cudaMalloc((void **)&h->EMF, n));
  CUDA_TRY(cudaMemsetAsync(h->FX, 0, 6 * n, h->stream));
  CUDA_TRY(cudaMemsetAsync(h->FY, 0, 6 * n, h->stream));
  CUDA_TRY(cudaMemsetAsync(h->EMF, 0, n, h->stream));
  CUDA_TRY(cudaMalloc((void **)&h->st, sizeof(StepState)));
  StepState st0{};
  st0.t_end = 1e300;
  CUDA_TRY(cudaMemcpyAsync(h->st, &st0, sizeof(st0), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}
}  // namespace
extern "C" {

int ppk_mhd2d_destroy(ppk_mhd2d *h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  for (double *p : {h->U[0], h->U[1], h->Q, h->S, h->FX, h->FY, h->EMF})
    if (p) cudaFree(p);
  if (h->st) cudaFree(h->st);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int ppk_mhd2d_upload(ppk_mhd2d *h, const double *u_host) {
  if (!h || !u_host) return fail(PPK_ERR_INVALID_ARGUMENT, "null argument");
  DeviceGuard guard(h->device);
  CUDA_TRY(cudaMemcpyAsync(h->cur(), u_host, (size_t)NBVAR * h->g.ncell * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

int ppk_mhd2d_download(ppk_mhd2d *h, double *u_host) {
  if (!h || !u_host) return fail(PPK_ERR_INVALID_ARGUMENT, "null argument");
  DeviceGuard guard(h->device);
  CUDA_TRY(cudaMemcpyAsync(u_host, h->cur(), (size_t)NBVAR * h->g.ncell * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

int ppk_mhd2d_set_time(ppk_mhd2d *h, double t, double t_end, long iteration) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  DeviceGuard guard(h->device);
  if (((iteration ^ h->host_iteration) & 1) != 0) std::swap(h->U[0], h->U[1]);
  StepState st{};
  st.t = t; st.t_end = t_end; st.iteration = iteration;
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaMemcpy(h->st, &st, sizeof(st), cudaMemcpyHostToDevice));
  h->host_iteration = iteration;
  return 0;
}

int ppk_mhd2d_get_time(ppk_mhd2d *h, double *t, double *dt, long *iteration) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  DeviceGuard guard(h->device);
  StepState st{};
  CUDA_TRY(cudaMemcpyAsync(&st, h->st, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (t) *t = st.t;
  if (dt) *dt = st.dt;
  if (iteration) *iteration = (long)st.iteration;
  return 0;
}

int ppk_mhd2d_make_boundaries(ppk_mhd2d *h) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  DeviceGuard guard(h->device);
  h->kt->boundary2d(h->g, h->cur(), 0, h->stream);
  h->kt->boundary2d(h->g, h->cur(), 1, h->stream);
  h->launches += 2;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int ppk_mhd2d_compute_dt(ppk_mhd2d *h, double *dt) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  DeviceGuard guard(h->device);
  h->kt->boundary2d(h->g, h->cur(), 0, h->stream);
  h->kt->boundary2d(h->g, h->cur(), 1, h->stream);
  h->kt->prim_dt2d(h->g, h->cur(), h->Q, h->st, h->stream);
  h->kt->finalize_dt(h->g, h->st, h->stream);
  h->launches += 4;
  CUDA_TRY(cudaGetLastError());
  return ppk_mhd2d_get_time(h, nullptr, dt, nullptr);
}

int ppk_mhd2d_step(ppk_mhd2d *h) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  DeviceGuard guard(h->device);
  const GridParams &g = h->g;
  cudaStream_t s = h->stream;
  double *Uin = h->cur(), *Uout = h->nxt();
  h->kt->boundary2d(g, Uin, 0, s);
  h->kt->boundary2d(g, Uin, 1, s);
  h->kt->prim_dt2d(g, Uin, h->Q, h->st, s);
  h->kt->finalize_dt(g, h->st, s);
  h->kt->trace2d(g, h->st, Uin, h->Q, h->S, s);
  h->kt->flux_emf2d(g, h->S, h->FX, h->FY, h->EMF, s);
  h->kt->update2d(g, h->st, Uin, Uout, h->FX, h->FY, h->EMF, s);
  h->kt->advance_time(h->st, s);
  h->launches += 8;
  h->host_iteration += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int ppk_mhd2d_run(ppk_mhd2d *h, int nsteps) {
  for (int s = 0; s < nsteps; ++s)
    if (int rc = ppk_mhd2d_step(h)) return rc;
  return 0;
}

int ppk_mhd2d_synchronize(ppk_mhd2d *h) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  DeviceGuard guard(h->device);
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

long ppk_mhd2d_launch_count(ppk_mhd2d *h) { return h ? h->launches : 0; }

int ppk_mhd3d_halo_plan(const ppk_mhd3d_params *p, int capacity, ppk_halo_msg *msgs) {
  if (!p || (capacity > 0 && !msgs)) return -1;
  if (p->mz <= 1) return 0;
  const long long gw = p->ghost_width;
  const long long isize = p->nx + 2 * gw, jsize = p->ny + 2 * gw, ksize = p->nz + 2 * gw;
  const long long plane = isize * jsize, ncell = plane * ksize, cnt = plane * gw;
  const bool lo_outer = p->rank_z == 0, hi_outer = p->rank_z == p->mz - 1;
  const bool exch_lo = !lo_outer || p->boundary_type[4] == PPK_BC_PERIODIC;
  const bool exch_hi = !hi_outer || p->boundary_type[5] == PPK_BC_PERIODIC;
  // ranks are laid out like MPI_Cart_create does for dims (mx, my, mz): z fastest (HydroParams.cpp:249-277)
  const int base = (p->rank_x * p->my + p->rank_y) * p->mz;
  const int zlo = base + (p->rank_z - 1 + p->mz) % p->mz, zhi = base + (p->rank_z + 1) % p->mz;
  int n = 0;
  auto add = [&](int peer, int is_send, int var, long long off) {
    if (n < capacity) msgs[n] = ppk_halo_msg{peer, is_send, var, var * ncell + off, cnt};
    ++n;
  };
  for (int v = 0; v < PPK_NBVAR; ++v) {
    // same order on every rank: (send down, recv from up, send up, recv from down) per variable
    if (exch_lo) add(zlo, 1, v, plane * gw);              // first interior planes k in [gw, 2gw) go down
    if (exch_hi) add(zhi, 0, v, plane * (p->nz + gw));    // upper ghost planes k in [nz+gw, nz+2gw)
    if (exch_hi) add(zhi, 1, v, plane * p->nz);           // last interior planes k in [nz, nz+gw) go up
    if (exch_lo) add(zlo, 0, v, 0);                       // lower ghost planes k in [0, gw)
  }
  return n;
}

int ppk_mhd3d_face_plan(const ppk_mhd3d_params *p, int dir, ppk_face_msg msgs[4]) {
  if (!p || !msgs || dir < 0 || dir > 1) return -1;
  const int m = dir == 0 ? p->mx : p->my, pos = dir == 0 ? p->rank_x : p->rank_y;
  if (m <= 1) return 0;
  const long long gw = p->ghost_width;
  const long long isize = p->nx + 2 * gw, jsize = p->ny + 2 * gw, ksize = p->nz + 2 * gw;
  const long long count = gw * (dir == 0 ? jsize : isize) * ksize * PPK_NBVAR;
  const bool exch_lo = pos != 0 || p->boundary_type[2 * dir] == PPK_BC_PERIODIC;
  const bool exch_hi = pos != m - 1 || p->boundary_type[2 * dir + 1] == PPK_BC_PERIODIC;
  auto rank_of = [&](int q) {  // global rank of the block at position q along dir
    const int cx = dir == 0 ? q : p->rank_x, cy = dir == 1 ? q : p->rank_y;
    return (cx * p->my + cy) * p->mz + p->rank_z;
  };
  const int lo = rank_of((pos - 1 + m) % m), hi = rank_of((pos + 1) % m);
  const int n_int = dir == 0 ? p->nx : p->ny;
  int n = 0;
  // same order on every rank (like the z plan): send down, receive from up, send up, receive from down
  if (exch_lo) msgs[n++] = ppk_face_msg{lo, 1, 0, (int)gw, count};            // first interior layers [gw, 2gw) go down
  if (exch_hi) msgs[n++] = ppk_face_msg{hi, 0, 1, (int)(n_int + gw), count};  // upper ghost layers [n+gw, n+2gw)
  if (exch_hi) msgs[n++] = ppk_face_msg{hi, 1, 1, n_int, count};              // last interior layers [n, n+gw) go up
  if (exch_lo) msgs[n++] = ppk_face_msg{lo, 0, 0, 0, count};                  // lower ghost layers [0, gw)
  return n;
}

int ppk_selftest_fastmath(int n, const double *x_host, double *rcp_out, double *sqrt_out, double *rsqrt_out) {
  if (n <= 0 || !x_host || !rcp_out || !sqrt_out || !rsqrt_out) return fail(PPK_ERR_INVALID_ARGUMENT, "bad argument");
  double *d = nullptr;
  CUDA_TRY(cudaMalloc((void **)&d, (size_t)4 * n * sizeof(double)));
  CUDA_TRY(cudaMemcpy(d, x_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  kernel_table_fast()->fastmath_selftest(n, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n, nullptr);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(rcp_out, d + n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(sqrt_out, d + 2 * (size_t)n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(rsqrt_out, d + 3 * (size_t)n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaFree(d));
  return 0;
}

int ppk_nccl_get_unique_id(void *out) {
  if (!out) return fail(PPK_ERR_INVALID_ARGUMENT, "null argument");
  if (int rc = load_nccl()) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  memcpy(out, &id, sizeof(id));
  return 0;
}

int ppk_mhd3d_comm_init(ppk_mhd3d *h, const void *unique_id, int nranks, int rank) {
  if (!h || !unique_id) return fail(PPK_ERR_INVALID_ARGUMENT, "null argument");
  const ppk_mhd3d_params &q = h->params;
  if (nranks != q.mx * q.my * q.mz || rank != (q.rank_x * q.my + q.rank_y) * q.mz + q.rank_z)
    return fail(PPK_ERR_INVALID_ARGUMENT, "communicator shape must match mx*my*mz and rank = (rank_x*my + rank_y)*mz + rank_z");
  if (int rc = load_nccl()) return rc;
  DeviceGuard guard(h->device);
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  NCCL_TRY(g_nccl.CommInitRank(&h->comm, nranks, id, rank));
  h->nranks = nranks;
  h->rank = rank;
  return 0;
}

int ppk_mhd3d_set_stream(ppk_mhd3d *h, void *cuda_stream) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  DeviceGuard guard(h->device);
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return 0;
}

int ppk_mhd3d_set_pipeline(ppk_mhd3d *h, int pipeline) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  if (pipeline < PPK_PIPELINE_UNFUSED || pipeline > PPK_PIPELINE_ORDERED) return fail(PPK_ERR_INVALID_ARGUMENT, "unknown pipeline");
  if (pipeline == PPK_PIPELINE_ORDERED && !h->tma) pipeline = PPK_PIPELINE_UNFUSED;  // (no TMA tiles on this grid: the same kernels, one by one)
  if (pipeline == PPK_PIPELINE_TILED && (!h->tma || !h->prod || h->params.mz != 1))
    return fail(PPK_ERR_UNSUPPORTED, "the tiled pipeline needs an even nx >= 32 (16-byte TMA rows) and a single slab (mz = 1)");
  DeviceGuard guard(h->device);
  if ((pipeline == PPK_PIPELINE_UNFUSED || pipeline == PPK_PIPELINE_TILED || pipeline == PPK_PIPELINE_ORDERED) && !h->F[0]) {  // flux / EMF arrays exist only for the unfused pipeline
    int rc = 0;
    const long long n = h->g.ncell;
    if ((rc = alloc_doubles(h, &h->F[0], NFLUX * n)) || (rc = alloc_doubles(h, &h->F[1], NFLUX * n)) ||
        (rc = alloc_doubles(h, &h->F[2], NFLUX * n)) || (!h->EMF && (rc = alloc_doubles(h, &h->EMF, NEMF * n))))
      return rc;
  }
  if (pipeline == PPK_PIPELINE_STREAMED && !h->EMF) {  // the streamed pipeline stores the EMFs but no flux
    if (int rc = alloc_doubles(h, &h->EMF, NEMF * h->g.ncell)) return rc;
  }
  h->pipeline = pipeline;
  return 0;
}

int ppk_mhd3d_get_pipeline(ppk_mhd3d *h) { return h ? h->pipeline : -1; }

int ppk_mhd3d_profile(ppk_mhd3d *h, int enable) {
  if (!h) return fail(PPK_ERR_INVALID_ARGUMENT, "null handle");
  DeviceGuard guard(h->device);
  if (int rc = collect_timings(h)) return rc;
  h->profile = enable != 0;
  if (h->profile) {  // origin of the timeline (ppk_mhd3d_kernel_timeline)
    if (!h->t0) CUDA_TRY(cudaEventCreate(&h->t0));
    h->timeline.clear();
    CUDA_TRY(cudaEventRecord(h->t0, h->stream));
  }
  return 0;
}

int ppk_mhd3d_kernel_timeline(ppk_mhd3d *h, int capacity, const char **names, int *on_comm_stream, double *start_ms, double *end_ms) {
  if (!h) return -1;
  DeviceGuard guard(h->device);
  if (collect_timings(h)) return -1;
  const int n = (int)h->timeline.size() < capacity ? (int)h->timeline.size() : capacity;
  for (int i = 0; i < n; ++i) {
    if (names) names[i] = kKernelNames[h->timeline[i].kind];
    if (on_comm_stream) on_comm_stream[i] = h->timeline[i].on_comm;
    if (start_ms) start_ms[i] = h->timeline[i].start_ms;
    if (end_ms) end_ms[i] = h->timeline[i].end_ms;
  }
  return n;
}

int ppk_mhd3d_kernel_times(ppk_mhd3d *h, int capacity, const char **names, double *ms, long *launches, int reset) {
  if (!h) return -1;
  DeviceGuard guard(h->device);
  if (collect_timings(h)) return -1;
  const int n = capacity < KK_COUNT ? capacity : KK_COUNT;
  for (int i = 0; i < n; ++i) {
    if (names) names[i] = kKernelNames[i];
    if (ms) ms[i] = h->acc_ms[i];
    if (launches) launches[i] = h->acc_n[i];
  }
  if (reset)
    for (int i = 0; i < KK_COUNT; ++i) { h->acc_ms[i] = 0; h->acc_n[i] = 0; }
  return n;
}

long ppk_mhd3d_launch_count(ppk_mhd3d *h) { return h ? h->launches : 0; }
long long ppk_mhd3d_device_bytes(ppk_mhd3d *h) { return h ? h->bytes : 0; }

int ppk_mhd3d_debug_array(ppk_mhd3d *h, const char *name, double *host_out, int *ncomp) {
  if (!h || !name) return fail(PPK_ERR_INVALID_ARGUMENT, "null argument");
  DeviceGuard guard(h->device);
  const std::string s(name);
  const double *src = nullptr;
  int nc = 0;
  if (s == "U") { src = h->cur(); nc = NBVAR; }
  else if (s == "U2") { src = h->nxt(); nc = NBVAR; }
  else if (s == "Q") { src = h->Q; nc = NBVAR; }
  else if (s == "ElecField") { src = h->E; nc = NELEC; }
  else if (s == "dbf") { src = h->DBF; nc = NDBF; }
  else if (s == "basis") { src = h->BASIS; nc = NBASIS; }
  else if (s == "Fluxes_x") { src = h->F[0]; nc = NFLUX; }
  else if (s == "Fluxes_y") { src = h->F[1]; nc = NFLUX; }
  else if (s == "Fluxes_z") { src = h->F[2]; nc = NFLUX; }
  else if (s == "Emf") { src = h->EMF; nc = NEMF; }
  else return fail(PPK_ERR_INVALID_ARGUMENT, "unknown array name " + s);
  if (!src) return fail(PPK_ERR_STATE, s + " has not been produced by the pipeline of this handle (ppk_mhd3d_set_pipeline)");
  if (h->pipeline == PPK_PIPELINE_TILED && (s == "Q" || s == "ElecField"))
    return fail(PPK_ERR_STATE, s + " never reaches device memory in the tiled pipeline (ppk_mhd3d_set_pipeline)");
  // x periodic: the launchers skip the column i = nx+gw of the x-fluxes and of the z- / y-EMFs (the update reads the
  // bit-identical column i = gw); complete the arrays for the caller
  if (s == "Fluxes_x") h->kt->wrap_x_column(h->g, h->F[0], NFLUX, h->stream);
  if (s == "Emf") h->kt->wrap_x_column(h->g, h->EMF, 2, h->stream);
  if (ncomp) *ncomp = nc;
  if (host_out) {
    CUDA_TRY(cudaMemcpyAsync(host_out, src, (size_t)nc * h->g.ncell * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

}  // extern "C"

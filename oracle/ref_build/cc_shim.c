/* LD_PRELOAD shim for running the reference's Kokkos-CUDA build on Blackwell. The reference vendors Kokkos 4.3, whose
 * occupancy helper (core/src/Cuda/Kokkos_Cuda_BlockSize_Deduction.hpp:28-43) throws on compute capability 10.x. The
 * reference sources stay untouched: this shim makes cudaGetDeviceProperties report 9.0 for a 10.x device, so Kokkos takes
 * its Hopper code path (same warp allocation granularity; kernels are JIT-compiled from the embedded compute_90 PTX).
 * Test infrastructure only (bench.py --with-ref-cuda, profiles/run_ref_cuda.sh). */
#define _GNU_SOURCE
#include <cuda_runtime.h>
#include <dlfcn.h>

typedef cudaError_t (*props_fn)(struct cudaDeviceProp *, int);

static cudaError_t forward(const char *name, struct cudaDeviceProp *p, int dev) {
  props_fn real = (props_fn)dlsym(RTLD_NEXT, name);
  if (!real) return cudaErrorUnknown;
  cudaError_t e = real(p, dev);
  if (e == cudaSuccess && p->major >= 10) { p->major = 9; p->minor = 0; }
  return e;
}
#undef cudaGetDeviceProperties
cudaError_t cudaGetDeviceProperties_v2(struct cudaDeviceProp *p, int dev) { return forward("cudaGetDeviceProperties_v2", p, dev); }
cudaError_t cudaGetDeviceProperties(struct cudaDeviceProp *p, int dev) { return forward("cudaGetDeviceProperties", p, dev); }

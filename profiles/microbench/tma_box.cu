// Standalone check of 4-D fp64 TMA box loads (cp.async.bulk.tensor.4d) for the box shapes the MHD kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_box tma_box.cu && ./tma_box
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int XB, int YB, int ZB, int NS>
__global__ void k_box(const __grid_constant__ CUtensorMap map, int x0, int y0, int z0, double *out) {
  extern __shared__ __align__(128) double sm[];
  __shared__ unsigned long long bar;
  constexpr int BOX = XB * YB * ZB;
  constexpr int SLOT = (BOX * 8 + 127) / 128 * 16;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(NS * BOX * 8) : "memory");
    for (int s = 0; s < NS; ++s)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                     smem_u32(sm + s * SLOT)),
                   "l"(&map), "r"(x0), "r"(y0), "r"(z0), "r"(2 * s + 1), "r"(smem_u32(&bar))
                   : "memory");
  }
  asm volatile(
    "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(
      smem_u32(&bar)),
    "r"(0)
    : "memory");
  for (int s = 0; s < NS; ++s)
    for (int e = threadIdx.x; e < BOX; e += blockDim.x) out[s * BOX + e] = sm[s * SLOT + e];
}

template <int XB, int YB, int ZB, int NS>
__global__ void k_box_g(const CUtensorMap *map, int x0, int y0, int z0, double *out) {
  extern __shared__ __align__(128) double sm[];
  __shared__ unsigned long long bar;
  constexpr int BOX = XB * YB * ZB;
  constexpr int SLOT = (BOX * 8 + 127) / 128 * 16;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(NS * BOX * 8) : "memory");
    for (int s = 0; s < NS; ++s)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                     smem_u32(sm + s * SLOT)),
                   "l"(map), "r"(x0), "r"(y0), "r"(z0), "r"(2 * s + 1), "r"(smem_u32(&bar))
                   : "memory");
  }
  asm volatile(
    "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(
      smem_u32(&bar)),
    "r"(0)
    : "memory");
  for (int s = 0; s < NS; ++s)
    for (int e = threadIdx.x; e < BOX; e += blockDim.x) out[s * BOX + e] = sm[s * SLOT + e];
}

__global__ void k_bulk1d(const double *src, double *out) {
  __shared__ __align__(128) double sm[256];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(2048) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm)), "l"(src),
                 "r"(2048), "r"(smem_u32(&bar))
                 : "memory");
  }
  asm volatile(
    "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(
      smem_u32(&bar)),
    "r"(0)
    : "memory");
  for (int e = threadIdx.x; e < 256; e += blockDim.x) out[e] = sm[e];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int g_mode = 0;
template <int XB, int YB, int ZB>
int run(EncodeTiledFn enc, double *d, int is, int js, int ks, int nc, const std::vector<double> &h) {
  constexpr int NS = 3, BOX = XB * YB * ZB, SLOT = (BOX * 8 + 127) / 128 * 16;
  CUtensorMap m;
  const int f = g_mode == 1 ? 2 : 1;  // mode 1: view the doubles as pairs of 32-bit words
  const cuuint64_t dims[4] = {(cuuint64_t)is * f, (cuuint64_t)js, (cuuint64_t)ks, (cuuint64_t)nc};
  const cuuint64_t strides[3] = {(cuuint64_t)is * 8, (cuuint64_t)is * js * 8, (cuuint64_t)is * js * ks * 8};
  const cuuint32_t box[4] = {(cuuint32_t)XB * f, YB, ZB, 1}, estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&m, g_mode == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, g_mode == 2 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("box %dx%dx%d: encode failed %d\n", XB, YB, ZB, (int)r); return 1; }
  double *out;
  cudaMalloc(&out, NS * BOX * 8);
  cudaFuncSetAttribute(k_box<XB, YB, ZB, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, NS * SLOT * 8);
  int bad = 0;
  for (int trial = 0; trial < 3; ++trial) {
    const int x0 = 2 + 6 * trial /* even: a box must start 16-byte aligned */, y0 = 2 + 3 * trial, z0 = 3 + trial;
    if (g_mode == 3) {
      CUtensorMap *dm;
      cudaMalloc(&dm, sizeof(CUtensorMap));
      cudaMemcpy(dm, &m, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
      cudaFuncSetAttribute(k_box_g<XB, YB, ZB, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, NS * SLOT * 8);
      k_box_g<XB, YB, ZB, NS><<<1, 128, NS * SLOT * 8>>>(dm, x0 * f, y0, z0, out);
    } else
    k_box<XB, YB, ZB, NS><<<1, 128, NS * SLOT * 8>>>(m, x0 * f, y0, z0, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("box %dx%dx%d: kernel error %s\n", XB, YB, ZB, cudaGetErrorString(e)); return 1; }
    std::vector<double> o(NS * BOX);
    cudaMemcpy(o.data(), out, NS * BOX * 8, cudaMemcpyDeviceToHost);
    for (int s = 0; s < NS; ++s)
      for (int z = 0; z < ZB; ++z)
        for (int y = 0; y < YB; ++y)
          for (int x = 0; x < XB; ++x) {
            const int gx = x0 + x, gy = y0 + y, gz = z0 + z, c = 2 * s + 1;
            const double want = (gx < is && gy < js && gz < ks) ? h[(size_t)gx + (size_t)is * (gy + (size_t)js * (gz + (size_t)ks * c))] : 0.0;
            if (o[s * BOX + x + XB * (y + YB * z)] != want) ++bad;
          }
  }
  printf("box %dx%dx%d: %s (%d mismatches)\n", XB, YB, ZB, bad ? "WRONG" : "ok", bad);
  cudaFree(out);
  return bad != 0;
}

int main(int argc, char **argv) {
  const int shape = argc > 1 ? atoi(argv[1]) : -1;
  g_mode = argc > 2 ? atoi(argv[2]) : 0;
  const int is = 70, js = 46, ks = 42, nc = 8;
  std::vector<double> h((size_t)is * js * ks * nc);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i * 0.5 + 1.0;
  double *d;
  cudaMalloc(&d, h.size() * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  int rc = 0;
  if (shape == 99) {
    double *o; cudaMalloc(&o, 2048);
    k_bulk1d<<<1, 128>>>(d, o);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<double> ho(256);
    cudaMemcpy(ho.data(), o, 2048, cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < 256; ++i) bad += ho[i] != h[i];
    printf("1-D bulk copy: %s, %d mismatches\n", cudaGetErrorString(e), bad);
    return 0;
  }
  if (shape < 0 || shape == 0) rc |= run<34, 8, 1>(enc, d, is, js, ks, nc, h);
  if (shape < 0 || shape == 1) rc |= run<32, 9, 1>(enc, d, is, js, ks, nc, h);
  if (shape < 0 || shape == 2) rc |= run<32, 4, 3>(enc, d, is, js, ks, nc, h);
  if (shape < 0 || shape == 3) rc |= run<34, 9, 1>(enc, d, is, js, ks, nc, h);
  if (shape < 0 || shape == 4) rc |= run<32, 5, 3>(enc, d, is, js, ks, nc, h);
  if (shape < 0 || shape == 5) rc |= run<34, 4, 3>(enc, d, is, js, ks, nc, h);
  if (shape == 6) rc |= run<32, 8, 1>(enc, d, is, js, ks, nc, h);
  if (shape == 7) rc |= run<16, 8, 1>(enc, d, is, js, ks, nc, h);
  return rc;
}

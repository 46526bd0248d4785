// minimal 2-D TMA probe: which (dtype, rank) combinations does this driver/GPU accept?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap map, unsigned bytes, float *out, int n) {
  extern __shared__ __align__(128) float sm[];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (RANK == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(sm)),
                   "l"(&map), "r"(0), "r"(0), "r"(smem_u32(&bar)) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(sm)),
                   "l"(&map), "r"(0), "r"(0), "r"(0), "r"(smem_u32(&bar)) : "memory");
    if (RANK == 4)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(sm)),
                   "l"(&map), "r"(0), "r"(0), "r"(0), "r"(0), "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int e = threadIdx.x; e < n; e += blockDim.x) out[e] = sm[e];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
  const int rank = atoi(argv[1]), f64 = atoi(argv[2]);
  const int es = f64 ? 8 : 4;
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const cuuint64_t dims[4] = {256, 64, 8, 4};
  const cuuint64_t strides[3] = {256ull * es, 256ull * 64 * es, 256ull * 64 * 8 * es};
  const cuuint32_t box[4] = {32, 8, 1, 1}, estr[4] = {1, 1, 1, 1};
  void *d; cudaMalloc(&d, 256 * 64 * 8 * 4 * es); cudaMemset(d, 0, 256 * 64 * 8 * 4 * es);
  CUtensorMap m;
  CUresult r = enc(&m, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("rank %d f64 %d: encode %d, query %d; ", rank, f64, (int)r, (int)q);
  float *out; cudaMalloc(&out, 32 * 8 * es);
  const unsigned bytes = 32 * 8 * es;
  if (rank == 2) k<2><<<1, 64, bytes>>>(m, bytes, out, 32 * 8 * es / 4);
  if (rank == 3) k<3><<<1, 64, bytes>>>(m, bytes, out, 32 * 8 * es / 4);
  if (rank == 4) k<4><<<1, 64, bytes>>>(m, bytes, out, 32 * 8 * es / 4);
  printf("kernel: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

// Experiment: semantics of cp.async.bulk.tensor.2d tile::gather4 with SWIZZLE_128B, for box rows = 1 and 4.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap map, const int* rows, float* out, int col0) {
  __shared__ __align__(1024) float buf[8 * 32];   // 8 rows x 128 B
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) buf[i] = -1.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(1024u) : "memory");
    for (int g = 0; g < 2; ++g)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
                   " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                   ::"r"(s32(buf + g * 128)), "l"(&map), "r"(s32(&bar)), "r"(col0), "r"(rows[4 * g]), "r"(rows[4 * g + 1]),
                     "r"(rows[4 * g + 2]), "r"(rows[4 * g + 3]) : "memory");
  }
  unsigned ok = 0; int spins = 0;
  while (!ok && spins++ < 1000000)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(s32(&bar)), "r"(0u) : "memory");
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = buf[i];
  if (threadIdx.x == 0) out[256] = ok ? 1.f : 0.f;
}
int main() {
  const int R = 64, D = 64;
  std::vector<float> h(R * D);
  for (int r = 0; r < R; ++r) for (int c = 0; c < D; ++c) h[r * D + c] = r * 100 + c;
  float* d; cudaMalloc(&d, R * D * 4); cudaMemcpy(d, h.data(), R * D * 4, cudaMemcpyHostToDevice);
  int hr[8] = {5, 17, 3, 40, 41, 9, 63, 0}; int* dr; cudaMalloc(&dr, 32); cudaMemcpy(dr, hr, 32, cudaMemcpyHostToDevice);
  float* out; cudaMalloc(&out, 257 * 4);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)p;
  for (int boxrows : {1, 4}) {
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)R}; cuuint64_t gstr[1] = {(cuuint64_t)D * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)boxrows}; cuuint32_t es[2] = {1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("boxrows=%d encode=%d\n", boxrows, (int)r);
    if (r != CUDA_SUCCESS) continue;
    cudaMemset(out, 0, 257 * 4);
    k<<<1, 128>>>(map, dr, out, 32);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> o(257); cudaMemcpy(o.data(), out, 257 * 4, cudaMemcpyDeviceToHost);
    printf("  barrier completed=%g\n", o[256]);
    for (int row = 0; row < 8; ++row) {
      printf("  smem row %d:", row);
      for (int ch = 0; ch < 8; ++ch) printf(" %6.0f", o[row * 32 + ch * 4]);   // first element of each 16-B chunk
      printf("\n");
    }
  }
  return 0;
}

// Experiment: minimal cta_group::2 tcgen05 pipeline (2SM TMA -> leader MMA M=256 N=256 K=32 tf32 -> multicast commit
// -> both CTAs read their TMEM half).  Prints max error against a host reference.  Stages can be disabled: argv[1].
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>
#include <cmath>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e_), #x); return 1; } } while (0)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ bool wait_bounded(uint64_t* bar, unsigned parity) {
  for (int i = 0; i < 2000000; ++i) if (try_wait(bar, parity)) return true;
  return false;
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
k(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, float* out, int* status, int mode) {
  extern __shared__ unsigned char raw[];
  unsigned char* sm = raw + ((1024u - (s32(raw) & 1023u)) & 1023u);
  unsigned char* sa = sm;                 // A: 128 rows x 128 B
  unsigned char* sb = sm + 16384;         // B half: 128 rows x 128 B
  uint64_t* full = (uint64_t*)(sm + 32768);
  uint64_t* tfull = full + 1;
  uint32_t* slot = (uint32_t*)(tfull + 1);
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    status[blockIdx.x * 8 + 0] = (int)s32(full);      // address as seen by this CTA
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(full)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(tfull)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) status[blockIdx.x * 8 + 1] = (int)tmem;
  if (warp == 0 && lane == 0 && mode >= 1) {
    if (rank == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(full)), "r"(65536u) : "memory");
    const uint32_t bar = s32(full) & 0xFEFFFFFFu;
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s32(sa)), "l"(&ta), "r"(bar), "r"(0), "r"((int)rank * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s32(sb)), "l"(&tb), "r"(bar), "r"(0), "r"((int)rank * 128) : "memory");
  }
  if (warp == 1 && lane == 0 && rank == 0 && mode >= 2) {
    const bool ok = wait_bounded(full, 0);
    status[blockIdx.x * 8 + 2] = ok ? 1 : -1;
    if (ok && mode >= 3) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
      for (int j = 0; j < 4; ++j) {
        const uint64_t da = (uint64_t)(((s32(sa) & 0x3FFFFu) >> 4) + 2 * j) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
        const uint64_t db = (uint64_t)(((s32(sb) & 0x3FFFFu) >> 4) + 2 * j) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(j ? 1u : 0u) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(s32(tfull)), "h"((unsigned short)3) : "memory");
    }
  }
  if (warp >= 2 && mode >= 3) {
    const bool ok = wait_bounded(tfull, 0);
    if (lane == 0) status[blockIdx.x * 8 + 3 + (warp - 2)] = ok ? 1 : -1;
    if (ok) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int q = warp & 3;
      for (int c0 = 0; c0 < 256; c0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tmem + ((uint32_t)(q * 32) << 16) + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) out[((size_t)(rank * 128 + q * 32 + lane)) * 256 + c0 + i] = __uint_as_float(r[i]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}
int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 3;
  const int M = 256, N = 256, K = 32;
  std::vector<float> A(M * K), B(N * K);
  for (int i = 0; i < M * K; ++i) A[i] = (float)((i * 7) % 13 - 6) * 0.25f;
  for (int i = 0; i < N * K; ++i) B[i] = (float)((i * 5) % 11 - 5) * 0.5f;
  float *dA, *dB, *dO; int* dS;
  CK(cudaMalloc(&dA, M * K * 4)); CK(cudaMalloc(&dB, N * K * 4)); CK(cudaMalloc(&dO, M * N * 4)); CK(cudaMalloc(&dS, 64));
  CK(cudaMemcpy(dA, A.data(), M * K * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), N * K * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0, M * N * 4)); CK(cudaMemset(dS, 0, 64));
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn fn = (EncodeTiledFn)p;
  CUtensorMap ta, tb;
  cuuint64_t gd[2] = {(cuuint64_t)K, (cuuint64_t)M}; cuuint64_t gs[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {32, 128}; cuuint32_t es[2] = {1, 1};
  if (fn(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode a failed\n"); return 1; }
  if (fn(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode b failed\n"); return 1; }
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000));
  k<<<2, 192, 40000>>>(ta, tb, dO, dS, mode);
  cudaError_t e = cudaDeviceSynchronize();
  printf("mode %d kernel: %s\n", mode, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  int st[16]; CK(cudaMemcpy(st, dS, 64, cudaMemcpyDeviceToHost));
  for (int c = 0; c < 2; ++c)
    printf("cta %d: full addr 0x%x tmem 0x%x fullwait %d epi %d %d %d %d\n", c, st[c * 8], st[c * 8 + 1], st[c * 8 + 2], st[c * 8 + 3],
           st[c * 8 + 4], st[c * 8 + 5], st[c * 8 + 6]);
  if (mode >= 3) {
    std::vector<float> O(M * N); CK(cudaMemcpy(O.data(), dO, M * N * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
      double ref = 0; for (int kk = 0; kk < K; ++kk) ref += (double)A[m * K + kk] * B[n * K + kk];
      maxerr = fmax(maxerr, fabs(ref - O[m * N + n]));
    }
    printf("max abs error vs host: %g (O[0]=%g O[last]=%g)\n", maxerr, O[0], O[M * N - 1]);
  }
  return 0;
}

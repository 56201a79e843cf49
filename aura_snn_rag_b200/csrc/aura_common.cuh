// Shared device/host helpers for libaura_hippo (sm_100a only).
//
// Nothing in here mirrors reference code: the reference (src/core/hippocampal.py) is
// PyTorch-eager; these are the building blocks of the hand-written replacement kernels:
//   * 64-bit ranking keys (score, then lower row first) used by every top-k in the library
//   * a warp-distributed sorted top-k list with warp-uniform insertion
//   * a block-wide bitonic sort over shared memory for the CTA / cross-CTA merges
//   * mbarrier + cp.async.bulk (TMA 1-D bulk copy) PTX wrappers for the HBM streaming pipeline
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/aura_hippo.h"

namespace aura {

typedef unsigned long long u64;
static constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------ error plumbing (host)
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);
int  sm_count();
int  max_smem_optin();
void note_launches(int n);  // kernels launched by this library in this process (aura_kernel_launches)
unsigned* trap_trace_device();   // mapped pinned words a watchdog writes before it traps (aura_debug_last_trap); may be null
int  env_int(const char* name, int dflt);  // tuning knob from the environment (callers cache it in a function-local static)

// scan_topk.cu: shared launcher of aura_scan_topk (probes == nullptr) and aura_ivf_search
size_t scan_workspace_bytes(int n_queries, int k);
int launch_scan(const void* rows, int dtype, long long n_rows, int d, const float* queries, int n_queries,
                const float* scale, const float* bias, int k, long long row_base, long long* out_idx, float* out_score,
                void* workspace, const long long* probes, int nprobe, int n_lists, const int* list_offsets,
                const int* list_rows, long long expected_rows, cudaStream_t st, int empty_ok);

// kmeans.cu: out[i] = ||x_i||^2 for n fp32 rows
void launch_row_sq_norms(const float* x, int n, int d, float* out, cudaStream_t st);

#define AURA_CUDA_OK(expr)                                                        \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) return ::aura::cuda_fail(_e, #expr);                   \
  } while (0)

#define AURA_REQUIRE(cond, code, ...)                                             \
  do {                                                                            \
    if (!(cond)) { ::aura::set_error(__VA_ARGS__); return (code); }               \
  } while (0)

// ------------------------------------------------------------------ ranking keys
// key = orderable(score) << 32 | (0xFFFFFFFF - row).  Larger key ranks first: higher score,
// and among equal scores the LOWER row (the library's stated tie rule; torch.topk leaves
// ties unspecified, hippocampal.py:307).  key == 0 is "empty slot".
__host__ __device__ __forceinline__ unsigned f32_orderable(float f) {
#ifdef __CUDA_ARCH__
  unsigned u = __float_as_uint(f);
#else
  union { float f; unsigned u; } c; c.f = f; unsigned u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float f32_from_orderable(unsigned o) {
  unsigned u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; unsigned u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ u64 make_key(float score, unsigned row) {
  return ((u64)f32_orderable(score) << 32) | (u64)(0xFFFFFFFFu - row);
}
__host__ __device__ __forceinline__ unsigned key_row(u64 key) { return 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull); }
__host__ __device__ __forceinline__ float key_score(u64 key) { return f32_from_orderable((unsigned)(key >> 32)); }

#ifdef __CUDACC__
// ------------------------------------------------------------------ warp-distributed top-k
// Capacity 32*KPL keys, sorted descending, blocked across lanes: position p = lane*KPL + j.
// insert() must be called by all 32 lanes with a warp-uniform key (key > thr).
template <int KPL>
struct WarpTopK {
  u64 e[KPL];
  u64 thr;  // key at the last position (warp-uniform): anything <= thr cannot enter

  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < KPL; ++j) e[j] = 0ull;
    thr = 0ull;
  }
  __device__ __forceinline__ void insert(u64 key, int lane) {
    int pos = 0;  // number of stored keys ranking before `key`
#pragma unroll
    for (int j = 0; j < KPL; ++j) pos += __popc(__ballot_sync(FULL, e[j] > key));
    const u64 carry = __shfl_up_sync(FULL, e[KPL - 1], 1);
    const int base = lane * KPL;
#pragma unroll
    for (int j = KPL - 1; j >= 0; --j) {
      const int p = base + j;
      const u64 prev = (j == 0) ? carry : e[j > 0 ? j - 1 : 0];
      if (p > pos) e[j] = prev;
      else if (p == pos) e[j] = key;
    }
    thr = __shfl_sync(FULL, e[KPL - 1], 31);
  }
  // key at sorted position p (warp-uniform p)
  __device__ __forceinline__ u64 at(int p) const {
    u64 v = 0ull;
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      const u64 t = __shfl_sync(FULL, e[j], p / KPL);
      if (j == p % KPL) v = t;
    }
    return v;
  }
  // dump all 32*KPL keys (sorted) to dst[0 .. 32*KPL)
  __device__ __forceinline__ void store(u64* dst, int lane) const {
#pragma unroll
    for (int j = 0; j < KPL; ++j) dst[lane * KPL + j] = e[j];
  }
};

// ------------------------------------------------------------------ block bitonic sort (descending)
// n2 must be a power of two; all threads of the block participate.
__device__ __forceinline__ void block_bitonic_sort_desc(u64* keys, int n2) {
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const u64 a = keys[lo], b = keys[hi];
        if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// ---- shared tail: CTA merge + cross-CTA final merge ------------------------------------------
template <int KPL>
__device__ __forceinline__ void publish_cta_topk(const WarpTopK<KPL>& tk, int warp, int lane, int n_warps,
                                                 u64* merge, int k, u64* dst) {
  constexpr int KC = 32 * KPL;
  const int n2 = next_pow2(n_warps * KC);
  __syncthreads();
  if (warp < n_warps) tk.store(merge + warp * KC, lane);
  for (int i = n_warps * KC + threadIdx.x; i < n2; i += blockDim.x) merge[i] = 0ull;
  block_bitonic_sort_desc(merge, n2);
  for (int i = threadIdx.x; i < k; i += blockDim.x) dst[i] = merge[i];
  __syncthreads();
}

// Merge n_lists sorted-or-not lists of k keys (list l at src + l*list_stride) into out (last CTA).
__device__ __forceinline__ void final_merge_write(const u64* src, int n_lists, size_t list_stride, int k, u64* merge,
                                                  int merge_cap, long long row_base, long long* out_idx,
                                                  float* out_score) {
  int have = 0, list = 0;
  while (list < n_lists) {
    const int lists_fit = max(1, (merge_cap - have) / k);
    const int take = min(lists_fit, n_lists - list);
    __syncthreads();
    for (int i = threadIdx.x; i < take * k; i += blockDim.x)
      merge[have + i] = src[(size_t)(list + i / k) * list_stride + (i % k)];
    const int n2 = next_pow2(max(have + take * k, 2));
    for (int i = have + take * k + threadIdx.x; i < n2; i += blockDim.x) merge[i] = 0ull;
    block_bitonic_sort_desc(merge, n2);
    list += take;
    have = min(k, n2);
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const u64 key = merge[i];
    out_idx[i] = key ? row_base + (long long)key_row(key) : -1ll;
    out_score[i] = key ? key_score(key) : -INFINITY;
  }
  __syncthreads();
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// ------------------------------------------------------------------ mbarrier / bulk-copy PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar` (TMA engine, SASS UBLKCP).
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

__device__ __forceinline__ float bf16_lo(unsigned packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi(unsigned packed) { return __uint_as_float(packed & 0xffff0000u); }
#endif  // __CUDACC__

}  // namespace aura

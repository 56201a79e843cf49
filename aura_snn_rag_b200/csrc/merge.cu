// aura_topk_merge - k-way merge of per-shard top-k blocks (the step after the NCCL all-gather
// of a row-sharded bank).  One CTA per query: load n_lists*k_in (score,idx) pairs,
// bitonic-sort the 64-bit ranking keys in shared memory, emit the best k_out.
// Indices here are GLOBAL int64 rows; ties are broken on the lower global row, which is the
// same rule every local top-k uses, so sharded results equal single-GPU results exactly.
#include "aura_common.cuh"

namespace aura {

// 96-bit ranking (score, idx64) does not fit a u64 key, so sort (orderable score, slot) keys and
// break score ties on the idx by a second compare inside the comparator.
__global__ void __launch_bounds__(256) topk_merge_kernel(const float* __restrict__ in_score,
                                                         const long long* __restrict__ in_idx, int n_in, int k_out,
                                                         float* __restrict__ out_score, long long* __restrict__ out_idx,
                                                         int n2) {
  extern __shared__ __align__(16) unsigned char smem[];
  u64* keys = reinterpret_cast<u64*>(smem);
  const float* sc = in_score + (size_t)blockIdx.x * n_in;
  const long long* ix = in_idx + (size_t)blockIdx.x * n_in;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    u64 key = 0ull;
    if (i < n_in && ix[i] >= 0) key = ((u64)f32_orderable(sc[i]) << 32) | (u64)(0xFFFFFFFFu - (unsigned)i);
    keys[i] = key;
  }
  // bitonic sort, descending by (score, then lower idx)
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const u64 a = keys[lo], b = keys[hi];
        bool a_before_b;  // does a rank ahead of b?
        const unsigned sa = (unsigned)(a >> 32), sb = (unsigned)(b >> 32);
        if (sa != sb) a_before_b = sa > sb;
        else if (a == 0ull || b == 0ull) a_before_b = a > b;
        else a_before_b = ix[key_row(a)] < ix[key_row(b)];
        if (a_before_b != desc) { keys[lo] = b; keys[hi] = a; }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k_out; i += blockDim.x) {
    const u64 key = i < n2 ? keys[i] : 0ull;
    out_score[(size_t)blockIdx.x * k_out + i] = key ? sc[key_row(key)] : -INFINITY;
    out_idx[(size_t)blockIdx.x * k_out + i] = key ? ix[key_row(key)] : -1ll;
  }
}

}  // namespace aura
using namespace aura;

extern "C" int aura_topk_merge(const float* in_score, const int64_t* in_idx, int n_queries, int n_lists, int k_in,
                               int k_out, float* out_score, int64_t* out_idx, void* stream) {
  AURA_REQUIRE(n_queries >= 0 && n_lists >= 1 && k_in >= 1 && k_out >= 1, AURA_ERR_INVALID_ARG,
               "aura_topk_merge: n_queries=%d n_lists=%d k_in=%d k_out=%d", n_queries, n_lists, k_in, k_out);
  if (n_queries == 0) return AURA_OK;
  AURA_REQUIRE(in_score && in_idx && out_score && out_idx, AURA_ERR_INVALID_ARG, "aura_topk_merge: null pointer");
  const long long n_in = (long long)n_lists * k_in;
  int n2 = 2;
  while (n2 < n_in) n2 <<= 1;
  AURA_REQUIRE((size_t)n2 * 8 <= (size_t)max_smem_optin() - 1024, AURA_ERR_UNSUPPORTED,
               "aura_topk_merge: n_lists*k_in=%lld does not fit shared memory", n_in);
  const size_t smem = (size_t)n2 * 8;
  AURA_CUDA_OK(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  topk_merge_kernel<<<n_queries, 256, smem, (cudaStream_t)stream>>>(in_score, reinterpret_cast<const long long*>(in_idx),
                                                                    (int)n_in, k_out, out_score,
                                                                    reinterpret_cast<long long*>(out_idx), n2);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}

// ---- sharded search plumbing: one payload per rank, one merge kernel -------------------------------------------------
// payload[b] = { idx[b][0..k) , score bits[b][0..k) , flag[b] } as int64: what a rank contributes to the all-gather
namespace aura {
__global__ void __launch_bounds__(256) pack_topk_kernel(const long long* __restrict__ idx, const float* __restrict__ score,
                                                        const int* __restrict__ flags, int n_queries, int k,
                                                        const long long* __restrict__ id_map, long long id_base,
                                                        long long* __restrict__ payload) {
  const int w = 2 * k + 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n_queries * w; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / w), j = (int)(i % w);
    long long v;
    if (j < k) { v = idx[(size_t)b * k + j]; if (v >= 0) v = id_map ? id_map[v] : v + id_base; }   // local row -> global id
    else if (j < 2 * k) v = (long long)__float_as_int(score[(size_t)b * k + (j - k)]);
    else v = flags ? (long long)flags[b] : 0ll;
    payload[i] = v;
  }
}

// gathered: [n_ranks][n_queries][2k+1]; one CTA per query merges the n_ranks*k pairs (score desc, lower global row on ties)
__global__ void __launch_bounds__(256) merge_packed_kernel(const long long* __restrict__ gathered, int n_ranks, int n_queries, int k,
                                                           float* __restrict__ out_score, long long* __restrict__ out_idx,
                                                           int* __restrict__ any_flag, int n2) {
  extern __shared__ __align__(16) unsigned char smem[];
  u64* keys = reinterpret_cast<u64*>(smem);
  const int b = blockIdx.x, w = 2 * k + 1, n_in = n_ranks * k;
  auto idx_at = [&](int i) { return gathered[((size_t)(i / k) * n_queries + b) * w + (i % k)]; };
  auto score_at = [&](int i) { return __int_as_float((int)gathered[((size_t)(i / k) * n_queries + b) * w + k + (i % k)]); };
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    u64 key = 0ull;
    if (i < n_in && idx_at(i) >= 0) key = ((u64)f32_orderable(score_at(i)) << 32) | (u64)(0xFFFFFFFFu - (unsigned)i);
    keys[i] = key;
  }
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const u64 x = keys[lo], y = keys[hi];
        bool x_first;
        const unsigned sx = (unsigned)(x >> 32), sy = (unsigned)(y >> 32);
        if (sx != sy) x_first = sx > sy;
        else if (x == 0ull || y == 0ull) x_first = x > y;
        else x_first = idx_at((int)key_row(x)) < idx_at((int)key_row(y));
        if (x_first != desc) { keys[lo] = y; keys[hi] = x; }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const u64 key = i < n2 ? keys[i] : 0ull;
    out_score[(size_t)b * k + i] = key ? score_at((int)key_row(key)) : -INFINITY;
    out_idx[(size_t)b * k + i] = key ? idx_at((int)key_row(key)) : -1ll;
  }
  if (threadIdx.x == 0 && any_flag) {
    int f = 0;
    for (int r = 0; r < n_ranks; ++r) f |= (int)gathered[((size_t)r * n_queries + b) * w + 2 * k] != 0;
    any_flag[b] = f;
  }
}
}  // namespace aura

extern "C" int aura_pack_topk(const int64_t* idx, const float* score, const int32_t* flags, int n_queries, int k,
                              const int64_t* id_map, int64_t id_base, int64_t* payload, void* stream) {
  AURA_REQUIRE(n_queries >= 1 && k >= 1 && idx && score && payload, AURA_ERR_INVALID_ARG, "aura_pack_topk: bad argument");
  const long long n = (long long)n_queries * (2 * k + 1);
  int g = (int)((n + 255) / 256);
  if (g > sm_count() * 8) g = sm_count() * 8;
  pack_topk_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(idx), score, flags, n_queries, k,
                                                        reinterpret_cast<const long long*>(id_map), (long long)id_base,
                                                        reinterpret_cast<long long*>(payload));
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}

extern "C" int aura_topk_merge_packed(const int64_t* gathered, int n_ranks, int n_queries, int k, float* out_score,
                                      int64_t* out_idx, int32_t* any_flag, void* stream) {
  AURA_REQUIRE(n_ranks >= 1 && n_queries >= 1 && k >= 1 && gathered && out_score && out_idx, AURA_ERR_INVALID_ARG,
               "aura_topk_merge_packed: bad argument");
  int n2 = 2;
  while (n2 < n_ranks * k) n2 <<= 1;
  AURA_REQUIRE((size_t)n2 * 8 <= (size_t)max_smem_optin() - 1024, AURA_ERR_UNSUPPORTED,
               "aura_topk_merge_packed: n_ranks*k=%d does not fit shared memory", n_ranks * k);
  const size_t smem = (size_t)n2 * 8;
  AURA_CUDA_OK(cudaFuncSetAttribute(merge_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  merge_packed_kernel<<<n_queries, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(gathered), n_ranks,
                                                                      n_queries, k, out_score,
                                                                      reinterpret_cast<long long*>(out_idx), any_flag, n2);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}


// ---- the same exchange without a library collective: every rank writes its payload straight into every peer's gather
// buffer over NVLink (peer-mapped symmetric memory) and raises a per-source flag there; the merge kernel of a rank waits
// for the flags of all sources and merges what has landed in its OWN memory.  One pack+scatter launch and one wait+merge
// launch per search, no host involvement, nothing the CUDA graph capture cannot record (the all-gather of
// B x k x 8 bytes per rank is latency-bound, 172 KB per rank on BASELINE config 2).
//   gather buffer of a rank: [2 slots][n_ranks][n_queries][2k+1] int64, then flags [2][n_ranks] uint32
// Slot = step parity.  A rank can only be one step ahead of a peer (its step s+1 merge needs the peer's step s+1 pack,
// which the peer's stream issues after its step s merge), so two slots are enough.  The step counters live in device
// memory and advance by one per launch, so the launches carry no per-step arguments (graph replay).
namespace aura {
static constexpr int AURA_MAX_RANKS = 16;
struct PeerPtrs { long long* buf[AURA_MAX_RANKS]; };

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) pack_scatter_kernel(const long long* __restrict__ idx, const float* __restrict__ score,
                                                           const int* __restrict__ flags, int n_queries, int k,
                                                           const long long* __restrict__ id_map, long long id_base,
                                                           const PeerPtrs peers, int rank, int n_ranks, size_t slot_elems,
                                                           unsigned* __restrict__ step_counter, unsigned* __restrict__ done_counter) {
  const unsigned step = *reinterpret_cast<volatile unsigned*>(step_counter);
  const unsigned slot = step & 1u;
  const int w = 2 * k + 1;
  const size_t total = (size_t)n_queries * w;
  const size_t dst0 = ((size_t)slot * n_ranks + rank) * total;               // my block inside every peer's buffer
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / w), j = (int)(i % w);
    long long v;
    if (j < k) { v = idx[(size_t)b * k + j]; if (v >= 0) v = id_map ? id_map[v] : v + id_base; }
    else if (j < 2 * k) v = (long long)__float_as_int(score[(size_t)b * k + (j - k)]);
    else v = flags ? (long long)flags[b] : 0ll;
    for (int r = 0; r < n_ranks; ++r) peers.buf[r][dst0 + i] = v;
  }
  __threadfence_system();                                 // this thread's peer stores are visible system-wide ...
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  // ... so the last CTA may tell every peer that rank `rank` has delivered step `step`
  if (threadIdx.x < n_ranks) {
    unsigned* pf = reinterpret_cast<unsigned*>(peers.buf[threadIdx.x] + 2 * slot_elems) + slot * n_ranks + rank;
    st_release_sys(pf, step + 1u);
  }
  if (threadIdx.x == 0) { *done_counter = 0u; *step_counter = step + 1u; }
}

__global__ void __launch_bounds__(256) merge_gathered_kernel(const long long* __restrict__ gather_buf, int n_ranks, int n_queries, int k,
                                                             size_t slot_elems, const unsigned* __restrict__ step_counter,
                                                             unsigned* __restrict__ done_counter, float* __restrict__ out_score,
                                                             long long* __restrict__ out_idx, int* __restrict__ any_flag, int n2) {
  extern __shared__ __align__(16) unsigned char smem[];
  u64* keys = reinterpret_cast<u64*>(smem);
  const unsigned step = *reinterpret_cast<const volatile unsigned*>(step_counter);
  const unsigned slot = step & 1u;
  const unsigned* flags = reinterpret_cast<const unsigned*>(gather_buf + 2 * slot_elems) + slot * n_ranks;
  if (threadIdx.x < n_ranks) {                            // wait until every source has delivered this step
    unsigned spins = 0;
    while (ld_acquire_sys(flags + threadIdx.x) != step + 1u) {
      if (++spins > (1u << 28)) { asm volatile("trap;"); }     // a peer that never arrives must not hang the GPU
      __nanosleep(32);
    }
  }
  __syncthreads();
  const long long* gathered = gather_buf + (size_t)slot * slot_elems;
  const int b = blockIdx.x, w = 2 * k + 1, n_in = n_ranks * k;
  auto idx_at = [&](int i) { return gathered[((size_t)(i / k) * n_queries + b) * w + (i % k)]; };
  auto score_at = [&](int i) { return __int_as_float((int)gathered[((size_t)(i / k) * n_queries + b) * w + k + (i % k)]); };
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    u64 key = 0ull;
    if (i < n_in && idx_at(i) >= 0) key = ((u64)f32_orderable(score_at(i)) << 32) | (u64)(0xFFFFFFFFu - (unsigned)i);
    keys[i] = key;
  }
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const u64 x = keys[lo], y = keys[hi];
        bool x_first;
        const unsigned sx = (unsigned)(x >> 32), sy = (unsigned)(y >> 32);
        if (sx != sy) x_first = sx > sy;
        else if (x == 0ull || y == 0ull) x_first = x > y;
        else x_first = idx_at((int)key_row(x)) < idx_at((int)key_row(y));
        if (x_first != desc) { keys[lo] = y; keys[hi] = x; }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const u64 key = i < n2 ? keys[i] : 0ull;
    out_score[(size_t)b * k + i] = key ? score_at((int)key_row(key)) : -INFINITY;
    out_idx[(size_t)b * k + i] = key ? idx_at((int)key_row(key)) : -1ll;
  }
  if (threadIdx.x == 0 && any_flag) {
    int f = 0;
    for (int r = 0; r < n_ranks; ++r) f |= (int)gathered[((size_t)r * n_queries + b) * w + 2 * k] != 0;
    any_flag[b] = f;
  }
  // the last CTA advances this rank's merge step
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(done_counter, 1u) == gridDim.x - 1) {
    *done_counter = 0u;
    *const_cast<unsigned*>(step_counter) = step + 1u;
  }
}
}  // namespace aura

extern "C" size_t aura_peer_gather_buffer_bytes(int n_ranks, int n_queries, int k) {
  if (n_ranks < 1 || n_ranks > AURA_MAX_RANKS || n_queries < 1 || k < 1) return 0;
  const size_t slot_elems = (size_t)n_ranks * n_queries * (2 * k + 1);
  return 2 * slot_elems * 8 + 2 * (size_t)n_ranks * 4 + 64;
}

extern "C" int aura_pack_scatter(const int64_t* idx, const float* score, const int32_t* flags, int n_queries, int k,
                                 const int64_t* id_map, int64_t id_base, void* const* peer_bufs_host, int rank, int n_ranks,
                                 uint32_t* counters, void* stream) {
  AURA_REQUIRE(n_queries >= 1 && k >= 1 && idx && score && peer_bufs_host && counters, AURA_ERR_INVALID_ARG, "aura_pack_scatter: bad argument");
  AURA_REQUIRE(n_ranks >= 1 && n_ranks <= AURA_MAX_RANKS && rank >= 0 && rank < n_ranks, AURA_ERR_INVALID_ARG,
               "aura_pack_scatter: rank %d of %d", rank, n_ranks);
  PeerPtrs peers;
  for (int r = 0; r < AURA_MAX_RANKS; ++r) peers.buf[r] = r < n_ranks ? reinterpret_cast<long long*>(peer_bufs_host[r]) : nullptr;
  const size_t total = (size_t)n_queries * (2 * k + 1);
  int g = (int)((total + 255) / 256);
  if (g > sm_count() * 2) g = sm_count() * 2;
  pack_scatter_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(idx), score, flags, n_queries, k,
                                                           reinterpret_cast<const long long*>(id_map), (long long)id_base, peers, rank,
                                                           n_ranks, (size_t)n_ranks * total, counters, counters + 1);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}

extern "C" int aura_merge_gathered(const void* gather_buf, int n_ranks, int n_queries, int k, uint32_t* counters, float* out_score,
                                   int64_t* out_idx, int32_t* any_flag, void* stream) {
  AURA_REQUIRE(n_ranks >= 1 && n_ranks <= AURA_MAX_RANKS && n_queries >= 1 && k >= 1 && gather_buf && counters && out_score && out_idx,
               AURA_ERR_INVALID_ARG, "aura_merge_gathered: bad argument");
  int n2 = 2;
  while (n2 < n_ranks * k) n2 <<= 1;
  AURA_REQUIRE((size_t)n2 * 8 <= (size_t)max_smem_optin() - 1024, AURA_ERR_UNSUPPORTED,
               "aura_merge_gathered: n_ranks*k=%d does not fit shared memory", n_ranks * k);
  const size_t smem = (size_t)n2 * 8;
  AURA_CUDA_OK(cudaFuncSetAttribute(merge_gathered_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t slot_elems = (size_t)n_ranks * n_queries * (2 * k + 1);
  merge_gathered_kernel<<<n_queries, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(gather_buf), n_ranks, n_queries, k,
                                                                        slot_elems, counters + 2, counters + 3, out_score,
                                                                        reinterpret_cast<long long*>(out_idx), any_flag, n2);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}

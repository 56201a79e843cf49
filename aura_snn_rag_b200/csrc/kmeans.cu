// Centroid index maintenance: K2 assign, K3 update, K1 online assign, CSR inverted lists.
// Replaces hippocampal.py:345-377 (rebuild_centroids: cdist + argmin + a Python loop of C masked
// means with a host sync each + a second cdist/argmin + a C-iteration count loop) and :218-230
// (per-write nearest-centroid + running mean).
//
// This file holds the CUDA-core (fp32 FMA) formulation: exact fp32 arithmetic, any d / C / M.
// The tcgen05 formulation of the assign GEMM lives in gemm_topk.cu (tc_assign); this one stays as the
// exact-fp32 path for small problems and for re-checking rows whose two best centroids are
// closer than the tensor-core rounding.
#include "tc_common.cuh"

namespace aura {

// ------------------------------------------------------------------ K2: assign (SIMT tiled)
// score(x, c) = ||c||^2 - 2 x.c   (||x||^2 is constant per row; same expansion torch.cdist uses
// for M or C > 25, hippocampal.py:358,370).  argmin over c, first minimum on ties (ATen CPU argmin).
static constexpr int AS_BM = 64, AS_BN = 64, AS_BK = 16, AS_THREADS = 256;

template <bool BF16>
__global__ void __launch_bounds__(AS_THREADS) kmeans_assign_kernel(const void* __restrict__ rows, long long n_rows, int d,
                                                                    const float* __restrict__ cent, int n_cent,
                                                                    const float* __restrict__ cent_sq,
                                                                    int* __restrict__ assign, float* __restrict__ cid_f32,
                                                                    int cid_stride, float* __restrict__ best_out) {
  __shared__ float As[AS_BK][AS_BM + 4];
  __shared__ float Bs[AS_BK][AS_BN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4x4 micro-tile each
  const long long row0 = (long long)blockIdx.x * AS_BM;
  float best[4];
  int arg[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { best[i] = INFINITY; arg[i] = 0x7fffffff; }

  for (int c0 = 0; c0 < n_cent; c0 += AS_BN) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < d; k0 += AS_BK) {
      // cooperative loads: 64 x 16 of each operand, 4 elements per thread
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int e = threadIdx.x + t * AS_THREADS;  // 0..1023
        const int r = e >> 4, kk = e & 15;
        const long long gr = row0 + r;
        float va = 0.f, vb = 0.f;
        if (gr < n_rows && k0 + kk < d) {
          va = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rows)[(size_t)gr * d + k0 + kk])
                    : reinterpret_cast<const float*>(rows)[(size_t)gr * d + k0 + kk];
        }
        if (c0 + r < n_cent && k0 + kk < d) vb = cent[(size_t)(c0 + r) * d + k0 + kk];
        As[kk][r] = va;
        Bs[kk][r] = vb;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < AS_BK; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + tx * 4 + j;
        if (c < n_cent) {
          const float s = fmaf(-2.f, acc[i][j], cent_sq[c]);
          if (s < best[i] || (s == best[i] && c < arg[i])) { best[i] = s; arg[i] = c; }
        }
      }
  }
  // reduce across the 16 threads (tx) that share rows: they are 16 consecutive lanes
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(FULL, best[i], o);
      const int oa = __shfl_xor_sync(FULL, arg[i], o);
      if (ob < best[i] || (ob == best[i] && oa < arg[i])) { best[i] = ob; arg[i] = oa; }
    }
    const long long gr = row0 + ty * 4 + i;
    if (tx == 0 && gr < n_rows) {
      assign[gr] = arg[i];
      if (cid_f32) cid_f32[(size_t)gr * cid_stride] = (float)arg[i];   // float ids, hippocampal.py:376
      if (best_out) best_out[gr] = best[i];
    }
  }
}

__global__ void __launch_bounds__(256) row_sq_norm_kernel(const float* __restrict__ x, int n, int d, float* __restrict__ out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  float ss = 0.f;
  for (int e = lane; e < d; e += 32) { const float v = x[(size_t)w * d + e]; ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  if (lane == 0) out[w] = ss;
}

// ------------------------------------------------------------------ CSR inverted lists (counting sort)
__global__ void __launch_bounds__(256) hist_kernel(const int* __restrict__ cid, long long n, int n_lists,
                                                   int* __restrict__ counts) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = cid[i];
    if (c >= 0 && c < n_lists) atomicAdd(&counts[c], 1);
  }
}
// single CTA exclusive scan of counts[0..n_lists) -> offsets[0..n_lists], cursor := offsets
__global__ void __launch_bounds__(1024) scan_offsets_kernel(const int* __restrict__ counts, int n_lists,
                                                            int* __restrict__ offsets, int* __restrict__ cursor) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n_lists; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n_lists ? counts[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, w, o); if (lane >= o) w += t; }
      warp_tot[lane] = w;  // inclusive
    }
    __syncthreads();
    const int excl = carry_s + (warp ? warp_tot[warp - 1] : 0) + inc - v;
    if (i < n_lists) { offsets[i] = excl; cursor[i] = excl; }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n_lists] = carry_s;
}
__global__ void __launch_bounds__(256) scatter_kernel(const int* __restrict__ cid, long long n, int n_lists,
                                                      int* __restrict__ cursor, int* __restrict__ list_rows) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = cid[i];
    if (c >= 0 && c < n_lists) list_rows[atomicAdd(&cursor[c], 1)] = (int)i;
  }
}

// ------------------------------------------------------------------ K3: per-list sums (fp64) + finalize
// One CTA per (list, 256-column chunk): thread = column, loop over the list's rows.  fp64
// accumulation makes the sum independent of the (unordered) CSR row order in all but
// astronomically rare rounding cases, and is at least as accurate as the fp32 mean of :363.
template <bool BF16>
__global__ void __launch_bounds__(256) list_sums_kernel(const void* __restrict__ rows, int d,
                                                        const int* __restrict__ list_offsets,
                                                        const int* __restrict__ list_rows, double* __restrict__ sums) {
  const int c = blockIdx.x;
  const int col = blockIdx.y * 256 + threadIdx.x;
  if (col >= d) return;
  const int b = list_offsets[c], e = list_offsets[c + 1];
  double acc = 0.0;
  int i = b;
  for (; i + 4 <= e; i += 4) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t off = (size_t)list_rows[i + u] * d + col;
      v[u] = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rows)[off])
                  : reinterpret_cast<const float*>(rows)[off];
    }
    acc += (double)v[0]; acc += (double)v[1]; acc += (double)v[2]; acc += (double)v[3];
  }
  for (; i < e; ++i) {
    const size_t off = (size_t)list_rows[i] * d + col;
    acc += (double)(BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rows)[off])
                         : reinterpret_cast<const float*>(rows)[off]);
  }
  sums[(size_t)c * d + col] = acc;
}

// centroid[c] = sums[c] / count[c]  when count[c] > 0, else left untouched (empty cluster keeps its
// sampled seed, hippocampal.py:362).  counts are the per-list sizes (possibly all-reduced).
__global__ void __launch_bounds__(256) finalize_centroids_kernel(const double* __restrict__ sums,
                                                                 const long long* __restrict__ counts, int n_cent, int d,
                                                                 float* __restrict__ cent) {
  const size_t n = (size_t)n_cent * d;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const long long cnt = counts[i / d];
    if (cnt > 0) cent[i] = (float)(sums[i] / (double)cnt);
  }
}

__global__ void offsets_to_counts_kernel(const int* __restrict__ offsets, int n_lists, long long* __restrict__ counts_i64,
                                         float* __restrict__ counts_f32) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_lists) {
    const int c = offsets[i + 1] - offsets[i];
    if (counts_i64) counts_i64[i] = c;
    if (counts_f32) counts_f32[i] = (float)c;
  }
}

template <bool BF16>
__global__ void __launch_bounds__(256) gather_seed_rows_kernel(const void* __restrict__ rows, int d,
                                                               const long long* __restrict__ seeds, int n_seeds,
                                                               float* __restrict__ cent) {
  const int s = blockIdx.x;
  if (s >= n_seeds) return;
  const size_t r = (size_t)seeds[s];
  for (int e = threadIdx.x; e < d; e += blockDim.x)
    cent[(size_t)s * d + e] = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rows)[r * d + e])
                                   : reinterpret_cast<const float*>(rows)[r * d + e];
}

// ------------------------------------------------------------------ K1: online assign (one write)
// dist_c = ||centroid_c - f||_2 (direct form, as torch.norm(centroids_view - features, dim=1), :223),
// c* = first argmin; count[c*] += 1; eta = 1/max(count,1); centroid[c*] = (1-eta) centroid + eta f.
// Grid of CTAs over centroids; the last CTA to finish reduces the per-CTA minima and applies the
// update, so one write = one launch and consecutive writes serialise on the stream (the sequential
// semantics of :218-230 are preserved exactly).
template <bool BF16>
__global__ void __launch_bounds__(256) online_assign_kernel(const void* __restrict__ rows, long long row, int d,
                                                            float* __restrict__ cent, int n_live,
                                                            float* __restrict__ counts, int* __restrict__ cid_i32,
                                                            float* __restrict__ cid_f32, int cid_stride,
                                                            u64* __restrict__ partial, unsigned* __restrict__ counter) {
  __shared__ u64 warp_best[8];
  __shared__ int s_last;
  __shared__ int s_arg;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 8 + warp, tw = gridDim.x * 8;
  u64 best = 0ull;  // max of key(-dist, c)  ==  min dist, lower c on ties
  for (int c = gw; c < n_live; c += tw) {
    float ss = 0.f;
    for (int e = lane; e < d; e += 32) {
      const float f = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rows)[(size_t)row * d + e])
                           : reinterpret_cast<const float*>(rows)[(size_t)row * d + e];
      const float t = cent[(size_t)c * d + e] - f;
      ss = fmaf(t, t, ss);
    }
    ss = warp_sum(ss);
    const u64 key = make_key(-sqrtf(ss), (unsigned)c);
    best = key > best ? key : best;
  }
  if (lane == 0) warp_best[warp] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    u64 b = 0ull;
    for (int w = 0; w < 8; ++w) b = warp_best[w] > b ? warp_best[w] : b;
    partial[blockIdx.x] = b;
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) {
    u64 b = 0ull;
    for (unsigned i = 0; i < gridDim.x; ++i) { const u64 p = partial[i]; b = p > b ? p : b; }
    const int c = (int)key_row(b);
    s_arg = c;
    counts[c] += 1.0f;                                   // :225
    cid_i32[row] = c;
    if (cid_f32) cid_f32[(size_t)row * cid_stride] = (float)c;  // :230
    *counter = 0u;
  }
  __syncthreads();
  const int c = s_arg;
  const float eta = 1.0f / fmaxf(counts[c], 1.0f);       // :227
  for (int e = threadIdx.x; e < d; e += blockDim.x) {
    const float f = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rows)[(size_t)row * d + e])
                         : reinterpret_cast<const float*>(rows)[(size_t)row * d + e];
    cent[(size_t)c * d + e] = __fadd_rn(__fmul_rn(1.0f - eta, cent[(size_t)c * d + e]), __fmul_rn(eta, f));  // :228, no contraction
  }
}

// ------------------------------------------------------------------ K1b: a run of writes in ONE cooperative launch
// The writes are inherently sequential (each may move a centroid the next one is compared with), so the one-launch-per-
// write form above is launch- and L2-latency bound (~30 us per write at 4096 x 1024).  Here a persistent grid keeps the
// centroids STATIONARY: CTA b owns the contiguous slice [b*per, (b+1)*per) and holds it in shared memory for the whole
// run (4096 x 1024 fp32 = 16 MB = 113 KB per SM).  Per write: every CTA scores its slice against the new row from shared
// memory, publishes its best key, one grid barrier, every CTA reads the 148 keys, and only the owner of the winner updates
// its slice - same arithmetic in the same order as online_assign_kernel, so the results are bit-identical.
struct OnlineRunArgs {
  const void* rows; long long first_row; int n_writes, d, n_live, per, in_smem;
  float* cent; float* counts; int* cid_i32; float* cid_f32; int cid_stride;
  u64* partial;                  // [2][gridDim.x] published slots, double-buffered by write parity, zeroed before the launch
  unsigned long long* barrier;   // unused (kept zero)
};

template <bool BF16>
__global__ void __launch_bounds__(256, 1) online_assign_run_kernel(const OnlineRunArgs a) {
  extern __shared__ float osm[];
  float* xs2 = osm;                      // [2][d] the row being written / the next one (prefetched)
  float* cs = osm + 2 * a.d;             // [per][d] this CTA's centroid slice (in_smem)
  __shared__ u64 warp_best[8];
  __shared__ int s_arg;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;
  const int c0 = blockIdx.x * a.per, c1 = min(a.n_live, c0 + a.per);
  auto load_row = [&](int w, float* dst) {
    const long long row = a.first_row + w;
    for (int e = threadIdx.x; e < a.d; e += blockDim.x)
      dst[e] = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.rows)[(size_t)row * a.d + e])
                    : reinterpret_cast<const float*>(a.rows)[(size_t)row * a.d + e];
  };
  if (a.in_smem)
    for (int i = threadIdx.x; i < (c1 - c0) * a.d; i += blockDim.x) cs[i] = a.cent[(size_t)c0 * a.d + i];
  load_row(0, xs2);
  __syncthreads();
  for (int w = 0; w < a.n_writes; ++w) {
    const long long row = a.first_row + w;
    const float* xs = xs2 + (size_t)(w & 1) * a.d;
    u64 best = 0ull;
    for (int c = c0 + warp; c < c1; c += 8) {
      const float* cp = a.in_smem ? cs + (size_t)(c - c0) * a.d : a.cent + (size_t)c * a.d;
      float ss = 0.f;
      for (int e = lane; e < a.d; e += 32) {
        const float t = (a.in_smem ? cp[e] : __ldcg(cp + e)) - xs[e];
        ss = fmaf(t, t, ss);
      }
      ss = warp_sum(ss);
      const u64 key = make_key(-sqrtf(ss), (unsigned)c);
      best = key > best ? key : best;
    }
    if (lane == 0) warp_best[warp] = best;
    __syncthreads();
    // Publish + barrier in one step: slot[parity][cta] = (orderable(-dist) : 32 | ~centroid : 16 | tag : 16).  A slot is
    // ready when its tag is this write's; parity double-buffering keeps a fast CTA from overwriting a slot a slow one
    // still polls (it cannot publish write w+2 before every CTA has published w+1, i.e. has finished reading w).
    const unsigned tag = (unsigned)((w >> 1) + 1) & 0xFFFFu;
    unsigned long long* slots = reinterpret_cast<unsigned long long*>(a.partial) + (size_t)(w & 1) * G;
    if (threadIdx.x == 0) {
      u64 b = 0ull;
      for (int i = 0; i < 8; ++i) b = warp_best[i] > b ? warp_best[i] : b;
      // b = 0 (no centroid in this slice can win: empty slice) still publishes a tagged, lowest-ranking slot
      const u64 packed = (b & 0xFFFFFFFF00000000ull) | ((b & 0xFFFFull) << 16) | tag;
      __stcg(slots + blockIdx.x, (unsigned long long)packed);
    }
    if (w + 1 < a.n_writes) load_row(w + 1, xs2 + (size_t)((w + 1) & 1) * a.d);     // overlaps the wait below
    u64 m = 0ull;
    for (int i = threadIdx.x; i < G; i += blockDim.x) {
      unsigned long long p;
      unsigned polls = 0;
      while ((((p = *reinterpret_cast<volatile unsigned long long*>(slots + i))) & 0xFFFFull) != tag)
        if (++polls > (1u << 27)) __trap();           // a lost CTA must not hang the GPU
      m = (u64)p > m ? (u64)p : m;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const u64 other = __shfl_xor_sync(0xffffffffu, m, o);
      m = other > m ? other : m;
    }
    __syncthreads();                                    // warp_best was read by thread 0 above
    if (lane == 0) warp_best[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      u64 b = 0ull;
      for (int i = 0; i < 8; ++i) b = warp_best[i] > b ? warp_best[i] : b;
      s_arg = (int)(0xFFFFu - (unsigned)((b >> 16) & 0xFFFFull));     // key_row(): row = 0xFFFFFFFF - low word
    }
    __syncthreads();
    const int c = s_arg;
    if (c >= c0 && c < c1) {                            // owner: :225-230
      const float cnt = a.counts[c] + 1.0f;
      const float eta = 1.0f / fmaxf(cnt, 1.0f);
      float* cp = a.in_smem ? cs + (size_t)(c - c0) * a.d : a.cent + (size_t)c * a.d;
      for (int e = threadIdx.x; e < a.d; e += blockDim.x) {
        const float old = a.in_smem ? cp[e] : __ldcg(cp + e);
        cp[e] = __fadd_rn(__fmul_rn(1.0f - eta, old), __fmul_rn(eta, xs[e]));
      }
      __syncthreads();                                  // everyone has read counts[c]
      if (threadIdx.x == 0) {
        a.counts[c] = cnt;
        a.cid_i32[row] = c;
        if (a.cid_f32) a.cid_f32[(size_t)row * a.cid_stride] = (float)c;
      }
    }
    __syncthreads();
  }
  if (a.in_smem)
    for (int i = threadIdx.x; i < (c1 - c0) * a.d; i += blockDim.x) a.cent[(size_t)c0 * a.d + i] = cs[i];
}

void launch_row_sq_norms(const float* x, int n, int d, float* out, cudaStream_t st) {
  row_sq_norm_kernel<<<(n + 7) / 8, 256, 0, st>>>(x, n, d, out);
  note_launches(1);
}

void launch_scan_offsets(const int* counts, int n, int* offsets, int* cursor, cudaStream_t st) {
  scan_offsets_kernel<<<1, 1024, 0, st>>>(counts, n, offsets, cursor);
  note_launches(1);
}

static int blocks_for(long long n, int per) {
  long long g = (n + per - 1) / per;
  const long long cap = (long long)sm_count() * 8;
  return (int)(g < 1 ? 1 : g > cap ? cap : g);
}

}  // namespace aura
using namespace aura;

extern "C" size_t aura_kmeans_assign_workspace_bytes(int64_t n_rows, int d, int dtype, int n_centroids) {
  if (n_centroids < 1) return 0;
  const size_t csq = ((size_t)n_centroids * 4 + 255) / 256 * 256;
  return csq + tc_assign_workspace_bytes(n_rows, d, dtype, n_centroids) + 256;
}

extern "C" int aura_kmeans_assign(const void* rows, int dtype, int64_t n_rows, int d, const float* centroids,
                                  int n_centroids, const float* row_inv_norm, int32_t* assign, float* cid_f32,
                                  int cid_stride, float* best_score, void* workspace, size_t workspace_bytes, void* stream) {
  AURA_REQUIRE(dtype == AURA_F32 || dtype == AURA_BF16, AURA_ERR_INVALID_ARG, "aura_kmeans_assign: bad dtype %d", dtype);
  AURA_REQUIRE(n_rows >= 0 && d >= 1 && n_centroids >= 1, AURA_ERR_INVALID_ARG,
               "aura_kmeans_assign: n_rows=%lld d=%d n_centroids=%d", (long long)n_rows, d, n_centroids);
  if (n_rows == 0) return AURA_OK;
  AURA_REQUIRE(rows && centroids && assign && workspace, AURA_ERR_INVALID_ARG, "aura_kmeans_assign: null pointer");
  AURA_REQUIRE(workspace_bytes >= aura_kmeans_assign_workspace_bytes(n_rows, d, dtype, n_centroids), AURA_ERR_WORKSPACE,
               "aura_kmeans_assign: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* csq = reinterpret_cast<float*>(workspace);
  launch_row_sq_norms(centroids, n_centroids, d, csq, st);
  // large problems: tcgen05 GEMM + exact fp32 re-score of the near-tied candidates (gemm_topk.cu);
  // small ones (and callers without per-row norms): exact fp32 SIMT tiles
  if (row_inv_norm != nullptr && tc_assign_supported(rows, dtype, n_rows, d, n_centroids)) {
    void* tws = reinterpret_cast<unsigned char*>(workspace) + ((size_t)n_centroids * 4 + 255) / 256 * 256;
    return tc_assign(rows, dtype, n_rows, d, centroids, n_centroids, csq, row_inv_norm, assign, cid_f32, cid_stride,
                     best_score, tws, st);
  }
  const long long g = (n_rows + AS_BM - 1) / AS_BM;
  AURA_REQUIRE(g < 0x7fffffffll, AURA_ERR_UNSUPPORTED, "aura_kmeans_assign: too many rows");
  if (dtype == AURA_BF16)
    kmeans_assign_kernel<true><<<(int)g, AS_THREADS, 0, st>>>(rows, n_rows, d, centroids, n_centroids, csq, assign,
                                                              cid_f32, cid_stride, best_score);
  else
    kmeans_assign_kernel<false><<<(int)g, AS_THREADS, 0, st>>>(rows, n_rows, d, centroids, n_centroids, csq, assign,
                                                               cid_f32, cid_stride, best_score);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}

extern "C" size_t aura_ivf_build_lists_workspace_bytes(int n_lists) { return (size_t)n_lists * 8 + 256; }

extern "C" int aura_ivf_build_lists(const int32_t* cid, int64_t n_rows, int n_lists, int32_t* list_offsets,
                                    int32_t* list_rows, void* workspace, size_t workspace_bytes, void* stream) {
  AURA_REQUIRE(n_rows >= 0 && n_rows < 0x7fffffffll && n_lists >= 1, AURA_ERR_INVALID_ARG,
               "aura_ivf_build_lists: n_rows=%lld n_lists=%d", (long long)n_rows, n_lists);
  AURA_REQUIRE(list_offsets && workspace && (n_rows == 0 || (cid && list_rows)), AURA_ERR_INVALID_ARG,
               "aura_ivf_build_lists: null pointer");
  AURA_REQUIRE(workspace_bytes >= aura_ivf_build_lists_workspace_bytes(n_lists), AURA_ERR_WORKSPACE,
               "aura_ivf_build_lists: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int* counts = reinterpret_cast<int*>(workspace);
  int* cursor = counts + n_lists;
  AURA_CUDA_OK(cudaMemsetAsync(counts, 0, (size_t)n_lists * 4, st));
  if (n_rows > 0) hist_kernel<<<blocks_for(n_rows, 256), 256, 0, st>>>(cid, n_rows, n_lists, counts);
  scan_offsets_kernel<<<1, 1024, 0, st>>>(counts, n_lists, list_offsets, cursor);
  if (n_rows > 0) scatter_kernel<<<blocks_for(n_rows, 256), 256, 0, st>>>(cid, n_rows, n_lists, cursor, list_rows);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(n_rows > 0 ? 3 : 1);
  return AURA_OK;
}

extern "C" int aura_kmeans_seed(const void* rows, int dtype, int d, const int64_t* seed_rows, int n_seeds,
                                float* centroids, void* stream) {
  AURA_REQUIRE(dtype == AURA_F32 || dtype == AURA_BF16, AURA_ERR_INVALID_ARG, "aura_kmeans_seed: bad dtype %d", dtype);
  if (n_seeds <= 0) return AURA_OK;
  AURA_REQUIRE(rows && seed_rows && centroids && d >= 1, AURA_ERR_INVALID_ARG, "aura_kmeans_seed: null pointer");
  if (dtype == AURA_BF16)
    gather_seed_rows_kernel<true><<<n_seeds, 256, 0, (cudaStream_t)stream>>>(rows, d, reinterpret_cast<const long long*>(seed_rows), n_seeds, centroids);
  else
    gather_seed_rows_kernel<false><<<n_seeds, 256, 0, (cudaStream_t)stream>>>(rows, d, reinterpret_cast<const long long*>(seed_rows), n_seeds, centroids);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}

extern "C" int aura_kmeans_list_sums(const void* rows, int dtype, int d, const int32_t* list_offsets,
                                     const int32_t* list_rows, int n_lists, double* sums, int64_t* counts, void* stream) {
  AURA_REQUIRE(dtype == AURA_F32 || dtype == AURA_BF16, AURA_ERR_INVALID_ARG, "aura_kmeans_list_sums: bad dtype %d", dtype);
  AURA_REQUIRE(n_lists >= 1 && d >= 1 && rows && list_offsets && list_rows && sums && counts, AURA_ERR_INVALID_ARG,
               "aura_kmeans_list_sums: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 g(n_lists, (d + 255) / 256);
  if (dtype == AURA_BF16) list_sums_kernel<true><<<g, 256, 0, st>>>(rows, d, list_offsets, list_rows, sums);
  else list_sums_kernel<false><<<g, 256, 0, st>>>(rows, d, list_offsets, list_rows, sums);
  offsets_to_counts_kernel<<<(n_lists + 255) / 256, 256, 0, st>>>(list_offsets, n_lists, reinterpret_cast<long long*>(counts), nullptr);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(2);
  return AURA_OK;
}

extern "C" int aura_kmeans_finalize(const double* sums, const int64_t* counts, int n_centroids, int d, float* centroids,
                                    void* stream) {
  AURA_REQUIRE(n_centroids >= 1 && d >= 1 && sums && counts && centroids, AURA_ERR_INVALID_ARG,
               "aura_kmeans_finalize: bad argument");
  finalize_centroids_kernel<<<blocks_for((long long)n_centroids * d, 256), 256, 0, (cudaStream_t)stream>>>(
      sums, reinterpret_cast<const long long*>(counts), n_centroids, d, centroids);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}

extern "C" int aura_ivf_list_counts(const int32_t* list_offsets, int n_lists, float* counts_f32, void* stream) {
  AURA_REQUIRE(n_lists >= 1 && list_offsets && counts_f32, AURA_ERR_INVALID_ARG, "aura_ivf_list_counts: bad argument");
  offsets_to_counts_kernel<<<(n_lists + 255) / 256, 256, 0, (cudaStream_t)stream>>>(list_offsets, n_lists, nullptr, counts_f32);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}

extern "C" size_t aura_online_assign_workspace_bytes(void) { return (size_t)sm_count() * 2 * 8 + 256; }

extern "C" int aura_online_assign(const void* rows, int dtype, int d, int64_t first_row, int n_writes, float* centroids,
                                  int n_live, float* counts, int32_t* cid_i32, float* cid_f32, int cid_stride,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  AURA_REQUIRE(dtype == AURA_F32 || dtype == AURA_BF16, AURA_ERR_INVALID_ARG, "aura_online_assign: bad dtype %d", dtype);
  AURA_REQUIRE(d >= 1 && n_live >= 1 && first_row >= 0 && n_writes >= 0, AURA_ERR_INVALID_ARG,
               "aura_online_assign: d=%d n_live=%d first_row=%lld n_writes=%d", d, n_live, (long long)first_row, n_writes);
  if (n_writes == 0) return AURA_OK;
  AURA_REQUIRE(rows && centroids && counts && cid_i32 && workspace, AURA_ERR_INVALID_ARG, "aura_online_assign: null pointer");
  AURA_REQUIRE(workspace_bytes >= aura_online_assign_workspace_bytes(), AURA_ERR_WORKSPACE,
               "aura_online_assign: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* counter = reinterpret_cast<unsigned*>(workspace);
  u64* partial = reinterpret_cast<u64*>(reinterpret_cast<unsigned char*>(workspace) + 256);
  AURA_CUDA_OK(cudaMemsetAsync(counter, 0, 8, st));
  // a run of writes: one cooperative launch with the centroids stationary in shared memory (AURA_ONLINE_RUN=0: off)
  static int coop = -1;
  if (coop < 0) {
    int dev = 0, v = 0;
    coop = (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && v) ? 1 : 0;
    if (const char* e = getenv("AURA_ONLINE_RUN")) coop = coop && atoi(e) != 0;
  }
  if (coop && n_writes >= 4 && n_live <= 0xFFFF) {    // 16-bit centroid field in the published slots
    AURA_CUDA_OK(cudaMemsetAsync(workspace, 0, aura_online_assign_workspace_bytes(), st));
    OnlineRunArgs a;
    a.rows = rows; a.first_row = first_row; a.n_writes = n_writes; a.d = d; a.n_live = n_live;
    a.cent = centroids; a.counts = counts; a.cid_i32 = cid_i32; a.cid_f32 = cid_f32; a.cid_stride = cid_stride;
    a.partial = partial; a.barrier = reinterpret_cast<unsigned long long*>(workspace);
    int G = sm_count();
    if (G > n_live) G = n_live;
    a.per = (n_live + G - 1) / G;
    G = (n_live + a.per - 1) / a.per;
    size_t smem = ((size_t)a.per * d + 2 * (size_t)d) * 4;
    a.in_smem = smem + 1024 <= (size_t)max_smem_optin();
    if (!a.in_smem) smem = 2 * (size_t)d * 4;
    if (smem + 1024 <= (size_t)max_smem_optin()) {
      void* kern = dtype == AURA_BF16 ? (void*)online_assign_run_kernel<true> : (void*)online_assign_run_kernel<false>;
      AURA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      void* params[] = {(void*)&a};
      AURA_CUDA_OK(cudaLaunchCooperativeKernel(kern, dim3(G), dim3(256), params, smem, st));
      note_launches(1);
      return AURA_OK;
    }
  }
  int grid = (n_live + 7) / 8;
  const int cap = sm_count() * 2;
  if (grid > cap) grid = cap;
  for (int w = 0; w < n_writes; ++w) {
    if (dtype == AURA_BF16)
      online_assign_kernel<true><<<grid, 256, 0, st>>>(rows, first_row + w, d, centroids, n_live, counts, cid_i32,
                                                       cid_f32 ? cid_f32 + (size_t)0 : nullptr, cid_stride, partial, counter);
    else
      online_assign_kernel<false><<<grid, 256, 0, st>>>(rows, first_row + w, d, centroids, n_live, counts, cid_i32,
                                                        cid_f32, cid_stride, partial, counter);
  }
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(n_writes);
  return AURA_OK;
}

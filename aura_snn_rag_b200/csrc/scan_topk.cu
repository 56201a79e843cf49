// K5 - exact scan: cosine + per-row affine combine + fused top-k, single query or a small
// query block (<= 8 per pass).  Replaces hippocampal.py:272-307 (F.normalize x2, gather,
// torch.mm, ~10 elementwise launches, torch.topk) and .tmp_infer_old.py:40-49.
//
// Roofline: HBM.  Algorithmic bytes per query pass = n_rows * d * sizeof(row element)
// (+8 B/row of scale/bias, <0.3 %).  Design:
//   * persistent grid, one CTA per SM, each CTA owns a contiguous range of row "stages";
//   * a producer thread streams stages HBM -> shared memory with 1-D TMA bulk copies
//     (cp.async.bulk, mbarrier complete_tx, L2 evict_first) through an S-deep ring, so the
//     bytes in flight per SM (S x ~24 KB) do not depend on registers or occupancy;
//   * 8 consumer warps read the stage with conflict-free 128-bit LDS, one warp per row,
//     RU rows at a time against the (pre-normalised) query block held in shared memory;
//   * every warp keeps a register-resident sorted top-k (WarpTopK) with a warp-uniform
//     threshold test, so the common case per row is one compare;
//   * warp lists -> CTA list (bitonic sort in the drained stage buffers) -> global partials;
//     the last CTA to finish merges all partials and writes the result (no second launch).
#include "aura_common.cuh"

namespace aura {

static constexpr int SCAN_NW = 8;                       // consumer warps
static constexpr int SCAN_THREADS = 32 * (SCAN_NW + 1);  // + 1 producer warp
static constexpr int SCAN_MAX_STAGES = 8;
static constexpr int SCAN_RU = 2;                        // rows per warp step
static constexpr int FINAL_MERGE_CAP = 16384;            // keys the last CTA can sort in smem

struct ScanArgs {
  const void* rows;
  long long n_rows;
  int d;
  const float* queries;
  int n_queries;
  const float* scale;
  const float* bias;
  int k;
  long long row_base;
  long long* out_idx;
  float* out_score;
  u64* partial;        // [n_qblocks][grid][QB][k]
  unsigned* counters;  // [n_qblocks]
  int rows_per_stage;  // multiple of SCAN_NW
  int n_stage_bufs;
  long long n_stages;  // ceil(n_rows / rows_per_stage)
  unsigned stage_bytes;
  unsigned q_bytes;    // QB * d * 4 rounded up to 128
  unsigned merge_keys; // capacity (keys) of the merge area, power of two
};

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

__device__ __forceinline__ void final_merge_write(const u64* src, int n_keys, int k, u64* merge, int merge_cap,
                                                  long long row_base, long long* out_idx, float* out_score) {
  // src: n_keys keys (grid lists of k); merge area holds merge_cap keys.  Chunked so that any
  // grid*k works: keep the running best k at the front, refill the rest, sort, repeat.
  int done = 0, have = 0;
  while (done < n_keys || have == 0) {
    const int room = merge_cap - have;
    const int take = min(room, n_keys - done);
    __syncthreads();
    for (int i = threadIdx.x; i < take; i += blockDim.x) merge[have + i] = src[done + i];
    const int n2 = next_pow2(max(have + take, 2));
    for (int i = have + take + threadIdx.x; i < n2; i += blockDim.x) merge[i] = 0ull;
    block_bitonic_sort_desc(merge, n2);
    done += take;
    have = min(k, n2);
    if (take == 0) break;
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const u64 key = merge[i];
    out_idx[i] = key ? row_base + (long long)key_row(key) : -1ll;
    out_score[i] = key ? key_score(key) : -INFINITY;
  }
  __syncthreads();
}

// ---- row . query-block dot products out of shared memory ---------------------------------------
template <int QB>
__device__ __forceinline__ void dot_rows_f32(const float4* __restrict__ r0, const float4* __restrict__ r1,
                                             const float4* __restrict__ qs, int d4, int lane,
                                             float (&acc)[SCAN_RU][QB]) {
#pragma unroll 4
  for (int c = lane; c < d4; c += 32) {
    const float4 x0 = r0[c], x1 = r1[c];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
      const float4 q = qs[qi * d4 + c];
      acc[0][qi] = fmaf(x0.x, q.x, fmaf(x0.y, q.y, fmaf(x0.z, q.z, fmaf(x0.w, q.w, acc[0][qi]))));
      acc[1][qi] = fmaf(x1.x, q.x, fmaf(x1.y, q.y, fmaf(x1.z, q.z, fmaf(x1.w, q.w, acc[1][qi]))));
    }
  }
}
// bf16 rows: one 16-byte chunk = 8 elements, matched with two float4 of the fp32 query
template <int QB>
__device__ __forceinline__ void dot_rows_bf16(const uint4* __restrict__ r0, const uint4* __restrict__ r1,
                                              const float4* __restrict__ qs, int d8, int lane,
                                              float (&acc)[SCAN_RU][QB]) {
#pragma unroll 2
  for (int c = lane; c < d8; c += 32) {
    const uint4 x0 = r0[c], x1 = r1[c];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
      const float4 qa = qs[qi * 2 * d8 + 2 * c], qb = qs[qi * 2 * d8 + 2 * c + 1];
      float a0 = acc[0][qi], a1 = acc[1][qi];
      a0 = fmaf(bf16_lo(x0.x), qa.x, a0); a0 = fmaf(bf16_hi(x0.x), qa.y, a0);
      a0 = fmaf(bf16_lo(x0.y), qa.z, a0); a0 = fmaf(bf16_hi(x0.y), qa.w, a0);
      a0 = fmaf(bf16_lo(x0.z), qb.x, a0); a0 = fmaf(bf16_hi(x0.z), qb.y, a0);
      a0 = fmaf(bf16_lo(x0.w), qb.z, a0); a0 = fmaf(bf16_hi(x0.w), qb.w, a0);
      a1 = fmaf(bf16_lo(x1.x), qa.x, a1); a1 = fmaf(bf16_hi(x1.x), qa.y, a1);
      a1 = fmaf(bf16_lo(x1.y), qa.z, a1); a1 = fmaf(bf16_hi(x1.y), qa.w, a1);
      a1 = fmaf(bf16_lo(x1.z), qb.x, a1); a1 = fmaf(bf16_hi(x1.z), qb.y, a1);
      a1 = fmaf(bf16_lo(x1.w), qb.z, a1); a1 = fmaf(bf16_hi(x1.w), qb.w, a1);
      acc[0][qi] = a0; acc[1][qi] = a1;
    }
  }
}

// ---- the streaming kernel ------------------------------------------------------------------------
template <bool BF16, int QB, int KPL>
__global__ void __launch_bounds__(SCAN_THREADS, 1) scan_topk_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + SCAN_MAX_STAGES;
  float* qs = reinterpret_cast<float*>(smem + 128);
  unsigned char* stage0 = smem + 128 + a.q_bytes;
  u64* merge = reinterpret_cast<u64*>(stage0);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qblk = blockIdx.y;
  const int q0 = qblk * QB;
  const int d = a.d;
  const int S = a.n_stage_bufs;
  const int R = a.rows_per_stage;
  const size_t row_bytes = (size_t)d * (BF16 ? 2 : 4);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], SCAN_NW); }
    fence_mbar_init();
  }
  // query block -> shared, normalised like F.normalize(q, dim=1) (hippocampal.py:273)
  for (int qi = warp; qi < QB; qi += SCAN_NW + 1) {
    const bool live = (q0 + qi) < a.n_queries;
    const float* q = a.queries + (size_t)(live ? q0 + qi : 0) * d;
    float ss = 0.f;
    for (int e = lane; e < d; e += 32) { const float v = live ? q[e] : 0.f; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    for (int e = lane; e < d; e += 32) qs[qi * d + e] = live ? q[e] / denom : 0.f;
  }
  __syncthreads();

  // contiguous range of stages for this CTA
  const long long s_begin = (a.n_stages * (long long)blockIdx.x) / gridDim.x;
  const long long s_end = (a.n_stages * (long long)(blockIdx.x + 1)) / gridDim.x;

  WarpTopK<KPL> tk[QB];
#pragma unroll
  for (int qi = 0; qi < QB; ++qi) tk[qi].init();

  if (warp == SCAN_NW) {
    // ===== producer: one thread keeps the ring full =====
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      const unsigned char* base = reinterpret_cast<const unsigned char*>(a.rows);
      long long it = 0;
      for (long long g = s_begin; g < s_end; ++g, ++it) {
        const int slot = (int)(it % S);
        const unsigned ph = (unsigned)((it / S) & 1);
        mbar_wait(&empty[slot], ph ^ 1u);
        const long long r0 = g * R;
        const int nrows = (int)min((long long)R, a.n_rows - r0);
        const unsigned bytes = (unsigned)(nrows * row_bytes);
        mbar_arrive_expect_tx(&full[slot], bytes);
        bulk_g2s(stage0 + (size_t)slot * a.stage_bytes, base + (size_t)r0 * row_bytes, bytes, &full[slot], pol);
      }
    }
  } else {
    // ===== consumers: one warp per row, SCAN_RU rows per step =====
    const int rpw = R / SCAN_NW;
    long long it = 0;
    for (long long g = s_begin; g < s_end; ++g, ++it) {
      const int slot = (int)(it % S);
      const unsigned ph = (unsigned)((it / S) & 1);
      const long long r0 = g * R;
      const int nrows = (int)min((long long)R, a.n_rows - r0);
      const int w_lo = warp * rpw;
      const int w_hi = min(w_lo + rpw, nrows);
      mbar_wait(&full[slot], ph);
      const unsigned char* st = stage0 + (size_t)slot * a.stage_bytes;
      for (int r = w_lo; r < w_hi; r += SCAN_RU) {
        const int ra = r, rb = min(r + 1, w_hi - 1);
        // per-row affine terms, requested early so the loads overlap the dot products
        float sc = 1.f, bi = 0.f;
        if (lane < SCAN_RU) {
          const long long gr = r0 + (lane == 0 ? ra : rb);
          sc = a.scale ? a.scale[gr] : 1.f;
          bi = a.bias ? a.bias[gr] : 0.f;
        }
        float acc[SCAN_RU][QB];
#pragma unroll
        for (int u = 0; u < SCAN_RU; ++u)
#pragma unroll
          for (int qi = 0; qi < QB; ++qi) acc[u][qi] = 0.f;
        if (BF16) {
          dot_rows_bf16<QB>(reinterpret_cast<const uint4*>(st + ra * row_bytes),
                            reinterpret_cast<const uint4*>(st + rb * row_bytes),
                            reinterpret_cast<const float4*>(qs), d >> 3, lane, acc);
        } else {
          dot_rows_f32<QB>(reinterpret_cast<const float4*>(st + ra * row_bytes),
                           reinterpret_cast<const float4*>(st + rb * row_bytes),
                           reinterpret_cast<const float4*>(qs), d >> 2, lane, acc);
        }
#pragma unroll
        for (int u = 0; u < SCAN_RU; ++u) {
          const float s_u = __shfl_sync(FULL, sc, u), b_u = __shfl_sync(FULL, bi, u);
          const int rr = (u == 0) ? ra : r + 1;
          const bool valid = rr < w_hi;  // warp-uniform
#pragma unroll
          for (int qi = 0; qi < QB; ++qi) {
            const float dot = warp_sum(acc[u][qi]);
            if (valid) {
              const u64 key = make_key(fmaf(dot, s_u, b_u), (unsigned)(r0 + rr));
              if (key > tk[qi].thr) tk[qi].insert(key, lane);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
  }

  // ---- CTA merge, publish, and (last CTA) final merge ----
  __shared__ int s_last;
  const int k = a.k;
  u64* my_partial = a.partial + ((size_t)qblk * gridDim.x + blockIdx.x) * QB * k;
#pragma unroll
  for (int qi = 0; qi < QB; ++qi) publish_cta_topk<KPL>(tk[qi], warp, lane, SCAN_NW, merge, k, my_partial + qi * k);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&a.counters[qblk], 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // gather list qi of every CTA: keys are strided [cta][QB][k] -> handle per query
  for (int qi = 0; qi < QB && q0 + qi < a.n_queries; ++qi) {
    // compact the strided lists into a dense tail of the partial area of CTA 0?  No: read strided.
    const u64* src = a.partial + (size_t)qblk * gridDim.x * QB * k + (size_t)qi * k;
    int have = 0;
    const int G = gridDim.x;
    int list = 0;
    while (list < G) {
      const int lists_fit = max(1, ((int)a.merge_keys - have) / k);
      const int take = min(lists_fit, G - list);
      __syncthreads();
      for (int i = threadIdx.x; i < take * k; i += blockDim.x) {
        const int l = list + i / k, j = i % k;
        merge[have + i] = src[(size_t)l * QB * k + j];
      }
      const int n2 = next_pow2(max(have + take * k, 2));
      for (int i = have + take * k + threadIdx.x; i < n2; i += blockDim.x) merge[i] = 0ull;
      block_bitonic_sort_desc(merge, n2);
      list += take;
      have = min(k, n2);
    }
    long long* oi = a.out_idx + (size_t)(q0 + qi) * k;
    float* os = a.out_score + (size_t)(q0 + qi) * k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      const u64 key = merge[i];
      oi[i] = key ? a.row_base + (long long)key_row(key) : -1ll;
      os[i] = key ? key_score(key) : -INFINITY;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) a.counters[qblk] = 0u;  // leave the workspace reusable
}

// ---- generic kernel: any d / alignment, rows read straight from global memory ----------------
template <bool BF16, int KPL>
__global__ void __launch_bounds__(256) scan_topk_generic_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  u64* merge = reinterpret_cast<u64*>(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = blockDim.x >> 5;
  const int q = blockIdx.y;
  const int d = a.d;
  const float* qv = a.queries + (size_t)q * d;
  float ss = 0.f;
  for (int e = lane; e < d; e += 32) ss = fmaf(qv[e], qv[e], ss);
  const float denom = fmaxf(sqrtf(warp_sum(ss)), 1e-12f);

  WarpTopK<KPL> tk;
  tk.init();
  const long long gw = (long long)blockIdx.x * nw + warp, tw = (long long)gridDim.x * nw;
  for (long long r = gw; r < a.n_rows; r += tw) {
    float acc = 0.f;
    if (BF16) {
      const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(a.rows) + (size_t)r * d;
      for (int e = lane; e < d; e += 32) acc = fmaf(__bfloat162float(x[e]), qv[e] / denom, acc);
    } else {
      const float* x = reinterpret_cast<const float*>(a.rows) + (size_t)r * d;
      for (int e = lane; e < d; e += 32) acc = fmaf(x[e], qv[e] / denom, acc);
    }
    const float dot = warp_sum(acc);
    const float sc = a.scale ? a.scale[r] : 1.f, bi = a.bias ? a.bias[r] : 0.f;
    const u64 key = make_key(fmaf(dot, sc, bi), (unsigned)r);
    if (key > tk.thr) tk.insert(key, lane);
  }
  __shared__ int s_last;
  const int k = a.k;
  publish_cta_topk<KPL>(tk, warp, lane, nw, merge, k, a.partial + ((size_t)q * gridDim.x + blockIdx.x) * k);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&a.counters[q], 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  final_merge_write(a.partial + (size_t)q * gridDim.x * k, gridDim.x * k, k, merge, (int)a.merge_keys, a.row_base,
                    a.out_idx + (size_t)q * k, a.out_score + (size_t)q * k);
  if (threadIdx.x == 0) a.counters[q] = 0u;
}

// ---- host side ---------------------------------------------------------------------------------------
struct ScanPlan {
  bool pipelined;
  int qb, kpl, grid, n_qblocks;
  int rows_per_stage, n_stage_bufs;
  unsigned stage_bytes, q_bytes, merge_keys;
  size_t smem;
  long long n_stages;
};

static int pick_qb(int n_queries) { return n_queries >= 8 ? 8 : n_queries >= 3 ? 4 : n_queries == 2 ? 2 : 1; }

static bool make_plan(long long n_rows, int d, int dtype, int n_queries, int k, ScanPlan* p) {
  const size_t row_bytes = (size_t)d * (dtype == AURA_BF16 ? 2 : 4);
  p->kpl = k <= 32 ? 1 : k <= 64 ? 2 : 4;
  const int sms = sm_count();
  const int smem_cap = max_smem_optin() - 1024;  // static smem + slack
  p->qb = pick_qb(n_queries);
  // keep (QB x KPL) register lists sane: wide lists only with narrow query blocks
  if (p->kpl == 4 && p->qb > 2) p->qb = 2;
  if (p->kpl == 2 && p->qb > 4) p->qb = 4;
  p->pipelined = (row_bytes % 16 == 0);
  if (p->pipelined) {
    p->q_bytes = (unsigned)(((size_t)p->qb * d * 4 + 127) / 128 * 128);
    const size_t budget = (size_t)smem_cap - 128 - p->q_bytes;
    int rpw = (int)(24576 / (SCAN_NW * row_bytes));
    if (rpw < 1) rpw = 1;
    if (rpw > 64) rpw = 64;
    p->rows_per_stage = rpw * SCAN_NW;
    p->stage_bytes = (unsigned)(((size_t)p->rows_per_stage * row_bytes + 127) / 128 * 128);
    if ((size_t)p->q_bytes + 128 > (size_t)smem_cap || budget < 2 * (size_t)p->stage_bytes) p->pipelined = false;
    else {
      int s = (int)(budget / p->stage_bytes);
      p->n_stage_bufs = s > SCAN_MAX_STAGES ? SCAN_MAX_STAGES : s;
    }
  }
  if (p->pipelined) {
    p->n_stages = (n_rows + p->rows_per_stage - 1) / p->rows_per_stage;
    long long g = p->n_stages < sms ? (p->n_stages > 0 ? p->n_stages : 1) : sms;
    p->grid = (int)g;
    p->n_qblocks = (n_queries + p->qb - 1) / p->qb;
    const size_t ring = (size_t)p->n_stage_bufs * p->stage_bytes;
    size_t mk = 1;  // largest power of two of keys that fits the ring, capped
    while (mk * 2 * 8 <= ring && mk * 2 <= (size_t)FINAL_MERGE_CAP) mk *= 2;
    // the CTA merge needs SCAN_NW * 32 * KPL keys
    const size_t need = (size_t)SCAN_NW * 32 * p->kpl;
    if (mk < need) { p->pipelined = false; }
    p->merge_keys = (unsigned)mk;
    p->smem = 128 + p->q_bytes + ring;
  }
  if (!p->pipelined) {
    p->qb = 1;
    p->n_qblocks = n_queries;
    long long warps_needed = (n_rows + 3) / 4;
    long long g = (warps_needed + 7) / 8;
    const long long gmax = (long long)sms * 4;
    p->grid = (int)(g < 1 ? 1 : g > gmax ? gmax : g);
    p->merge_keys = 8192;
    p->smem = (size_t)p->merge_keys * 8;
    p->n_stages = 0;
  }
  return true;
}

template <bool BF16, int QB>
static cudaError_t launch_pipelined(const ScanPlan& p, const ScanArgs& a, cudaStream_t st) {
  void (*kern)(ScanArgs) = nullptr;
  switch (p.kpl) {
    case 1: kern = scan_topk_kernel<BF16, QB, 1>; break;
    case 2: kern = scan_topk_kernel<BF16, (QB > 4 ? 4 : QB), 2>; break;
    default: kern = scan_topk_kernel<BF16, (QB > 2 ? 2 : QB), 4>; break;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  if (e != cudaSuccess) return e;
  kern<<<dim3(p.grid, p.n_qblocks), SCAN_THREADS, p.smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace aura

using namespace aura;

extern "C" size_t aura_scan_topk_workspace_bytes(int64_t n_rows, int d, int n_queries, int k) {
  if (n_queries < 1 || k < 1 || d < 1) return 0;
  // upper bound over both plans: partial keys [n_queries rounded up to 8][grid<=4*SMs][k] + counters
  const size_t grid = (size_t)sm_count() * 4;
  const size_t nq = ((size_t)n_queries + 7) / 8 * 8;
  (void)n_rows;
  return nq * grid * (size_t)k * 8 + nq * 4 + 256;
}

extern "C" int aura_scan_topk(const void* rows, int dtype, int64_t n_rows, int d, const float* queries, int n_queries,
                              const float* scale, const float* bias, int k, int64_t row_base, int64_t* out_idx,
                              float* out_score, void* workspace, size_t workspace_bytes, void* stream) {
  AURA_REQUIRE(dtype == AURA_F32 || dtype == AURA_BF16, AURA_ERR_INVALID_ARG, "aura_scan_topk: bad dtype %d", dtype);
  AURA_REQUIRE(n_rows >= 0 && n_rows < 0xFFFFFFFFll, AURA_ERR_INVALID_ARG, "aura_scan_topk: n_rows %lld out of range",
               (long long)n_rows);
  AURA_REQUIRE(d >= 1 && n_queries >= 1, AURA_ERR_INVALID_ARG, "aura_scan_topk: d=%d n_queries=%d", d, n_queries);
  AURA_REQUIRE(k >= 1 && k <= AURA_MAX_K, AURA_ERR_INVALID_ARG, "aura_scan_topk: k=%d not in [1,%d]", k, AURA_MAX_K);
  AURA_REQUIRE(queries && out_idx && out_score && (rows || n_rows == 0), AURA_ERR_INVALID_ARG,
               "aura_scan_topk: null pointer");
  AURA_REQUIRE(workspace && workspace_bytes >= aura_scan_topk_workspace_bytes(n_rows, d, n_queries, k),
               AURA_ERR_WORKSPACE, "aura_scan_topk: workspace too small (%zu bytes)", workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  ScanPlan p;
  make_plan(n_rows, d, dtype, n_queries, k, &p);

  const size_t nq8 = ((size_t)n_queries + 7) / 8 * 8;
  unsigned* counters = reinterpret_cast<unsigned*>(workspace);
  u64* partial = reinterpret_cast<u64*>(reinterpret_cast<unsigned char*>(workspace) + (nq8 * 4 + 255) / 256 * 256);
  AURA_CUDA_OK(cudaMemsetAsync(counters, 0, nq8 * 4, st));

  ScanArgs a;
  a.rows = rows; a.n_rows = n_rows; a.d = d; a.queries = queries; a.n_queries = n_queries;
  a.scale = scale; a.bias = bias; a.k = k; a.row_base = row_base;
  a.out_idx = reinterpret_cast<long long*>(out_idx); a.out_score = out_score;
  a.partial = partial; a.counters = counters;
  a.rows_per_stage = p.rows_per_stage; a.n_stage_bufs = p.n_stage_bufs; a.n_stages = p.n_stages;
  a.stage_bytes = p.stage_bytes; a.q_bytes = p.q_bytes; a.merge_keys = p.merge_keys;

  cudaError_t e;
  if (p.pipelined) {
    const bool bf = dtype == AURA_BF16;
    switch (p.qb) {
      case 1: e = bf ? launch_pipelined<true, 1>(p, a, st) : launch_pipelined<false, 1>(p, a, st); break;
      case 2: e = bf ? launch_pipelined<true, 2>(p, a, st) : launch_pipelined<false, 2>(p, a, st); break;
      case 4: e = bf ? launch_pipelined<true, 4>(p, a, st) : launch_pipelined<false, 4>(p, a, st); break;
      default: e = bf ? launch_pipelined<true, 8>(p, a, st) : launch_pipelined<false, 8>(p, a, st); break;
    }
  } else {
    void (*kern)(ScanArgs);
    const bool bf = dtype == AURA_BF16;
    if (p.kpl == 1) kern = bf ? scan_topk_generic_kernel<true, 1> : scan_topk_generic_kernel<false, 1>;
    else if (p.kpl == 2) kern = bf ? scan_topk_generic_kernel<true, 2> : scan_topk_generic_kernel<false, 2>;
    else kern = bf ? scan_topk_generic_kernel<true, 4> : scan_topk_generic_kernel<false, 4>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e == cudaSuccess) {
      kern<<<dim3(p.grid, n_queries), 256, p.smem, st>>>(a);
      e = cudaGetLastError();
    }
  }
  AURA_CUDA_OK(e);
  return AURA_OK;
}

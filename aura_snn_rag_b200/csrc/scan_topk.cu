// K5 - streaming scan: cosine + per-row affine combine + fused top-k.
//   direct mode   : every live row of the bank        (exact path, hippocampal.py:272-307,
//                                                       .tmp_infer_old.py:40-49)
//   indirect mode : rows of the probed inverted lists  (centroid path fine stage, :264-279,
//                                                       with CSR lists replacing the P mask passes)
//
// Roofline: HBM.  Algorithmic bytes per query pass = rows_scanned * d * sizeof(element)
// (+8 B/row of scale/bias, <0.3 %).  Design:
//   * persistent grid, one CTA per SM, each CTA owns a contiguous range of row "stages";
//   * a producer warp streams stages HBM -> shared memory with 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx, L2 evict_first) through an S-deep ring, so the
//     bytes in flight per SM (~190 KB) do not depend on registers or occupancy.  The per-row
//     scale/bias terms ride in the same stage (two more bulk copies) so consumers never wait
//     on a dependent global load;
//   * NW consumer warps read the stage with conflict-free 128-bit LDS, one warp per row, RU rows
//     at a time against the (pre-normalised) query block held in shared memory;
//   * every warp keeps a register-resident sorted top-k (WarpTopK) with a warp-uniform
//     threshold test, so the common case per row is one compare;
//   * warp lists -> CTA list (bitonic sort in the drained ring) -> global partials; the last CTA
//     to finish merges all partials and writes the result (single launch per query block).
#include <stdlib.h>

#include "aura_common.cuh"

namespace aura {

static constexpr int SCAN_MAX_STAGES = 8;
static constexpr int FINAL_MERGE_CAP = 16384;  // keys the last CTA can sort in smem

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
  int rows_per_stage;  // multiple of NW (and of 4)
  int n_stage_bufs;
  unsigned stage_bytes;   // row area of one stage, multiple of 128
  unsigned q_bytes;       // QB * d * 4 rounded up to 128
  unsigned merge_keys;    // capacity (keys) of the merge area, power of two
  int terms_bulk;         // scale/bias are 16-byte aligned: stage them with bulk copies
  // indirect mode (inverted lists)
  const long long* probes;  // [n_queries][nprobe] list ids (<0 or >= n_lists: empty)
  int nprobe;
  int n_lists;
  const int* list_offsets;  // [n_lists + 1]
  const int* list_rows;     // [n_rows_in_lists]
  int interleave;           // stage g of CTA b = b + it*grid (1) or a contiguous range (0)
  int empty_ok;             // indirect mode: a query whose probed lists are all empty returns no result (row-sharded
                            // callers) instead of scanning every row (hippocampal.py:269-270)
};

// ---- row . query-block dot products out of shared memory ---------------------------------------
template <int QB, int RU>
__device__ __forceinline__ void dot_rows_f32(const float4* const (&rp)[RU], const float4* __restrict__ qs, int d4,
                                             int lane, float (&acc)[RU][QB]) {
#pragma unroll 2
  for (int c = lane; c < d4; c += 32) {
    float4 x[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) x[u] = rp[u][c];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
      const float4 q = qs[qi * d4 + c];
#pragma unroll
      for (int u = 0; u < RU; ++u)
        acc[u][qi] = fmaf(x[u].x, q.x, fmaf(x[u].y, q.y, fmaf(x[u].z, q.z, fmaf(x[u].w, q.w, acc[u][qi]))));
    }
  }
}
// bf16 rows: one 16-byte chunk = 8 elements, matched with two float4 of the fp32 query
template <int QB, int RU>
__device__ __forceinline__ void dot_rows_bf16(const uint4* const (&rp)[RU], const float4* __restrict__ qs, int d8,
                                              int lane, float (&acc)[RU][QB]) {
#pragma unroll 2
  for (int c = lane; c < d8; c += 32) {
    uint4 x[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) x[u] = rp[u][c];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
      const float4 qa = qs[qi * 2 * d8 + 2 * c], qb = qs[qi * 2 * d8 + 2 * c + 1];
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        float a = acc[u][qi];
        a = fmaf(bf16_lo(x[u].x), qa.x, a); a = fmaf(bf16_hi(x[u].x), qa.y, a);
        a = fmaf(bf16_lo(x[u].y), qa.z, a); a = fmaf(bf16_hi(x[u].y), qa.w, a);
        a = fmaf(bf16_lo(x[u].z), qb.x, a); a = fmaf(bf16_hi(x[u].z), qb.y, a);
        a = fmaf(bf16_lo(x[u].w), qb.z, a); a = fmaf(bf16_hi(x[u].w), qb.w, a);
        acc[u][qi] = a;
      }
    }
  }
}

// ---- the streaming kernel ------------------------------------------------------------------------
// Shared memory map (dynamic):
//   [0,128)            full[8], empty[8] mbarriers
//   [128, +q_bytes)    normalised query block
//   then per stage s:  rows (stage_bytes) | scale R*4 | bias R*4 | rowid R*4        (ring)
//   then (INDIRECT)    pre[nprobe+1], lbase[nprobe]
template <bool BF16, int QB, int KPL, int NW, int RU, bool INDIRECT>
__global__ void __launch_bounds__(32 * (NW + 1), 1) scan_topk_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + SCAN_MAX_STAGES;
  float* qs = reinterpret_cast<float*>(smem + 128);
  unsigned char* ring = smem + 128 + a.q_bytes;
  u64* merge = reinterpret_cast<u64*>(ring);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qblk = blockIdx.y;
  const int q0 = qblk * QB;
  const int d = a.d;
  const int S = a.n_stage_bufs;
  const int R = a.rows_per_stage;
  const size_t row_bytes = (size_t)d * (BF16 ? 2 : 4);
  const size_t slot_bytes = (size_t)a.stage_bytes + (size_t)R * 12;
  int* pre = reinterpret_cast<int*>(ring + (size_t)S * slot_bytes);  // INDIRECT only
  int* lbase = pre + (a.nprobe + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
    fence_mbar_init();
  }
  // query block -> shared, normalised like F.normalize(q, dim=1) (hippocampal.py:273)
  for (int qi = warp; qi < QB; qi += NW + 1) {
    const bool live = (q0 + qi) < a.n_queries;
    const float* q = a.queries + (size_t)(live ? q0 + qi : 0) * d;
    float ss = 0.f;
    for (int e = lane; e < d; e += 32) { const float v = live ? q[e] : 0.f; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    for (int e = lane; e < d; e += 32) qs[qi * d + e] = live ? q[e] / denom : 0.f;
  }
  long long total = a.n_rows;  // positions to scan
  bool mapped = false;         // positions go through the inverted lists
  if (INDIRECT) {
    if (warp == 0) {
      // prefix sums of the probed lists' lengths (nprobe <= 128: 4 per lane)
      int run = 0;
      for (int b = 0; b < a.nprobe; b += 32) {
        const int p = b + lane;
        int len = 0, base = 0;
        if (p < a.nprobe) {
          const long long c = a.probes[(size_t)qblk * a.nprobe + p];
          if (c >= 0 && c < a.n_lists) { base = a.list_offsets[c]; len = a.list_offsets[c + 1] - base; }
          lbase[p] = base;
        }
        int inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
        if (p < a.nprobe) pre[p + 1] = run + inc;
        run += __shfl_sync(FULL, inc, 31);
      }
      if (lane == 0) pre[0] = 0;
    }
    __syncthreads();
    const int t = pre[a.nprobe];
    if (t > 0 || a.empty_ok) { total = t; mapped = true; }  // no candidates -> all live rows, hippocampal.py:269-270
  } else {
    __syncthreads();
  }

  // contiguous range of stages for this CTA
  const long long n_stages = (total + R - 1) / R;
  const bool il = a.interleave != 0;
  const long long s_begin = il ? (long long)blockIdx.x : (n_stages * (long long)blockIdx.x) / gridDim.x;
  const long long s_end = il ? n_stages : (n_stages * (long long)(blockIdx.x + 1)) / gridDim.x;
  const long long s_step = il ? (long long)gridDim.x : 1;

  WarpTopK<KPL> tk[QB];
#pragma unroll
  for (int qi = 0; qi < QB; ++qi) tk[qi].init();

  if (warp == NW) {
    // ===== producer warp =====
    const uint64_t pol = l2_policy_evict_first();
    const unsigned char* base = reinterpret_cast<const unsigned char*>(a.rows);
    int slot = 0; unsigned ph = 0;
    for (long long g = s_begin; g < s_end; g += s_step, slot = (slot + 1 == S ? 0 : slot + 1), ph ^= (slot == 0)) {
      unsigned char* st = ring + (size_t)slot * slot_bytes;
      float* st_scale = reinterpret_cast<float*>(st + a.stage_bytes);
      float* st_bias = st_scale + R;
      int* st_rowid = reinterpret_cast<int*>(st_bias + R);
      const long long p0 = g * R;
      const int nrows = (int)min((long long)R, total - p0);
      if (lane == 0) mbar_wait(&empty[slot], ph ^ 1u);
      __syncwarp();
      if (!mapped) {
        if (lane == 0) {
          // rows + the 16-byte-aligned part of scale/bias in three bulk copies
          const int n4 = a.terms_bulk ? (nrows & ~3) : 0;
          unsigned bytes = (unsigned)(nrows * row_bytes);
          if (a.scale && n4) bytes += (unsigned)n4 * 4u;
          if (a.bias && n4) bytes += (unsigned)n4 * 4u;
          mbar_arrive_expect_tx(&full[slot], bytes);
          bulk_g2s(st, base + (size_t)p0 * row_bytes, (unsigned)(nrows * row_bytes), &full[slot], pol);
          if (a.scale && n4) bulk_g2s(st_scale, a.scale + p0, (unsigned)n4 * 4u, &full[slot], pol);
          if (a.bias && n4) bulk_g2s(st_bias, a.bias + p0, (unsigned)n4 * 4u, &full[slot], pol);
        }
      } else {
        // inverted lists: one bulk copy per row, issued by up to 32 lanes at a time
        for (int b = 0; b < nrows; b += 32) {
          const int j = b + lane;
          if (j < nrows) {
            const int pos = (int)(p0 + j);
            int lo = 0, hi = a.nprobe;  // last p with pre[p] <= pos
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pre[mid] <= pos) lo = mid; else hi = mid; }
            st_rowid[j] = a.list_rows[lbase[lo] + (pos - pre[lo])];
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(&full[slot], (unsigned)(nrows * row_bytes));
        __syncwarp();
        for (int b = 0; b < nrows; b += 32) {
          const int j = b + lane;
          if (j < nrows)
            bulk_g2s(st + (size_t)j * row_bytes, base + (size_t)st_rowid[j] * row_bytes, (unsigned)row_bytes,
                     &full[slot], pol);
        }
      }
    }
  } else {
    // ===== consumers: one warp per row, RU rows per step =====
    const int rpw = R / NW;  // <= 32
    int slot = 0; unsigned ph = 0;
    for (long long g = s_begin; g < s_end; g += s_step, slot = (slot + 1 == S ? 0 : slot + 1), ph ^= (slot == 0)) {
      const unsigned char* st = ring + (size_t)slot * slot_bytes;
      const float* st_scale = reinterpret_cast<const float*>(st + a.stage_bytes);
      const float* st_bias = st_scale + R;
      const int* st_rowid = reinterpret_cast<const int*>(st_bias + R);
      const long long p0 = g * R;
      const int nrows = (int)min((long long)R, total - p0);
      const int w_lo = warp * rpw;
      const int w_hi = min(w_lo + rpw, nrows);
      mbar_wait(&full[slot], ph);
      // this warp's per-row terms: lane j <-> row w_lo + j
      float sc = 1.f, bi = 0.f;
      unsigned rid = 0;
      if (w_lo + lane < w_hi) {
        const int j = w_lo + lane;
        if (mapped) {
          rid = (unsigned)st_rowid[j];
          sc = a.scale ? a.scale[rid] : 1.f;
          bi = a.bias ? a.bias[rid] : 0.f;
        } else {
          rid = (unsigned)(p0 + j);
          if (a.terms_bulk && j < (nrows & ~3)) {
            sc = a.scale ? st_scale[j] : 1.f;
            bi = a.bias ? st_bias[j] : 0.f;
          } else {  // unaligned tail of the bank: plain loads
            sc = a.scale ? a.scale[rid] : 1.f;
            bi = a.bias ? a.bias[rid] : 0.f;
          }
        }
      }
      for (int r = w_lo; r < w_hi; r += RU) {
        float acc[RU][QB];
#pragma unroll
        for (int u = 0; u < RU; ++u)
#pragma unroll
          for (int qi = 0; qi < QB; ++qi) acc[u][qi] = 0.f;
        if (BF16) {
          const uint4* rp[RU];
#pragma unroll
          for (int u = 0; u < RU; ++u) rp[u] = reinterpret_cast<const uint4*>(st + (size_t)min(r + u, w_hi - 1) * row_bytes);
          dot_rows_bf16<QB, RU>(rp, reinterpret_cast<const float4*>(qs), d >> 3, lane, acc);
        } else {
          const float4* rp[RU];
#pragma unroll
          for (int u = 0; u < RU; ++u) rp[u] = reinterpret_cast<const float4*>(st + (size_t)min(r + u, w_hi - 1) * row_bytes);
          dot_rows_f32<QB, RU>(rp, reinterpret_cast<const float4*>(qs), d >> 2, lane, acc);
        }
        // all RU*QB warp reductions first (independent shuffle chains overlap), then the rare insert path
        float dots[RU][QB];
#pragma unroll
        for (int u = 0; u < RU; ++u)
#pragma unroll
          for (int qi = 0; qi < QB; ++qi) dots[u][qi] = warp_sum(acc[u][qi]);
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          const int src = min(r + u, w_hi - 1) - w_lo;
          const float s_u = __shfl_sync(FULL, sc, src), b_u = __shfl_sync(FULL, bi, src);
          const unsigned rid_u = __shfl_sync(FULL, rid, src);
          if ((r + u) < w_hi) {  // warp-uniform
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) {
              const u64 key = make_key(fmaf(dots[u][qi], s_u, b_u), rid_u);
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
  for (int qi = 0; qi < QB; ++qi) publish_cta_topk<KPL>(tk[qi], warp, lane, NW, merge, k, my_partial + qi * k);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&a.counters[qblk], 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int qi = 0; qi < QB && q0 + qi < a.n_queries; ++qi) {
    final_merge_write(a.partial + (size_t)qblk * gridDim.x * QB * k + (size_t)qi * k, gridDim.x, (size_t)QB * k, k,
                      merge, (int)a.merge_keys, a.row_base, a.out_idx + (size_t)(q0 + qi) * k,
                      a.out_score + (size_t)(q0 + qi) * k);
  }
  if (threadIdx.x == 0) a.counters[qblk] = 0u;  // leave the workspace reusable
}

// ---- generic kernel: any d / alignment, rows read straight from global memory ----------------
template <bool BF16, int KPL>
__global__ void __launch_bounds__(256) scan_topk_generic_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  u64* merge = reinterpret_cast<u64*>(smem);
  __shared__ int s_total;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = blockDim.x >> 5;
  const int q = blockIdx.y;
  const int d = a.d;
  const float* qv = a.queries + (size_t)q * d;
  float ss = 0.f;
  for (int e = lane; e < d; e += 32) ss = fmaf(qv[e], qv[e], ss);
  const float denom = fmaxf(sqrtf(warp_sum(ss)), 1e-12f);

  // indirect mode: total candidates over the probed lists
  long long total = a.n_rows;
  bool mapped = false;
  if (a.probes) {
    if (threadIdx.x == 0) {
      int t = 0;
      for (int p = 0; p < a.nprobe; ++p) {
        const long long c = a.probes[(size_t)q * a.nprobe + p];
        if (c >= 0 && c < a.n_lists) t += a.list_offsets[c + 1] - a.list_offsets[c];
      }
      s_total = t;
    }
    __syncthreads();
    if (s_total > 0 || a.empty_ok) { total = s_total; mapped = true; }
  }

  WarpTopK<KPL> tk;
  tk.init();
  const long long gw = (long long)blockIdx.x * nw + warp, tw = (long long)gridDim.x * nw;
  for (long long pos = gw; pos < total; pos += tw) {
    long long r = pos;
    if (mapped) {
      long long rem = pos;
      for (int p = 0; p < a.nprobe; ++p) {
        const long long c = a.probes[(size_t)q * a.nprobe + p];
        if (c < 0 || c >= a.n_lists) continue;
        const int b = a.list_offsets[c], len = a.list_offsets[c + 1] - b;
        if (rem < len) { r = a.list_rows[b + rem]; break; }
        rem -= len;
      }
    }
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
  final_merge_write(a.partial + (size_t)q * gridDim.x * k, gridDim.x, (size_t)k, k, merge, (int)a.merge_keys,
                    a.row_base, a.out_idx + (size_t)q * k, a.out_score + (size_t)q * k);
  if (threadIdx.x == 0) a.counters[q] = 0u;
}

// ---- host side ---------------------------------------------------------------------------------------
static constexpr int SCAN_NW = 8;
static constexpr int SCAN_RU = 2;

struct ScanPlan {
  bool pipelined;
  int qb, kpl, grid, n_qblocks;
  int rows_per_stage, n_stage_bufs;
  unsigned stage_bytes, q_bytes, merge_keys;
  size_t smem;
};

static int pick_qb(int n_queries) { return n_queries >= 8 ? 8 : n_queries >= 3 ? 4 : n_queries == 2 ? 2 : 1; }

// expected_rows: rows one query scans (n_rows for direct mode, the candidate estimate for lists)
static void make_plan(long long expected_rows, int d, int dtype, int n_queries, int k, bool indirect, int nprobe,
                      ScanPlan* p) {
  const size_t row_bytes = (size_t)d * (dtype == AURA_BF16 ? 2 : 4);
  p->kpl = k <= 32 ? 1 : k <= 64 ? 2 : 4;
  const int sms = sm_count();
  const int smem_cap = max_smem_optin() - 1024;  // static smem + slack
  p->qb = indirect ? 1 : pick_qb(n_queries);
  // keep (QB x KPL) register lists sane: wide lists only with narrow query blocks
  if (p->kpl == 4 && p->qb > 2) p->qb = 2;
  if (p->kpl == 2 && p->qb > 4) p->qb = 4;
  p->pipelined = (row_bytes % 16 == 0);
  if (p->pipelined) {
    p->q_bytes = (unsigned)(((size_t)p->qb * d * 4 + 127) / 128 * 128);
    const size_t extra = indirect ? ((size_t)(2 * nprobe + 1) * 4 + 127) / 128 * 128 : 0;
    // rows per warp per stage: a power of two, ~48 KB stages, but at least 2 (one-row stages starve the
    // consumers: measured 4.1 TB/s against 6.5 on B200)
    int rpw = 1;
    const size_t stage_budget = dtype == AURA_BF16 ? 65536 : 49152;   // bf16 rows: 4 rows per warp step want rpw % 4 == 0
    while (rpw * 2 <= 32 && (size_t)(rpw * 2) * SCAN_NW * row_bytes <= stage_budget) rpw *= 2;
    if (rpw < 2 && (size_t)2 * SCAN_NW * row_bytes * 2 + 8192 <= (size_t)smem_cap) rpw = 2;
    { static const int v = env_int("AURA_SCAN_RPW", 0); if (v) rpw = v; }   // tuning knob (experiments only)
    if (rpw < 1) rpw = 1;
    if (rpw > 32) rpw = 32;
    p->rows_per_stage = rpw * SCAN_NW;
    p->stage_bytes = (unsigned)(((size_t)p->rows_per_stage * row_bytes + 127) / 128 * 128);
    const size_t slot = (size_t)p->stage_bytes + (size_t)p->rows_per_stage * 12;
    const size_t fixed = 128 + (size_t)p->q_bytes + extra;
    if (fixed + 2 * slot > (size_t)smem_cap) p->pipelined = false;
    else {
      int s = (int)(((size_t)smem_cap - fixed) / slot);
      p->n_stage_bufs = s > SCAN_MAX_STAGES ? SCAN_MAX_STAGES : s;
      // ~150 KB in flight per SM saturates HBM; a deeper ring only took L1 away and measured slower (6.48 vs 6.75 TB/s)
      while (p->n_stage_bufs > 3 && (size_t)p->n_stage_bufs * slot > 163840) --p->n_stage_bufs;
      { static const int v = env_int("AURA_SCAN_STAGES", 0); if (v >= 2 && v <= s && v <= SCAN_MAX_STAGES) p->n_stage_bufs = v; }
      const size_t ring = (size_t)p->n_stage_bufs * slot;
      size_t mk = 1;  // largest power of two of keys that fits the ring, capped
      while (mk * 2 * 8 <= ring && mk * 2 <= (size_t)FINAL_MERGE_CAP) mk *= 2;
      if (mk < (size_t)SCAN_NW * 32 * p->kpl) p->pipelined = false;  // CTA merge needs NW*32*KPL keys
      p->merge_keys = (unsigned)mk;
      p->smem = fixed + ring;
      const long long n_stages = (expected_rows + p->rows_per_stage - 1) / p->rows_per_stage;
      // a CTA should own at least ~2 stages, otherwise launch/merge overhead dominates
      long long g = (n_stages + 1) / 2;
      p->grid = (int)(g < 1 ? 1 : g > sms ? sms : g);
      { static const int v = env_int("AURA_SCAN_GRID", 0); if (v >= 1 && v <= sms) p->grid = v; }
      p->n_qblocks = (n_queries + p->qb - 1) / p->qb;
    }
  }
  if (!p->pipelined) {
    p->qb = 1;
    p->n_qblocks = n_queries;
    const long long g = (expected_rows + 31) / 32;
    const long long gmax = (long long)sms * 4;
    p->grid = (int)(g < 1 ? 1 : g > gmax ? gmax : g);
    p->merge_keys = 8192;
    p->smem = (size_t)p->merge_keys * 8;
  }
}

template <bool BF16, int QB, bool INDIRECT>
static cudaError_t launch_pipelined(const ScanPlan& p, const ScanArgs& a, cudaStream_t st) {
  void (*kern)(ScanArgs) = nullptr;
  // bf16 rows carry half the bytes per element: 4 rows per step keep enough loads / FMAs in flight per warp
  const bool ru4 = BF16 && QB <= 2 && (p.rows_per_stage / SCAN_NW) % 4 == 0 && env_int("AURA_SCAN_RU2", 0) == 0;
  switch (p.kpl) {
    case 1: kern = ru4 ? scan_topk_kernel<BF16, QB, 1, SCAN_NW, (BF16 && QB <= 2 ? 4 : SCAN_RU), INDIRECT>
                       : scan_topk_kernel<BF16, QB, 1, SCAN_NW, SCAN_RU, INDIRECT>; break;
    case 2: kern = scan_topk_kernel<BF16, (QB > 4 ? 4 : QB), 2, SCAN_NW, SCAN_RU, INDIRECT>; break;
    default: kern = scan_topk_kernel<BF16, (QB > 2 ? 2 : QB), 4, SCAN_NW, SCAN_RU, INDIRECT>; break;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  if (e != cudaSuccess) return e;
  kern<<<dim3(p.grid, p.n_qblocks), 32 * (SCAN_NW + 1), p.smem, st>>>(a);
  return cudaGetLastError();
}

size_t scan_workspace_bytes(int n_queries, int k) {
  // upper bound over both plans: partial keys [n_queries rounded up to 8][grid<=4*SMs][k] + counters
  const size_t grid = (size_t)sm_count() * 4;
  const size_t nq = ((size_t)n_queries + 7) / 8 * 8;
  return nq * grid * (size_t)k * 8 + nq * 4 + 256;
}

// Shared launcher for aura_scan_topk (direct) and aura_ivf_search (indirect).
int launch_scan(const void* rows, int dtype, long long n_rows, int d, const float* queries, int n_queries,
                const float* scale, const float* bias, int k, long long row_base, long long* out_idx, float* out_score,
                void* workspace, const long long* probes, int nprobe, int n_lists, const int* list_offsets,
                const int* list_rows, long long expected_rows, cudaStream_t st, int empty_ok) {
  const bool indirect = probes != nullptr;
  ScanPlan p;
  make_plan(indirect ? expected_rows : n_rows, d, dtype, n_queries, k, indirect, nprobe, &p);
  if ((reinterpret_cast<uintptr_t>(rows) & 15) != 0 && p.pipelined) {  // bulk copies need 16-byte aligned rows
    p.pipelined = false; p.qb = 1; p.n_qblocks = n_queries; p.grid = sm_count() * 4; p.merge_keys = 8192;
    p.smem = (size_t)p.merge_keys * 8;
  }

  const size_t nq8 = ((size_t)n_queries + 7) / 8 * 8;
  unsigned* counters = reinterpret_cast<unsigned*>(workspace);
  u64* partial = reinterpret_cast<u64*>(reinterpret_cast<unsigned char*>(workspace) + (nq8 * 4 + 255) / 256 * 256);
  AURA_CUDA_OK(cudaMemsetAsync(counters, 0, nq8 * 4, st));

  ScanArgs a;
  a.rows = rows; a.n_rows = n_rows; a.d = d; a.queries = queries; a.n_queries = n_queries;
  a.scale = scale; a.bias = bias; a.k = k; a.row_base = row_base;
  a.out_idx = out_idx; a.out_score = out_score;
  a.partial = partial; a.counters = counters;
  a.rows_per_stage = p.rows_per_stage; a.n_stage_bufs = p.n_stage_bufs;
  a.stage_bytes = p.stage_bytes; a.q_bytes = p.q_bytes; a.merge_keys = p.merge_keys;
  a.terms_bulk = ((reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0;
  a.interleave = env_int("AURA_SCAN_INTERLEAVE", 0);
  a.empty_ok = empty_ok;
  a.probes = probes; a.nprobe = nprobe; a.n_lists = n_lists; a.list_offsets = list_offsets; a.list_rows = list_rows;

  cudaError_t e;
  const bool bf = dtype == AURA_BF16;
  if (p.pipelined && indirect) {
    e = bf ? launch_pipelined<true, 1, true>(p, a, st) : launch_pipelined<false, 1, true>(p, a, st);
  } else if (p.pipelined) {
    switch (p.qb) {
      case 1: e = bf ? launch_pipelined<true, 1, false>(p, a, st) : launch_pipelined<false, 1, false>(p, a, st); break;
      case 2: e = bf ? launch_pipelined<true, 2, false>(p, a, st) : launch_pipelined<false, 2, false>(p, a, st); break;
      case 4: e = bf ? launch_pipelined<true, 4, false>(p, a, st) : launch_pipelined<false, 4, false>(p, a, st); break;
      default: e = bf ? launch_pipelined<true, 8, false>(p, a, st) : launch_pipelined<false, 8, false>(p, a, st); break;
    }
  } else {
    void (*kern)(ScanArgs);
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
  note_launches(1);
  return AURA_OK;
}

}  // namespace aura

using namespace aura;

extern "C" size_t aura_scan_topk_workspace_bytes(int64_t n_rows, int d, int n_queries, int k) {
  (void)n_rows; (void)d;
  if (n_queries < 1 || k < 1) return 0;
  return scan_workspace_bytes(n_queries, k);
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
  AURA_REQUIRE(workspace && workspace_bytes >= scan_workspace_bytes(n_queries, k), AURA_ERR_WORKSPACE,
               "aura_scan_topk: workspace too small (%zu bytes)", workspace_bytes);
  return launch_scan(rows, dtype, n_rows, d, queries, n_queries, scale, bias, k, row_base,
                     reinterpret_cast<long long*>(out_idx), out_score, workspace, nullptr, 0, 0, nullptr, nullptr, n_rows,
                     (cudaStream_t)stream, 0);
}

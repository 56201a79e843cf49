// aura_ivf_search_batch - list-major fine stage of the centroid index for a block of queries
// (hippocampal.py:257-307 for B queries at once).
//
// The per-query path (ivf.cu + scan_topk.cu) streams nprobe lists per query: B * nprobe * |list| * d bytes.  At
// BASELINE config 4 (10M x 1024 fp32, 4096 lists, nprobe 32, B = 4096) that is 1.3 TB per batch, although every list is
// wanted by ~32 queries.  Here the (query, probe) pairs are counting-sorted BY LIST and every list is scored once
// against the group of queries that probe it - a ragged grouped GEMM on the tensor cores:
//   item  = (list c, tile of <= 128 of its queries, chunk of <= 2048 of its rows)
//   A     = the queries of the group, gathered by row id from the normalised query block
//   B     = the rows of the list chunk, gathered by row id straight from the bank (CSR order), so the inverted
//           lists stay an index (int32 row ids) and the bank keeps the reference's insertion order.
//           Gathers are 16-byte cp.async (LDGSTS) copies issued by 4 producer warps into the SWIZZLE_128B layout
//           the UMMA descriptors expect (8 lanes cover one 128-byte row slab; chunk j of row r lands at
//           r*128 + ((j ^ (r & 7)) << 4)); a stage's mbarrier receives each thread's arrival when its copies have
//           landed (cp.async.mbarrier.arrive.noinc), the MMA thread adds fence.proxy.async.  (TMA tile::gather4 gives
//           the same layout - scripts/exp/gather4_test.cu - but measured only ~5 GB/s per SM on B200: 57 ms per C4
//           batch against 6 ms of list bytes at the HBM roofline.)
//           Optional list-major mode (rows_by_list: a resident copy of the bank in CSR order made by
//           aura_ivf_pack_lists): the B tile is one TMA box per k-block instead of 2048 gathered pieces.
//   MMA / TMEM / epilogue as in gemm_topk.cu (tf32 from an fp32 bank, bf16 from a bf16 bank, per-row top-32 in
//   registers), one partial list per (pair, chunk)
// and a finish kernel merges a query's partial lists, re-scores the 32 best in exact fp32 and certifies the top-k
// (same rule as aura_batch_topk).  Algorithmic bytes per batch = bytes of the probed lists, each read once.
#include <stdlib.h>

#include "tc_common.cuh"

namespace aura {

static constexpr int IB_CH_TILES = 8;                       // column tiles per item
static constexpr int IB_CH_ROWS = IB_CH_TILES * GT_BN;      // 2048 list rows per item
static constexpr int IB_MERGE_CAP = 2048;                   // keys the finish kernel sorts at a time
static constexpr int IB_THREADS = 288;                      // warp 0 MMA, warps 1-4 epilogue, warps 5-8 gather producers

static constexpr int IB_MAX_CHUNKS = 16;                    // chunks per list: long lists get longer chunks, not more of them

// chunk geometry of a list of `len` rows: n_ch chunks of ch_rows rows (a multiple of the 256-row tile)
__host__ __device__ __forceinline__ void ib_chunks(int len, int& n_ch, int& ch_rows) {
  n_ch = (len + IB_CH_ROWS - 1) / IB_CH_ROWS;
  if (n_ch > IB_MAX_CHUNKS) n_ch = IB_MAX_CHUNKS;
  if (n_ch < 1) { n_ch = 0; ch_rows = IB_CH_ROWS; return; }
  ch_rows = ((len + n_ch - 1) / n_ch + GT_BN - 1) / GT_BN * GT_BN;
  n_ch = (len + ch_rows - 1) / ch_rows;
}

struct IvfBatchArgs {
  int n_lists, nprobe, k_blocks, n_stages, cap_items;
  const int* list_offsets; const int* list_rows;   // CSR of the bank
  const int* q_off;         // [n_lists + 1]  pairs per list, exclusive scan
  const int* pair_of_pos;   // [B * nprobe]   pair id (b * nprobe + p) at list-sorted position
  const int4* items;        // [cap_items]    {list, query tile, chunk, 0}
  const int* n_items;       // device scalar
  const float* scale; const float* bias;           // per bank row
  const int* pbase;         // [n_lists + 1]  first partial list of every list (exclusive scan of queries x chunks)
  int cap_plists;
  u64* partial;             // [cap_plists][GT_L]: one sorted list per (pair, chunk), at pbase[c] + chunk * nq_c + rel
  int use_gthr, spread;     // tuning switches (experiments)
  int list_major;           // the bank copy behind the kernel's tensor map holds every list contiguously (CSR order)
  int l2_hint;              // bit 0: single-tile lists evict-first, bit 1: multi-tile lists evict-last, bit 2: query rows evict-last
  int sync_polls;           // sibling CTAs of a cluster wait at most this many polls for each other per tile (0: never)
  const u64* ceil_keys;     // per query: only keys strictly below are eligible (multi-round top-k), may be null
  unsigned* gthr;           // [B] orderable lower bound of every query's final L-th best score (atomicMax)
};

// group position i (0..127) <-> A-tile row / TMEM lane: consecutive positions go to different epilogue warps, so a
// group of ~32 queries keeps all four warps busy instead of filling warp 0 only
__device__ __forceinline__ int ib_row_of_pos(int i, int spread) { return spread ? (i & 3) * 32 + (i >> 2) : i; }
__device__ __forceinline__ int ib_pos_of_row(int r, int spread) { return spread ? (r & 31) * 4 + (r >> 5) : r; }

// 16-byte global -> shared async copy; src_bytes = 0 writes zeros (K tail past the end of a row)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, unsigned src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// same with an L2 eviction policy
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, unsigned src_bytes, uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(dst), "l"(src), "r"(src_bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool TF32, bool CEIL>
__global__ void __launch_bounds__(IB_THREADS, 1)
ivf_gemm_kernel(const __grid_constant__ CUtensorMap tmap_lm, const unsigned char* __restrict__ qmat,
                const unsigned char* __restrict__ bank, const int row_pitch, const IvfBatchArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int S = a.n_stages;
  unsigned char* ring = smem;
  float2* sbuf = reinterpret_cast<float2*>(ring + (size_t)S * GT_STAGE_BYTES);   // [2][GT_BN] (scale, bias)
  int* rid_s = reinterpret_cast<int*>(sbuf + 2 * GT_BN);                         // [2][GT_BN] bank row of each column
  uint64_t* full = reinterpret_cast<uint64_t*>(rid_s + 2 * GT_BN);
  uint64_t* empty = full + GT_MAX_STAGES;
  uint64_t* tfull = empty + GT_MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint32_t* sync_ctr = tmem_slot + 1;      // tiles announced by the sibling CTAs of this cluster (monotonic)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cs = (int)tc::cluster_nctarank(), crank = (int)tc::cluster_ctarank();
  const int cid = blockIdx.x / cs, ncl = gridDim.x / cs;
  constexpr int ELEMS_PER_SLAB = TF32 ? 32 : 64;

  if (threadIdx.x == 0) {
    // full: one asynchronous arrival per producer thread, plus the expect-tx arrival of the TMA box in list-major mode
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], a.list_major ? 129 : 128); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
    *sync_ctr = 0u;
    if (a.list_major) tc::tma_prefetch_desc(&tmap_lm);
    fence_mbar_init();
  }
  if (warp == 0) { tc::tmem_alloc(tmem_slot, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();     // sibling counters are written remotely: every CTA of the cluster has initialised its own
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // a work table or partial-list area that does not fit processes nothing (the finish kernel hands every query back)
  const int n_items = (*a.n_items <= a.cap_items && a.pbase[a.n_lists] <= a.cap_plists) ? *a.n_items : 0;

  if (warp >= 5) {
    // ===================== gather producers: 128 threads, 16-byte cp.async into the swizzled stage =====================
    const int pt = threadIdx.x - 160;               // 0..127
    const int prow = pt >> 3, pj = pt & 7;          // this thread copies chunk pj of rows prow + 16*i
    const uint32_t dst_off = (uint32_t)(prow * GT_SLAB + ((pj ^ (prow & 7)) << 4));
    const uint32_t ring_u32 = smem_u32(ring);
    int stage = 0; unsigned phase = 0, sib_expected = 0u;
    // L2 residency: rows of a list probed by a single query tile are never needed again (evict first); rows of a list
    // with several query tiles are re-read by the sibling items and query rows by every tile (evict last)
    const uint64_t pol_stream = (a.l2_hint & 1) ? l2_policy_evict_first() : l2_policy_evict_normal();
    const uint64_t pol_keep = (a.l2_hint & 2) ? l2_policy_evict_last() : l2_policy_evict_normal();
    const uint64_t pol_a = (a.l2_hint & 4) ? l2_policy_evict_last() : l2_policy_evict_normal();
    for (int grp = cid; grp * cs < n_items; grp += ncl) {
      const int item = grp * cs + crank;
      if (item >= n_items) continue;
      const int4 it = a.items[item];
      if (it.x < 0) continue;          // padding slot
      const int qb = a.q_off[it.x], nq = a.q_off[it.x + 1] - qb;
      const int a0 = qb + it.y * GT_BM, n_a = min(GT_BM, nq - it.y * GT_BM);
      const int lb = a.list_offsets[it.x], len = a.list_offsets[it.x + 1] - lb;
      int n_ch_, ch_rows_;
      ib_chunks(len, n_ch_, ch_rows_);
      const int r0 = it.z * ch_rows_, r1 = min(len, r0 + ch_rows_);
      const unsigned char* qsrc[8];   // A-tile rows prow + 16*i (rows past the group re-load a valid query, masked later)
#pragma unroll
      for (int i = 0; i < 8; ++i)
        qsrc[i] = qmat + (size_t)(a.pair_of_pos[a0 + min(ib_pos_of_row(prow + 16 * i, a.spread), n_a - 1)] / a.nprobe) * row_pitch + pj * 16;
      // siblings: the other CTAs of this cluster whose item is another query tile of the same list chunk.  Each announces
      // every tile it starts to the others and waits (bounded - this is pacing, not correctness) until they have started
      // it too, so the chunk is fetched from HBM once and the other copies hit L2 while the lines are still resident.
      unsigned sib_mask = 0u; int n_sib = 0;
      for (int j = 0; j < cs; ++j)
        if (j != crank && grp * cs + j < n_items) {
          const int4 o = a.items[grp * cs + j];
          if (o.x == it.x && o.z == it.z) { sib_mask |= 1u << j; ++n_sib; }
        }
      bool in_step = n_sib > 0 && a.sync_polls > 0;
      const uint64_t pol_b = nq > GT_BM ? pol_keep : pol_stream;
      for (int cr = r0; cr < r1; cr += GT_BN) {
        if (n_sib > 0) {
          if (pt == 0)
            for (int j = 0; j < cs; ++j)
              if ((sib_mask >> j) & 1u) tc::cluster_red_inc(sync_ctr, j);
          sib_expected += (unsigned)n_sib;
          if (in_step) {
            int polls = 0;
            while ((int)(*reinterpret_cast<volatile uint32_t*>(sync_ctr) - sib_expected) < 0) {
              if (++polls > a.sync_polls) { in_step = false; break; }     // the sibling is far behind: stop waiting for it
              __nanosleep(64);
            }
          }
        }
        int rb[16];                   // bank rows of B-tile rows prow + 16*i (unused with the list-major copy)
#pragma unroll
        for (int i = 0; i < 16; ++i) rb[i] = a.list_major ? 0 : a.list_rows[lb + min(cr + prow + 16 * i, r1 - 1)];
        if (a.list_major) {
          // the bank has a resident copy with every list contiguous: the B tile is one TMA box per k-block (256
          // consecutive rows, streamed from HBM like the exact-search kernel); only the query rows are gathered
          for (int kb = 0; kb < a.k_blocks; ++kb) {
            tc::mbar_wait_guarded(&empty[stage], phase ^ 1u);
            unsigned char* sp = ring + (size_t)stage * GT_STAGE_BYTES;
            if (pt == 0) {
              mbar_arrive_expect_tx(&full[stage], (unsigned)(GT_STAGE_BYTES - GT_A_BYTES));
              tc::tma_load_2d(sp + GT_A_BYTES, &tmap_lm, kb * ELEMS_PER_SLAB, lb + cr, &full[stage], pol_b);
            }
            const unsigned sa0 = ring_u32 + (uint32_t)stage * GT_STAGE_BYTES + dst_off;
            const size_t koff = (size_t)kb * GT_SLAB;
            const bool in = (int)koff + pj * 16 < row_pitch;
            const size_t ko1 = in ? koff : 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) cp_async16(sa0 + i * 16 * GT_SLAB, qsrc[i] + ko1, in ? 16u : 0u, pol_a);
            cp_async_arrive_noinc(&full[stage]);
            if (++stage == S) { stage = 0; phase ^= 1u; }
          }
          continue;
        }
        // k-blocks go in PAIRS into two consecutive stages: the two 128-byte pieces of a row are adjacent in memory
        // and are requested back to back, so DRAM serves them from one open page (256 B per row visit, not 128 B)
        for (int kb = 0; kb < a.k_blocks; kb += 2) {
          const int nkb = min(2, a.k_blocks - kb);
          int st[2]; unsigned sa[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            st[h] = stage;
            if (h < nkb) {
              tc::mbar_wait_guarded(&empty[stage], phase ^ 1u);
              sa[h] = ring_u32 + (uint32_t)stage * GT_STAGE_BYTES + dst_off;
              if (++stage == S) { stage = 0; phase ^= 1u; }
            }
          }
          const size_t koff0 = (size_t)kb * GT_SLAB;
          bool in_row[2]; size_t ko[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            in_row[h] = h < nkb && (int)(koff0 + h * GT_SLAB) + pj * 16 < row_pitch;
            ko[h] = in_row[h] ? koff0 + h * GT_SLAB : 0;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h)
              if (h < nkb) cp_async16(sa[h] + i * 16 * GT_SLAB, qsrc[i] + ko[h], in_row[h] ? 16u : 0u, pol_a);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const unsigned char* src = bank + (size_t)rb[i] * row_pitch + pj * 16;
#pragma unroll
            for (int h = 0; h < 2; ++h)
              if (h < nkb) cp_async16(sa[h] + GT_A_BYTES + i * 16 * GT_SLAB, src + ko[h], in_row[h] ? 16u : 0u, pol_b);
          }
          // completion is signalled asynchronously: the mbarrier of each stage receives this thread's arrival when all
          // of its copies issued so far have landed (cp.async.mbarrier.arrive.noinc), so the thread never blocks on its
          // own loads and up to a full ring of stages stays in flight
#pragma unroll
          for (int h = 0; h < 2; ++h)
            if (h < nkb) cp_async_arrive_noinc(&full[st[h]]);
        }
      }
    }
    cp_async_wait<0>();
  } else if (warp == 0) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(TF32 ? 2 : 1, GT_BM, GT_BN);
      int stage = 0; unsigned phase = 0;
      unsigned tile_n = 0;
      for (int grp = cid; grp * cs < n_items; grp += ncl) {
        const int item = grp * cs + crank;
        if (item >= n_items) continue;
        const int4 it = a.items[item];
        if (it.x < 0) continue;        // padding slot
        const int len = a.list_offsets[it.x + 1] - a.list_offsets[it.x];
        int n_ch_, ch_rows_;
      ib_chunks(len, n_ch_, ch_rows_);
      const int r0 = it.z * ch_rows_, r1 = min(len, r0 + ch_rows_);
        for (int cr = r0; cr < r1; cr += GT_BN, ++tile_n) {
          const unsigned acc = tile_n & 1u, acc_phase = (tile_n >> 1) & 1u;
          tc::mbar_wait_guarded(&tempty[acc], acc_phase ^ 1u);
          tc::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * GT_BN;
          for (int kb = 0; kb < a.k_blocks; ++kb) {
            tc::mbar_wait_guarded(&full[stage], phase);
            fence_proxy_async();                      // cp.async wrote the stage through the generic proxy; the MMA reads it through the async proxy
            tc::tc_fence_after();
            const unsigned char* sa = ring + (size_t)stage * GT_STAGE_BYTES;
            const uint64_t da = tc::make_smem_desc_sw128(sa);
            const uint64_t db = tc::make_smem_desc_sw128(sa + GT_A_BYTES);
#pragma unroll
            for (int j = 0; j < GT_SLAB / 32; ++j)
              tc::umma<TF32>(d_tmem, da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), idesc, (kb | j) != 0 ? 1u : 0u);
            tc::umma_commit(&empty[stage]);
            if (++stage == S) { stage = 0; phase ^= 1u; }
          }
          tc::umma_commit(&tfull[acc]);
        }
      }
    }
  } else {
    // ===================== epilogue: thread = one query of the group =====================
    const int quarter = warp & 3;
    const int te = quarter * 32 + lane;
    const int et = threadIdx.x - 32;
    unsigned tile_n = 0;
    for (int grp = cid; grp * cs < n_items; grp += ncl) {
      const int item = grp * cs + crank;
      if (item >= n_items) continue;
      const int4 it = a.items[item];
      if (it.x < 0) continue;          // padding slot
      const int nq = a.q_off[it.x + 1] - a.q_off[it.x];
      const int n_a = min(GT_BM, nq - it.y * GT_BM);
      const int lb = a.list_offsets[it.x], len = a.list_offsets[it.x + 1] - lb;
      int n_ch_, ch_rows_;
      ib_chunks(len, n_ch_, ch_rows_);
      const int r0 = it.z * ch_rows_, r1 = min(len, r0 + ch_rows_);
      u64 e[GT_L];
#pragma unroll
      for (int s = 0; s < GT_L; ++s) e[s] = 0ull;
      const int my_pos = ib_pos_of_row(te, a.spread);
      const bool live = my_pos < n_a;
      float thr = live ? -INFINITY : INFINITY;
      const int my_query = live ? a.pair_of_pos[a.q_off[it.x] + it.y * GT_BM + my_pos] / a.nprobe : 0;
      unsigned* my_gthr = a.gthr + my_query;
      unsigned published = 0u;
      const u64 ceil_key = (CEIL && live) ? a.ceil_keys[my_query] : ~0ull;
      const float ceil_score = ceil_key == ~0ull ? INFINITY : key_score(ceil_key);
      if (ceil_key == 0ull) thr = INFINITY;
      for (int cr = r0; cr < r1; cr += GT_BN, ++tile_n) {
        const unsigned acc = tile_n & 1u, acc_phase = (tile_n >> 1) & 1u;
        if (live && a.use_gthr) {   // other CTAs scoring other lists of this query may already have raised the bar
          const unsigned g = *reinterpret_cast<volatile unsigned*>(my_gthr);
          if (g != 0u) thr = fmaxf(thr, f32_from_orderable(g));
        }
        float2* sb = sbuf + acc * GT_BN;
        int* rs = rid_s + acc * GT_BN;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = et + h * 128;
          float2 t;
          int rid = 0;
          if (cr + c < r1) {
            rid = a.list_rows[lb + cr + c];
            t.x = a.scale ? a.scale[rid] : 1.f;
            t.y = a.bias ? a.bias[rid] : 0.f;
          } else { t.x = 0.f; t.y = __int_as_float(0x7fc00000); }
          sb[c] = t;
          rs[c] = rid;
        }
        tc::named_bar_sync(1, 128);
        tc::mbar_wait_guarded(&tfull[acc], acc_phase);
        tc::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * GT_BN;
#pragma unroll 1
        for (int c0 = 0; c0 < GT_BN; c0 += 32) {
          float v[32];
          tc::tmem_ld_32x32(taddr + c0, v);
          unsigned mask = 0u;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float2 t = sb[c0 + j];
            const float sj = fmaf(v[j], t.x, t.y);
            mask |= (sj >= thr && (!CEIL || sj <= ceil_score)) ? (1u << j) : 0u;
          }
          while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1u;
            const float2 t = sb[c0 + j];
            const u64 key = make_key(fmaf(select32(v, j), t.x, t.y), (unsigned)rs[c0 + j]);
            if (key > e[GT_L - 1] && (!CEIL || key < ceil_key)) {
              list_insert_sorted<GT_L>(e, key);
              if (e[GT_L - 1] != 0ull) thr = fmaxf(thr, key_score(e[GT_L - 1]));
            }
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        if (live && a.use_gthr && e[GT_L - 1] != 0ull) {   // a full list: its L-th best bounds the query's final L-th best from below
          const unsigned o = (unsigned)(e[GT_L - 1] >> 32);
          if (o > published) { atomicMax(my_gthr, o); published = o; }
        }
      }
      if (live) {
        u64* dst = a.partial + ((size_t)a.pbase[it.x] + (size_t)it.z * nq + it.y * GT_BM + my_pos) * GT_L;
#pragma unroll
        for (int s = 0; s < GT_L; ++s) dst[s] = e[s];
      }
    }
  }
  __syncthreads();
  tc::cluster_sync_all();     // no CTA leaves while a sibling may still write its counter
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

// ---- work-table construction (all on device, no host sync) ----------------------------------------------
__global__ void __launch_bounds__(256) ib_pair_hist_kernel(const long long* __restrict__ probes, int n_pairs, int n_lists,
                                                           int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  const long long c = probes[i];
  if (c >= 0 && c < n_lists) atomicAdd(&counts[c], 1);
}
__global__ void __launch_bounds__(256) ib_pair_scatter_kernel(const long long* __restrict__ probes, int n_pairs, int n_lists,
                                                              int* __restrict__ cursor, int* __restrict__ pair_of_pos,
                                                              int* __restrict__ pos_of_pair) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  const long long c = probes[i];
  int pos = -1;
  if (c >= 0 && c < n_lists) { pos = atomicAdd(&cursor[c], 1); pair_of_pos[pos] = i; }
  pos_of_pair[i] = pos;
}
__global__ void __launch_bounds__(256) ib_item_count_kernel(const int* __restrict__ q_off, const int* __restrict__ list_offsets,
                                                            int n_lists, int cs, int* __restrict__ items_c, int* __restrict__ plists_c) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_lists) return;
  const int nq = q_off[c + 1] - q_off[c], len = list_offsets[c + 1] - list_offsets[c];
  int n_ch, ch_rows;
  ib_chunks(len, n_ch, ch_rows);
  // two sections: lists probed by more than one query tile first, their tile count padded to the cluster size so
  // that the tiles of one chunk ("siblings") fall into one aligned group of cs items; single-tile lists after them
  const int n_qt = (nq + GT_BM - 1) / GT_BM;
  const bool heavy = n_qt > 1;
  items_c[c] = heavy ? (n_qt + cs - 1) / cs * cs * n_ch : 0;
  items_c[n_lists + c] = heavy ? 0 : n_qt * n_ch;
  plists_c[c] = nq * n_ch;
}
__global__ void __launch_bounds__(256) ib_item_fill_kernel(const int* __restrict__ item_base, const int* __restrict__ q_off,
                                                           const int* __restrict__ list_offsets, int n_lists, int cap,
                                                           int4* __restrict__ items, int* __restrict__ n_items, int chunk_major) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= n_lists) return;
  int base = item_base[c], cnt = item_base[c + 1] - base;
  if (cnt == 0) { base = item_base[n_lists + c]; cnt = item_base[n_lists + c + 1] - base; }
  int n_ch, ch_rows;
  ib_chunks(list_offsets[c + 1] - list_offsets[c], n_ch, ch_rows);
  const int n_qt = (q_off[c + 1] - q_off[c] + GT_BM - 1) / GT_BM;
  const int n_qt_pad = n_ch > 0 ? cnt / n_ch : 1;     // > n_qt in the padded section: the extra slots are skipped (x = -1)
  for (int i = lane; i < cnt; i += 32)
    if (base + i < cap) {  // chunk-major: the query tiles of one chunk are adjacent and run at the same time
      const int qt = chunk_major ? i % n_qt_pad : i / n_ch, ch = chunk_major ? i / n_qt_pad : i % n_ch;
      items[base + i] = make_int4(qt < n_qt ? c : -1, qt, ch, 0);
    }
  if (c == n_lists - 1 && lane == 0) *n_items = item_base[2 * n_lists];
}

// ---- finish: merge the partial lists of one query, exact re-score, certify -----------------------------
struct IvfFinishArgs {
  const long long* probes; const int* pos_of_pair; const int* q_off; const int* item_base; const int* list_offsets;
  const int* n_items; int cap_items, nprobe, n_lists;
  const int* pbase; int cap_plists;
  const u64* partial;
  const void* rows; int bf16; int d; const float* qn; const float* scale; const float* bias; float eps;
  int k; long long row_base; int spread, chunk_major;
  u64* cand; u64* ceil_out; int round; int* force_flag;   // multi-round mode (see gemm_topk.cu)
  int empty_ok;             // a query without candidates is a valid empty result (row-sharded callers), not a hand-back
  long long* out_idx; float* out_score; int* uncertain;
};

__global__ void __launch_bounds__(128) ivf_finish_kernel(const IvfFinishArgs f) {
  __shared__ u64 keys[IB_MERGE_CAP];
  __shared__ u64 ex[GT_MAX_L];
  const int b = blockIdx.x;
  const bool overflow = *f.n_items > f.cap_items || f.pbase[f.n_lists] > f.cap_plists;
  // Partial lists of this query: one per (probe, chunk of the probed list), each sorted and zero-padded.  Two passes, a
  // warp per probe so that the dependent table reads of different probes overlap:
  //   1. the list HEADS; the GT_L-th largest head is a floor - GT_L keys (those heads) are at or above it, so nothing
  //      below it can be among the best GT_L;
  //   2. every list's keys >= floor (at most GT_L lists x GT_L keys) are buffered and sorted.
  __shared__ const u64* lptr[AURA_MAX_NPROBE];   // first partial list of probe p
  __shared__ int lstride[AURA_MAX_NPROBE];       // distance between the chunks' lists, in keys
  __shared__ int lcnt[AURA_MAX_NPROBE];          // chunks (0: probe has no list)
  __shared__ int n_heads, n_kept;
  __shared__ u64 floor_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { n_heads = 0; n_kept = 0; }
  for (int p = threadIdx.x; p < f.nprobe; p += blockDim.x) {
    int cnt = 0, stride = 0;
    const u64* ptr = nullptr;
    const long long c = overflow ? -1 : f.probes[(size_t)b * f.nprobe + p];
    if (c >= 0 && c < f.n_lists) {
      const int len = f.list_offsets[c + 1] - f.list_offsets[c];
      if (len > 0) {
        int n_ch, ch_rows;
        ib_chunks(len, n_ch, ch_rows);
        const int rel = f.pos_of_pair[(size_t)b * f.nprobe + p] - f.q_off[c];
        const int nq = f.q_off[c + 1] - f.q_off[c];
        cnt = n_ch; stride = nq * GT_L;
        ptr = f.partial + ((size_t)f.pbase[c] + rel) * GT_L;
      }
    }
    lptr[p] = ptr; lstride[p] = stride; lcnt[p] = cnt;
  }
  __syncthreads();
  for (int p = warp; p < f.nprobe; p += 4)
    for (int j = lane; j < lcnt[p]; j += 32) {
      const u64 head = lptr[p][(size_t)j * lstride[p]];
      if (head != 0ull) keys[atomicAdd(&n_heads, 1)] = head;     // <= nprobe * IB_MAX_CHUNKS = IB_MERGE_CAP heads
    }
  __syncthreads();
  const int nh = n_heads;
  const int h2 = max(64, next_pow2(nh));
  for (int i = nh + threadIdx.x; i < h2; i += blockDim.x) keys[i] = 0ull;
  block_bitonic_sort_desc(keys, h2);
  if (threadIdx.x == 0) floor_s = nh >= GT_L ? keys[GT_L - 1] : 0ull;
  __syncthreads();
  const u64 floor_key = floor_s;
  __syncthreads();                                               // heads consumed: the buffer now takes the survivors
  for (int p = warp; p < f.nprobe; p += 4)
    for (int j = 0; j < lcnt[p]; ++j) {
      const u64 key = lptr[p][(size_t)j * lstride[p] + lane];
      const unsigned m = __ballot_sync(0xffffffffu, key != 0ull && key >= floor_key);
      if (m == 0u) continue;
      int base = 0;
      if (lane == 0) base = atomicAdd(&n_kept, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if ((m >> lane) & 1u) keys[base + __popc(m & ((1u << lane) - 1u))] = key;
    }
  __syncthreads();
  const int n = n_kept;
  const int n2 = max(64, next_pow2(n));       // sort only what was buffered
  for (int i = n + threadIdx.x; i < n2; i += blockDim.x) keys[i] = 0ull;
  block_bitonic_sort_desc(keys, n2);
  if (f.cand != nullptr) {
    for (int i = threadIdx.x; i < GT_L; i += blockDim.x) f.cand[(size_t)b * GT_MAX_L + f.round * GT_L + i] = keys[i];
    if (threadIdx.x == 0) {
      f.ceil_out[b] = keys[GT_L - 1];
      if (f.round == 0) f.force_flag[b] = (overflow || (keys[0] == 0ull && !f.empty_ok)) ? 1 : 0;
    }
    return;
  }
  RescoreArgs ra;
  ra.rows = f.rows; ra.bf16 = f.bf16; ra.d = f.d; ra.q = f.qn + (size_t)b * f.d; ra.scale = f.scale; ra.bias = f.bias;
  ra.eps = f.eps; ra.k = f.k; ra.L = GT_L; ra.row_base = f.row_base;
  ra.out_idx = f.out_idx + (size_t)b * f.k; ra.out_score = f.out_score + (size_t)b * f.k;
  ra.uncertain = f.uncertain + b;
  rescore_and_write(keys, n2, ex, ra);
  // no candidate at all (every probed list empty -> the reference scans all rows, hippocampal.py:269-270) or a
  // work table that did not fit: hand the query back to the per-query path
  if (threadIdx.x == 0 && (overflow || (keys[0] == 0ull && !f.empty_ok))) f.uncertain[b] = 1;
}

static size_t a256(size_t v) { return (v + 255) / 256 * 256; }
static int ib_cap_items(int n_queries, int nprobe, int n_lists) {
  const long long pairs = (long long)n_queries * nprobe;
  const long long heavy = pairs / 128 < n_lists ? pairs / 128 : n_lists;       // lists that can hold more than one query tile
  long long cap = 16ll * n_lists + pairs / 8 + 48 * heavy + 4096;               // 16 B per item; 48 = padding to a cluster of 4
  if (cap > 2000000) cap = 2000000;
  return (int)cap;
}
// partial lists (256 B each): one per (pair, chunk of its list)
static int ib_cap_plists(int n_queries, int nprobe) {
  const long long cap = (long long)IB_MAX_CHUNKS * n_queries * nprobe + 4096;   // cannot overflow: <= IB_MAX_CHUNKS per pair
  return (int)(cap > 64000000 ? 64000000 : cap);
}

struct IbLayout {
  size_t probes, counts, q_off, cursor, pair_of_pos, pos_of_pair, items_c, item_base, n_items, items, qn, qb, partial, coarse, gthr, cand, ceil, force, plists_c, pbase, total;
};
static IbLayout ib_layout(int n_queries, int d, int n_lists, int nprobe, int cap, size_t coarse_bytes) {
  IbLayout L;
  size_t o = 0;
  const size_t pairs = (size_t)n_queries * nprobe;
  L.probes = o; o += a256(pairs * 8);
  L.counts = o; o += a256((size_t)n_lists * 4);
  L.q_off = o; o += a256((size_t)(n_lists + 1) * 4);
  L.cursor = o; o += a256((size_t)n_lists * 8);
  L.pair_of_pos = o; o += a256(pairs * 4);
  L.pos_of_pair = o; o += a256(pairs * 4);
  L.items_c = o; o += a256((size_t)n_lists * 8);
  L.item_base = o; o += a256((size_t)(2 * n_lists + 1) * 4);
  L.n_items = o; o += 256;
  L.items = o; o += a256((size_t)cap * 16);
  L.qn = o; o += a256((size_t)n_queries * d * 4);
  L.qb = o; o += a256((size_t)n_queries * d * 2);
  L.partial = o; o += a256((size_t)ib_cap_plists(n_queries, nprobe) * GT_L * 8);
  L.plists_c = o; o += a256((size_t)n_lists * 4);
  L.pbase = o; o += a256((size_t)(n_lists + 1) * 4);
  L.coarse = o; o += a256(coarse_bytes);
  L.gthr = o; o += a256((size_t)n_queries * 4);
  L.cand = o; o += a256((size_t)n_queries * GT_MAX_L * 8);
  L.ceil = o; o += a256((size_t)n_queries * 8);
  L.force = o; o += a256((size_t)n_queries * 4);
  L.total = o;
  return L;
}

// ivf.cu
size_t ivf_coarse_ws_bytes(int n_queries, int d, int n_cent, int nprobe);
int ivf_run_coarse(const float* queries, int n_queries, int d, const float* centroids, int n_cent, int nprobe,
                   long long* probes, void* workspace, cudaStream_t st);
// gemm_topk.cu
int launch_cand_rescore(const u64* cand, int n_cand, const void* rows, int bf16, int d, const float* qn, const float* scale,
                        const float* bias, float eps, int k, long long row_base, long long* out_idx, float* out_score,
                        int* uncertain, int n_queries, cudaStream_t st, const int* force_flag);
void launch_normalize_queries(const float* q, int n, int d, float* qn, __nv_bfloat16* qb, cudaStream_t st);

}  // namespace aura
using namespace aura;

extern "C" size_t aura_ivf_search_batch_workspace_bytes(int n_queries, int d, int n_centroid_rows, int nprobe) {
  if (n_queries < 1 || d < 1 || n_centroid_rows < 1 || nprobe < 1) return 0;
  const int cap = ib_cap_items(n_queries, nprobe, n_centroid_rows);
  return ib_layout(n_queries, d, n_centroid_rows, nprobe, cap, ivf_coarse_ws_bytes(n_queries, d, n_centroid_rows, nprobe)).total;
}

extern "C" int aura_ivf_search_batch(const void* rows, int dtype, int64_t n_rows, int d, const float* queries, int n_queries,
                                     const float* centroids, int n_centroid_rows, int nprobe, const int32_t* list_offsets,
                                     const int32_t* list_rows, const void* rows_by_list, const float* scale, const float* bias,
                                     int k, int64_t row_base, int flags, float eps, int64_t* out_idx, float* out_score,
                                     int32_t* out_uncertain, void* workspace, size_t workspace_bytes, void* stream) {
  AURA_REQUIRE(dtype == AURA_F32 || dtype == AURA_BF16, AURA_ERR_INVALID_ARG, "aura_ivf_search_batch: bad dtype %d", dtype);
  AURA_REQUIRE(n_rows >= 1 && n_rows < 0x7fffffffll && d >= 1 && n_queries >= 1 && n_centroid_rows >= 1, AURA_ERR_INVALID_ARG,
               "aura_ivf_search_batch: n_rows=%lld d=%d n_queries=%d n_centroid_rows=%d", (long long)n_rows, d, n_queries,
               n_centroid_rows);
  AURA_REQUIRE(nprobe >= 1 && nprobe <= AURA_MAX_NPROBE && nprobe <= n_centroid_rows, AURA_ERR_INVALID_ARG,
               "aura_ivf_search_batch: nprobe=%d", nprobe);
  AURA_REQUIRE(k >= 1 && k + 14 <= GT_MAX_L, AURA_ERR_UNSUPPORTED, "aura_ivf_search_batch: k=%d too large (max %d)", k, GT_MAX_L - 14);
  const bool bf16 = dtype == AURA_BF16;
  const int eb = bf16 ? 2 : 4;
  AURA_REQUIRE(((size_t)d * eb) % 16 == 0 && (d % 4) == 0 && (reinterpret_cast<uintptr_t>(rows) & 15) == 0, AURA_ERR_UNSUPPORTED,
               "aura_ivf_search_batch: rows must be 16-byte aligned with a 16-byte multiple pitch (d=%d)", d);
  AURA_REQUIRE(rows && queries && centroids && list_offsets && list_rows && out_idx && out_score && out_uncertain && workspace,
               AURA_ERR_INVALID_ARG, "aura_ivf_search_batch: null pointer");
  AURA_REQUIRE(workspace_bytes >= aura_ivf_search_batch_workspace_bytes(n_queries, d, n_centroid_rows, nprobe),
               AURA_ERR_WORKSPACE, "aura_ivf_search_batch: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int cap = ib_cap_items(n_queries, nprobe, n_centroid_rows);
  const IbLayout L = ib_layout(n_queries, d, n_centroid_rows, nprobe, cap, ivf_coarse_ws_bytes(n_queries, d, n_centroid_rows, nprobe));
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  long long* probes = reinterpret_cast<long long*>(ws + L.probes);
  int* counts = reinterpret_cast<int*>(ws + L.counts);
  int* q_off = reinterpret_cast<int*>(ws + L.q_off);
  int* cursor = reinterpret_cast<int*>(ws + L.cursor);
  int* pair_of_pos = reinterpret_cast<int*>(ws + L.pair_of_pos);
  int* pos_of_pair = reinterpret_cast<int*>(ws + L.pos_of_pair);
  int* items_c = reinterpret_cast<int*>(ws + L.items_c);
  int* item_base = reinterpret_cast<int*>(ws + L.item_base);
  int* n_items = reinterpret_cast<int*>(ws + L.n_items);
  int4* items = reinterpret_cast<int4*>(ws + L.items);
  float* qn = reinterpret_cast<float*>(ws + L.qn);
  __nv_bfloat16* qb = bf16 ? reinterpret_cast<__nv_bfloat16*>(ws + L.qb) : nullptr;
  u64* partial = reinterpret_cast<u64*>(ws + L.partial);
  unsigned* gthr = reinterpret_cast<unsigned*>(ws + L.gthr);
  AURA_CUDA_OK(cudaMemsetAsync(gthr, 0, (size_t)n_queries * 4, st));

  int rc = ivf_run_coarse(queries, n_queries, d, centroids, n_centroid_rows, nprobe, probes, ws + L.coarse, st);
  if (rc != AURA_OK) return rc;
  launch_normalize_queries(queries, n_queries, d, qn, qb, st);
  int chunk_major = 1;
  // AURA_IVF_CLUSTER=2|4 launches the CTAs in clusters: the query tiles of one list chunk go to the CTAs of one cluster,
  // which pace each other tile by tile so the chunk is fetched from HBM once.  Measured at BASELINE config 4: within 2 %
  // of the plain launch (the few lists probed by thousands of queries have 32 sibling tiles, not 2-4), so it is off.
  static const int env_cs = env_int("AURA_IVF_CLUSTER", 1), env_order = env_int("AURA_IVF_ORDER", 1);
  static const int env_gthr = env_int("AURA_IVF_GTHR", 1), env_spread = env_int("AURA_IVF_SPREAD", 1);
  static const int env_l2 = env_int("AURA_IVF_L2HINT", 3), env_polls = env_int("AURA_IVF_SYNC_POLLS", 512);
  static const int env_grid = env_int("AURA_IVF_GRID", 0);
  int cs = env_cs;
  if (cs != 1 && cs != 2 && cs != 4) cs = 1;
  chunk_major = env_order;
  const int n_pairs = n_queries * nprobe;
  AURA_CUDA_OK(cudaMemsetAsync(counts, 0, (size_t)n_centroid_rows * 4, st));
  ib_pair_hist_kernel<<<(n_pairs + 255) / 256, 256, 0, st>>>(probes, n_pairs, n_centroid_rows, counts);
  launch_scan_offsets(counts, n_centroid_rows, q_off, cursor, st);
  ib_pair_scatter_kernel<<<(n_pairs + 255) / 256, 256, 0, st>>>(probes, n_pairs, n_centroid_rows, cursor, pair_of_pos, pos_of_pair);
  int* plists_c = reinterpret_cast<int*>(ws + L.plists_c);
  int* pbase = reinterpret_cast<int*>(ws + L.pbase);
  ib_item_count_kernel<<<(n_centroid_rows + 255) / 256, 256, 0, st>>>(q_off, list_offsets, n_centroid_rows, cs, items_c, plists_c);
  launch_scan_offsets(items_c, 2 * n_centroid_rows, item_base, cursor, st);
  launch_scan_offsets(plists_c, n_centroid_rows, pbase, cursor, st);
  ib_item_fill_kernel<<<(n_centroid_rows * 32 + 255) / 256, 256, 0, st>>>(item_base, q_off, list_offsets, n_centroid_rows, cap, items, n_items, chunk_major);
  note_launches(4);

  IvfBatchArgs a;
  a.n_lists = n_centroid_rows; a.nprobe = nprobe; a.cap_items = cap;
  const int elems = GT_SLAB / eb;
  a.k_blocks = (d + elems - 1) / elems;
  const size_t fixed = 2 * GT_BN * 8 + 2 * GT_BN * 4 + (2 * GT_MAX_STAGES + 4) * 8 + 16;
  int stages = (int)(((size_t)max_smem_optin() - 1024 - fixed) / GT_STAGE_BYTES);
  if (stages > GT_MAX_STAGES) stages = GT_MAX_STAGES;
  AURA_REQUIRE(stages >= 4, AURA_ERR_UNSUPPORTED, "aura_ivf_search_batch: needs 4 pipeline stages of shared memory");
  a.n_stages = stages;
  a.list_offsets = list_offsets; a.list_rows = list_rows; a.q_off = q_off; a.pair_of_pos = pair_of_pos;
  a.items = items; a.n_items = n_items; a.scale = scale; a.bias = bias; a.partial = partial; a.gthr = gthr; a.pbase = pbase; a.cap_plists = ib_cap_plists(n_queries, nprobe);
  a.use_gthr = env_gthr; a.spread = env_spread; a.l2_hint = env_l2; a.sync_polls = env_polls;
  const size_t smem = (size_t)stages * GT_STAGE_BYTES + fixed + 1024;
  CUtensorMap tmap_lm;
  memset(&tmap_lm, 0, sizeof(tmap_lm));
  a.list_major = 0;
  if (rows_by_list != nullptr) {
    AURA_REQUIRE((reinterpret_cast<uintptr_t>(rows_by_list) & 15) == 0, AURA_ERR_UNSUPPORTED, "aura_ivf_search_batch: rows_by_list must be 16-byte aligned");
    const int trc = encode_tmap_2d(&tmap_lm, rows_by_list, eb, bf16, n_rows, d, GT_BN);
    if (trc != AURA_OK) return trc;
    a.list_major = 1;
  }
  typedef void (*IvfKern)(const CUtensorMap, const unsigned char*, const unsigned char*, int, const IvfBatchArgs);
  IvfKern kern0 = bf16 ? ivf_gemm_kernel<false, false> : ivf_gemm_kernel<true, false>;    // first round: no ceiling
  IvfKern kern1 = bf16 ? ivf_gemm_kernel<false, true> : ivf_gemm_kernel<true, true>;
  AURA_CUDA_OK(cudaFuncSetAttribute(kern0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  AURA_CUDA_OK(cudaFuncSetAttribute(kern1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  u64* cand = reinterpret_cast<u64*>(ws + L.cand);
  u64* ceil_buf = reinterpret_cast<u64*>(ws + L.ceil);
  int* force = reinterpret_cast<int*>(ws + L.force);
  const int rounds = (k + 14 + GT_L - 1) / GT_L;      // rounds of 32 candidates (see aura_batch_topk)

  IvfFinishArgs f;
  f.probes = probes; f.pos_of_pair = pos_of_pair; f.q_off = q_off; f.item_base = item_base; f.list_offsets = list_offsets;
  f.n_items = n_items; f.cap_items = cap; f.nprobe = nprobe; f.n_lists = n_centroid_rows; f.partial = partial;
  f.pbase = pbase; f.cap_plists = a.cap_plists;
  f.rows = rows; f.bf16 = bf16 ? 1 : 0; f.d = d; f.qn = qn; f.scale = scale; f.bias = bias; f.eps = eps; f.k = k;
  f.row_base = row_base; f.spread = a.spread; f.chunk_major = chunk_major; f.out_idx = reinterpret_cast<long long*>(out_idx); f.out_score = out_score; f.uncertain = out_uncertain;
  f.cand = rounds > 1 ? cand : nullptr; f.ceil_out = ceil_buf; f.round = 0; f.force_flag = force;
  f.empty_ok = (flags & AURA_IVF_EMPTY_OK) ? 1 : 0;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute cattr[1];
  cattr[0].id = cudaLaunchAttributeClusterDimension;
  cattr[0].val.clusterDim.x = cs; cattr[0].val.clusterDim.y = 1; cattr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3(sm_count() / cs * cs); cfg.blockDim = dim3(IB_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cfg.attrs = cattr; cfg.numAttrs = 1;
  if (cs > 1) {            // some GPCs cannot seat every cluster: size the persistent grid to what is co-resident
    int n_cl = 0;
    if (cudaOccupancyMaxActiveClusters(&n_cl, kern0, &cfg) == cudaSuccess && n_cl > 0 && n_cl * cs < (int)cfg.gridDim.x)
      cfg.gridDim = dim3(n_cl * cs);
    (void)cudaGetLastError();
  }
  if (env_grid >= cs) cfg.gridDim = dim3(env_grid / cs * cs);
  for (int r = 0; r < rounds; ++r) {
    if (r > 0) AURA_CUDA_OK(cudaMemsetAsync(gthr, 0, (size_t)n_queries * 4, st));
    a.ceil_keys = r ? ceil_buf : nullptr;
    AURA_CUDA_OK(cudaLaunchKernelEx(&cfg, r ? kern1 : kern0, tmap_lm, reinterpret_cast<const unsigned char*>(bf16 ? (const void*)qb : (const void*)qn),
                                    reinterpret_cast<const unsigned char*>(rows), d * eb, a));
    f.round = r;
    ivf_finish_kernel<<<n_queries, 128, 0, st>>>(f);
    AURA_CUDA_OK(cudaGetLastError());
    note_launches(2);
  }
  if (rounds > 1)
    return launch_cand_rescore(cand, rounds * GT_L, rows, bf16 ? 1 : 0, d, qn, scale, bias, eps, k, row_base,
                               reinterpret_cast<long long*>(out_idx), out_score, out_uncertain, n_queries, st, force);
  return AURA_OK;
}


/* Diagnostics: number of work items the last aura_ivf_search_batch call on this workspace generated, and the capacity
 * of its work table (a call whose items exceed the capacity hands every query back to the per-query path). */
// ---- list-major resident copy: out[p] = rows[list_rows[p]] (one warp per row, 16-byte pieces) ----
__global__ void __launch_bounds__(256) ivf_pack_lists_kernel(const uint4* __restrict__ rows, const int* __restrict__ list_rows,
                                                             long long n, int vec_per_row, uint4* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < n; p += (long long)gridDim.x * (blockDim.x >> 5)) {
    const uint4* src = rows + (size_t)list_rows[p] * vec_per_row;
    uint4* dst = out + (size_t)p * vec_per_row;
    for (int v = lane; v < vec_per_row; v += 32) __stcs(dst + v, __ldcs(src + v));
  }
}

extern "C" int aura_ivf_pack_lists(const void* rows, int dtype, int d, const int32_t* list_rows, int64_t n_listed,
                                   void* rows_by_list, void* stream) {
  AURA_REQUIRE(dtype == AURA_F32 || dtype == AURA_BF16, AURA_ERR_INVALID_ARG, "aura_ivf_pack_lists: bad dtype %d", dtype);
  AURA_REQUIRE(rows && list_rows && rows_by_list && d >= 1 && n_listed >= 0, AURA_ERR_INVALID_ARG, "aura_ivf_pack_lists: bad argument");
  const size_t pitch = (size_t)d * (dtype == AURA_BF16 ? 2 : 4);
  AURA_REQUIRE(pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(rows) & 15) == 0 && (reinterpret_cast<uintptr_t>(rows_by_list) & 15) == 0,
               AURA_ERR_UNSUPPORTED, "aura_ivf_pack_lists: rows must be 16-byte aligned with a 16-byte multiple pitch (d=%d)", d);
  if (n_listed == 0) return AURA_OK;
  ivf_pack_lists_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(rows), list_rows, (long long)n_listed,
                                                                             (int)(pitch / 16), reinterpret_cast<uint4*>(rows_by_list));
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}

extern "C" int aura_ivf_search_batch_items(const void* workspace, int n_queries, int d, int n_centroid_rows, int nprobe,
                                           int32_t* items_out, int32_t* host_cap, void* stream) {
  AURA_REQUIRE(workspace && items_out && host_cap, AURA_ERR_INVALID_ARG, "aura_ivf_search_batch_items: null pointer");
  const int cap = ib_cap_items(n_queries, nprobe, n_centroid_rows);
  const IbLayout L = ib_layout(n_queries, d, n_centroid_rows, nprobe, cap, ivf_coarse_ws_bytes(n_queries, d, n_centroid_rows, nprobe));
  AURA_CUDA_OK(cudaMemcpyAsync(items_out, reinterpret_cast<const unsigned char*>(workspace) + L.n_items, 4,
                               cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  *host_cap = cap;
  return AURA_OK;
}

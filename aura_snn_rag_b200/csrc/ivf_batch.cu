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

// L = entries of the per-(query, chunk) register list: GT_L, or GT_L_SMALL when k <= 10 (the list is the epilogue's cost);
// the partial lists in memory keep a stride of GT_L keys either way (unused entries are zero).
template <bool TF32, bool CEIL, int L>
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
      u64 e[L];
#pragma unroll
      for (int s = 0; s < L; ++s) e[s] = 0ull;
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
            if (key > e[L - 1] && (!CEIL || key < ceil_key)) {
              list_insert_sorted<L>(e, key);
              if (e[L - 1] != 0ull) thr = fmaxf(thr, key_score(e[L - 1]));
            }
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        if (live && a.use_gthr && e[L - 1] != 0ull) {   // a full list: its L-th best bounds the query's final L-th best from below
          const unsigned o = (unsigned)(e[L - 1] >> 32);
          if (o > published) { atomicMax(my_gthr, o); published = o; }
        }
      }
      if (live) {
        u64* dst = a.partial + ((size_t)a.pbase[it.x] + (size_t)it.z * nq + it.y * GT_BM + my_pos) * GT_L;
#pragma unroll
        for (int s = 0; s < GT_L; ++s) dst[s] = s < L ? e[s < L ? s : 0] : 0ull;
      }
    }
  }
  __syncthreads();
  tc::cluster_sync_all();     // no CTA leaves while a sibling may still write its counter
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

// ---- seeded start bound of the list-major fine stages ---------------------------------------------------
// The fine-stage kernels select behind a per-query threshold that only becomes useful once SOME item has collected L keys
// of the query; until then every item appends / inserts whatever it sees (C5 shard: 6.5 k appended keys per query for a
// shortlist of 128, and the appends - not the tensor pipe - were a fifth of the batch).  This kernel scores, in exact fp32,
// the first L rows of the query's nearest probed lists (probe order = nearest centroid first): any L candidate rows give a
// lower bound of the final L-th best score - min over them - and rows of the nearest list are the likeliest to sit near
// the top, so on clustered data the bound already rejects every row of the far lists.  The tensor-core score of a row
// differs from its exact score by at most the certification bound eps, hence gthr[b] = min - eps.
__global__ void __launch_bounds__(128) ivf_seed_gthr_kernel(const void* __restrict__ rows, int bf16, int d, const float* __restrict__ qn,
                                                            const long long* __restrict__ probes, int nprobe, int n_lists,
                                                            const int* __restrict__ list_offsets, const int* __restrict__ list_rows,
                                                            const float* __restrict__ scale, const float* __restrict__ bias,
                                                            float eps, const float* __restrict__ eps_q, int L,
                                                            unsigned* __restrict__ gthr) {
  __shared__ int s_row[GT_MAX_L];
  __shared__ int s_n;
  __shared__ float s_min[4];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {            // one warp walks the probes: positions [taken, taken + len) of the candidate list
    int taken = 0;
    for (int p = 0; p < nprobe && taken < L; ++p) {
      const long long c = probes[(size_t)b * nprobe + p];
      if (c < 0 || c >= n_lists) continue;
      const int lb = list_offsets[c], len = min(list_offsets[c + 1] - lb, L - taken);
      for (int i = lane; i < len; i += 32) s_row[taken + i] = list_rows[lb + i];
      taken += len;
    }
    if (lane == 0) s_n = taken;
  }
  __syncthreads();
  if (s_n < L) return;               // fewer than L candidates in all probed lists: no bound (every one of them is kept anyway)
  const float* q = qn + (size_t)b * d;
  float worst = INFINITY;
  for (int i = warp; i < L; i += 4) {
    const unsigned r = (unsigned)s_row[i];
    const float dot = warp_sum(exact_dot_partial(rows, bf16, d, r, q, lane));
    worst = fminf(worst, fmaf(dot, scale ? scale[r] : 1.f, bias ? bias[r] : 0.f));
  }
  if (lane == 0) s_min[warp] = worst;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float m = fminf(fminf(s_min[0], s_min[1]), fminf(s_min[2], s_min[3])) - (eps_q ? eps_q[b] : eps);
    // a non-finite bound (NaN rows) seeds nothing; 0 means "no bound" to the kernels, so a bound that maps to 0 is dropped too
    if (m == m && m > -INFINITY) gthr[b] = f32_orderable(m);
  }
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
  int L;                    // entries the GEMM kernel keeps per partial list (<= GT_L)
  const float* eps_q;       // per-query certification bound (bf16 list-major shadow of an fp32 bank), overrides eps
  u64* cand; u64* ceil_out; int round; int* force_flag;   // multi-round mode (see gemm_topk.cu)
  int empty_ok;             // a query without candidates is a valid empty result (row-sharded callers), not a hand-back
  const unsigned* gthr;     // final shared bounds of the GEMM pass (seed included), part of the completeness floor
  long long* out_idx; float* out_score; int* uncertain;
};

__global__ void __launch_bounds__(128) ivf_finish_kernel(const IvfFinishArgs f) {
  __shared__ u64 keys[IB_MERGE_CAP];
  __shared__ u64 ex[GT_DEEP];
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
  if (threadIdx.x == 0) floor_s = nh >= f.L ? keys[f.L - 1] : 0ull;
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
  // completeness floor of keys[] (second chance of rescore_and_write): rows absent from keys[] were rejected against a
  // list tail, the shared / seeded bound of the query, or the head floor above
  __shared__ unsigned s_floor;
  if (threadIdx.x == 0) {
    unsigned fl = (unsigned)(floor_key >> 32);
    if (f.gthr != nullptr) fl = max(fl, f.gthr[b]);
    s_floor = fl;
  }
  __syncthreads();
  {
    unsigned fl = 0u;
    for (int p = warp; p < f.nprobe; p += 4)
      for (int j = lane; j < lcnt[p]; j += 32) fl = max(fl, (unsigned)(lptr[p][(size_t)j * lstride[p] + (f.L - 1)] >> 32));
    if (fl != 0u) atomicMax(&s_floor, fl);
  }
  __syncthreads();
  RescoreArgs ra;
  ra.out_mul = 1.f;
  ra.rows = f.rows; ra.bf16 = f.bf16; ra.d = f.d; ra.q = f.qn + (size_t)b * f.d; ra.scale = f.scale; ra.bias = f.bias;
  ra.deep = 1; ra.floor_score = s_floor ? f32_from_orderable(s_floor) : -INFINITY;
  ra.eps = f.eps_q ? f.eps_q[b] : f.eps; ra.k = f.k; ra.L = f.L; ra.row_base = f.row_base;
  ra.out_idx = f.out_idx + (size_t)b * f.k; ra.out_score = f.out_score + (size_t)b * f.k;
  ra.uncertain = f.uncertain + b;
  rescore_and_write(keys, n2, ex, ra);
  // no candidate at all (every probed list empty -> the reference scans all rows, hippocampal.py:269-270) or a
  // work table that did not fit: hand the query back to the per-query path
  if (threadIdx.x == 0 && (overflow || (keys[0] == 0ull && !f.empty_ok))) f.uncertain[b] = 1;
}

// =====================================================================================================================
// List-major fine stage, rows-as-M formulation with ONE-PASS selection (round 2).
//
//   item  = (list c, group of <= NQ of the queries probing it, chunk of its rows)
//   A (M) = 128 bank rows of the list per tile (TMA box from the list-major copy, or gathered by row id with cp.async)
//   B (N) = the query group, N = 16..NQ columns (gathered from the normalised query block, 128 B x N per k-block)
//   D     = [128 rows x N queries] fp32 in TMEM, two accumulator stages; two k-blocks (256 B of every row) per pipeline
//           stage so that a row visit is 256 contiguous bytes in both modes.
// Against the queries-as-M kernel above: a list probed by 16-32 queries fills the M = 128 tile with rows instead of
// padding, the gathered operand shrinks from 16 KB to 2-4 KB per k-block, and the epilogue work per tile is
// proportional to the live queries.  Three launches cover the lists by group size (NQ = 32 / 64 / 128) so that the ring
// is as deep as the operands allow (5 x 40 KB, 4 x 48 KB, 3 x 64 KB).
//
// Selection (any k <= 114 in one pass).  Epilogue thread = bank row; for every query column it compares the score with
// the query's threshold thr[q] (shared memory) and appends survivors (64-bit ranking keys) to the query's candidate
// buffer of this CTA (global memory, L2-resident; counters in shared memory).  After every tile, one warp per query
// - raises thr[q] to the bound other CTAs published for the query (gthr, atomicMax) and
// - compacts a buffer that grew past Lc keys: exact selection of its L best keys by bisection on the key bits
//   (registers only), thr[q] = L-th best score, which is also published to gthr.
// A buffer holds <= Lc keys when a tile starts and a tile appends <= 128, so capacity Lc + 128 never overflows.  When
// the item ends, its (at most L) keys at or above gthr are written to a result slot that the query links to; the finish
// kernel gathers the slots of a query, takes the L best approximate keys - exactly the set the multi-round scheme
// enumerated, because thresholds are always scores of an L-th best key of a SUBSET of the candidates and ties pass -
// re-scores them in exact fp32 and certifies.
// =====================================================================================================================
static constexpr int IR_THREADS = 288;       // warp 0 MMA, warps 1-4 epilogue, warps 5-8 producers
static constexpr int IR_BM = 128;            // list rows per tile (TMEM lanes)
static constexpr int IR_MAX_STAGES = 6;
static constexpr int IR_MAX_NQ = 256;       // queries per item of the heaviest section (TMEM: 2 accumulator stages x 256 columns)
static constexpr int IR_MAX_KPL = 12;        // candidate buffer capacity / 32
static constexpr int IR_SECTIONS = 3;

struct IvfRowsArgs {
  int n_lists, nprobe, k_blocks, n_stages;
  int nq_max;                 // 32 / 64 / 128: queries per item of this launch (TMEM accumulator stride, buffer geometry)
  int gmax;                   // query-group size of lists probed by more than 64 queries (IR_MAX_NQ, or what fits resident)
  int L, Lc, capq;            // shortlist length, compaction trigger, capacity of a candidate buffer
  int section, cap_items;
  const int* item_base;       // [IR_SECTIONS * n_lists + 1] exclusive scan of the per-(section, list) item counts
  const int4* items;          // {list, query group, chunk, 0}
  const int* list_offsets; const int* list_rows;
  const int* q_off; const int* pair_of_pos;
  const float* scale; const float* bias;
  int list_major, l2_hint;
  unsigned* gthr;             // [B] orderable lower bound of every query's final L-th best score
  u64* cbuf;                  // [grid][nq_max][capq]
  // results of the items: one slot of <= L keys per (item, query) that kept anything, allocated from a global counter;
  // query b lists its slots in qslots[b][0 .. qn[b])
  u64* slot_keys; int* slot_cnt; int* n_slots; int cap_slots;
  int* qslots; int* qn; int qs_max; int* qflag;
  unsigned* trace;            // watchdog trace words (mapped pinned host memory), may be null
};

// chunk geometry of the rows-as-M path: a list is cut into chunks only for load balance, and every (chunk, query) pair
// costs a result slot, so lists probed by many queries - whose items are tensor-bound and long anyway - get longer chunks
__host__ __device__ __forceinline__ void ir_chunks(int len, int nq, int& n_ch, int& ch_rows) {
  int mult = nq / 32;
  mult = mult < 1 ? 1 : mult > 8 ? 8 : mult;
  const int target = IB_CH_ROWS * mult;                     // 2048 .. 16384 rows per chunk
  // a list probed by many queries is already split into query groups: fewer, longer chunks there (every query that
  // probes it pays one result slot per chunk)
  const int groups = nq <= 64 ? 1 : (nq + 127) / 128;
  const int max_ch = groups >= 8 ? 4 : groups >= 2 ? 8 : IB_MAX_CHUNKS;
  n_ch = (len + target - 1) / target;
  if (n_ch > max_ch) n_ch = max_ch;
  if (n_ch < 1) { n_ch = 0; ch_rows = target; return; }
  ch_rows = ((len + n_ch - 1) / n_ch + GT_BN - 1) / GT_BN * GT_BN;
  n_ch = (len + ch_rows - 1) / ch_rows;
}
__host__ __device__ __forceinline__ int ir_section_of(int nq) { return nq <= 32 ? 0 : nq <= 64 ? 1 : 2; }
__host__ __device__ __forceinline__ int ir_groups_of(int nq, int gmax) { return nq <= 64 ? 1 : (nq + gmax - 1) / gmax; }

// exact selection inside one candidate buffer by the whole warp: keep its L largest keys (in place, front of the
// buffer), return the L-th largest key.  n <= 32 * IR_MAX_KPL keys, n >= L.
__device__ __forceinline__ u64 ir_warp_compact(u64* buf, int n, int L, int lane) {
  u64 kk[IR_MAX_KPL];
#pragma unroll
  for (int j = 0; j < IR_MAX_KPL; ++j) { const int i = lane + 32 * j; kk[j] = i < n ? buf[i] : 0ull; }
  // bisection on the score bits: largest t with |{hi >= t}| >= L
  unsigned lo = 0u, hi = 0xFFFFFFFFu;
  while (lo < hi) {
    const unsigned mid = lo + ((hi - lo) >> 1) + 1u;
    int c = 0;
#pragma unroll
    for (int j = 0; j < IR_MAX_KPL; ++j) c += ((unsigned)(kk[j] >> 32) >= mid) ? 1 : 0;
    c = __reduce_add_sync(FULL, c);
    if (c >= L) lo = mid; else hi = mid - 1u;
  }
  const unsigned t = lo;
  int c_gt = 0, c_eq = 0;
#pragma unroll
  for (int j = 0; j < IR_MAX_KPL; ++j) {
    const unsigned h = (unsigned)(kk[j] >> 32);
    c_gt += h > t ? 1 : 0;
    c_eq += (h == t && kk[j] != 0ull) ? 1 : 0;
  }
  c_gt = __reduce_add_sync(FULL, c_gt);
  c_eq = __reduce_add_sync(FULL, c_eq);
  unsigned u = 0u;                               // keys with the threshold score: keep the (L - c_gt) lowest rows
  if (c_gt + c_eq > L) {
    const int want = L - c_gt;
    unsigned l2 = 0u, h2 = 0xFFFFFFFFu;
    while (l2 < h2) {
      const unsigned mid = l2 + ((h2 - l2) >> 1) + 1u;
      int c = 0;
#pragma unroll
      for (int j = 0; j < IR_MAX_KPL; ++j) c += ((unsigned)(kk[j] >> 32) == t && (unsigned)kk[j] >= mid) ? 1 : 0;
      c = __reduce_add_sync(FULL, c);
      if (c >= want) l2 = mid; else h2 = mid - 1u;
    }
    u = l2;
  }
  __syncwarp();
  int base = 0;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < IR_MAX_KPL; ++j) {
    const unsigned h = (unsigned)(kk[j] >> 32);
    const bool keep = kk[j] != 0ull && (h > t || (h == t && (unsigned)kk[j] >= u));
    const unsigned m = __ballot_sync(FULL, keep);
    if (keep) buf[base + __popc(m & lt)] = kk[j];
    base += __popc(m);
  }
  __syncwarp();
  return ((u64)t << 32) | (u64)u;
}

// RB ("resident B", bf16 list-major copy only): the query group of an item stays in shared memory for all of the item's
// row tiles - k_blocks x [gmax x 128 B] swizzled slabs loaded once per item by one producer warp - and the ring carries
// row tiles only.  The lists every query probes are re-read from L2 once per query group and the groups' operand once
// per row tile: with both operands streamed, an M128 x N128 tile moves 392 KB through the SM per 3072 tensor clocks and
// the kernel ran at ~24 % tensor-pipe utilisation, bound by the L2 -> SM path (ncu: r02_c5_rows).  Resident queries
// take the query operand and its 16-byte cp.async gathers out of the tile loop (see the measurement at the launch site).
template <bool TF32, bool RB>
__global__ void __launch_bounds__(IR_THREADS, 1)
ivf_rows_kernel(const __grid_constant__ CUtensorMap tmap_lm, const unsigned char* __restrict__ qmat,
                const unsigned char* __restrict__ bank, const int row_pitch, const IvfRowsArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int S = a.n_stages, NQ = a.nq_max;
  const unsigned b_bytes = (unsigned)(RB ? a.gmax : NQ) * GT_SLAB;    // one k-block of the query group
  const unsigned stage_bytes = RB ? 2u * GT_A_BYTES : 2u * (GT_A_BYTES + b_bytes);   // [A k0][A k1]([B k0][B k1])
  unsigned char* ring = smem;
  unsigned char* resb = ring + (size_t)S * stage_bytes;            // RB: [k_blocks][gmax x 128 B]
  float* thr_s = reinterpret_cast<float*>(resb + (RB ? (size_t)a.k_blocks * b_bytes : 0));   // [IR_MAX_NQ]
  int* cnt_s = reinterpret_cast<int*>(thr_s + IR_MAX_NQ);
  int* base_s = cnt_s + IR_MAX_NQ;              // keys in the buffer right after its last compaction
  int* qid_s = base_s + IR_MAX_NQ;
  int* qrow_s = qid_s + IR_MAX_NQ;              // the producers' copy of the group's query rows
  uint64_t* full = reinterpret_cast<uint64_t*>(qrow_s + IR_MAX_NQ);
  uint64_t* empty = full + IR_MAX_STAGES;
  uint64_t* tfull = empty + IR_MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* bfull = tempty + 2;                 // RB: the item's query slabs have landed (32 cp.async arrivals)
  uint64_t* bfree = bfull + 1;                  // RB: the MMAs of the item have finished reading them
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfree + 1);
  // need_s[tile parity] == tile number + 1: some candidate buffer needs compaction after this tile.  Written during the
  // tile's pass, read after the barrier that ends it; the same word is next written two tiles later, i.e. after the
  // barrier of the tile in between, which every reader of this tile has passed - no reset, no race.
  volatile unsigned* need_s = reinterpret_cast<volatile unsigned*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ELEMS_PER_SLAB = TF32 ? 32 : 64;
  const uint32_t tmem_cols = NQ <= 32 ? 64u : NQ <= 64 ? 128u : NQ <= 128 ? 256u : 512u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], RB ? 1 : a.list_major ? 129 : 128); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
    mbar_init(bfull, 32); mbar_init(bfree, 1);
    need_s[0] = 0u; need_s[1] = 0u;
    if (a.list_major) tc::tma_prefetch_desc(&tmap_lm);
    fence_mbar_init();
  }
  if (warp == 0) { tc::tmem_alloc(tmem_slot, tmem_cols); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // this launch's section of the work table; a table that did not fit processes nothing (the finish kernel hands
  // every query back)
  const int total_items = a.item_base[IR_SECTIONS * a.n_lists];
  const int item_lo = a.item_base[a.section * a.n_lists];
  const int item_hi = total_items <= a.cap_items ? a.item_base[(a.section + 1) * a.n_lists] : item_lo;

#define IR_DECODE_ITEM(item)                                                                                   \
  const int4 it = a.items[item];                                                                               \
  const int qb = a.q_off[it.x], nq = a.q_off[it.x + 1] - qb;                                                    \
  const int n_qg = ir_groups_of(nq, a.gmax);                                                                   \
  const int g_lo = (int)(((long long)nq * it.y) / n_qg), g_hi = (int)(((long long)nq * (it.y + 1)) / n_qg);     \
  const int a0 = qb + g_lo, n_live = g_hi - g_lo;                                                              \
  const int n_pad = max(16, (n_live + 15) & ~15);                                                              \
  const int lb = a.list_offsets[it.x], len = a.list_offsets[it.x + 1] - lb;                                    \
  int n_ch_, ch_rows_;                                                                                         \
  ir_chunks(len, nq, n_ch_, ch_rows_);                                                                         \
  const int r0 = it.z * ch_rows_, r1 = min(len, r0 + ch_rows_);

  if (RB && warp >= 5) {
    // ===================== producer (resident B): one warp; lane 0 also streams the row tiles by TMA =====================
    if (warp == 5) {
      const uint32_t resb_u32 = smem_u32(resb);
      const uint64_t pol_stream = (a.l2_hint & 1) ? l2_policy_evict_first() : l2_policy_evict_normal();
      const uint64_t pol_keep = (a.l2_hint & 2) ? l2_policy_evict_last() : l2_policy_evict_normal();
      const uint64_t pol_q = (a.l2_hint & 4) ? l2_policy_evict_last() : l2_policy_evict_normal();
      int stage = 0; unsigned phase = 0, item_n = 0;
      const int pj = lane & 7, prow = lane >> 3;      // piece = 16 bytes (chunk pj) of B rows prow + 4*i
      for (int item = item_lo + (int)blockIdx.x; item < item_hi; item += (int)gridDim.x, ++item_n) {
        IR_DECODE_ITEM(item)
        tc::mbar_wait_traced(bfree, (item_n & 1u) ^ 1u, a.trace, 1u);          // the previous item's MMAs are done with the slabs
        for (int r = prow; r < n_pad; r += 4) {
          const int src_row = min(r, n_live - 1);                  // columns past the group re-load a valid query (threshold +inf)
          const unsigned char* src = qmat + (size_t)(a.pair_of_pos[a0 + src_row] / a.nprobe) * row_pitch + pj * 16;
          const uint32_t dst = resb_u32 + (uint32_t)(r * GT_SLAB + ((pj ^ (r & 7)) << 4));
          for (int kb = 0; kb < a.k_blocks; ++kb) {
            const bool in = kb * GT_SLAB + pj * 16 < row_pitch;
            cp_async16(dst + (uint32_t)kb * b_bytes, src + (in ? (size_t)kb * GT_SLAB : 0), in ? 16u : 0u, pol_q);
          }
        }
        cp_async_arrive_noinc(bfull);
        if (lane == 0) {
          const uint64_t pol_b = n_qg > 1 ? pol_keep : pol_stream;   // rows of a list with several query groups are re-read from L2
          for (int cr = r0; cr < r1; cr += IR_BM) {
            for (int kb = 0; kb < a.k_blocks; kb += 2) {
              const int nkb = min(2, a.k_blocks - kb);
              tc::mbar_wait_traced(&empty[stage], phase ^ 1u, a.trace, 2u);
              mbar_arrive_expect_tx(&full[stage], (unsigned)(nkb * GT_A_BYTES));
              unsigned char* spg = ring + (size_t)stage * stage_bytes;
              for (int h = 0; h < nkb; ++h)
                tc::tma_load_2d(spg + h * GT_A_BYTES, &tmap_lm, (kb + h) * ELEMS_PER_SLAB, lb + cr, &full[stage], pol_b);
              if (++stage == S) { stage = 0; phase ^= 1u; }
            }
          }
        }
        __syncwarp();
      }
      cp_async_wait<0>();
    }
  } else if (warp >= 5) {
    // ===================== producers: 128 threads =====================
    const int pt = threadIdx.x - 160;
    const int prow = pt >> 3, pj = pt & 7;          // A gather: chunk pj of tile rows prow + 16*i
    const uint32_t a_dst = (uint32_t)(prow * GT_SLAB + ((pj ^ (prow & 7)) << 4));
    const uint32_t ring_u32 = smem_u32(ring);
    const uint64_t pol_stream = (a.l2_hint & 1) ? l2_policy_evict_first() : l2_policy_evict_normal();
    const uint64_t pol_keep = (a.l2_hint & 2) ? l2_policy_evict_last() : l2_policy_evict_normal();
    const uint64_t pol_q = (a.l2_hint & 4) ? l2_policy_evict_last() : l2_policy_evict_normal();
    int stage = 0; unsigned phase = 0;
    for (int item = item_lo + (int)blockIdx.x; item < item_hi; item += (int)gridDim.x) {
      IR_DECODE_ITEM(item)
      // query rows of the group, for the gathers below: a table in shared memory, each producer warp filling exactly the
      // entries it reads (rows 4w + j + 16*i), so a __syncwarp is all the hand-shake it needs.
      // Piece p = pt + 128*i of a k-block covers 16 bytes (chunk pj) of B row prow + 16*i.
      __syncwarp();                                 // this warp is done with the previous item's entries
      for (int idx = lane; idx < (n_pad >> 2); idx += 32) {
        const int r = (prow & ~3) + (idx & 3) + 16 * (idx >> 2);
        qrow_s[r] = a.pair_of_pos[a0 + min(r, n_live - 1)] / a.nprobe;   // columns past the group re-load a valid query (threshold +inf)
      }
      __syncwarp();
      const int n_pieces = n_pad >> 4;
      const uint64_t pol_b = n_qg > 1 ? pol_keep : pol_stream;   // rows of a list with several query groups are re-read from L2
      for (int cr = r0; cr < r1; cr += IR_BM) {
        int rb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rb[i] = a.list_major ? 0 : a.list_rows[lb + min(cr + prow + 16 * i, r1 - 1)];
        for (int kb = 0; kb < a.k_blocks; kb += 2) {
          const int nkb = min(2, a.k_blocks - kb);
          tc::mbar_wait_traced(&empty[stage], phase ^ 1u, a.trace, 3u);
          const uint32_t sp = ring_u32 + (uint32_t)stage * stage_bytes;
          bool in_row[2]; size_t ko[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            in_row[h] = h < nkb && (int)((kb + h) * GT_SLAB) + pj * 16 < row_pitch;
            ko[h] = in_row[h] ? (size_t)(kb + h) * GT_SLAB : 0;
          }
          if (a.list_major) {
            if (pt == 0) {
              mbar_arrive_expect_tx(&full[stage], (unsigned)(nkb * GT_A_BYTES));
              unsigned char* spg = ring + (size_t)stage * stage_bytes;
              for (int h = 0; h < nkb; ++h)
                tc::tma_load_2d(spg + h * GT_A_BYTES, &tmap_lm, (kb + h) * ELEMS_PER_SLAB, lb + cr, &full[stage], pol_b);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const unsigned char* src = bank + (size_t)rb[i] * row_pitch + pj * 16;
#pragma unroll
              for (int h = 0; h < 2; ++h)
                if (h < nkb) cp_async16(sp + h * GT_A_BYTES + a_dst + i * 16 * GT_SLAB, src + ko[h], in_row[h] ? 16u : 0u, pol_b);
            }
          }
          for (int i = 0; i < n_pieces; ++i) {
            const int row = prow + 16 * i;
            const unsigned char* src = qmat + (size_t)qrow_s[row] * row_pitch + pj * 16;
            const uint32_t dst = sp + 2 * GT_A_BYTES + (uint32_t)(row * GT_SLAB + ((pj ^ (row & 7)) << 4));
#pragma unroll
            for (int h = 0; h < 2; ++h)
              if (h < nkb) cp_async16(dst + h * b_bytes, src + ko[h], in_row[h] ? 16u : 0u);   // no L2 hint: see below
          }
          // (the query gathers carry no L2 cache hint: with `.L2::cache_hint` inside this run-time loop the kernel died with
          // "illegal instruction" on the B200 - any producer variant but the original 8-way unrolled one did - and the
          // hint was evict_normal, i.e. the default, anyway)
          cp_async_arrive_noinc(&full[stage]);
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
    cp_async_wait<0>();
  } else if (warp == 0) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0; unsigned phase = 0, tile_n = 0, item_n = 0;
      for (int item = item_lo + (int)blockIdx.x; item < item_hi; item += (int)gridDim.x, ++item_n) {
        IR_DECODE_ITEM(item)
        (void)a0;
        const uint32_t idesc = tc::make_idesc(TF32 ? 2 : 1, IR_BM, n_pad);
        if (RB) {
          tc::mbar_wait_traced(bfull, item_n & 1u, a.trace, 4u);    // this item's query slabs are resident
          fence_proxy_async();
        }
        for (int cr = r0; cr < r1; cr += IR_BM, ++tile_n) {
          const unsigned acc = tile_n & 1u, acc_phase = (tile_n >> 1) & 1u;
          tc::mbar_wait_traced(&tempty[acc], acc_phase ^ 1u, a.trace, 5u);
          tc::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * (uint32_t)NQ;
          for (int kb = 0; kb < a.k_blocks; kb += 2) {
            const int nkb = min(2, a.k_blocks - kb);
            tc::mbar_wait_traced(&full[stage], phase, a.trace, 6u);
            if (!RB) fence_proxy_async();           // cp.async wrote through the generic proxy; the MMA reads through the async proxy
            tc::tc_fence_after();
            const unsigned char* sp = ring + (size_t)stage * stage_bytes;
            for (int h = 0; h < nkb; ++h) {
              const uint64_t da = tc::make_smem_desc_sw128(sp + h * GT_A_BYTES);
              const uint64_t db = tc::make_smem_desc_sw128(RB ? resb + (size_t)(kb + h) * b_bytes : sp + 2 * GT_A_BYTES + h * b_bytes);
#pragma unroll
              for (int j = 0; j < GT_SLAB / 32; ++j)
                tc::umma<TF32>(d_tmem, da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), idesc, ((kb + h) | j) != 0 ? 1u : 0u);
            }
            tc::umma_commit(&empty[stage]);
            if (++stage == S) { stage = 0; phase ^= 1u; }
          }
          tc::umma_commit(&tfull[acc]);
        }
        if (RB) tc::umma_commit(bfree);             // every MMA that reads this item's slabs has completed when this arrives
      }
    }
  } else {
    // ===================== epilogue: thread = one bank row of the tile =====================
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const int te = quarter * 32 + lane;
    const int et = threadIdx.x - 32;                // 0..127
    const int wi = warp - 1;                        // 0..3: this warp maintains the queries q = wi (mod 4)
    u64* my_cbuf = a.cbuf + (size_t)blockIdx.x * NQ * a.capq;
    const int L = a.L, Lc = a.Lc, capq = a.capq;
    unsigned tile_n = 0;
    for (int item = item_lo + (int)blockIdx.x; item < item_hi; item += (int)gridDim.x) {
      IR_DECODE_ITEM(item)
      tc::named_bar_sync(1, 128);                   // the previous item's flush is complete
      for (int e = et; e < NQ; e += 128) {
        const bool live = e < n_live;
        const int b = live ? a.pair_of_pos[a0 + e] / a.nprobe : 0;
        qid_s[e] = b; cnt_s[e] = 0; base_s[e] = 0;
        float t = INFINITY;                         // dead columns select nothing
        if (live) { const unsigned g = *reinterpret_cast<volatile unsigned*>(a.gthr + b); t = g ? f32_from_orderable(g) : -INFINITY; }
        thr_s[e] = t;
      }
      // per-row terms of the first tile; every later tile's are fetched while the previous one is processed
      bool valid = r0 + te < r1;
      int rid = valid ? a.list_rows[lb + r0 + te] : 0;
      float sc = valid ? (a.scale ? a.scale[rid] : 1.f) : 0.f;
      float bi = valid ? (a.bias ? a.bias[rid] : 0.f) : 0.f;
      tc::named_bar_sync(1, 128);
      int tiles_since_refresh = 0;
      for (int cr = r0; cr < r1; cr += IR_BM, ++tile_n) {
        const unsigned acc = tile_n & 1u, acc_phase = (tile_n >> 1) & 1u;
        const bool valid_n = cr + IR_BM + te < r1;
        const int rid_n = valid_n ? a.list_rows[lb + cr + IR_BM + te] : 0;
        tc::mbar_wait_traced(&tfull[acc], acc_phase, a.trace, 7u);
        tc::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * (uint32_t)NQ;
        // 32 columns per TMEM load, the next load in flight while the current columns are compared (the epilogue, not the
        // tensor pipe, set the pace of this kernel: ncu r3b - 40 % of its samples sat in 16-column load -> wait -> compare
        // round trips).  Columns past n_pad hold stale accumulators; their thresholds are +inf.
        uint32_t va[32], vb[32];
        auto scan32 = [&](const uint32_t (&vr)[32], int c0) {
          unsigned mask = 0u;
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(thr_s + c0 + j4);
            mask |= (fmaf(__uint_as_float(vr[j4 + 0]), sc, bi) >= t4.x) ? (1u << (j4 + 0)) : 0u;
            mask |= (fmaf(__uint_as_float(vr[j4 + 1]), sc, bi) >= t4.y) ? (1u << (j4 + 1)) : 0u;
            mask |= (fmaf(__uint_as_float(vr[j4 + 2]), sc, bi) >= t4.z) ? (1u << (j4 + 2)) : 0u;
            mask |= (fmaf(__uint_as_float(vr[j4 + 3]), sc, bi) >= t4.w) ? (1u << (j4 + 3)) : 0u;
          }
          if (!valid) mask = 0u;
          if (c0 + 32 > n_pad) mask &= (1u << (n_pad - c0)) - 1u;      // n_pad is a multiple of 16: never trust stale columns
          while (mask) {                            // rare once the thresholds are warm
            const int j = __ffs(mask) - 1;
            mask &= mask - 1u;
            float vj = 0.f;
#pragma unroll
            for (int u = 0; u < 32; ++u) vj = (u == j) ? __uint_as_float(vr[u]) : vj;
            const u64 key = make_key(fmaf(vj, sc, bi), (unsigned)rid);
            const int pos = atomicAdd(&cnt_s[c0 + j], 1);
            if (pos < capq) my_cbuf[(size_t)(c0 + j) * capq + pos] = key;
            if (pos + 1 == L || pos >= Lc) need_s[acc] = tile_n + 1u;   // a buffer reached L keys for the first time, or outgrew Lc
          }
        };
        tc::tmem_ld_32x32_issue(taddr, va);
        tc::tmem_ld_fence(va);
#pragma unroll 1
        for (int c0 = 0; c0 < n_pad; c0 += 64) {
          const bool more_b = c0 + 32 < n_pad, more_a = c0 + 64 < n_pad;
          if (more_b) tc::tmem_ld_32x32_issue(taddr + c0 + 32, vb);
          scan32(va, c0);
          if (more_b) {
            tc::tmem_ld_fence(vb);
            if (more_a) tc::tmem_ld_32x32_issue(taddr + c0 + 64, va);
            scan32(vb, c0 + 32);
            if (more_a) tc::tmem_ld_fence(va);
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        const float sc_n = valid_n ? (a.scale ? a.scale[rid_n] : 1.f) : 0.f;
        const float bi_n = valid_n ? (a.bias ? a.bias[rid_n] : 0.f) : 0.f;
        tc::named_bar_sync(1, 128);                 // every append of this tile is visible
        const bool refresh = ++tiles_since_refresh >= 4;      // pull the bounds other CTAs published every few tiles
        if (need_s[acc] == tile_n + 1u || refresh) {
          tiles_since_refresh = 0;
          // lane l looks after the queries q = wi + 4*l (+ 128 for the second half of a 256-query item)
          for (int q0 = 0; q0 < n_live; q0 += 128) {
            const int q = q0 + wi + 4 * lane;
            bool need = false;
            unsigned g = 0u;
            if (q < n_live) {
              g = *reinterpret_cast<volatile unsigned*>(a.gthr + qid_s[q]);
              const int c = cnt_s[q];
              if (c > capq) { a.qflag[qid_s[q]] = 1; cnt_s[q] = capq; }    // cannot happen by construction; never silently drop
              need = c > Lc || (base_s[q] < L && c >= L);
              if (!need && g != 0u) thr_s[q] = fmaxf(thr_s[q], f32_from_orderable(g));
            }
            unsigned todo = __ballot_sync(FULL, need);
            while (todo) {
              const int src = __ffs(todo) - 1;
              todo &= todo - 1u;
              const int qq = q0 + wi + 4 * src;
              const unsigned gq = __shfl_sync(FULL, g, src);
              const u64 kth = ir_warp_compact(my_cbuf + (size_t)qq * capq, min(cnt_s[qq], capq), L, lane);
              if (lane == 0) {
                const unsigned o = (unsigned)(kth >> 32);
                cnt_s[qq] = L; base_s[qq] = L;
                thr_s[qq] = fmaxf(thr_s[qq], f32_from_orderable(o > gq ? o : gq));
                if (o > gq) atomicMax(a.gthr + qid_s[qq], o);     // L keys of this item alone are at or above it
              }
            }
          }
          tc::named_bar_sync(1, 128);
        }
        valid = valid_n; rid = rid_n; sc = sc_n; bi = bi_n;
      }
      // ---- item done: the keys of every query that can still matter go to a result slot ----
      for (int q = wi; q < n_live; q += 4) {
        const int b = qid_s[q];
        u64* buf = my_cbuf + (size_t)q * capq;
        int c = min(cnt_s[q], capq);
        if (c > L || (c == L && base_s[q] < L)) {
          const u64 kth = ir_warp_compact(buf, c, L, lane);
          c = L;
          if (lane == 0) atomicMax(a.gthr + b, (unsigned)(kth >> 32));
        }
        __syncwarp();
        const unsigned g = *reinterpret_cast<volatile unsigned*>(a.gthr + b);
        u64 kk[4];
        unsigned m[4];
        int total = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = lane + 32 * j;
          kk[j] = i < c ? buf[i] : 0ull;
          m[j] = __ballot_sync(FULL, kk[j] != 0ull && (unsigned)(kk[j] >> 32) >= g);
          total += __popc(m[j]);
        }
        if (total > 0) {
          int slot = 0;
          if (lane == 0) {
            slot = atomicAdd(a.n_slots, 1);
            const int at = slot < a.cap_slots ? atomicAdd(a.qn + b, 1) : a.qs_max;
            if (at < a.qs_max) { a.qslots[(size_t)b * a.qs_max + at] = slot; a.slot_cnt[slot] = total; }
            else { a.qflag[b] = 1; slot = -1; }     // slot table or the query's slot list is full: hand the query back
          }
          slot = __shfl_sync(FULL, slot, 0);
          if (slot >= 0) {
            u64* dst = a.slot_keys + (size_t)slot * L;
            int off = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if ((m[j] >> lane) & 1u) dst[off + __popc(m[j] & ((1u << lane) - 1u))] = kk[j];
              off += __popc(m[j]);
            }
          }
        }
      }
    }
  }
#undef IR_DECODE_ITEM
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, tmem_cols); }
}

// ---- work table of the rows-as-M kernel: three sections by queries per list ----
__global__ void __launch_bounds__(256) ir_item_count_kernel(const int* __restrict__ q_off, const int* __restrict__ list_offsets,
                                                            int n_lists, int gmax, int* __restrict__ items_c) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_lists) return;
  const int nq = q_off[c + 1] - q_off[c], len = list_offsets[c + 1] - list_offsets[c];
  int n_ch, ch_rows;
  ir_chunks(len, nq, n_ch, ch_rows);
  const int sec = ir_section_of(nq);
  const int cnt = (nq > 0 && len > 0) ? ir_groups_of(nq, gmax) * n_ch : 0;
  for (int s = 0; s < IR_SECTIONS; ++s) items_c[s * n_lists + c] = s == sec ? cnt : 0;
}
__global__ void __launch_bounds__(256) ir_item_fill_kernel(const int* __restrict__ item_base, const int* __restrict__ q_off,
                                                           int n_lists, int cap, int gmax, int4* __restrict__ items) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= n_lists) return;
  const int nq = q_off[c + 1] - q_off[c];
  const int sec = ir_section_of(nq);
  const int base = item_base[sec * n_lists + c], cnt = item_base[sec * n_lists + c + 1] - base;
  const int n_qg = ir_groups_of(nq, gmax);
  for (int i = lane; i < cnt; i += 32)
    if (base + i < cap) items[base + i] = make_int4(c, i % n_qg, i / n_qg, 0);   // the groups of one chunk run side by side (L2)
}

// ---- finish: the result slots of one query -> its L best approximate keys -> exact re-score + certification ----
static constexpr int IR_QS_MAX = 1024;       // result slots one query can link (8 x AURA_MAX_NPROBE)
static constexpr int IR_HIST = 1024;
struct IvfSlotFinishArgs {
  const u64* slot_keys; const int* slot_cnt; int cap_slots; const int* qslots; const int* qn; int qs_max; const int* qflag;
  const unsigned* gthr; const int* item_base; int n_lists, cap_items;
  const void* rows; int bf16; int d; const float* qn_vec; const float* scale; const float* bias; float eps;
  const float* eps_q;
  int k, L; long long row_base; int empty_ok;
  long long* out_idx; float* out_score; int* uncertain;
};
__global__ void __launch_bounds__(128) ivf_slot_finish_kernel(const IvfSlotFinishArgs f) {
  __shared__ u64 keys[IB_MERGE_CAP];
  __shared__ u64 ex[GT_DEEP];
  __shared__ int slot_id[IR_QS_MAX];
  __shared__ int slot_n[IR_QS_MAX];
  __shared__ int hist[IR_HIST];
  __shared__ int n_kept;
  __shared__ unsigned s_max, s_t1;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bool overflow = f.item_base[IR_SECTIONS * f.n_lists] > f.cap_items || f.qflag[b] != 0;
  const int ns = overflow ? 0 : min(f.qn[b], f.qs_max);
  const unsigned g = f.gthr[b];
  if (threadIdx.x == 0) { n_kept = 0; s_max = 0u; }
  for (int i = threadIdx.x; i < ns; i += blockDim.x) {
    const int sl = f.qslots[(size_t)b * f.qs_max + i];
    slot_id[i] = sl; slot_n[i] = min(f.slot_cnt[sl], f.L);
  }
  __syncthreads();
  // pass 1: everything at or above the bound the items agreed on (the max over items of an item's own L-th best)
  unsigned my_max = 0u;
  for (int i = warp; i < ns; i += 4) {
    const u64* src = f.slot_keys + (size_t)slot_id[i] * f.L;
    for (int j = lane; j < slot_n[i]; j += 32) {
      const u64 key = src[j];
      const unsigned o = (unsigned)(key >> 32);
      if (o >= g) { const int pos = atomicAdd(&n_kept, 1); if (pos < IB_MERGE_CAP) keys[pos] = key; my_max = max(my_max, o); }
    }
  }
  atomicMax(&s_max, my_max);
  __syncthreads();
  int kept = n_kept;
  bool refined = false;
  if (kept > IB_MERGE_CAP) {
    refined = true;
    // the neighbours are spread over many lists, so no single item's L-th best is a tight bound: refine it with a
    // histogram of the scores between the bound and the best score, then collect again
    const unsigned long long range = (unsigned long long)(s_max - g) + 1ull;
    for (int i = threadIdx.x; i < IR_HIST; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = warp; i < ns; i += 4) {
      const u64* src = f.slot_keys + (size_t)slot_id[i] * f.L;
      for (int j = lane; j < slot_n[i]; j += 32) {
        const unsigned o = (unsigned)(src[j] >> 32);
        if (o >= g) atomicAdd(&hist[(int)(((unsigned long long)(o - g) * IR_HIST) / range)], 1);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = 0, t = IR_HIST - 1;
      for (; t > 0; --t) { acc += hist[t]; if (acc >= f.L) break; }      // bins >= t hold at least L keys (or t == 0: all of them)
      // key in a bin >= t  <=>  (o - g) * IR_HIST >= t * range
      s_t1 = g + (unsigned)(((unsigned long long)t * range + IR_HIST - 1) / IR_HIST);
      n_kept = 0;
    }
    __syncthreads();
    const unsigned t1 = s_t1;
    for (int i = warp; i < ns; i += 4) {
      const u64* src = f.slot_keys + (size_t)slot_id[i] * f.L;
      for (int j = lane; j < slot_n[i]; j += 32) {
        const u64 key = src[j];
        if ((unsigned)(key >> 32) >= t1) { const int pos = atomicAdd(&n_kept, 1); if (pos < IB_MERGE_CAP) keys[pos] = key; }
      }
    }
    __syncthreads();
    kept = n_kept;
    if (kept > IB_MERGE_CAP) { overflow = true; kept = 0; }            // thousands of keys inside one score bin: hand back
  }
  const int n2 = max(64, next_pow2(kept));
  for (int i = kept + threadIdx.x; i < n2; i += blockDim.x) keys[i] = 0ull;
  block_bitonic_sort_desc(keys, n2);
  RescoreArgs ra;
  ra.out_mul = 1.f;
  ra.rows = f.rows; ra.bf16 = f.bf16; ra.d = f.d; ra.q = f.qn_vec + (size_t)b * f.d; ra.scale = f.scale; ra.bias = f.bias;
  // every key at or above the final bound of the query (or of the histogram refinement) is in keys[]: second chance allowed
  ra.deep = 1;
  { const unsigned fl = refined ? max(g, s_t1) : g; ra.floor_score = fl ? f32_from_orderable(fl) : -INFINITY; }
  ra.eps = f.eps_q ? f.eps_q[b] : f.eps; ra.k = f.k; ra.L = f.L; ra.row_base = f.row_base;
  ra.out_idx = f.out_idx + (size_t)b * f.k; ra.out_score = f.out_score + (size_t)b * f.k;
  ra.uncertain = f.uncertain + b;
  rescore_and_write(keys, n2, ex, ra);
  // no candidate at all (every probed list empty -> the reference scans all rows, hippocampal.py:269-270, unless the
  // caller is a row shard) or a table that did not fit: hand the query back to the per-query path
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
  L.force = o; o += 2 * a256((size_t)n_queries * 4 + 256);     // force flags, then the per-query certification bounds
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
void launch_normalize_queries_eps(const float* q, int n, int d, float* qn, __nv_bfloat16* qb, const float* relerr, float unit,
                                  float* eps_q, cudaStream_t st);

// diagnostics of the last batch call on a workspace: {items, result slots used, most slots linked by one query, queries
// flagged inside the kernel, queries without any slot}
__global__ void __launch_bounds__(256) ir_stats_kernel(const int* __restrict__ item_base, int n_lists, const int* __restrict__ n_slots,
                                                       const int* __restrict__ qn, const int* __restrict__ qflag, int n_queries,
                                                       int* __restrict__ out) {
  __shared__ int s_max, s_flag, s_empty;
  if (threadIdx.x == 0) { s_max = 0; s_flag = 0; s_empty = 0; }
  __syncthreads();
  for (int b = threadIdx.x; b < n_queries; b += blockDim.x) {
    atomicMax(&s_max, qn[b]);
    if (qflag[b]) atomicAdd(&s_flag, 1);
    if (qn[b] == 0) atomicAdd(&s_empty, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) { out[0] = item_base[IR_SECTIONS * n_lists]; out[1] = *n_slots; out[2] = s_max; out[3] = s_flag; out[4] = s_empty; }
}

// ---- host side of the rows-as-M path -----------------------------------------------------------------------------
static int ir_list_len(int k) {
  // shortlist length.  Up to 32: the lengths of the queries-as-M kernel.  Beyond: k + 28 in steps of 32, capped at the
  // 128 keys the re-score holds - the one-pass selection costs the same for any length (bisection over the candidate
  // buffer), and the margin beyond k is what certifies a result on near-tie heavy data (k = 100 keeps 128, not 114)
  if (k + 14 <= GT_L_SMALL) return GT_L_SMALL;
  if (k + 14 <= GT_L) return GT_L;
  const int want = (k + 28 + GT_L - 1) / GT_L * GT_L;
  return want < GT_MAX_L ? want : GT_MAX_L;
}
static constexpr int IR_CBUF_KEYS = IR_MAX_NQ * 32 * IR_MAX_KPL;   // candidate-buffer keys per CTA, largest geometry

struct IrLayout {
  size_t probes, counts, q_off, cursor, pair_of_pos, pos_of_pair, items_c, item_base, items, qn, qb, eps_q, coarse, gthr, qcnt, qflag,
      n_slots, qslots, slot_cnt, slot_keys, cbuf, total;
  int cap_slots;
};
static IrLayout ir_layout(int n_queries, int d, int n_lists, int nprobe, int cap, size_t coarse_bytes, int L) {
  IrLayout Lo;
  size_t o = 0;
  const size_t pairs = (size_t)n_queries * nprobe;
  Lo.probes = o; o += a256(pairs * 8);
  Lo.counts = o; o += a256((size_t)n_lists * 4);
  Lo.q_off = o; o += a256((size_t)(n_lists + 1) * 4);
  Lo.cursor = o; o += a256((size_t)(IR_SECTIONS * n_lists + 1) * 4);
  Lo.pair_of_pos = o; o += a256(pairs * 4);
  Lo.pos_of_pair = o; o += a256(pairs * 4);
  Lo.items_c = o; o += a256((size_t)IR_SECTIONS * n_lists * 4);
  Lo.item_base = o; o += a256((size_t)(IR_SECTIONS * n_lists + 1) * 4);
  Lo.items = o; o += a256((size_t)cap * 16);
  Lo.qn = o; o += a256((size_t)n_queries * d * 4);
  Lo.qb = o; o += a256((size_t)n_queries * d * 2);
  Lo.eps_q = o; o += a256((size_t)n_queries * 4);
  Lo.coarse = o; o += a256(coarse_bytes);
  Lo.gthr = o; o += a256((size_t)n_queries * 4);          // gthr, qcnt, qflag, n_slots are cleared with one memset
  Lo.qcnt = o; o += a256((size_t)n_queries * 4);
  Lo.qflag = o; o += a256((size_t)n_queries * 4);
  Lo.n_slots = o; o += 256;
  Lo.qslots = o; o += a256((size_t)n_queries * IR_QS_MAX * 4);
  // one result slot per (item, query) that keeps anything: ~1 per (query, probe) plus one per extra chunk of a long list
  Lo.cap_slots = (int)(8 * pairs + 4096 < 32000000 ? 8 * pairs + 4096 : 32000000);
  Lo.slot_cnt = o; o += a256((size_t)Lo.cap_slots * 4);
  Lo.slot_keys = o; o += a256((size_t)Lo.cap_slots * L * 8);
  Lo.cbuf = o; o += a256((size_t)sm_count() * IR_CBUF_KEYS * 8);
  Lo.total = o;
  return Lo;
}

static size_t ir_workspace_bytes(int n_queries, int d, int n_lists, int nprobe, int k) {
  const int cap = ib_cap_items(n_queries, nprobe, n_lists);
  return ir_layout(n_queries, d, n_lists, nprobe, cap, ivf_coarse_ws_bytes(n_queries, d, n_lists, nprobe), ir_list_len(k)).total;
}

static int ivf_rows_search(const void* rows, bool bank_bf16, long long n_rows, int d, const float* queries, int n_queries,
                           const float* centroids, int n_lists, int nprobe, const int* list_offsets, const int* list_rows,
                           const void* rows_by_list, bool lm_shadow, bool measured, const float* shadow_relerr, const float* scale,
                           const float* bias, int k, long long row_base, int flags, float eps, long long* out_idx,
                           float* out_score, int* out_uncertain, void* workspace, cudaStream_t st) {
  // lm_shadow: the list-major copy is a bf16 shadow of an fp32 bank - the tensor cores read bf16, the re-score reads fp32
  const bool bf16 = bank_bf16 || lm_shadow;
  const int eb = bf16 ? 2 : 4;
  const int cap = ib_cap_items(n_queries, nprobe, n_lists);
  const IrLayout Lo = ir_layout(n_queries, d, n_lists, nprobe, cap, ivf_coarse_ws_bytes(n_queries, d, n_lists, nprobe), ir_list_len(k));
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  long long* probes = reinterpret_cast<long long*>(ws + Lo.probes);
  int* counts = reinterpret_cast<int*>(ws + Lo.counts);
  int* q_off = reinterpret_cast<int*>(ws + Lo.q_off);
  int* cursor = reinterpret_cast<int*>(ws + Lo.cursor);
  int* pair_of_pos = reinterpret_cast<int*>(ws + Lo.pair_of_pos);
  int* pos_of_pair = reinterpret_cast<int*>(ws + Lo.pos_of_pair);
  int* items_c = reinterpret_cast<int*>(ws + Lo.items_c);
  int* item_base = reinterpret_cast<int*>(ws + Lo.item_base);
  int4* items = reinterpret_cast<int4*>(ws + Lo.items);
  float* qn = reinterpret_cast<float*>(ws + Lo.qn);
  __nv_bfloat16* qb = bf16 ? reinterpret_cast<__nv_bfloat16*>(ws + Lo.qb) : nullptr;
  unsigned* gthr = reinterpret_cast<unsigned*>(ws + Lo.gthr);
  AURA_CUDA_OK(cudaMemsetAsync(gthr, 0, Lo.qslots - Lo.gthr, st));     // gthr, qcnt, qflag, n_slots

  int rc = ivf_run_coarse(queries, n_queries, d, centroids, n_lists, nprobe, probes, ws + Lo.coarse, st);
  if (rc != AURA_OK) return rc;
  float* eps_q = measured ? reinterpret_cast<float*>(ws + Lo.eps_q) : nullptr;
  if (eps_q) launch_normalize_queries_eps(queries, n_queries, d, qn, qb, lm_shadow ? shadow_relerr : nullptr, eps, eps_q, st);
  else launch_normalize_queries(queries, n_queries, d, qn, qb, st);
  static const int env_seed = env_int("AURA_IVF_SEED", 1);
  if (env_seed) {
    ivf_seed_gthr_kernel<<<n_queries, 128, 0, st>>>(rows, bank_bf16 ? 1 : 0, d, qn, probes, nprobe, n_lists, list_offsets, list_rows,
                                                    scale, bias, eps, eps_q, ir_list_len(k), gthr);
    note_launches(1);
  }
  const int n_pairs = n_queries * nprobe;
  AURA_CUDA_OK(cudaMemsetAsync(counts, 0, (size_t)n_lists * 4, st));
  ib_pair_hist_kernel<<<(n_pairs + 255) / 256, 256, 0, st>>>(probes, n_pairs, n_lists, counts);
  launch_scan_offsets(counts, n_lists, q_off, cursor, st);
  ib_pair_scatter_kernel<<<(n_pairs + 255) / 256, 256, 0, st>>>(probes, n_pairs, n_lists, cursor, pair_of_pos, pos_of_pair);
  // geometry first: the query-group size of the heavy lists depends on whether their operand can stay resident
  static const int env_l2 = env_int("AURA_IVF_L2HINT", 3), env_grid = env_int("AURA_IVF_GRID", 0);
  static const int env_stages = env_int("AURA_IVF_STAGES", 0), env_rb = env_int("AURA_IVF_RB", 0);
  const int elems = GT_SLAB / eb;
  const int k_blocks = (d + elems - 1) / elems;
  const size_t fixed = 5 * IR_MAX_NQ * 4 + (2 * IR_MAX_STAGES + 6) * 8 + 16;
  const size_t smem_cap = (size_t)max_smem_optin() - 1024;
  // resident-query mode (bf16 operands, list-major copy; opt-in AURA_IVF_RB=1): ring of 3 x 32 KB row-tile stages +
  // k_blocks slabs of gmax queries (d = 768: 80).  Measured on one C5 shard: 12.28 ms for the heavy section against
  // 10.8 ms with both operands streamed - the groups shrink from 128 to 80 queries, so the rows are streamed 1.6x as
  // often and only 20 % of the L2 -> SM traffic is saved, while a ring of 96 KB covers less latency.  Kept for d <= 512,
  // where all 128 queries fit.
  // query-group size of the heavy lists: 256 (M128 x N256 tiles move 18 B per (row, query) pair through the L2 -> SM path,
  // M128 x N128 tiles 24 B; that path, not HBM or the tensor pipe, bounds those lists - DESIGN 4.6); AURA_IVF_GMAX=128 for A/B
  static const int env_gmax = env_int("AURA_IVF_GMAX", IR_MAX_NQ);
  int gmax = env_gmax == 128 ? 128 : IR_MAX_NQ;
  bool rb = false;
  if (env_rb && bf16 && rows_by_list != nullptr && smem_cap > fixed + 3 * 2 * GT_A_BYTES) {
    int fit = (int)((smem_cap - fixed - 3 * 2 * GT_A_BYTES) / ((size_t)k_blocks * GT_SLAB)) & ~15;
    if (fit > 128) fit = 128;
    if (fit >= 64) { rb = true; gmax = fit; }
  }
  ir_item_count_kernel<<<(n_lists + 255) / 256, 256, 0, st>>>(q_off, list_offsets, n_lists, gmax, items_c);
  launch_scan_offsets(items_c, IR_SECTIONS * n_lists, item_base, cursor, st);
  ir_item_fill_kernel<<<(n_lists * 32 + 255) / 256, 256, 0, st>>>(item_base, q_off, n_lists, cap, gmax, items);
  note_launches(4);

  IvfRowsArgs a;
  a.n_lists = n_lists; a.nprobe = nprobe; a.cap_items = cap; a.gmax = gmax;
  a.k_blocks = k_blocks;
  a.L = ir_list_len(k);
  a.Lc = a.L * 2 < 64 ? 64 : a.L * 2;
  a.capq = a.Lc + IR_BM;
  a.item_base = item_base; a.items = items; a.list_offsets = list_offsets; a.list_rows = list_rows;
  a.q_off = q_off; a.pair_of_pos = pair_of_pos; a.scale = scale; a.bias = bias;
  a.l2_hint = env_l2; a.gthr = gthr; a.cbuf = reinterpret_cast<u64*>(ws + Lo.cbuf);
  a.slot_keys = reinterpret_cast<u64*>(ws + Lo.slot_keys); a.slot_cnt = reinterpret_cast<int*>(ws + Lo.slot_cnt);
  a.n_slots = reinterpret_cast<int*>(ws + Lo.n_slots); a.cap_slots = Lo.cap_slots;
  a.qslots = reinterpret_cast<int*>(ws + Lo.qslots); a.qn = reinterpret_cast<int*>(ws + Lo.qcnt);
  a.qs_max = IR_QS_MAX; a.qflag = reinterpret_cast<int*>(ws + Lo.qflag);
  a.trace = trap_trace_device();
  CUtensorMap tmap_lm;
  memset(&tmap_lm, 0, sizeof(tmap_lm));
  a.list_major = 0;
  if (rows_by_list != nullptr) {
    AURA_REQUIRE((reinterpret_cast<uintptr_t>(rows_by_list) & 15) == 0, AURA_ERR_UNSUPPORTED, "aura_ivf_search_batch: rows_by_list must be 16-byte aligned");
    const int trc = encode_tmap_2d(&tmap_lm, rows_by_list, eb, bf16, n_rows, d, IR_BM);
    if (trc != AURA_OK) return trc;
    a.list_major = 1;
  }
  typedef void (*IrKern)(const CUtensorMap, const unsigned char*, const unsigned char*, int, const IvfRowsArgs);
  IrKern kern = bf16 ? ivf_rows_kernel<false, false> : ivf_rows_kernel<true, false>;
  IrKern kern_rb = ivf_rows_kernel<false, true>;
  int grid = sm_count();
  if (env_grid >= 1 && env_grid <= grid) grid = env_grid;
  const int nq_of_section[IR_SECTIONS] = {32, 64, rb ? 128 : gmax};
  size_t smem_max = 0, smem_of[IR_SECTIONS];
  int stages_of[IR_SECTIONS];
  for (int s = 0; s < IR_SECTIONS; ++s) {
    const bool rb_s = rb && s == IR_SECTIONS - 1;
    const size_t resident = rb_s ? (size_t)k_blocks * gmax * GT_SLAB : 0;
    const size_t stage_bytes = rb_s ? 2 * (size_t)GT_A_BYTES : 2 * ((size_t)GT_A_BYTES + (size_t)nq_of_section[s] * GT_SLAB);
    int stages = (int)((smem_cap - fixed - resident) / stage_bytes);
    if (stages > IR_MAX_STAGES) stages = IR_MAX_STAGES;
    if (env_stages >= 2 && env_stages < stages) stages = env_stages;
    AURA_REQUIRE(stages >= 2, AURA_ERR_UNSUPPORTED, "aura_ivf_search_batch: needs 2 pipeline stages of shared memory");
    stages_of[s] = stages;
    smem_of[s] = (size_t)stages * stage_bytes + resident + fixed + 1024;
    if (!rb_s && smem_of[s] > smem_max) smem_max = smem_of[s];
  }
  AURA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  if (rb) AURA_CUDA_OK(cudaFuncSetAttribute(kern_rb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of[IR_SECTIONS - 1]));
  for (int s = IR_SECTIONS - 1; s >= 0; --s) {       // the lists probed by the most queries first
    const bool rb_s = rb && s == IR_SECTIONS - 1;
    a.section = s; a.nq_max = nq_of_section[s]; a.n_stages = stages_of[s];
    (rb_s ? kern_rb : kern)<<<grid, IR_THREADS, smem_of[s], st>>>(tmap_lm, reinterpret_cast<const unsigned char*>(bf16 ? (const void*)qb : (const void*)qn),
                                                                   reinterpret_cast<const unsigned char*>(rows), d * eb, a);
    AURA_CUDA_OK(cudaGetLastError());
  }
  IvfSlotFinishArgs f;
  f.slot_keys = a.slot_keys; f.slot_cnt = a.slot_cnt; f.cap_slots = a.cap_slots; f.qslots = a.qslots; f.qn = a.qn;
  f.qs_max = a.qs_max; f.qflag = a.qflag; f.gthr = gthr;
  f.item_base = item_base; f.n_lists = n_lists; f.cap_items = cap;
  f.rows = rows; f.bf16 = bank_bf16 ? 1 : 0; f.d = d; f.qn_vec = qn; f.scale = scale; f.bias = bias; f.eps = eps;
  f.eps_q = eps_q;
  f.k = k; f.L = a.L; f.row_base = row_base; f.empty_ok = (flags & AURA_IVF_EMPTY_OK) ? 1 : 0;
  f.out_idx = out_idx; f.out_score = out_score; f.uncertain = out_uncertain;
  ivf_slot_finish_kernel<<<n_queries, 128, 0, st>>>(f);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(IR_SECTIONS + 1);
  return AURA_OK;
}

}  // namespace aura
using namespace aura;

extern "C" size_t aura_ivf_search_batch_workspace_bytes(int n_queries, int d, int n_centroid_rows, int nprobe, int k) {
  if (n_queries < 1 || d < 1 || n_centroid_rows < 1 || nprobe < 1 || k < 1) return 0;
  const int cap = ib_cap_items(n_queries, nprobe, n_centroid_rows);
  const size_t old_bytes = ib_layout(n_queries, d, n_centroid_rows, nprobe, cap, ivf_coarse_ws_bytes(n_queries, d, n_centroid_rows, nprobe)).total;
  const size_t new_bytes = ir_workspace_bytes(n_queries, d, n_centroid_rows, nprobe, k);
  return old_bytes > new_bytes ? old_bytes : new_bytes;
}

extern "C" int aura_ivf_search_batch(const void* rows, int dtype, int64_t n_rows, int d, const float* queries, int n_queries,
                                     const float* centroids, int n_centroid_rows, int nprobe, const int32_t* list_offsets,
                                     const int32_t* list_rows, const void* rows_by_list, int lm_dtype, const float* shadow_relerr,
                                     const float* scale, const float* bias, int k, int64_t row_base, int flags, float eps,
                                     int64_t* out_idx, float* out_score, int32_t* out_uncertain, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  AURA_REQUIRE(dtype == AURA_F32 || dtype == AURA_BF16, AURA_ERR_INVALID_ARG, "aura_ivf_search_batch: bad dtype %d", dtype);
  AURA_REQUIRE(n_rows >= 1 && n_rows < 0x7fffffffll && d >= 1 && n_queries >= 1 && n_centroid_rows >= 1, AURA_ERR_INVALID_ARG,
               "aura_ivf_search_batch: n_rows=%lld d=%d n_queries=%d n_centroid_rows=%d", (long long)n_rows, d, n_queries,
               n_centroid_rows);
  AURA_REQUIRE(nprobe >= 1 && nprobe <= AURA_MAX_NPROBE && nprobe <= n_centroid_rows, AURA_ERR_INVALID_ARG,
               "aura_ivf_search_batch: nprobe=%d", nprobe);
  AURA_REQUIRE(k >= 1 && k + 14 <= GT_MAX_L, AURA_ERR_UNSUPPORTED, "aura_ivf_search_batch: k=%d too large (max %d)", k, GT_MAX_L - 14);
  const bool bank_bf16 = dtype == AURA_BF16;
  AURA_REQUIRE(rows_by_list == nullptr || lm_dtype == dtype || (lm_dtype == AURA_BF16 && dtype == AURA_F32), AURA_ERR_INVALID_ARG,
               "aura_ivf_search_batch: the list-major copy must have the bank's dtype or be a bf16 shadow of an fp32 bank");
  // a bf16 list-major SHADOW of an fp32 bank: tensor cores and HBM see bf16 (half the bytes, twice the rate), the finish
  // kernel re-scores from the fp32 rows; eps is then the score-per-cosine unit and the bound is measured per query from
  // shadow_relerr (as aura_batch_topk)
  const bool lm_shadow = rows_by_list != nullptr && lm_dtype == AURA_BF16 && dtype == AURA_F32;
  AURA_REQUIRE(!lm_shadow || ((d % 8) == 0 && shadow_relerr != nullptr), AURA_ERR_INVALID_ARG,
               "aura_ivf_search_batch: a bf16 list-major shadow needs d %% 8 == 0 and its rounding-error scalar");
  const bool bf16 = bank_bf16 || lm_shadow;          // element type the tensor cores read
  // measured certification bound: eps is the score-per-cosine unit and the bound is computed per query from the ACTUAL
  // rounding errors - of the query alone for a bf16 bank (its rows are exact operands), of query and rows for a shadow
  const bool measured = lm_shadow || (bank_bf16 && (flags & AURA_IVF_MEASURED_EPS) != 0);
  const int eb = bf16 ? 2 : 4;
  AURA_REQUIRE(((size_t)d * (bank_bf16 ? 2 : 4)) % 16 == 0 && (d % 4) == 0 && (reinterpret_cast<uintptr_t>(rows) & 15) == 0, AURA_ERR_UNSUPPORTED,
               "aura_ivf_search_batch: rows must be 16-byte aligned with a 16-byte multiple pitch (d=%d)", d);
  AURA_REQUIRE(rows && queries && centroids && list_offsets && list_rows && out_idx && out_score && out_uncertain && workspace,
               AURA_ERR_INVALID_ARG, "aura_ivf_search_batch: null pointer");
  AURA_REQUIRE(workspace_bytes >= aura_ivf_search_batch_workspace_bytes(n_queries, d, n_centroid_rows, nprobe, k),
               AURA_ERR_WORKSPACE, "aura_ivf_search_batch: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  // Two formulations of the list-major fine stage.  Shortlists of <= 32 keys (k <= 18) fit the queries-as-M kernel's
  // register lists and it is the faster one there (C4: 10.1 ms against 12.8 ms per batch); longer shortlists (k up to 114)
  // would need one full pass per 32 keys, so they take the rows-as-M kernel with its one-pass selection (one C5 shard,
  // k = 100: 16 ms against 46 ms).  AURA_IVF_ROWS=1|0 forces one of them (test switch, read once per call: the
  // parity tests run every shape through both).
  const int forced = env_int("AURA_IVF_ROWS", -1);
  // (the multi-round form of the queries-as-M kernel has no per-query bound: a measured bound with k > 18 always goes rows-as-M)
  const bool rows_path = (forced >= 0 ? forced != 0 : k + 14 > GT_L) || (measured && k + 14 > GT_L);
  if (rows_path)
    return ivf_rows_search(rows, bank_bf16, n_rows, d, queries, n_queries, centroids, n_centroid_rows, nprobe, list_offsets, list_rows,
                           rows_by_list, lm_shadow, measured, shadow_relerr, scale, bias, k, row_base, flags, eps,
                           reinterpret_cast<long long*>(out_idx), out_score, out_uncertain, workspace, st);
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
  float* eps_q = measured ? reinterpret_cast<float*>(ws + L.force + a256((size_t)n_queries * 4 + 256)) : nullptr;
  if (eps_q) launch_normalize_queries_eps(queries, n_queries, d, qn, qb, lm_shadow ? shadow_relerr : nullptr, eps, eps_q, st);
  else launch_normalize_queries(queries, n_queries, d, qn, qb, st);
  static const int env_seed = env_int("AURA_IVF_SEED", 1);
  if (env_seed) {                    // start bound of round 0 (later rounds clear gthr: their keys lie below a ceiling)
    static const int env_shadow_small0 = env_int("AURA_IVF_SHADOW_SMALL", 0);
    const int seed_L = (k + 14 <= GT_L_SMALL && (!lm_shadow || env_shadow_small0)) ? GT_L_SMALL : GT_L;
    ivf_seed_gthr_kernel<<<n_queries, 128, 0, st>>>(rows, bank_bf16 ? 1 : 0, d, qn, probes, nprobe, n_centroid_rows, list_offsets,
                                                    list_rows, scale, bias, eps, eps_q, seed_L, gthr);
    note_launches(1);
  }
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
  // k <= 10: 24-entry lists, one round (a bf16 shadow keeps 32: its measured bound is about as wide as the TF32 worst
  // case while the bf16 scores scatter more, and every uncertified query costs a per-query scan)
  static const int env_shadow_small = env_int("AURA_IVF_SHADOW_SMALL", 0);
  const bool small = k + 14 <= GT_L_SMALL && (!lm_shadow || env_shadow_small);
  IvfKern kern0 = small ? (bf16 ? ivf_gemm_kernel<false, false, GT_L_SMALL> : ivf_gemm_kernel<true, false, GT_L_SMALL>)
                        : (bf16 ? ivf_gemm_kernel<false, false, GT_L> : ivf_gemm_kernel<true, false, GT_L>);    // first round: no ceiling
  IvfKern kern1 = bf16 ? ivf_gemm_kernel<false, true, GT_L> : ivf_gemm_kernel<true, true, GT_L>;
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
  f.rows = rows; f.bf16 = bank_bf16 ? 1 : 0; f.d = d; f.qn = qn; f.scale = scale; f.bias = bias; f.eps = eps; f.k = k;
  f.eps_q = eps_q;
  f.L = small ? GT_L_SMALL : GT_L;
  f.row_base = row_base; f.spread = a.spread; f.chunk_major = chunk_major; f.out_idx = reinterpret_cast<long long*>(out_idx); f.out_score = out_score; f.uncertain = out_uncertain;
  f.cand = rounds > 1 ? cand : nullptr; f.ceil_out = ceil_buf; f.round = 0; f.force_flag = force;
  f.empty_ok = (flags & AURA_IVF_EMPTY_OK) ? 1 : 0;
  f.gthr = gthr;
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
    return launch_cand_rescore(cand, rounds * GT_L, rows, bank_bf16 ? 1 : 0, d, qn, scale, bias, eps, k, row_base,
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

namespace aura {
// fp32 bank -> bf16 list-major shadow: one warp per row, conversion + the row's relative rounding error (max-reduced)
__global__ void __launch_bounds__(256) ivf_pack_lists_bf16_kernel(const float* __restrict__ rows, const int* __restrict__ list_rows,
                                                                  long long n, int d, __nv_bfloat16* __restrict__ out,
                                                                  float* __restrict__ relerr_max) {
  const int lane = threadIdx.x & 31;
  float worst = 0.f;
  for (long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < n; p += (long long)gridDim.x * (blockDim.x >> 5)) {
    const float* src = rows + (size_t)list_rows[p] * d;
    __nv_bfloat16* dst = out + (size_t)p * d;
    float ee = 0.f, ss = 0.f;
    for (int e = lane; e < d; e += 32) {
      const float v = src[e];
      const __nv_bfloat16 b = __float2bfloat16_rn(v);
      dst[e] = b;
      const float t = __bfloat162float(b) - v;
      ee = fmaf(t, t, ee); ss = fmaf(v, v, ss);
    }
    ee = warp_sum(ee); ss = warp_sum(ss);
    if (ss > 0.f) worst = fmaxf(worst, sqrtf(ee / ss));
  }
  if (relerr_max != nullptr && lane == 0 && worst > 0.f) atomicMax(reinterpret_cast<int*>(relerr_max), __float_as_int(worst * 1.0001f));
}
}  // namespace aura

extern "C" int aura_ivf_pack_lists(const void* rows, int dtype, int d, const int32_t* list_rows, int64_t n_listed,
                                   void* rows_by_list, int out_dtype, float* relerr_max, void* stream) {
  AURA_REQUIRE(dtype == AURA_F32 || dtype == AURA_BF16, AURA_ERR_INVALID_ARG, "aura_ivf_pack_lists: bad dtype %d", dtype);
  AURA_REQUIRE(out_dtype == dtype || (out_dtype == AURA_BF16 && dtype == AURA_F32), AURA_ERR_INVALID_ARG,
               "aura_ivf_pack_lists: the copy has the bank's dtype or is a bf16 shadow of an fp32 bank");
  AURA_REQUIRE(rows && list_rows && rows_by_list && d >= 1 && n_listed >= 0, AURA_ERR_INVALID_ARG, "aura_ivf_pack_lists: bad argument");
  const size_t pitch = (size_t)d * (dtype == AURA_BF16 ? 2 : 4);
  AURA_REQUIRE(pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(rows) & 15) == 0 && (reinterpret_cast<uintptr_t>(rows_by_list) & 15) == 0,
               AURA_ERR_UNSUPPORTED, "aura_ivf_pack_lists: rows must be 16-byte aligned with a 16-byte multiple pitch (d=%d)", d);
  if (n_listed == 0) return AURA_OK;
  if (out_dtype != dtype)
    ivf_pack_lists_bf16_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(rows), list_rows, (long long)n_listed, d,
                                                                                 reinterpret_cast<__nv_bfloat16*>(rows_by_list), relerr_max);
  else
    ivf_pack_lists_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(rows), list_rows, (long long)n_listed,
                                                                            (int)(pitch / 16), reinterpret_cast<uint4*>(rows_by_list));
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}

extern "C" int aura_ivf_search_batch_items(const void* workspace, int n_queries, int d, int n_centroid_rows, int nprobe, int k,
                                           int32_t* items_out, int32_t* host_cap, void* stream) {
  AURA_REQUIRE(workspace && items_out && host_cap && k >= 1, AURA_ERR_INVALID_ARG, "aura_ivf_search_batch_items: null pointer");
  const int cap = ib_cap_items(n_queries, nprobe, n_centroid_rows);
  const IrLayout L = ir_layout(n_queries, d, n_centroid_rows, nprobe, cap, ivf_coarse_ws_bytes(n_queries, d, n_centroid_rows, nprobe), ir_list_len(k));
  const unsigned char* ws = reinterpret_cast<const unsigned char*>(workspace);
  ir_stats_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const int*>(ws + L.item_base), n_centroid_rows,
                                                       reinterpret_cast<const int*>(ws + L.n_slots),
                                                       reinterpret_cast<const int*>(ws + L.qcnt),
                                                       reinterpret_cast<const int*>(ws + L.qflag), n_queries, items_out);
  AURA_CUDA_OK(cudaGetLastError());
  *host_cap = cap;
  return AURA_OK;
}

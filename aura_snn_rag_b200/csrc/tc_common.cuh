// tcgen05 / TMEM / TMA (tensor-map) PTX wrappers for the dense-contraction kernels (sm_100a).
// Bit layouts follow the PTX ISA "tcgen05" matrix / instruction descriptor tables.
#pragma once

#include <cuda.h>  // CUtensorMap (types only; the encode entry point is fetched through cudart)

#include "aura_common.cuh"

namespace aura {
namespace tc {

// ---- watchdog-guarded mbarrier wait: a protocol bug traps instead of hanging the GPU ----
__device__ __forceinline__ void mbar_wait_guarded(uint64_t* bar, unsigned parity) {
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) { asm volatile("trap;"); }
  }
}

// same, leaving a trace: {tag, block, thread, parity} in mapped pinned host memory (aura_debug_last_trap)
__device__ __forceinline__ void mbar_wait_traced(uint64_t* bar, unsigned parity, unsigned* trace, unsigned tag) {
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) {
      if (trace != nullptr) {
        trace[0] = tag; trace[1] = blockIdx.x; trace[2] = threadIdx.x; trace[3] = parity;
        __threadfence_system();
      }
      asm volatile("trap;");
    }
  }
}

// ---- TMA 2-D tiled load, completion on an mbarrier, with an L2 cache-policy hint ----
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- TMA gather4: 4 rows (box = [128 bytes of K, 1 row]) of a 2-D tensor land as 4 consecutive 128-byte smem
// rows, swizzled by the destination row like a tiled SWIZZLE_128B box (verified on B200, scripts/exp/gather4_test.cu)
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3,
                                            uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "l"(policy)
      : "memory");
}

// ---- TMEM allocation (one warp, .sync.aligned) ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- shared-memory matrix descriptor: K-major operand, 128-byte swizzle ----
// rows are 128 B apart inside an 8-row group (the TMA SWIZZLE_128B box layout), 8-row groups are
// 1024 B apart (SBO); start address / offsets are in 16-byte units; version 1 = sm_100.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(const void* smem_ptr) {
  const uint64_t addr = (uint64_t)((smem_u32(smem_ptr) & 0x3FFFFu) >> 4);
  return addr | (1ull << 16) /* LBO (ignored for swizzled K-major) */ | (64ull << 32) /* SBO = 1024 B */ |
         (1ull << 46) /* descriptor version */ | (2ull << 61) /* SWIZZLE_128B */;
}

// ---- instruction descriptor: D fp32, A/B both `fmt` (1 = bf16, 2 = tf32), K-major A and B ----
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int m, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <bool TF32>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// all MMAs issued so far by this thread arrive on `bar` when they complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- cta_group::2 (CTA pair) variants: one MMA spans two SMs, each CTA holds its own 128 A rows and half of the B tile ----
static constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> rank 0's copy
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// add 1 to the same-offset shared-memory word of CTA `rank` of this cluster
__device__ __forceinline__ void cluster_red_inc(uint32_t* local_word, int rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_word)), "r"(rank));
  asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], 1;" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// both CTAs of the pair issue this; the bytes are counted on the LEADER's mbarrier (same smem offset, rank bit cleared)
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <bool TF32>
__device__ __forceinline__ void umma_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// MMAs issued so far arrive on `bar` in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((unsigned short)3) : "memory");
}
// arrive on the LEADER CTA's copy of `bar` (works from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(raddr) : "r"(smem_u32(bar)));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}

// ---- TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns ----
// (tcgen05.ld / wait::ld are .sync.aligned: every lane of the warp must execute them together.  The epilogues run
// lane-divergent insertion loops between loads and nothing but the compiler's choice of reconvergence point put the lanes
// back together - an unrelated change elsewhere in a kernel moved it and the load faulted as an illegal instruction.  So
// every wrapper reconverges explicitly.)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  __syncwarp();
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Split form for software pipelining: issue the load of the NEXT 32 columns, work on the current ones, then fence.  The
// fence names the destination registers as read-write operands, so no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld_32x32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  __syncwarp();
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_fence(uint32_t (&r)[32]) {
  __syncwarp();
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// this warp's 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  __syncwarp();
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void named_bar_sync(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

}  // namespace tc

// ---- tile geometry and epilogue helpers shared by gemm_topk.cu and ivf_batch.cu ----
static constexpr int GT_BM = 128;            // A rows per tile (TMEM lanes)
static constexpr int GT_BN = 256;            // B rows per tile (TMEM columns per accumulator stage)
static constexpr int GT_SLAB = 128;          // bytes of K per pipeline stage (one swizzle atom row)
static constexpr int GT_A_BYTES = GT_BM * GT_SLAB;   // 16 KB
static constexpr int GT_B_BYTES = GT_BN * GT_SLAB;   // 32 KB
static constexpr int GT_STAGE_BYTES = GT_A_BYTES + GT_B_BYTES;
static constexpr int GT_MAX_STAGES = 4;
static constexpr int GT_THREADS = 192;       // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
static constexpr int GT_L = 32;             // per-row list length kept by the epilogue (registers)
static constexpr int GT_L_ASSIGN = 4;       // list length of the nearest-centroid variant
static constexpr int GT_L_SMALL = 24;       // exact search with k <= 10 (k + 14 margin)
static constexpr int GT_L_WIDE = 48;        // bf16 shadow shortlist of an fp32 bank: the wider rounding bound needs more margin
static constexpr int GT_MAX_L = 128;           // most candidates a query can carry into the exact re-score (4 rounds x 32)
static constexpr int GT_MAX_ROUNDS = GT_MAX_L / GT_L;


// v[j] for a run-time j without spilling v[] to local memory: 5-level select tree
__device__ __forceinline__ float select32(const float (&v)[32], int j) {
  float a[16], b[8], c[4];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 16) ? v[i + 16] : v[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 8) ? a[i + 8] : a[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[i + 4] : b[i];
  const float d0 = (j & 2) ? c[2] : c[0], d1 = (j & 2) ? c[3] : c[1];
  return (j & 1) ? d1 : d0;
}

__device__ __forceinline__ float select16(const float (&v)[16], int j) {
  float b[8], c[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 8) ? v[i + 8] : v[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[i + 4] : b[i];
  const float d0 = (j & 2) ? c[2] : c[0], d1 = (j & 2) ? c[3] : c[1];
  return (j & 1) ? d1 : d0;
}

// thread-private top-GT_L list, sorted descending, entirely in registers.  All compares are independent
// (key > e[i] is monotone in i), so an insertion is ~GT_L predicated moves with no dependent chain.
template <int L>
__device__ __forceinline__ void list_insert_sorted(u64 (&e)[L], u64 key) {
#pragma unroll
  for (int i = L - 1; i >= 1; --i) {
    const bool ci = key > e[i], cp = key > e[i - 1];
    e[i] = ci ? (cp ? e[i - 1] : key) : e[i];
  }
  e[0] = key > e[0] ? key : e[0];
}

// Same list, for streams whose keys arrive in ASCENDING ROW order (the dense kernels walk the bank's columns left to
// right): a key that ties an entry on the score has the higher row and ranks below it, so comparing the 32 score bits
// alone reproduces the 64-bit order at half the compare cost.
template <int L>
__device__ __forceinline__ void list_insert_sorted_asc(u64 (&e)[L], u64 key) {
  const unsigned kh = (unsigned)(key >> 32);
#pragma unroll
  for (int i = L - 1; i >= 1; --i) {
    const bool ci = kh > (unsigned)(e[i] >> 32), cp = kh > (unsigned)(e[i - 1] >> 32);
    e[i] = ci ? (cp ? e[i - 1] : key) : e[i];
  }
  e[0] = kh > (unsigned)(e[0] >> 32) ? key : e[0];
}

// Per-lane partial of dot(bank row, fp32 query) in exactly the operation order of scan_topk.cu (lane-strided 128-bit
// chunks, fmaf nest); warp_sum() of the result is bit-identical to the scan kernel's dot product.
__device__ __forceinline__ float exact_dot_partial(const void* rows, int bf16, int d, unsigned row, const float* q, int lane) {
  float acc = 0.f;
  const float4* q4 = reinterpret_cast<const float4*>(q);
  if (bf16) {
    const uint4* x8 = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(rows) + (size_t)row * d);
    for (int ch = lane; ch < (d >> 3); ch += 32) {
      const uint4 x = x8[ch];
      const float4 qa = q4[2 * ch], qb = q4[2 * ch + 1];
      acc = fmaf(bf16_lo(x.x), qa.x, acc); acc = fmaf(bf16_hi(x.x), qa.y, acc);
      acc = fmaf(bf16_lo(x.y), qa.z, acc); acc = fmaf(bf16_hi(x.y), qa.w, acc);
      acc = fmaf(bf16_lo(x.z), qb.x, acc); acc = fmaf(bf16_hi(x.z), qb.y, acc);
      acc = fmaf(bf16_lo(x.w), qb.z, acc); acc = fmaf(bf16_hi(x.w), qb.w, acc);
    }
  } else {
    const float4* x4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(rows) + (size_t)row * d);
    for (int ch = lane; ch < (d >> 2); ch += 32) {
      const float4 x = x4[ch], qq = q4[ch];
      acc = fmaf(x.x, qq.x, fmaf(x.y, qq.y, fmaf(x.z, qq.z, fmaf(x.w, qq.w, acc))));
    }
  }
  return acc;
}


// ---- shared tail of the finish kernels: exact fp32 re-score of the L best tensor-core candidates + certification ----
// keys[0..n2) sorted descending (approximate keys), ex = GT_DEEP scratch keys in shared memory; 128 threads.
//
// Second chance.  A result is certified when its exact k-th best beats (approximate score of the first candidate NOT
// re-scored) + eps.  With only the L best re-scored that candidate is the (L+1)-th, and on near-tie heavy data the margin
// of L - k places is sometimes not enough - the caller then pays a full exact scan for the query.  But the finish kernels
// hold MORE than L candidates: every row whose approximate score exceeds `floor_score` (the largest threshold any list /
// item ever rejected against: list tails, shared bounds, sampled or seeded start bounds) is present in keys[].  So a
// query that fails with L re-scores up to GT_DEEP (256) of them and certifies against the first key beyond, or against
// floor_score itself when everything above it was re-scored.  Only flagged queries take this path.
static constexpr int GT_DEEP = 256;
struct RescoreArgs {
  const void* rows; int bf16; int d;
  const float* q;              // this query, normalised fp32
  const float* scale; const float* bias; float eps;
  int k, L; long long row_base;
  long long* out_idx; float* out_score; int* uncertain;   // already offset to this query
  float out_mul;               // written score = exact score * out_mul (all-pairs: 1 / ||row_i||); ranking is unaffected
  int deep;                    // second chance allowed: keys[] is complete above floor_score
  float floor_score;
};
// exact keys of candidates [0, n_cand) -> ex[0, cap) (zero beyond n_cand / for empty candidates); each warp takes every
// 4th candidate, four at a time so that their row loads overlap
__device__ __forceinline__ void rescore_range(const u64* keys, int n_cand, int cap, u64* ex, const RescoreArgs& f) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = warp; c0 < cap; c0 += 16) {
    float part[4];
    unsigned rowv[4];
    bool live[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + 4 * u;
      live[u] = c < n_cand && keys[c] != 0ull;
      rowv[u] = live[u] ? key_row(keys[c]) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) part[u] = live[u] ? exact_dot_partial(f.rows, f.bf16, f.d, rowv[u], f.q, lane) : 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + 4 * u;
      u64 key = 0ull;
      if (live[u]) {
        const float dot = warp_sum(part[u]);
        const float sc = f.scale ? f.scale[rowv[u]] : 1.f, bi = f.bias ? f.bias[rowv[u]] : 0.f;
        key = make_key(fmaf(dot, sc, bi), rowv[u]);
      }
      if (lane == 0 && c < cap) ex[c] = key;
    }
  }
}
__device__ __forceinline__ void rescore_write_topk(const u64* ex, int cap, const RescoreArgs& f) {
  for (int i = threadIdx.x; i < f.k; i += blockDim.x) {
    const u64 key = i < cap ? ex[i] : 0ull;
    f.out_idx[i] = key ? f.row_base + (long long)key_row(key) : -1ll;
    f.out_score[i] = key ? key_score(key) * f.out_mul : -INFINITY;
  }
}
__device__ __forceinline__ void rescore_and_write(const u64* keys, int n2, u64* ex, const RescoreArgs& f) {
  __shared__ int s_flag, s_known;
  const int n_cand = min(f.L, n2);
  rescore_range(keys, n_cand, GT_MAX_L, ex, f);
  block_bitonic_sort_desc(ex, GT_MAX_L);
  rescore_write_topk(ex, GT_MAX_L, f);
  if (threadIdx.x == 0) {
    // rows outside the shortlist have approximate score <= s_L (the L-th best approximate score), hence exact
    // score <= s_L + eps: the result is certified when the exact k-th best beats that
    int flag = 0;
    if (f.L - 1 < n2 && keys[f.L - 1] != 0ull) {
      const float bound = key_score(keys[f.L - 1]) + f.eps;
      const u64 kth = ex[min(f.k, GT_MAX_L) - 1];
      if (kth == 0ull || !(key_score(kth) > bound)) flag = 1;
    }
    int known = 0;
    if (flag && f.deep) {
      // second chance: how many candidates are known above the completeness floor (counted up to GT_DEEP + 1)
      int lo = 0, hi = min(n2, GT_DEEP + 1);                      // first index whose key is empty or at / below the floor
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (keys[mid] != 0ull && key_score(keys[mid]) > f.floor_score) lo = mid + 1; else hi = mid;
      }
      known = lo;
    }
    s_flag = flag; s_known = known;
    if (f.uncertain) *f.uncertain = flag;
  }
  __syncthreads();
  // two depths: most queries that fail with L pass with 64 candidates (a CTA that re-scores 256 rows is the tail of the
  // whole finish launch), the rest go to GT_DEEP
  int done = f.L;
  for (int round = 0; round < 2; ++round) {
    const int cap = round == 0 ? 64 : GT_DEEP;
    const int known = s_known;
    const int depth = min(known, cap);
    if (s_flag == 0 || depth <= done) { __syncthreads(); continue; }
    __syncthreads();                                              // everyone has read s_flag before it is rewritten
    rescore_range(keys, depth, cap, ex, f);
    block_bitonic_sort_desc(ex, cap);
    rescore_write_topk(ex, cap, f);
    if (threadIdx.x == 0) {
      // every row not re-scored is either keys[depth...] (approximate score <= that of keys[depth]) or absent from
      // keys[] (approximate score <= floor_score)
      const float bound = (known > depth ? key_score(keys[depth]) : f.floor_score) + f.eps;
      const u64 kth = ex[min(f.k, cap) - 1];
      const int flag = (kth == 0ull || !(key_score(kth) > bound)) ? 1 : 0;
      s_flag = flag;
      if (f.uncertain) *f.uncertain = flag;
    }
    done = depth;
    __syncthreads();
  }
}

// ---- host: tensor-map encode through the runtime's driver entry point (no -lcuda link dependency) ----
// 2-D row-major matrix [n_rows, d] of `elem_bytes`-byte elements; box = [128 bytes of K, box_rows], SWIZZLE_128B.
int encode_tmap_2d(CUtensorMap* map, const void* base, int elem_bytes, bool bf16, long long n_rows, int d, int box_rows);


// kmeans.cu: exclusive scan of counts[0..n) -> offsets[0..n] (and cursor := offsets), single CTA
void launch_scan_offsets(const int* counts, int n, int* offsets, int* cursor, cudaStream_t st);

// gemm_topk.cu: nearest-centroid assignment / coarse probe selection on the tcgen05 kernel
bool tc_assign_supported(const void* rows, int dtype, long long n_rows, int d, int n_cent);
size_t tc_assign_workspace_bytes(long long n_rows, int d, int dtype, int n_cent);
int tc_assign(const void* rows, int dtype, long long n_rows, int d, const float* cent, int n_cent, const float* csq,
              const float* inv_norm, int* assign, float* cid_f32, int cid_stride, float* best_out, void* workspace,
              cudaStream_t st);
bool tc_coarse_supported(const float* queries, int n_queries, int d, const float* cent, int n_cent, int nprobe);
size_t tc_coarse_workspace_bytes(int n_queries, int d, int n_cent, int nprobe);
int tc_coarse(const float* queries, int n_queries, int d, const float* cent, int n_cent, const float* csq, int nprobe,
              long long* probes, void* workspace, cudaStream_t st);

}  // namespace aura

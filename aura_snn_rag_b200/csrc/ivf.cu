// Centroid-path query (hippocampal.py:257-307): coarse probe selection + fine scan of the probed
// inverted lists.
//   K4 coarse : dist[b][c] = ||centroid_c - q_b||_2 in the direct (difference) form torch.norm uses
//               (:261), over every row of the centroid buffer, then the nprobe nearest (:262).
//               CTA = 8 queries held in shared memory x all centroid rows, one warp per centroid row
//               (each centroid row is read once per 8 queries, 128-bit loads when d % 4 == 0).
//   fine      : scan_topk.cu in indirect mode (CSR lists replace the mask passes of :264-268).
#include "tc_common.cuh"

namespace aura {

static constexpr int CO_QT = 8;  // queries per CTA

__global__ void __launch_bounds__(256) coarse_dist_kernel(const float* __restrict__ queries, int n_queries, int d,
                                                          const float* __restrict__ cent, int n_cent,
                                                          float* __restrict__ dist /* [n_queries][n_cent] */) {
  extern __shared__ __align__(16) float qs[];  // [CO_QT][d]
  const int q0 = blockIdx.x * CO_QT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < CO_QT * d; i += blockDim.x) {
    const int qi = i / d, e = i - qi * d;
    qs[i] = (q0 + qi < n_queries) ? queries[(size_t)(q0 + qi) * d + e] : 0.f;
  }
  __syncthreads();
  const int c_per_cta = (n_cent + gridDim.y - 1) / gridDim.y;
  const int c_lo = blockIdx.y * c_per_cta, c_hi = min(n_cent, c_lo + c_per_cta);
  const bool vec = (d & 3) == 0 && ((reinterpret_cast<uintptr_t>(cent) & 15) == 0);
  for (int c = c_lo + warp; c < c_hi; c += 8) {
    float acc[CO_QT];
#pragma unroll
    for (int qi = 0; qi < CO_QT; ++qi) acc[qi] = 0.f;
    const float* cr = cent + (size_t)c * d;
    if (vec) {
      const float4* c4 = reinterpret_cast<const float4*>(cr);
      const float4* q4 = reinterpret_cast<const float4*>(qs);
      const int d4 = d >> 2;
      for (int e = lane; e < d4; e += 32) {
        const float4 cv = c4[e];
#pragma unroll
        for (int qi = 0; qi < CO_QT; ++qi) {
          const float4 qv = q4[qi * d4 + e];
          const float t0 = cv.x - qv.x, t1 = cv.y - qv.y, t2 = cv.z - qv.z, t3 = cv.w - qv.w;
          acc[qi] = fmaf(t0, t0, fmaf(t1, t1, fmaf(t2, t2, fmaf(t3, t3, acc[qi]))));
        }
      }
    } else {
      for (int e = lane; e < d; e += 32) {
        const float cv = cr[e];
#pragma unroll
        for (int qi = 0; qi < CO_QT; ++qi) { const float t = cv - qs[qi * d + e]; acc[qi] = fmaf(t, t, acc[qi]); }
      }
    }
#pragma unroll
    for (int qi = 0; qi < CO_QT; ++qi) {
      const float s = warp_sum(acc[qi]);
      if (lane == 0 && q0 + qi < n_queries) dist[(size_t)(q0 + qi) * n_cent + c] = sqrtf(s);
    }
  }
}

// one CTA (4 warps) per query: top-nprobe nearest centroid rows, nearest first, lower row on ties
template <int KPL>
__global__ void __launch_bounds__(128) coarse_select_kernel(const float* __restrict__ dist, int n_cent, int nprobe,
                                                            long long* __restrict__ probes) {
  __shared__ u64 merge[4 * 32 * KPL];
  __shared__ u64 best[32 * KPL];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* dq = dist + (size_t)blockIdx.x * n_cent;
  WarpTopK<KPL> tk;
  tk.init();
  for (int base = warp * 32; base < n_cent; base += 128) {
    const int c = base + lane;
    const u64 key = c < n_cent ? make_key(-dq[c], (unsigned)c) : 0ull;
    unsigned pending = __ballot_sync(FULL, key > tk.thr);
    while (pending) {
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
      const u64 kx = __shfl_sync(FULL, key, src);
      if (kx > tk.thr) tk.insert(kx, lane);
    }
  }
  publish_cta_topk<KPL>(tk, warp, lane, 4, merge, nprobe, best);
  for (int i = threadIdx.x; i < nprobe; i += blockDim.x) {
    const u64 key = best[i];
    probes[(size_t)blockIdx.x * nprobe + i] = key ? (long long)key_row(key) : -1ll;
  }
}

static size_t coarse_ws(int n_queries, int n_cent) { return ((size_t)n_queries * n_cent * 4 + 255) / 256 * 256; }
// the tensor-core formulation needs csq[n_cent] + its own scratch; the workspace is sized for whichever is larger
static size_t coarse_ws_any(int n_queries, int d, int n_cent, int nprobe) {
  const size_t simt = coarse_ws(n_queries, n_cent);
  const size_t tcw = ((size_t)n_cent * 4 + 255) / 256 * 256 + tc_coarse_workspace_bytes(n_queries, d, n_cent, nprobe);
  return simt > tcw ? simt : tcw;
}

static int run_coarse(const float* queries, int n_queries, int d, const float* centroids, int n_cent, int nprobe,
                      long long* probes, void* workspace, cudaStream_t st) {
  if (tc_coarse_supported(queries, n_queries, d, centroids, n_cent, nprobe)) {
    float* csq = reinterpret_cast<float*>(workspace);
    launch_row_sq_norms(centroids, n_cent, d, csq, st);
    return tc_coarse(queries, n_queries, d, centroids, n_cent, csq, nprobe, probes,
                     reinterpret_cast<unsigned char*>(workspace) + ((size_t)n_cent * 4 + 255) / 256 * 256, st);
  }
  float* dist = reinterpret_cast<float*>(workspace);
  const size_t smem = (size_t)CO_QT * d * 4;
  AURA_REQUIRE(smem <= (size_t)max_smem_optin() - 1024, AURA_ERR_UNSUPPORTED, "aura_ivf_coarse: d=%d too large", d);
  AURA_CUDA_OK(cudaFuncSetAttribute(coarse_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int gx = (n_queries + CO_QT - 1) / CO_QT;
  // split the centroid rows over gridDim.y so that small batches still fill the machine
  int gy = (2 * sm_count() + gx - 1) / gx;
  const int gy_max = (n_cent + 7) / 8;
  if (gy > gy_max) gy = gy_max;
  if (gy < 1) gy = 1;
  coarse_dist_kernel<<<dim3(gx, gy), 256, smem, st>>>(queries, n_queries, d, centroids, n_cent, dist);
  if (nprobe <= 32) coarse_select_kernel<1><<<n_queries, 128, 0, st>>>(dist, n_cent, nprobe, probes);
  else if (nprobe <= 64) coarse_select_kernel<2><<<n_queries, 128, 0, st>>>(dist, n_cent, nprobe, probes);
  else coarse_select_kernel<4><<<n_queries, 128, 0, st>>>(dist, n_cent, nprobe, probes);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(2);
  return AURA_OK;
}

// used by ivf_batch.cu
size_t ivf_coarse_ws_bytes(int n_queries, int d, int n_cent, int nprobe) { return coarse_ws_any(n_queries, d, n_cent, nprobe); }
int ivf_run_coarse(const float* queries, int n_queries, int d, const float* centroids, int n_cent, int nprobe,
                   long long* probes, void* workspace, cudaStream_t st) {
  return run_coarse(queries, n_queries, d, centroids, n_cent, nprobe, probes, workspace, st);
}

}  // namespace aura
using namespace aura;

extern "C" size_t aura_ivf_coarse_workspace_bytes(int n_queries, int d, int n_centroid_rows, int nprobe) {
  if (n_queries < 1 || n_centroid_rows < 1 || d < 1 || nprobe < 1) return 0;
  return coarse_ws_any(n_queries, d, n_centroid_rows, nprobe);
}

extern "C" int aura_ivf_coarse(const float* queries, int n_queries, int d, const float* centroids, int n_centroid_rows,
                               int nprobe, int64_t* probes, void* workspace, size_t workspace_bytes, void* stream) {
  AURA_REQUIRE(n_queries >= 1 && d >= 1 && n_centroid_rows >= 1, AURA_ERR_INVALID_ARG,
               "aura_ivf_coarse: n_queries=%d d=%d n_centroid_rows=%d", n_queries, d, n_centroid_rows);
  AURA_REQUIRE(nprobe >= 1 && nprobe <= AURA_MAX_NPROBE && nprobe <= n_centroid_rows, AURA_ERR_INVALID_ARG,
               "aura_ivf_coarse: nprobe=%d not in [1,min(%d,%d)]", nprobe, AURA_MAX_NPROBE, n_centroid_rows);
  AURA_REQUIRE(queries && centroids && probes && workspace, AURA_ERR_INVALID_ARG, "aura_ivf_coarse: null pointer");
  AURA_REQUIRE(workspace_bytes >= coarse_ws_any(n_queries, d, n_centroid_rows, nprobe), AURA_ERR_WORKSPACE,
               "aura_ivf_coarse: workspace too small");
  return run_coarse(queries, n_queries, d, centroids, n_centroid_rows, nprobe, reinterpret_cast<long long*>(probes),
                    workspace, (cudaStream_t)stream);
}

extern "C" size_t aura_ivf_search_workspace_bytes(int n_queries, int d, int n_centroid_rows, int nprobe, int k) {
  if (n_queries < 1 || n_centroid_rows < 1 || k < 1 || d < 1 || nprobe < 1) return 0;
  return coarse_ws_any(n_queries, d, n_centroid_rows, nprobe) + ((size_t)n_queries * AURA_MAX_NPROBE * 8 + 255) / 256 * 256 +
         scan_workspace_bytes(n_queries, k);
}

extern "C" int aura_ivf_search(const void* rows, int dtype, int64_t n_rows, int d, const float* queries, int n_queries,
                               const float* centroids, int n_centroid_rows, int nprobe, const int32_t* list_offsets,
                               const int32_t* list_rows, const float* scale, const float* bias, int k, int64_t row_base,
                               int flags, int64_t* out_idx, float* out_score, int64_t* out_probes, void* workspace,
                               size_t workspace_bytes, void* stream) {
  AURA_REQUIRE(dtype == AURA_F32 || dtype == AURA_BF16, AURA_ERR_INVALID_ARG, "aura_ivf_search: bad dtype %d", dtype);
  AURA_REQUIRE(n_rows >= 1 && n_rows < 0x7fffffffll && d >= 1 && n_queries >= 1 && n_centroid_rows >= 1,
               AURA_ERR_INVALID_ARG, "aura_ivf_search: n_rows=%lld d=%d n_queries=%d n_centroid_rows=%d",
               (long long)n_rows, d, n_queries, n_centroid_rows);
  AURA_REQUIRE(nprobe >= 1 && nprobe <= AURA_MAX_NPROBE && nprobe <= n_centroid_rows, AURA_ERR_INVALID_ARG,
               "aura_ivf_search: nprobe=%d not in [1,min(%d,%d)]", nprobe, AURA_MAX_NPROBE, n_centroid_rows);
  AURA_REQUIRE(k >= 1 && k <= AURA_MAX_K, AURA_ERR_INVALID_ARG, "aura_ivf_search: k=%d not in [1,%d]", k, AURA_MAX_K);
  AURA_REQUIRE(rows && queries && centroids && list_offsets && list_rows && out_idx && out_score && workspace,
               AURA_ERR_INVALID_ARG, "aura_ivf_search: null pointer");
  AURA_REQUIRE(workspace_bytes >= aura_ivf_search_workspace_bytes(n_queries, d, n_centroid_rows, nprobe, k), AURA_ERR_WORKSPACE,
               "aura_ivf_search: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  long long* probes = out_probes ? reinterpret_cast<long long*>(out_probes)
                                 : reinterpret_cast<long long*>(ws + coarse_ws_any(n_queries, d, n_centroid_rows, nprobe));
  const int rc = run_coarse(queries, n_queries, d, centroids, n_centroid_rows, nprobe, probes, ws, st);
  if (rc != AURA_OK) return rc;
  void* scan_ws = ws + coarse_ws_any(n_queries, d, n_centroid_rows, nprobe) + ((size_t)n_queries * AURA_MAX_NPROBE * 8 + 255) / 256 * 256;
  // expected candidates per query: nprobe lists of average length (plan sizing only)
  long long expect = (long long)((double)n_rows * nprobe / n_centroid_rows) + 1;
  return launch_scan(rows, dtype, n_rows, d, queries, n_queries, scale, bias, k, row_base,
                     reinterpret_cast<long long*>(out_idx), out_score, scan_ws, probes, nprobe, n_centroid_rows,
                     list_offsets, list_rows, expect, st, (flags & AURA_IVF_EMPTY_OK) ? 1 : 0);
}

/* libaura_hippo - C ABI of the B200-native episodic-memory retrieval path.
 *
 * The reference (auralmn/aura-snn-rag) has NO native / FFI interface: its boundary is the
 * Python class HippocampalFormation (src/core/hippocampal.py:31-377).  Each entry point
 * below names the reference statement(s) it replaces; the Python mirror of that class
 * (aura_snn_rag_b200/hippocampal.py) calls these through ctypes with tensor.data_ptr()
 * and the current CUDA stream handle.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - no entry point allocates: scratch comes from the caller (`*_workspace_bytes`);
 *   - no entry point synchronises the device; all work is enqueued on `stream`;
 *   - return value: 0 = ok, <0 = AURA_ERR_*; the message is in aura_last_error_string();
 *   - nothing throws across the boundary; no global mutable state except the
 *     thread-local error string and cached device attributes;
 *   - rows are row-major, contiguous (row stride == d); row indices are < 2^32 - 1;
 *   - top-k order: score descending, ties broken by LOWER row index; missing results
 *     (fewer than k candidates) are idx = -1, score = -inf.
 */
#ifndef AURA_HIPPO_H_
#define AURA_HIPPO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AURA_HIPPO_VERSION 100 /* 0.1.0 */

enum {
  AURA_OK = 0,
  AURA_ERR_INVALID_ARG = -1,
  AURA_ERR_UNSUPPORTED = -2,
  AURA_ERR_WORKSPACE = -3,
  AURA_ERR_CUDA = -4
};

/* element type of the memory bank rows */
enum { AURA_F32 = 0, AURA_BF16 = 1 };

#define AURA_MAX_K 128        /* largest k of any fused top-k */
#define AURA_MAX_NPROBE 128   /* largest nprobe of aura_ivf_search */

int aura_version(void);
const char* aura_last_error_string(void);

/* ---- per-row terms maintained at write time / per query -------------------------------- */

/* inv_norm[i] = 1 / max(||rows[i]||_2, 1e-12): the denominator F.normalize applies to every
 * bank row on every query (hippocampal.py:278), computed once per write instead. */
int aura_row_inv_norms(const void* rows, int dtype, int64_t n_rows, int d, float* inv_norm, void* stream);

/* Per-row affine terms of the combined score (hippocampal.py:282-303):
 *   scale[i] = 0.5 * strength_i * inv_norm[i]
 *   bias[i]  = (0.3 * spatial_i + 0.2 * exp(-(now - ts_i) / 3600)) * strength_i
 * with metadata[i] = {strength, timestamp, centroid id, 0} (hippocampal.py:97-99,215),
 * spatial_i = 1/(1+||loc_i - query_loc||) when query_loc != NULL else 0, so that
 *   combined_i = cos(q, row_i) * 0.5 * strength_i + bias_i = dot(q/||q||, row_i) * scale[i] + bias[i].
 * `now` is the fp32-rounded wall clock (the reference subtracts in fp32, :296). */
int aura_row_terms(const float* metadata, const float* locations, int spatial_dims, const float* query_loc,
                   float now, const float* inv_norm, int64_t n_rows, float* scale, float* bias, void* stream);

/* strength *= (1 - rate) over the live rows (hippocampal.py:334). */
int aura_decay_strength(float* metadata, int64_t n_rows, float rate, void* stream);

/* ---- exact scan: cosine + combine + top-k (hippocampal.py:272-307; .tmp_infer_old.py:40-49) ----
 * For each of n_queries fp32 queries [n_queries, d]:
 *   score_i = dot(q / max(||q||,1e-12), rows[i]) * scale[i] + (bias ? bias[i] : 0),  i < n_rows
 * and the k best (idx int64 [n_queries,k], score fp32 [n_queries,k]) are written.
 * Pure cosine (SimpleHippocampus.retrieve): scale = inv_norm, bias = NULL.
 * `row_base` is added to every returned index (row-sharded banks). 1 <= k <= AURA_MAX_K. */
size_t aura_scan_topk_workspace_bytes(int64_t n_rows, int d, int n_queries, int k);
int aura_scan_topk(const void* rows, int dtype, int64_t n_rows, int d, const float* queries, int n_queries,
                   const float* scale, const float* bias, int k, int64_t row_base, int64_t* out_idx,
                   float* out_score, void* workspace, size_t workspace_bytes, void* stream);

/* ---- k-way merge of per-shard top-k blocks (after the NCCL all-gather) ----------
 * in_score/in_idx: [n_queries, n_lists * k_in] (any order inside a row); out: [n_queries, k_out]. */
int aura_topk_merge(const float* in_score, const int64_t* in_idx, int n_queries, int n_lists, int k_in,
                    int k_out, float* out_score, int64_t* out_idx, void* stream);

/* Gather bank rows of a result block: out[b, j, :] = rows[idx[b, j]] (zeros when idx < 0), as fp32.
 * Replaces the per-result id_to_idx lookup + row copy of memory_augmented_layer.py:124-128. */
int aura_gather_rows(const void* rows, int dtype, int d, const int64_t* idx, int64_t n_idx, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AURA_HIPPO_H_ */

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

#define AURA_HIPPO_VERSION 300 /* 0.3.0: aura_ivf_search_batch / aura_ivf_pack_lists take the element type of the
                                  list-major copy (bf16 shadow of an fp32 bank), AURA_IVF_MEASURED_EPS */

enum {
  AURA_OK = 0,
  AURA_ERR_INVALID_ARG = -1,
  AURA_ERR_UNSUPPORTED = -2,
  AURA_ERR_WORKSPACE = -3,
  AURA_ERR_CUDA = -4
};

/* element type of the memory bank rows */
enum { AURA_F32 = 0, AURA_BF16 = 1 };

/* flags of aura_ivf_search / aura_ivf_search_batch */
#define AURA_IVF_EMPTY_OK 1   /* a query whose probed lists hold no LOCAL row returns no result (idx -1, score -inf)
                                 instead of scanning every row: row-sharded callers apply hippocampal.py:269-270 to
                                 the merged result, not per shard */
#define AURA_IVF_MEASURED_EPS 2 /* aura_ivf_search_batch over a bf16 bank: `eps` is the score-per-cosine unit
                                 (max |scale_r| * ||r||) and the certification bound is measured per query from the
                                 rounding error of the bf16 query copy (the bank rows are exact tensor-core operands) -
                                 about 1.7x tighter than the worst-case 2^-9 bound */

#define AURA_MAX_K 128        /* largest k of any fused top-k */
#define AURA_MAX_NPROBE 128   /* largest nprobe of aura_ivf_search */

int aura_version(void);
const char* aura_last_error_string(void);
/* number of CUDA kernels this library has launched in this process (bench accounting only) */
uint64_t aura_kernel_launches(void);
/* {tag, block, thread, parity} of the last watchdog trap of an instrumented kernel wait (the words live in mapped pinned
 * host memory, so they can be read after the trap has killed the CUDA context); all zero if none fired.
 * ivf_rows_kernel tags: 1 / 2 resident-mode producer (slabs free / ring slot free), 3 producer ring slot free, 4 MMA
 * slabs resident, 5 MMA accumulator drained, 6 MMA operands landed, 7 epilogue accumulator complete. */
int aura_debug_last_trap(uint32_t out[4]);

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

/* ---- memory-bank write (hippocampal.py:207-215) --------------------------------------------
 * Append/overwrite n_new consecutive bank rows starting at first_row from fp32 `features`
 * [n_new, d] (device): converts to the bank dtype, writes locations[row] = location (may be NULL
 * -> zeros), metadata[row] = {1.0, timestamp, -1, 0} and inv_norm[row]. One launch. */
int aura_bank_write(void* rows, int dtype, int d, int64_t first_row, int n_new, const float* features,
                    float* locations, int spatial_dims, const float* location, float* metadata, float timestamp,
                    float* inv_norm, void* stream);

/* ---- centroid index: build (hippocampal.py:345-377) -------------------------------------------
 * aura_kmeans_seed     centroids[s] = rows[seed_rows[s]]                      (:354-355; the caller
 *                      supplies randperm(M)[:k], CPU and CUDA generators differ)
 * aura_kmeans_assign   assign[i] = argmin_c ||rows[i] - centroid_c||  via ||c||^2 - 2 x.c, first
 *                      minimum on ties (:358-359,:370-371); optionally also writes the id as float
 *                      into cid_f32[i*cid_stride] (metadata[:,2], :376) and the best score.  Exact fp32 SIMT
 *                      tiles for small problems; when row_inv_norm (1/||row||, may be NULL) is given and
 *                      n_rows*n_centroids >= 1e8 the scores come from a tcgen05 GEMM (tf32 / bf16) and the
 *                      4 best candidates per row that lie within the rounding band 2^-7*||x||*max||c|| of
 *                      the best are re-scored in exact fp32 (exact whenever at most 4 centroids are that close).
 * aura_ivf_build_lists counting sort of the assignments into CSR inverted lists (replaces the
 *                      P float-equality mask passes + nonzero of :264-268).
 * aura_kmeans_list_sums / aura_kmeans_finalize   per-cluster mean, empty clusters keep their seed
 *                      (:360-363); sums are fp64 so a sharded build can all-reduce them.
 * aura_ivf_list_counts counts[c] = |list c| as fp32 (:372-374). */
int aura_kmeans_seed(const void* rows, int dtype, int d, const int64_t* seed_rows, int n_seeds, float* centroids,
                     void* stream);
size_t aura_kmeans_assign_workspace_bytes(int64_t n_rows, int d, int dtype, int n_centroids);
int aura_kmeans_assign(const void* rows, int dtype, int64_t n_rows, int d, const float* centroids, int n_centroids,
                       const float* row_inv_norm, int32_t* assign, float* cid_f32, int cid_stride, float* best_score,
                       void* workspace, size_t workspace_bytes, void* stream);
size_t aura_ivf_build_lists_workspace_bytes(int n_lists);
int aura_ivf_build_lists(const int32_t* cid, int64_t n_rows, int n_lists, int32_t* list_offsets, int32_t* list_rows,
                         void* workspace, size_t workspace_bytes, void* stream);
int aura_kmeans_list_sums(const void* rows, int dtype, int d, const int32_t* list_offsets, const int32_t* list_rows,
                          int n_lists, double* sums, int64_t* counts, void* stream);
int aura_kmeans_finalize(const double* sums, const int64_t* counts, int n_centroids, int d, float* centroids,
                         void* stream);
int aura_ivf_list_counts(const int32_t* list_offsets, int n_lists, float* counts_f32, void* stream);

/* ---- centroid index: one-shot writes (hippocampal.py:218-230) ------------------------------
 * For each of the n_writes rows first_row .. first_row+n_writes-1, IN ORDER:
 *   c* = argmin_{c < n_live} ||centroid_c - row||_2 (direct form, first minimum);
 *   counts[c*] += 1; eta = 1/max(counts[c*],1); centroid_c* = (1-eta) centroid_c* + eta row;
 *   cid_i32[row] = c*; cid_f32[row*cid_stride] = c*.
 * Sequential semantics are preserved (write i sees the centroid moved by write i-1). */
size_t aura_online_assign_workspace_bytes(void);
int aura_online_assign(const void* rows, int dtype, int d, int64_t first_row, int n_writes, float* centroids,
                       int n_live, float* counts, int32_t* cid_i32, float* cid_f32, int cid_stride, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ---- centroid index: query (hippocampal.py:257-307) ------------------------------------------
 * aura_ivf_coarse: probes[b, 0..nprobe) = the nprobe centroid rows nearest to query b by
 *   ||centroid_c - q||_2 over ALL n_centroid_rows rows of the buffer (zeroed tail rows included,
 *   as :261 does), nearest first, ties to the lower row (:262).  Blocks of >= 64 queries are scored as one
 *   TF32 tcgen05 GEMM (2 q.c - ||c||^2, nprobe <= 32); smaller blocks in exact fp32 difference form.
 * aura_ivf_search: coarse + scan of the probed inverted lists with the same score/top-k as
 *   aura_scan_topk.  A query whose probed lists are all empty scans every row (:269-270) unless
 *   flags & AURA_IVF_EMPTY_OK.
 *   Returned indices are bank rows (the reference returns candidate-local positions, :307-317:
 *   a documented bug this library does not reproduce). */
size_t aura_ivf_coarse_workspace_bytes(int n_queries, int d, int n_centroid_rows, int nprobe);
int aura_ivf_coarse(const float* queries, int n_queries, int d, const float* centroids, int n_centroid_rows, int nprobe,
                    int64_t* probes, void* workspace, size_t workspace_bytes, void* stream);
size_t aura_ivf_search_workspace_bytes(int n_queries, int d, int n_centroid_rows, int nprobe, int k);
int aura_ivf_search(const void* rows, int dtype, int64_t n_rows, int d, const float* queries, int n_queries,
                    const float* centroids, int n_centroid_rows, int nprobe, const int32_t* list_offsets,
                    const int32_t* list_rows, const float* scale, const float* bias, int k, int64_t row_base, int flags,
                    int64_t* out_idx, float* out_score, int64_t* out_probes, void* workspace, size_t workspace_bytes,
                    void* stream);

/* Block-of-queries form of aura_ivf_search: the (query, probe) pairs are sorted by list and every probed list is
 * scored ONCE per group of <= 128 queries probing it (ragged grouped GEMM on tcgen05; query rows and bank rows gathered
 * by row id with 16-byte cp.async into the swizzled operand layout), so a batch reads each probed list once instead of
 * once per query.  rows_by_list (may be NULL): a resident copy of the bank in CSR order made by aura_ivf_pack_lists for
 * these list_offsets / list_rows - list tiles are then streamed as TMA boxes instead of gathered.  Results: exact fp32
 * scores of the best candidates (same re-score + certification as aura_batch_topk, read from `rows`).  lm_dtype is the
 * element type of rows_by_list: the bank's, or AURA_BF16 for a bf16 SHADOW of an fp32 bank - the list tiles are then
 * half the bytes and run at the bf16 tensor rate, `eps` is the score-per-cosine unit and the certification bound is
 * measured per query from shadow_relerr (device scalar kept by aura_ivf_pack_lists), exactly as in aura_batch_topk
 * (flags & AURA_IVF_MEASURED_EPS asks for the measured bound over a bf16 bank);
 * out_uncertain[b] = 1 hands query b back to aura_ivf_search (uncertified result, no candidates, or work table
 * overflow).  k <= 114, d*sizeof(elem) % 16 == 0. */
size_t aura_ivf_search_batch_workspace_bytes(int n_queries, int d, int n_centroid_rows, int nprobe, int k);
int aura_ivf_search_batch(const void* rows, int dtype, int64_t n_rows, int d, const float* queries, int n_queries,
                          const float* centroids, int n_centroid_rows, int nprobe, const int32_t* list_offsets,
                          const int32_t* list_rows, const void* rows_by_list, int lm_dtype, const float* shadow_relerr,
                          const float* scale, const float* bias, int k, int64_t row_base, int flags, float eps,
                          int64_t* out_idx, float* out_score, int32_t* out_uncertain, void* workspace,
                          size_t workspace_bytes, void* stream);

/* rows_by_list[p] = rows[list_rows[p]] for p < n_listed (pitch d; out_dtype = dtype, or AURA_BF16 for a bf16 shadow of an
 * fp32 bank, in which case relerr_max - DEVICE float, zero it first - is raised to the largest relative rounding error
 * of the rows packed): the list-major resident copy of the bank
 * (the layout an inverted-file index normally stores; the reference keeps insertion order only, hippocampal.py:211).
 * Rebuild it whenever aura_ivf_build_lists ran. */
int aura_ivf_pack_lists(const void* rows, int dtype, int d, const int32_t* list_rows, int64_t n_listed, void* rows_by_list,
                        int out_dtype, float* relerr_max, void* stream);

/* diagnostics of the last aura_ivf_search_batch call on `workspace`, written to the DEVICE words items_out[0..5) by an
 * enqueued kernel (like every entry point it does not synchronise): {work items, result slots used, most slots linked by one
 * query, queries flagged inside the kernel, queries without any candidate}; host_cap = work-table capacity */
int aura_ivf_search_batch_items(const void* workspace, int n_queries, int d, int n_centroid_rows, int nprobe, int k,
                                int32_t* items_out, int32_t* host_cap, void* stream);

/* ---- batched exact search on the tensor cores (the batch the reference loops over one query at a
 * time, memory_augmented_layer.py:113-128; score of hippocampal.py:272-307) -------------------
 * Same result contract as aura_scan_topk.  tcgen05 (tf32 from an fp32 bank, bf16 from a bf16 bank) scores
 * every (query,row) pair and keeps a per-query shortlist; the shortlist is re-scored in exact fp32 and the
 * result certified: out_uncertain[b] = 0 means the top-k of query b is provably the exact fp32 top-k given
 * that tensor-core scores are within `eps` (score units) of the exact ones; 1 means the caller must re-run
 * query b through aura_scan_topk.  A query whose first re-score (the shortlist) cannot be certified is given a second
 * chance inside the call: up to 256 of the candidates the pass is known to hold completely are re-scored before the
 * flag is raised.  Needs d*sizeof(elem) % 16 == 0 and k <= 114.
 * shadow_bf16 (may be NULL): a bf16 copy of an fp32 bank (aura_rows_to_bf16, same row order).  The shortlist pass then
 * runs on the copy (kind::f16: half the bytes, twice the tensor rate, 24 / 32 / 48 candidates per query for k <= 10 / 18 / 34) and the re-score
 * still reads the fp32 rows, so certified results are the same exact fp32 top-k; `eps` must then bound the bf16 rounding
 * (2^-7 per unit of |scale * ||row|||) - or, with shadow_relerr (DEVICE scalar >= ||bf16(r) - r|| / ||r|| over the bank rows,
 * maintained by aura_rows_to_bf16), `eps` is just that unit, max |scale_r| * ||r||, and the bound is measured per query:
 * unit * ((1 + e_q) * relerr + e_q + 1e-4) with e_q the rounding error norm of the normalised query. */
size_t aura_batch_topk_workspace_bytes(int64_t n_rows, int d, int dtype, int n_queries, int k);
int aura_batch_topk(const void* rows, int dtype, int64_t n_rows, int d, const float* queries, int n_queries,
                    const float* scale, const float* bias, int k, int64_t row_base, float eps, const void* shadow_bf16,
                    const float* shadow_relerr, int64_t* out_idx, float* out_score, int32_t* out_uncertain, void* workspace,
                    size_t workspace_bytes, void* stream);
/* out_bf16[i] = bf16(rows[i]) (round to nearest even) for n_rows rows: builds / refreshes the shadow above.  relerr_max
 * (DEVICE float, may be NULL; zero it before the first call) is raised to the largest relative rounding error
 * ||bf16(r) - r|| / ||r|| of the rows converted. */
int aura_rows_to_bf16(const float* rows, int64_t n_rows, int d, void* out_bf16, float* relerr_max, void* stream);

/* ---- cognitive map: all-pairs cosine + top-k neighbours per row, self excluded ---------------------
 * (training_recipes.md:292-308; README.md:39,64 - documented upstream, never implemented.)
 * For the A block rows[a_row_first, a_row_first+n_a_rows): out_idx/out_score [n_a_rows, k] = the k rows of
 * the whole bank with the largest cos(row_i, row_j) = dot * inv_norm[i] * inv_norm[j], j != i, best first.
 * The A block argument is what a row-sharded build passes per GPU. k <= 64. */
size_t aura_allpairs_topk_workspace_bytes(int64_t n_a_rows, int64_t n_rows, int d, int dtype, int k);
int aura_allpairs_topk(const void* rows, int dtype, int64_t n_rows, int d, int64_t a_row_first, int64_t n_a_rows,
                       const float* inv_norm, int k, int64_t* out_idx, float* out_score, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ---- k-way merge of per-shard top-k blocks (after the NCCL all-gather) ----------
 * in_score/in_idx: [n_queries, n_lists * k_in] (any order inside a row); out: [n_queries, k_out]. */
int aura_topk_merge(const float* in_score, const int64_t* in_idx, int n_queries, int n_lists, int k_in,
                    int k_out, float* out_score, int64_t* out_idx, void* stream);

/* Sharded search, one collective per step: aura_pack_topk writes what a rank contributes to the all-gather,
 * payload[b] = { id[b][0..k), score bits[b][0..k), flag[b] } as int64 [n_queries, 2k+1], where id = id_map[idx] when id_map
 * is given (local row -> global memory id of a shard that took online writes), else idx + id_base; idx < 0 stays -1.
 * aura_topk_merge_packed merges the rank-major gathered block [n_ranks, n_queries, 2k+1] (same order rule as
 * aura_topk_merge) and ORs the flags. */
int aura_pack_topk(const int64_t* idx, const float* score, const int32_t* flags, int n_queries, int k,
                   const int64_t* id_map, int64_t id_base, int64_t* payload, void* stream);
int aura_topk_merge_packed(const int64_t* gathered, int n_ranks, int n_queries, int k, float* out_score, int64_t* out_idx,
                           int32_t* any_flag, void* stream);

/* The same exchange over peer memory instead of a library collective (NVLink / NVSwitch, symmetric allocation): rank r
 * owns a gather buffer of aura_peer_gather_buffer_bytes() bytes that every peer has mapped (zero it once);
 * aura_pack_scatter builds this rank's payload and stores it into EVERY rank's buffer (peer_bufs_host[r], host array of
 * device pointers in rank order) and raises its flag there; aura_merge_gathered waits, on the device, until every rank's
 * flag for the current step is up and merges what landed in the local buffer.  `counters`: 4 zero-initialised device
 * words private to this rank (the step numbers live there, so consecutive calls - and CUDA-graph replays - need no
 * changing argument).  Every rank must issue the same sequence of pack / merge calls. */
size_t aura_peer_gather_buffer_bytes(int n_ranks, int n_queries, int k);
int aura_pack_scatter(const int64_t* idx, const float* score, const int32_t* flags, int n_queries, int k,
                      const int64_t* id_map, int64_t id_base, void* const* peer_bufs_host, int rank, int n_ranks,
                      uint32_t* counters, void* stream);
int aura_merge_gathered(const void* gather_buf, int n_ranks, int n_queries, int k, uint32_t* counters, float* out_score,
                        int64_t* out_idx, int32_t* any_flag, void* stream);

/* Gather bank rows of a result block: out[b, j, :] = rows[idx[b, j]] (zeros when idx < 0), as fp32.
 * Replaces the per-result id_to_idx lookup + row copy of memory_augmented_layer.py:124-128. */
int aura_gather_rows(const void* rows, int dtype, int d, const int64_t* idx, int64_t n_idx, float* out, void* stream);

/* Memory-injection context of MemoryAugmentedLayer.inject_memories ("concat" / "gate" modes,
 * memory_augmented_layer.py:185-188,192-193), fused with the row gather:
 *   w[b, :] = softmax(score[b, :]) with a missing result (idx < 0) entering as score 0 / zero row - the zero padding of
 *   retrieve_memories (:113-130) - and context[b, :] = sum_j w[b, j] * rows[idx[b, j]]   (fp32, [n_queries, d]).
 * weights (may be NULL) receives w [n_queries, k].  k <= AURA_MAX_K. */
int aura_gather_context(const void* rows, int dtype, int d, const int64_t* idx, const float* score, int n_queries, int k,
                        float* context, float* weights, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AURA_HIPPO_H_ */

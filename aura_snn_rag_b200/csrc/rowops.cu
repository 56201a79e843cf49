// Per-row terms of the memory bank: inverse norms, score affine terms, strength decay, row gather.
// These are the O(M) elementwise pieces of hippocampal.py:272-303,334 kept OUT of the per-query
// feature traffic: 8 B/row instead of re-normalising d*4 B/row on every query.
#include "aura_common.cuh"

namespace aura {

// sum of squares of bank row r, one warp, fixed lane-strided order: the ONLY arithmetic that produces inv_norm, so the
// value written at insert time (aura_bank_write) and a later recomputation (aura_row_inv_norms) agree bit for bit
template <bool BF16>
__device__ __forceinline__ float row_sumsq_warp(const void* rows, long long r, int d, int lane) {
  float ss = 0.f;
  if (BF16) {
    const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(rows) + (size_t)r * d;
    if ((d & 7) == 0) {
      const uint4* x8 = reinterpret_cast<const uint4*>(x);
      for (int c = lane; c < (d >> 3); c += 32) {
        const uint4 v = x8[c];
        const unsigned u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float a = bf16_lo(u[i]), b = bf16_hi(u[i]); ss = fmaf(a, a, fmaf(b, b, ss)); }
      }
    } else {
      for (int e = lane; e < d; e += 32) { const float v = __bfloat162float(x[e]); ss = fmaf(v, v, ss); }
    }
  } else {
    const float* x = reinterpret_cast<const float*>(rows) + (size_t)r * d;
    if ((d & 3) == 0) {
      const float4* x4 = reinterpret_cast<const float4*>(x);
      for (int c = lane; c < (d >> 2); c += 32) {
        const float4 v = x4[c];
        ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
      }
    } else {
      for (int e = lane; e < d; e += 32) ss = fmaf(x[e], x[e], ss);
    }
  }
  return warp_sum(ss);
}

template <bool BF16>
__global__ void __launch_bounds__(256) inv_norm_kernel(const void* rows, long long n_rows, int d, float* inv_norm) {
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = w; r < n_rows; r += nw) {
    const float ss = row_sumsq_warp<BF16>(rows, r, d, lane);
    if (lane == 0) inv_norm[r] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps, hippocampal.py:278
  }
}

__global__ void __launch_bounds__(256) row_terms_kernel(const float4* __restrict__ metadata,
                                                        const float* __restrict__ locations, int sd,
                                                        const float* __restrict__ query_loc, float now,
                                                        const float* __restrict__ inv_norm, long long n,
                                                        float* __restrict__ scale, float* __restrict__ bias) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float4 md = metadata[i];  // {strength, timestamp, centroid id, 0}
    float spatial = 0.f;
    if (query_loc != nullptr) {     // hippocampal.py:287-289
      float ss = 0.f;
      for (int j = 0; j < sd; ++j) { const float t = locations[(size_t)i * sd + j] - query_loc[j]; ss = fmaf(t, t, ss); }
      spatial = 1.0f / (1.0f + sqrtf(ss));
    }
    const float age = now - md.y;                       // fp32 subtraction, :296
    const float temporal = expf(-age / 3600.0f);        // :297
    scale[i] = 0.5f * md.x * inv_norm[i];
    bias[i] = (0.3f * spatial + 0.2f * temporal) * md.x;  // :301-303
  }
}

__global__ void decay_kernel(float* metadata, long long n, float keep) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    metadata[i * 4] *= keep;
}

template <bool BF16>
__global__ void __launch_bounds__(256) gather_rows_kernel(const void* rows, int d, const long long* idx, long long n_idx,
                                                          float* out) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_idx) return;
  const long long r = idx[w];
  float* o = out + (size_t)w * d;
  if (r < 0) { for (int e = lane; e < d; e += 32) o[e] = 0.f; return; }
  if (BF16) {
    const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(rows) + (size_t)r * d;
    for (int e = lane; e < d; e += 32) o[e] = __bfloat162float(x[e]);
  } else {
    const float* x = reinterpret_cast<const float*>(rows) + (size_t)r * d;
    for (int e = lane; e < d; e += 32) o[e] = x[e];
  }
}

// One CTA per new row: convert + store the features, compute the inverse norm from the values AS
// STORED (bf16 banks normalise their rounded rows), write location and metadata {1, t, -1, 0}.
template <bool BF16>
__global__ void __launch_bounds__(256) bank_write_kernel(void* rows, int d, long long first_row,
                                                         const float* __restrict__ feats, float* locations, int sd,
                                                         const float* __restrict__ location, float* metadata,
                                                         float timestamp, float* inv_norm) {
  const long long r = first_row + blockIdx.x;
  const float* f = feats + (size_t)blockIdx.x * d;
  for (int e = threadIdx.x; e < d; e += blockDim.x) {
    if (BF16) reinterpret_cast<__nv_bfloat16*>(rows)[(size_t)r * d + e] = __float2bfloat16_rn(f[e]);
    else reinterpret_cast<float*>(rows)[(size_t)r * d + e] = f[e];
  }
  __syncthreads();                                   // the stored row is visible to the whole CTA
  if (threadIdx.x < 32) {                            // norm of the values AS STORED (bf16 banks normalise their rounded rows)
    const float ss = row_sumsq_warp<BF16>(rows, r, d, threadIdx.x);
    if (threadIdx.x == 0) {
      inv_norm[r] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
      reinterpret_cast<float4*>(metadata)[r] = make_float4(1.0f, timestamp, -1.0f, 0.0f);  // :215,:232
    }
  }
  if (locations != nullptr && threadIdx.x < sd)
    locations[(size_t)r * sd + threadIdx.x] = location ? location[threadIdx.x] : 0.f;  // :212
}

static int grid_for(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * 16;
  return (int)(g < 1 ? 1 : g > cap ? cap : g);
}

}  // namespace aura
using namespace aura;

extern "C" int aura_row_inv_norms(const void* rows, int dtype, int64_t n_rows, int d, float* inv_norm, void* stream) {
  AURA_REQUIRE(dtype == AURA_F32 || dtype == AURA_BF16, AURA_ERR_INVALID_ARG, "aura_row_inv_norms: bad dtype %d", dtype);
  AURA_REQUIRE(n_rows >= 0 && d >= 1, AURA_ERR_INVALID_ARG, "aura_row_inv_norms: n_rows=%lld d=%d", (long long)n_rows, d);
  if (n_rows == 0) return AURA_OK;
  AURA_REQUIRE(rows && inv_norm, AURA_ERR_INVALID_ARG, "aura_row_inv_norms: null pointer");
  const int g = grid_for(n_rows, 8);
  if (dtype == AURA_BF16) inv_norm_kernel<true><<<g, 256, 0, (cudaStream_t)stream>>>(rows, n_rows, d, inv_norm);
  else inv_norm_kernel<false><<<g, 256, 0, (cudaStream_t)stream>>>(rows, n_rows, d, inv_norm);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}

extern "C" int aura_row_terms(const float* metadata, const float* locations, int spatial_dims, const float* query_loc,
                              float now, const float* inv_norm, int64_t n_rows, float* scale, float* bias, void* stream) {
  AURA_REQUIRE(n_rows >= 0, AURA_ERR_INVALID_ARG, "aura_row_terms: n_rows=%lld", (long long)n_rows);
  if (n_rows == 0) return AURA_OK;
  AURA_REQUIRE(metadata && inv_norm && scale && bias, AURA_ERR_INVALID_ARG, "aura_row_terms: null pointer");
  AURA_REQUIRE(query_loc == nullptr || (locations != nullptr && spatial_dims >= 1), AURA_ERR_INVALID_ARG,
               "aura_row_terms: query_loc given without locations");
  AURA_REQUIRE((reinterpret_cast<uintptr_t>(metadata) & 15) == 0, AURA_ERR_INVALID_ARG,
               "aura_row_terms: metadata must be 16-byte aligned");
  row_terms_kernel<<<grid_for(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(metadata), locations, spatial_dims, query_loc, now, inv_norm, n_rows, scale, bias);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}

extern "C" int aura_decay_strength(float* metadata, int64_t n_rows, float rate, void* stream) {
  AURA_REQUIRE(n_rows >= 0, AURA_ERR_INVALID_ARG, "aura_decay_strength: n_rows=%lld", (long long)n_rows);
  if (n_rows == 0) return AURA_OK;
  AURA_REQUIRE(metadata, AURA_ERR_INVALID_ARG, "aura_decay_strength: null pointer");
  decay_kernel<<<grid_for(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(metadata, n_rows, 1.0f - rate);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}

extern "C" int aura_gather_rows(const void* rows, int dtype, int d, const int64_t* idx, int64_t n_idx, float* out,
                                void* stream) {
  AURA_REQUIRE(dtype == AURA_F32 || dtype == AURA_BF16, AURA_ERR_INVALID_ARG, "aura_gather_rows: bad dtype %d", dtype);
  if (n_idx <= 0) return AURA_OK;
  AURA_REQUIRE(rows && idx && out && d >= 1, AURA_ERR_INVALID_ARG, "aura_gather_rows: null pointer / d");
  const int g = (int)((n_idx + 7) / 8);
  if (dtype == AURA_BF16)
    gather_rows_kernel<true><<<g, 256, 0, (cudaStream_t)stream>>>(rows, d, reinterpret_cast<const long long*>(idx), n_idx, out);
  else
    gather_rows_kernel<false><<<g, 256, 0, (cudaStream_t)stream>>>(rows, d, reinterpret_cast<const long long*>(idx), n_idx, out);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}

// ---- memory injection context (memory_augmented_layer.py:185-188,192-193): for every query b,
//   w = softmax(scores[b, 0..k))  (a missing result counts with score 0 and a zero row, exactly as the zero-padded
//   tensors of retrieve_memories, :113-130, enter the reference's softmax),  context[b] = sum_j w_j * rows[idx[b, j]]
// fused with the row gather: the [B, k, d] feature block is never materialised.  One CTA per query.
namespace aura {
template <bool BF16>
__global__ void __launch_bounds__(256) gather_context_kernel(const void* __restrict__ rows, int d, const long long* __restrict__ idx,
                                                             const float* __restrict__ score, int k, float* __restrict__ context,
                                                             float* __restrict__ weights) {
  __shared__ float w_s[AURA_MAX_K];
  __shared__ long long r_s[AURA_MAX_K];
  const int b = blockIdx.x;
  if (threadIdx.x < 32) {
    // softmax over k <= 128 scores by one warp, in the order torch's softmax uses: max, exp(x - max), sum, divide
    float v[AURA_MAX_K / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < AURA_MAX_K / 32; ++j) {
      const int i = threadIdx.x + 32 * j;
      long long r = -1; float sc = -INFINITY;
      if (i < k) { r = idx[(size_t)b * k + i]; sc = r >= 0 ? score[(size_t)b * k + i] : 0.f; r_s[i] = r; }
      v[j] = sc;
      mx = fmaxf(mx, sc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < AURA_MAX_K / 32; ++j) { v[j] = (threadIdx.x + 32 * j) < k ? expf(v[j] - mx) : 0.f; sum += v[j]; }
    sum = warp_sum(sum);
#pragma unroll
    for (int j = 0; j < AURA_MAX_K / 32; ++j) {
      const int i = threadIdx.x + 32 * j;
      if (i < k) { const float w = v[j] / sum; w_s[i] = w; if (weights) weights[(size_t)b * k + i] = w; }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < d; e += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < k; ++j) {
      const long long r = r_s[j];
      if (r < 0) continue;
      const float x = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rows)[(size_t)r * d + e])
                           : reinterpret_cast<const float*>(rows)[(size_t)r * d + e];
      acc = fmaf(w_s[j], x, acc);
    }
    context[(size_t)b * d + e] = acc;
  }
}
}  // namespace aura

extern "C" int aura_gather_context(const void* rows, int dtype, int d, const int64_t* idx, const float* score, int n_queries,
                                   int k, float* context, float* weights, void* stream) {
  AURA_REQUIRE(dtype == AURA_F32 || dtype == AURA_BF16, AURA_ERR_INVALID_ARG, "aura_gather_context: bad dtype %d", dtype);
  if (n_queries <= 0) return AURA_OK;
  AURA_REQUIRE(rows && idx && score && context && d >= 1 && k >= 1 && k <= AURA_MAX_K, AURA_ERR_INVALID_ARG,
               "aura_gather_context: null pointer / d=%d k=%d", d, k);
  if (dtype == AURA_BF16)
    gather_context_kernel<true><<<n_queries, 256, 0, (cudaStream_t)stream>>>(rows, d, reinterpret_cast<const long long*>(idx), score, k, context, weights);
  else
    gather_context_kernel<false><<<n_queries, 256, 0, (cudaStream_t)stream>>>(rows, d, reinterpret_cast<const long long*>(idx), score, k, context, weights);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}

extern "C" int aura_bank_write(void* rows, int dtype, int d, int64_t first_row, int n_new, const float* features,
                               float* locations, int spatial_dims, const float* location, float* metadata,
                               float timestamp, float* inv_norm, void* stream) {
  AURA_REQUIRE(dtype == AURA_F32 || dtype == AURA_BF16, AURA_ERR_INVALID_ARG, "aura_bank_write: bad dtype %d", dtype);
  AURA_REQUIRE(d >= 1 && first_row >= 0 && n_new >= 0 && spatial_dims >= 0 && spatial_dims <= 256, AURA_ERR_INVALID_ARG,
               "aura_bank_write: d=%d first_row=%lld n_new=%d spatial_dims=%d", d, (long long)first_row, n_new,
               spatial_dims);
  if (n_new == 0) return AURA_OK;
  AURA_REQUIRE(rows && features && metadata && inv_norm, AURA_ERR_INVALID_ARG, "aura_bank_write: null pointer");
  AURA_REQUIRE((reinterpret_cast<uintptr_t>(metadata) & 15) == 0, AURA_ERR_INVALID_ARG,
               "aura_bank_write: metadata must be 16-byte aligned");
  if (dtype == AURA_BF16)
    bank_write_kernel<true><<<n_new, 256, 0, (cudaStream_t)stream>>>(rows, d, first_row, features, locations,
                                                                     spatial_dims, location, metadata, timestamp, inv_norm);
  else
    bank_write_kernel<false><<<n_new, 256, 0, (cudaStream_t)stream>>>(rows, d, first_row, features, locations,
                                                                      spatial_dims, location, metadata, timestamp, inv_norm);
  AURA_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return AURA_OK;
}

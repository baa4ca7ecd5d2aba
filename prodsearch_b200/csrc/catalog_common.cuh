// Shared definitions of the catalog-scoring (G5) translation units.
#pragma once
#include "psb_common.cuh"

namespace psb {

constexpr int kKMax = 128;  // largest supported k (reference writes cutoff = 100, trainer.py:140)

// (s, id) ranks before (ts, tid): higher score, ties -> lower item id
__device__ __forceinline__ bool beats(float s, int64_t id, float ts, int64_t tid) {
  return s > ts || (s == ts && id < tid);
}

// strict total order with a position tie-break (only empty slots share (score, id))
__device__ __forceinline__ bool before(float su, int64_t iu, int pu, float se, int64_t ie, int pe) {
  return su > se || (su == se && (iu < ie || (iu == ie && pu < pe)));
}

// canonical fp32 score recurrence (see catalog_topk.cu header)
__device__ __forceinline__ float canonical_dot(const float* __restrict__ q, const float* __restrict__ e, int d) {
  float acc = 0.f;
  const float4* e4 = reinterpret_cast<const float4*>(e);   // rows are 16-byte aligned, d % 4 == 0
  for (int c = 0; c < d; c += 4) {
    const float4 v = __ldg(e4 + (c >> 2));
    acc = fmaf(q[c], v.x, acc);
    acc = fmaf(q[c + 1], v.y, acc);
    acc = fmaf(q[c + 2], v.z, acc);
    acc = fmaf(q[c + 3], v.w, acc);
  }
  return acc;
}

int64_t exact_workspace_bytes(int64_t m, int64_t n_items);
int catalog_topk_exact(const float* queries, int64_t m, const float* table, int64_t n_items, int64_t d,
                       const float* bias, int64_t k, int64_t id_base, int64_t id_stride, void* workspace,
                       int64_t workspace_bytes, int64_t* out_ids, float* out_scores, cudaStream_t s);
int64_t tc_workspace_bytes(int64_t m, int64_t n_items, int64_t d, int64_t k);
int catalog_topk_tc(const float* queries, int64_t m, const float* table, int64_t n_items, int64_t d,
                    const float* bias, int64_t k, int64_t id_base, int64_t id_stride, const float* max_row_sqnorm,
                    void* workspace, int64_t workspace_bytes, int64_t* out_ids, float* out_scores, cudaStream_t s);
int table_max_row_sqnorm(const float* table, int64_t rows, int64_t d, float* out, cudaStream_t s);
int64_t tc16_workspace_bytes(int64_t m, int64_t n_items, int64_t d, int64_t k);
int catalog_prepare_f16(const float* table, int64_t n_items, int64_t d, void* table_f16, float* stats, cudaStream_t s);
int catalog_topk_tc16(const float* queries, int64_t m, const float* table, const void* table_f16, const float* stats,
                      int64_t n_items, int64_t d, const float* bias, int64_t k, int64_t id_base, int64_t id_stride,
                      void* workspace, int64_t workspace_bytes, int64_t* out_ids, float* out_scores, cudaStream_t s);

}  // namespace psb

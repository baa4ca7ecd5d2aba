// G5 (exact mode) : full-catalog scoring + top-k with reference (fp32) arithmetic on CUDA cores,
// plus the shard merge.  The tcgen05 / TMA shortlist path lives in catalog_tc.cu and reuses the
// canonical score and the ordering rule defined here.
//
// Canonical score of (query q, item e):  acc = 0; for c in 0..d-1: acc = fmaf(q[c], e[c], acc);
// score = acc + bias.  Every kernel that produces a final score uses exactly this recurrence so
// EXACT and TC modes return bit-identical (id, score) lists.
// Ordering rule: descending score, ties by ascending item id (SURVEY.md 0.7 contract).
#include <float.h>

#include "catalog_common.cuh"

namespace psb {

// ---------------------------------------------------------------------------------------------
// scores[q, j] for q < m, j < cn (chunk of items [c0, c0+cn)): register-tiled fp32 "GEMM" whose
// per-output accumulation order is c ascending (the canonical recurrence).
// Block tile 64 queries x 128 items, 256 threads, 4 x 8 outputs per thread, K staged 16 at a time.
// ---------------------------------------------------------------------------------------------
constexpr int kTQ = 64, kTI = 128, kTK = 16;

__global__ void __launch_bounds__(256)
exact_scores_kernel(const float* __restrict__ Q, int m, const float* __restrict__ E, int64_t c0, int cn,
                    int d, const float* __restrict__ bias, float* __restrict__ S, int ld_s) {
  __shared__ float sq[kTK][kTQ + 4];
  __shared__ float se[kTK][kTI + 4];
  const int tq = threadIdx.x >> 4;   // 0..15 -> queries tq*4 .. +3
  const int ti = threadIdx.x & 15;   // 0..15 -> items ti*8 .. +7  (strided by 16 below)
  const int q0 = blockIdx.y * kTQ;
  const int i0 = blockIdx.x * kTI;
  float acc[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  for (int k0 = 0; k0 < d; k0 += kTK) {
    // stage Q[q0.., k0..] and E[i0.., k0..] transposed into shared memory
    for (int t = threadIdx.x; t < kTQ * kTK; t += 256) {
      const int r = t / kTK, c = t % kTK;
      sq[c][r] = (q0 + r < m && k0 + c < d) ? Q[static_cast<int64_t>(q0 + r) * d + k0 + c] : 0.f;
    }
    for (int t = threadIdx.x; t < kTI * kTK; t += 256) {
      const int r = t / kTK, c = t % kTK;
      se[c][r] = (i0 + r < cn && k0 + c < d) ? E[(c0 + i0 + r) * d + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < kTK; ++c) {
      float a[4], b[8];
#pragma unroll
      for (int x = 0; x < 4; ++x) a[x] = sq[c][tq * 4 + x];
#pragma unroll
      for (int y = 0; y < 8; ++y) b[y] = se[c][ti + 16 * y];
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 8; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int q = q0 + tq * 4 + x;
    if (q >= m) continue;
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      const int i = i0 + ti + 16 * y;
      if (i < cn) S[static_cast<int64_t>(q) * ld_s + i] = acc[x][y] + (bias != nullptr ? bias[c0 + i] : 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Streaming exact top-k of one row: one CTA per query keeps a sorted best[kKMax] list in global
// memory across chunks and rank-merges the (few) scores that beat its current k-th entry.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
row_select_kernel(const float* __restrict__ S, int ld_s, int64_t c0, int cn, int k,
                  float* __restrict__ best_s, int32_t* __restrict__ best_i) {
  __shared__ float bs[kKMax];
  __shared__ int32_t bi[kKMax];
  __shared__ float cs[kKMax];
  __shared__ int32_t ci[kKMax];
  __shared__ float ns[kKMax];
  __shared__ int32_t ni[kKMax];
  __shared__ int ccount;
  __shared__ int scan_sm[8];
  const int q = blockIdx.x;
  const float* row = S + static_cast<int64_t>(q) * ld_s;
  if (threadIdx.x < kKMax) {
    bs[threadIdx.x] = best_s[q * kKMax + threadIdx.x];
    bi[threadIdx.x] = best_i[q * kKMax + threadIdx.x];
  }
  if (threadIdx.x == 0) ccount = 0;
  __syncthreads();
  for (int base = 0; base < cn; base += 256) {
    const int j = base + threadIdx.x;
    // current k-th entry (empty slots hold (-inf, INT_MAX) and lose to everything)
    const float ts = bs[k - 1];
    const int32_t tid_ = bi[k - 1];
    bool pass = false;
    float s = 0.f;
    int32_t id = 0;
    if (j < cn) {
      s = row[j];
      id = static_cast<int32_t>(c0 + j);
      pass = beats(s, id, ts, tid_);
    }
    // deterministic compaction of the passing scores (ascending j)
    const unsigned bal = __ballot_sync(kFull, pass);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) scan_sm[wid] = __popc(bal);
    __syncthreads();
    int nbefore = 0, total = 0;
    for (int w2 = 0; w2 < 8; ++w2) {
      const int c = scan_sm[w2];
      if (w2 < wid) nbefore += c;
      total += c;
    }
    if (total == 0) {
      __syncthreads();
      continue;
    }
    // candidates are merged in pieces of at most kKMax
    const int my = nbefore + __popc(bal & ((1u << lane) - 1u));
    for (int piece = 0; piece < total; piece += kKMax) {
      if (pass && my >= piece && my < piece + kKMax) {
        cs[my - piece] = s;
        ci[my - piece] = id;
      }
      const int pc = min(kKMax, total - piece);
      __syncthreads();
      // rank merge of best[0..kKMax) and cand[0..pc): element e goes to position #elements before it
      const int t = threadIdx.x;
      if (t < kKMax + pc) {
        const bool from_best = t < kKMax;
        const float es = from_best ? bs[t] : cs[t - kKMax];
        const int32_t ei = from_best ? bi[t] : ci[t - kKMax];
        int rank = 0;
        for (int u = 0; u < kKMax; ++u)
          if (u != t && before(bs[u], bi[u], u, es, ei, t)) ++rank;
        for (int u = 0; u < pc; ++u)
          if (kKMax + u != t && before(cs[u], ci[u], kKMax + u, es, ei, t)) ++rank;
        if (rank < kKMax) {
          ns[rank] = es;
          ni[rank] = ei;
        }
      }
      __syncthreads();
      if (t < kKMax) {
        bs[t] = ns[t];
        bi[t] = ni[t];
      }
      __syncthreads();
    }
  }
  if (threadIdx.x < kKMax) {
    best_s[q * kKMax + threadIdx.x] = bs[threadIdx.x];
    best_i[q * kKMax + threadIdx.x] = bi[threadIdx.x];
  }
}

__global__ void init_best_kernel(float* best_s, int32_t* best_i, int64_t n) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t < n) {
    best_s[t] = -INFINITY;
    best_i[t] = 0x7fffffff;
  }
}

__global__ void write_topk_kernel(const float* __restrict__ best_s, const int32_t* __restrict__ best_i, int m,
                                  int k, int64_t id_base, int64_t id_stride, int64_t* __restrict__ out_ids,
                                  float* __restrict__ out_scores) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * k) return;
  const int q = t / k, r = t % k;
  const int32_t id = best_i[q * kKMax + r];
  const bool empty = id == 0x7fffffff;
  out_ids[t] = empty ? -1 : id_base + static_cast<int64_t>(id) * id_stride;
  out_scores[t] = empty ? -INFINITY : best_s[q * kKMax + r];
}

// ---------------------------------------------------------------------------------------------
// Shard merge: ids/scores [g, m, k] -> top-k of the g*k candidates of each row (same rule, global ids).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
topk_merge_kernel(const int64_t* __restrict__ ids, const float* __restrict__ scores, int g, int m, int k,
                  int64_t* __restrict__ out_ids, float* __restrict__ out_scores) {
  extern __shared__ unsigned char sm_raw[];
  float* cs = reinterpret_cast<float*>(sm_raw);
  int64_t* ci = reinterpret_cast<int64_t*>(cs + ((g * k + 1) / 2) * 2);
  const int q = blockIdx.x;
  const int n = g * k;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const int sh = t / k, r = t % k;
    const int64_t src = (static_cast<int64_t>(sh) * m + q) * k + r;
    cs[t] = scores[src];
    ci[t] = ids[src];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const float es = cs[t];
    const int64_t ei = ci[t];
    if (ei < 0) continue;  // empty slot of a short shard list
    int rank = 0;
    for (int u = 0; u < n; ++u) {
      const int64_t ui = ci[u];
      if (u == t || ui < 0) continue;
      if (cs[u] > es || (cs[u] == es && (ui < ei || (ui == ei && u < t)))) ++rank;
    }
    if (rank < k) {
      out_ids[static_cast<int64_t>(q) * k + rank] = ei;
      out_scores[static_cast<int64_t>(q) * k + rank] = es;
    }
  }
}

__global__ void fill_empty_kernel(int64_t* ids, float* scores, int64_t n) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t < n) {
    ids[t] = -1;
    scores[t] = -INFINITY;
  }
}

constexpr int kChunk = 32768;

int64_t exact_workspace_bytes(int64_t m, int64_t n_items) {
  const int64_t ch = n_items < kChunk ? n_items : kChunk;
  return 256 + m * ((ch + 3) / 4 * 4) * 4 + m * kKMax * 8 + 512;
}

int catalog_topk_exact(const float* queries, int64_t m, const float* table, int64_t n_items, int64_t d,
                       const float* bias, int64_t k, int64_t id_base, int64_t id_stride, void* workspace,
                       int64_t workspace_bytes, int64_t* out_ids, float* out_scores, cudaStream_t s) {
  if (workspace_bytes < exact_workspace_bytes(m, n_items)) return PSB_E_WORKSPACE;
  const int64_t ch = n_items < kChunk ? n_items : kChunk;
  const int ld = static_cast<int>((ch + 3) / 4 * 4);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  float* S = reinterpret_cast<float*>(ws);
  float* best_s = reinterpret_cast<float*>(ws + ((m * ld * 4 + 255) / 256) * 256);
  int32_t* best_i = reinterpret_cast<int32_t*>(best_s + m * kKMax);
  int st;
  PSB_PROF("init_best_kernel", s);
  init_best_kernel<<<static_cast<int>((m * kKMax + 255) / 256), 256, 0, s>>>(best_s, best_i, m * kKMax);
  if ((st = launch_status()) != PSB_OK) return st;
  for (int64_t c0 = 0; c0 < n_items; c0 += ch) {
    const int cn = static_cast<int>(n_items - c0 < ch ? n_items - c0 : ch);
    dim3 grid((cn + kTI - 1) / kTI, static_cast<unsigned>((m + kTQ - 1) / kTQ));
    PSB_PROF("exact_scores_kernel", s);
    exact_scores_kernel<<<grid, 256, 0, s>>>(queries, static_cast<int>(m), table, c0, cn, static_cast<int>(d),
                                             bias, S, ld);
    if ((st = launch_status()) != PSB_OK) return st;
    PSB_PROF("row_select_kernel", s);
    row_select_kernel<<<static_cast<int>(m), 256, 0, s>>>(S, ld, c0, cn, static_cast<int>(k), best_s, best_i);
    if ((st = launch_status()) != PSB_OK) return st;
  }
  PSB_PROF("write_topk_kernel", s);
  write_topk_kernel<<<static_cast<int>((m * k + 255) / 256), 256, 0, s>>>(
      best_s, best_i, static_cast<int>(m), static_cast<int>(k), id_base, id_stride, out_ids, out_scores);
  return launch_status();
}

}  // namespace psb

using namespace psb;

extern "C" int64_t psb_catalog_topk_workspace_bytes(int64_t m, int64_t n_items, int64_t d, int64_t k,
                                                    int32_t mode) {
  if (m <= 0 || n_items <= 0 || d <= 0 || k <= 0) return PSB_E_ARG;
  const int64_t ex = exact_workspace_bytes(m, n_items);
  if (mode == PSB_TOPK_EXACT) return ex;
  if (mode == PSB_TOPK_TC) return ex + tc_workspace_bytes(m, n_items, d, k);
  if (mode == PSB_TOPK_TC16) return ex + tc16_workspace_bytes(m, n_items, d, k) + 256;
  return PSB_E_UNSUPPORTED;
}

extern "C" int psb_catalog_prepare_f16(const float* table, int64_t n_items, int64_t d, void* table_f16, float* stats,
                                       psb_stream_t stream) {
  int st = check_table_args(table, n_items, d);
  if (st != PSB_OK) return st;
  if (table_f16 == nullptr || stats == nullptr) return PSB_E_ARG;
  if (misaligned16(table_f16) || misaligned16(stats)) return PSB_E_ALIGN;
  return catalog_prepare_f16(table, n_items, d, table_f16, stats, static_cast<cudaStream_t>(stream));
}

extern "C" int psb_catalog_topk_f16(const float* queries, int64_t m, const float* table, const void* table_f16,
                                    const float* stats, int64_t n_items, int64_t d, const float* bias, int64_t k,
                                    int64_t id_base, int64_t id_stride, void* workspace, int64_t workspace_bytes,
                                    int64_t* out_ids, float* out_scores, psb_stream_t stream) {
  if (queries == nullptr || table == nullptr || table_f16 == nullptr || stats == nullptr || workspace == nullptr ||
      out_ids == nullptr || out_scores == nullptr || m <= 0 || n_items <= 0 || n_items >= (1ll << 31) || m >= (1 << 24))
    return PSB_E_ARG;
  if (d <= 0 || (d & 3) != 0 || d > 512 || k <= 0 || k > kKMax) return PSB_E_DIM;
  if (misaligned16(queries) || misaligned16(table) || misaligned16(table_f16) || misaligned16(workspace))
    return PSB_E_ALIGN;
  return catalog_topk_tc16(queries, m, table, table_f16, stats, n_items, d, bias, k, id_base, id_stride, workspace,
                           workspace_bytes, out_ids, out_scores, static_cast<cudaStream_t>(stream));
}

extern "C" int psb_catalog_topk(const float* queries, int64_t m, const float* table, int64_t n_items,
                                int64_t d, const float* bias, int64_t k, int64_t id_base, int64_t id_stride,
                                int32_t mode, const float* max_row_sqnorm, void* workspace,
                                int64_t workspace_bytes, int64_t* out_ids, float* out_scores,
                                psb_stream_t stream) {
  if (queries == nullptr || table == nullptr || workspace == nullptr || out_ids == nullptr ||
      out_scores == nullptr || m <= 0 || n_items <= 0 || n_items >= (1ll << 31) || m >= (1 << 24))
    return PSB_E_ARG;
  if (d <= 0 || (d & 3) != 0 || d > 512 || k <= 0 || k > kKMax) return PSB_E_DIM;
  if (misaligned16(queries) || misaligned16(table) || misaligned16(workspace)) return PSB_E_ALIGN;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (mode == PSB_TOPK_EXACT)
    return catalog_topk_exact(queries, m, table, n_items, d, bias, k, id_base, id_stride, workspace,
                              workspace_bytes, out_ids, out_scores, s);
  if (mode == PSB_TOPK_TC)
    return catalog_topk_tc(queries, m, table, n_items, d, bias, k, id_base, id_stride, max_row_sqnorm, workspace,
                           workspace_bytes, out_ids, out_scores, s);
  return PSB_E_UNSUPPORTED;
}

extern "C" int psb_table_max_row_sqnorm(const float* table, int64_t rows, int64_t d, float* out,
                                        psb_stream_t stream) {
  int st = check_table_args(table, rows, d);
  if (st != PSB_OK) return st;
  if (out == nullptr) return PSB_E_ARG;
  return table_max_row_sqnorm(table, rows, d, out, static_cast<cudaStream_t>(stream));
}

extern "C" int psb_topk_merge(const int64_t* ids, const float* scores, int64_t g, int64_t m, int64_t k,
                              int64_t* out_ids, float* out_scores, psb_stream_t stream) {
  if (ids == nullptr || scores == nullptr || out_ids == nullptr || out_scores == nullptr || g <= 0 || m <= 0 ||
      k <= 0)
    return PSB_E_ARG;
  if (g * k > 4096) return PSB_E_DIM;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PSB_PROF("fill_empty_kernel", s);
  fill_empty_kernel<<<static_cast<int>((m * k + 255) / 256), 256, 0, s>>>(out_ids, out_scores, m * k);
  int st = launch_status();
  if (st != PSB_OK) return st;
  const size_t smem = static_cast<size_t>((g * k + 1) / 2 * 2) * 4 + static_cast<size_t>(g * k) * 8;
  PSB_PROF("topk_merge_kernel", s);
  topk_merge_kernel<<<static_cast<int>(m), 256, smem, s>>>(ids, scores, static_cast<int>(g), static_cast<int>(m),
                                                           static_cast<int>(k), out_ids, out_scores);
  return launch_status();
}

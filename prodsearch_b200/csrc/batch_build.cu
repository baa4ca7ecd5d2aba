// N4: on-device batch construction for the item-transformer (TEM) path, target ranks for the
// metrics, and the TREC ranklist writer.
//
// The reference assembles every batch with per-sample Python list comprehensions over nested lists
// (data/item_pv_dataloader.py:85-105,:122-143,:31-49).  Here the corpus relations live in HBM as flat CSR arrays
// (psb_corpus_t) and ONE launch builds target ids, query words and the padded purchase history of a whole batch:
// one warp per sample, lane-parallel over the user's review sequence, order-preserving ballot compaction.
// Integer / byte work only: HBM- (really latency-) bound, ~ (|user sequence| * 9 + wq * 16 + L * 8) bytes per sample.
#include <stdio.h>

#include "psb_common.cuh"

namespace psb {

// 32-bit integer hash ("lowbias32" finaliser) of (seed, sample, position): the random-subset mode
// (fix_train_review=False, item_pv_dataloader.py:98-101) keeps the hist_limit candidates with the smallest keys.
// oracle/batches.py restates it in numpy.
__host__ __device__ __forceinline__ uint32_t subset_key(uint32_t seed, uint32_t sample, uint32_t pos) {
  uint32_t h = seed ^ (sample * 0x9E3779B1u) ^ (pos * 0x85EBCA77u);
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}

struct BatchArgs {
  psb_corpus_t C;
  const int64_t* review_idx;
  const int64_t* user_idx;
  const int64_t* item_idx;
  const int64_t* query_idx;
  const uint32_t* query_pick;
  int64_t B;
  int hist_limit;
  int mode;
  uint32_t seed;
  int64_t item_pad;
  int64_t* target_prod_idxs;
  int64_t* query_idx_out;
  int64_t* query_word_idxs;
  int64_t* u_item_idxs;
  int32_t* hist_len;
  int32_t* err_flag;
};

__global__ void __launch_bounds__(256)
build_item_batch_kernel(const __grid_constant__ BatchArgs A) {
  const psb_corpus_t& C = A.C;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t b = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); b < A.B; b += nwarps) {
    const int64_t review = A.review_idx[b];
    const bool review_ok = review >= 0 && review < C.n_reviews;
    int64_t user = A.user_idx != nullptr ? A.user_idx[b] : (review_ok ? C.review_user[review] : -1);
    int64_t item = A.item_idx != nullptr ? A.item_idx[b] : (review_ok ? C.review_item[review] : -1);
    bool bad = user < 0 || user >= C.n_users || item < 0 || item >= C.n_items ||
               (!review_ok && (A.user_idx == nullptr || A.item_idx == nullptr || A.mode == PSB_HIST_SEQ));
    // ---- query: given (test batches) or one of the item's queries picked by a supplied random word
    int64_t q = -1;
    if (!bad) {
      if (A.query_idx != nullptr) {
        q = A.query_idx[b];
      } else {
        const int64_t q0 = C.item_query_off[item], q1 = C.item_query_off[item + 1];
        if (q1 > q0) q = C.item_query[q0 + static_cast<int64_t>(A.query_pick[b] % static_cast<uint32_t>(q1 - q0))];
      }
      if (q < 0 || q >= C.n_queries) bad = true;
    }
    int64_t* qw = A.query_word_idxs + b * C.wq;
    int64_t* hist = A.u_item_idxs + b * A.hist_limit;
    if (bad) {  // warp-uniform: flag it and emit an all-pad sample
      if (lane == 0) {
        if (A.err_flag != nullptr) *A.err_flag = 1;
        A.target_prod_idxs[b] = A.item_pad;
        if (A.query_idx_out != nullptr) A.query_idx_out[b] = -1;
        A.hist_len[b] = 0;
      }
      for (int j = lane; j < C.wq; j += 32) qw[j] = C.word_pad;
      for (int j = lane; j < A.hist_limit; j += 32) hist[j] = A.item_pad;
      continue;
    }
    if (lane == 0) {
      A.target_prod_idxs[b] = item;
      if (A.query_idx_out != nullptr) A.query_idx_out[b] = q;
    }
    for (int j = lane; j < C.wq; j += 32) qw[j] = C.query_words[q * C.wq + j];

    // ---- purchase history (get_user_review_idxs, item_pv_dataloader.py:85-105)
    const int64_t s0 = C.user_seq_off[user];
    const int n_seq = static_cast<int>(C.user_seq_off[user + 1] - s0);
    const int32_t* seq = C.user_seq + s0;
    const int L = A.hist_limit;
    int n_out = 0;
    if (A.mode == PSB_HIST_SEQ) {
      // the (up to) L reviews right before this one in the user's time-ordered sequence
      const int loc = min(max(C.review_uloc[review], 0), n_seq);
      const int first = max(loc - L, 0);
      n_out = loc - first;
      for (int j = lane; j < n_out; j += 32) hist[j] = C.review_item[seq[first + j]];
    } else {
      // candidates: the user's reviews of the training split, except this one, in sequence order
      int n_cand = 0;
      for (int p0 = 0; p0 < n_seq; p0 += 32) {
        const int p = p0 + lane;
        bool c = false;
        if (p < n_seq) {
          const int32_t r = seq[p];
          c = C.review_in_set[r] != 0 && r != review;
        }
        n_cand += __popc(__ballot_sync(kFull, c));
      }
      const bool subset = A.mode == PSB_HIST_RANDOM && n_cand > L;
      const int skip = (!subset && n_cand > L) ? n_cand - L : 0;  // fix=True keeps the LAST L candidates
      int seen = 0;
      for (int p0 = 0; p0 < n_seq; p0 += 32) {
        const int p = p0 + lane;
        int32_t r = -1;
        bool c = false;
        if (p < n_seq) {
          r = seq[p];
          c = C.review_in_set[r] != 0 && r != review;
        }
        const unsigned cm = __ballot_sync(kFull, c);
        bool keep = c && (seen + __popc(cm & lt)) >= skip;
        if (subset) {
          // rank of this candidate's (key, position) among all candidates; keep the L smallest
          const uint32_t myk = subset_key(A.seed, static_cast<uint32_t>(b), static_cast<uint32_t>(p));
          int smaller = 0;
          if (c) {
            for (int t = 0; t < n_seq; ++t) {
              const int32_t rt = seq[t];
              if (C.review_in_set[rt] != 0 && rt != review) {
                const uint32_t kt = subset_key(A.seed, static_cast<uint32_t>(b), static_cast<uint32_t>(t));
                smaller += (kt < myk || (kt == myk && t < p)) ? 1 : 0;
              }
            }
          }
          keep = c && smaller < L;
        }
        const unsigned km = __ballot_sync(kFull, keep);
        if (keep) hist[n_out + __popc(km & lt)] = C.review_item[r];
        n_out += __popc(km);
        seen += __popc(cm);
      }
    }
    for (int j = n_out + lane; j < L; j += 32) hist[j] = A.item_pad;
    if (lane == 0) A.hist_len[b] = n_out;
  }
}

// ---- N3: review-transformer test batches ------------------------------------------------------------------
struct ReviewBatchArgs {
  psb_corpus_t C;
  const int64_t* review_idx;
  const int64_t* user_idx;
  const int64_t* candi;
  int64_t B, n_cand;
  int u_limit, i_limit, mode;
  int64_t review_pad, user_pad, item_pad, seg_pad;
  int64_t *ridxs, *seg, *users, *items;
  int32_t* seq_len;
  int32_t* err_flag;
};

// Order-preserving emission of the LAST `limit` flagged entries of seq[0..n) (flag = training-split review other
// than `skip_review`), or of the plain range [first, n) when `plain`.  Calls emit(j, review) with j = 0.. in order.
template <typename Emit>
__device__ __forceinline__ int emit_tail(const int32_t* __restrict__ seq, int n, const uint8_t* __restrict__ in_set,
                                         int64_t skip_review, int limit, bool plain, int plain_first, int lane,
                                         Emit&& emit) {
  const unsigned lt = (1u << lane) - 1u;
  if (plain) {
    const int cnt = n - plain_first;
    for (int j = lane; j < cnt; j += 32) emit(j, seq[plain_first + j]);
    return cnt;
  }
  int n_cand = 0;
  for (int p0 = 0; p0 < n; p0 += 32) {
    const int p = p0 + lane;
    bool c = false;
    if (p < n) {
      const int32_t r = seq[p];
      c = in_set[r] != 0 && r != skip_review;
    }
    n_cand += __popc(__ballot_sync(kFull, c));
  }
  const int skip = n_cand > limit ? n_cand - limit : 0;
  int seen = 0, n_out = 0;
  for (int p0 = 0; p0 < n; p0 += 32) {
    const int p = p0 + lane;
    int32_t r = -1;
    bool c = false;
    if (p < n) {
      r = seq[p];
      c = in_set[r] != 0 && r != skip_review;
    }
    const unsigned cm = __ballot_sync(kFull, c);
    const bool keep = c && (seen + __popc(cm & lt)) >= skip;
    const unsigned km = __ballot_sync(kFull, keep);
    if (keep) emit(n_out + __popc(km & lt), r);
    n_out += __popc(km);
    seen += __popc(cm);
  }
  return n_out;
}

__global__ void __launch_bounds__(256)
build_review_test_batch_kernel(const __grid_constant__ ReviewBatchArgs A) {
  const psb_corpus_t& C = A.C;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  const int W = A.u_limit + A.i_limit;
  const bool seq_mode = A.mode == PSB_HIST_SEQ;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); e < A.B * A.n_cand;
       e += nwarps) {
    const int64_t b = e / A.n_cand;
    const int64_t cand = A.candi[e];
    int64_t* ridx = A.ridxs + e * W;
    int64_t* seg = A.seg + e * (W + 1);
    int64_t* usr = A.users + e * (W + 1);
    int64_t* itm = A.items + e * (W + 1);
    const int64_t review = A.review_idx[b];
    const int64_t user = A.user_idx[b];
    const bool review_ok = review >= 0 && review < C.n_reviews;
    const bool bad = user < 0 || user >= C.n_users || cand >= C.n_items || (seq_mode && !review_ok);
    if (cand < 0 || bad) {  // padded candidate slot (or invalid ids): an all-pad row
      if (bad && lane == 0 && A.err_flag != nullptr) *A.err_flag = 1;
      for (int j = lane; j < W; j += 32) ridx[j] = A.review_pad;
      for (int j = lane; j <= W; j += 32) {
        seg[j] = A.seg_pad;
        usr[j] = A.user_pad;
        itm[j] = A.item_pad;
      }
      if (lane == 0 && A.seq_len != nullptr) A.seq_len[e] = 0;
      continue;
    }
    if (lane == 0) {
      seg[0] = 0;
      usr[0] = A.user_pad;
      itm[0] = A.item_pad;
    }
    // user part
    const int64_t us0 = C.user_seq_off[user];
    const int un = static_cast<int>(C.user_seq_off[user + 1] - us0);
    int u_first = 0;
    if (seq_mode) {
      const int loc = min(max(C.review_uloc[review], 0), un);
      u_first = max(loc - A.u_limit, 0);
    }
    const int nu = emit_tail(C.user_seq + us0, seq_mode ? min(max(C.review_uloc[review], 0), un) : un, C.review_in_set,
                             review, A.u_limit, seq_mode, u_first, lane, [&](int j, int32_t r) {
                               ridx[j] = r;
                               seg[1 + j] = 1;
                               usr[1 + j] = user;
                               itm[1 + j] = C.review_item[r];
                             });
    // candidate item part
    const int64_t is0 = C.item_seq_off[cand];
    const int in = static_cast<int>(C.item_seq_off[cand + 1] - is0);
    const int32_t* iseq = C.item_seq + is0;
    int i_end = in, i_first = 0;
    if (seq_mode) {  // dataset.bisect_right on the time stamps (prod_search_dataset.py:135-152)
      const int64_t ts = C.review_time[review];
      int lo = 0, hi = in;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ts < C.review_time[iseq[mid]]) hi = mid; else lo = mid + 1;
      }
      i_end = lo;
      i_first = max(i_end - A.i_limit, 0);
    }
    const int ni = emit_tail(iseq, i_end, C.review_in_set, -1, A.i_limit, seq_mode, i_first, lane,
                             [&](int j, int32_t r) {
                               ridx[nu + j] = r;
                               seg[1 + nu + j] = 2;
                               usr[1 + nu + j] = C.review_user[r];
                               itm[1 + nu + j] = cand;
                             });
    const int n = nu + ni;
    for (int j = n + lane; j < W; j += 32) ridx[j] = A.review_pad;
    for (int j = 1 + n + lane; j <= W; j += 32) {
      seg[j] = A.seg_pad;
      usr[j] = A.user_pad;
      itm[j] = A.item_pad;
    }
    if (lane == 0 && A.seq_len != nullptr) A.seq_len[e] = n;
  }
}

// rank[i] = 1 + position of target[i] in ids[i, :k], 0 when it is not in the list (calc_metrics,
// trainer.py:171-186, evaluated on the fused top-k lists instead of a full argsort).
__global__ void __launch_bounds__(256)
target_rank_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ target, int64_t m, int k,
                   int32_t* __restrict__ rank) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); i < m; i += nwarps) {
    const int64_t t = target[i];
    int found = 0;
    for (int j0 = 0; j0 < k && found == 0; j0 += 32) {
      const int j = j0 + lane;
      const unsigned hit = __ballot_sync(kFull, j < k && ids[i * k + j] == t);
      if (hit != 0u) found = j0 + __ffs(hit);
    }
    if (lane == 0) rank[i] = found;
  }
}

}  // namespace psb

using namespace psb;

extern "C" uint32_t psb_subset_key(uint32_t seed, uint32_t sample, uint32_t pos) { return subset_key(seed, sample, pos); }

extern "C" int psb_build_item_batch(const psb_corpus_t* corpus, const int64_t* review_idx, const int64_t* user_idx,
                                    const int64_t* item_idx, const int64_t* query_idx, const uint32_t* query_pick,
                                    int64_t batch, int64_t hist_limit, int32_t mode, uint32_t seed, int64_t item_pad,
                                    int64_t* target_prod_idxs, int64_t* query_idx_out, int64_t* query_word_idxs,
                                    int64_t* u_item_idxs, int32_t* hist_len, int32_t* err_flag, psb_stream_t stream) {
  if (corpus == nullptr || batch < 0 || hist_limit <= 0 || hist_limit > (1 << 20)) return PSB_E_ARG;
  if (mode != PSB_HIST_SEQ && mode != PSB_HIST_LAST && mode != PSB_HIST_RANDOM) return PSB_E_ARG;
  if (batch == 0) return PSB_OK;
  const psb_corpus_t& C = *corpus;
  if (C.review_user == nullptr || C.review_item == nullptr || C.review_in_set == nullptr ||
      C.user_seq_off == nullptr || C.user_seq == nullptr || C.query_words == nullptr || C.n_reviews <= 0 ||
      C.n_users <= 0 || C.n_items <= 0 || C.n_queries <= 0 || C.wq <= 0)
    return PSB_E_ARG;
  if (mode == PSB_HIST_SEQ && C.review_uloc == nullptr) return PSB_E_ARG;
  if (query_idx == nullptr && (query_pick == nullptr || C.item_query_off == nullptr || C.item_query == nullptr))
    return PSB_E_ARG;
  if (review_idx == nullptr || target_prod_idxs == nullptr || query_word_idxs == nullptr || u_item_idxs == nullptr ||
      hist_len == nullptr)
    return PSB_E_ARG;
  BatchArgs A;
  A.C = C;
  A.review_idx = review_idx;
  A.user_idx = user_idx;
  A.item_idx = item_idx;
  A.query_idx = query_idx;
  A.query_pick = query_pick;
  A.B = batch;
  A.hist_limit = static_cast<int>(hist_limit);
  A.mode = mode;
  A.seed = seed;
  A.item_pad = item_pad;
  A.target_prod_idxs = target_prod_idxs;
  A.query_idx_out = query_idx_out;
  A.query_word_idxs = query_word_idxs;
  A.u_item_idxs = u_item_idxs;
  A.hist_len = hist_len;
  A.err_flag = err_flag;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PSB_PROF("build_item_batch_kernel", s);
  build_item_batch_kernel<<<grid_for(batch, 8), 256, 0, s>>>(A);
  return launch_status();
}

extern "C" int psb_build_review_test_batch(const psb_corpus_t* corpus, const int64_t* review_idx,
                                           const int64_t* user_idx, const int64_t* candi_prod_idxs, int64_t batch,
                                           int64_t n_cand, int64_t u_limit, int64_t i_limit, int32_t mode,
                                           int64_t review_pad, int64_t user_pad, int64_t item_pad, int64_t seg_pad,
                                           int64_t* ridxs, int64_t* seg, int64_t* users, int64_t* items,
                                           int32_t* seq_len, int32_t* err_flag, psb_stream_t stream) {
  if (corpus == nullptr || batch < 0 || n_cand < 0 || u_limit < 0 || i_limit < 0 || u_limit + i_limit <= 0 ||
      u_limit + i_limit > (1 << 20))
    return PSB_E_ARG;
  if (mode != PSB_HIST_SEQ && mode != PSB_HIST_LAST) return PSB_E_ARG;
  if (batch == 0 || n_cand == 0) return PSB_OK;
  const psb_corpus_t& C = *corpus;
  if (C.review_user == nullptr || C.review_item == nullptr || C.review_in_set == nullptr ||
      C.user_seq_off == nullptr || C.user_seq == nullptr || C.item_seq_off == nullptr || C.item_seq == nullptr ||
      C.n_reviews <= 0 || C.n_users <= 0 || C.n_items <= 0)
    return PSB_E_ARG;
  if (mode == PSB_HIST_SEQ && (C.review_uloc == nullptr || C.review_time == nullptr)) return PSB_E_ARG;
  if (review_idx == nullptr || user_idx == nullptr || candi_prod_idxs == nullptr || ridxs == nullptr ||
      seg == nullptr || users == nullptr || items == nullptr)
    return PSB_E_ARG;
  ReviewBatchArgs A;
  A.C = C;
  A.review_idx = review_idx;
  A.user_idx = user_idx;
  A.candi = candi_prod_idxs;
  A.B = batch;
  A.n_cand = n_cand;
  A.u_limit = static_cast<int>(u_limit);
  A.i_limit = static_cast<int>(i_limit);
  A.mode = mode;
  A.review_pad = review_pad;
  A.user_pad = user_pad;
  A.item_pad = item_pad;
  A.seg_pad = seg_pad;
  A.ridxs = ridxs;
  A.seg = seg;
  A.users = users;
  A.items = items;
  A.seq_len = seq_len;
  A.err_flag = err_flag;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PSB_PROF("build_review_test_batch_kernel", s);
  build_review_test_batch_kernel<<<grid_for(batch * n_cand, 8), 256, 0, s>>>(A);
  return launch_status();
}

extern "C" int psb_target_rank(const int64_t* ids, const int64_t* target, int64_t m, int64_t k, int32_t* rank,
                               psb_stream_t stream) {
  if (m < 0 || k <= 0 || k > (1 << 24)) return PSB_E_ARG;
  if (m == 0) return PSB_OK;
  if (ids == nullptr || target == nullptr || rank == nullptr) return PSB_E_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PSB_PROF("target_rank_kernel", s);
  target_rank_kernel<<<grid_for(m, 8), 256, 0, s>>>(ids, target, m, static_cast<int>(k), rank);
  return launch_status();
}

// Host side of the eval path: the TREC run file of Trainer.test (trainer.py:158-169), one line per
// (query, rank): "%s_%d Q0 %s %d %f ReviewTransformer\n".  All pointers are HOST pointers.
extern "C" int64_t psb_write_ranklist(const char* path, const char* const* user_ids, const int64_t* user_idx,
                                      const int64_t* query_idx, const char* const* product_ids,
                                      const int64_t* ids, const float* scores, int64_t m, int64_t k, int64_t cutoff,
                                      int32_t append) {
  if (path == nullptr || user_ids == nullptr || user_idx == nullptr || query_idx == nullptr ||
      product_ids == nullptr || m < 0 || k <= 0 || (m > 0 && (ids == nullptr || scores == nullptr)))
    return PSB_E_ARG;
  FILE* f = fopen(path, append != 0 ? "a" : "w");
  if (f == nullptr) return PSB_E_ARG;
  static char iobuf[1 << 20];
  setvbuf(f, iobuf, _IOFBF, sizeof(iobuf));
  const int64_t top = cutoff < k ? cutoff : k;
  int64_t lines = 0;
  for (int64_t i = 0; i < m; ++i) {
    const char* uid = user_ids[user_idx[i]];
    for (int64_t r = 0; r < top; ++r) {
      const int64_t id = ids[i * k + r];
      if (id < 0) break;  // fewer than k candidates
      fprintf(f, "%s_%lld Q0 %s %lld %f ReviewTransformer\n", uid, static_cast<long long>(query_idx[i]),
              product_ids[id], static_cast<long long>(r + 1), static_cast<double>(scores[i * k + r]));
      ++lines;
    }
  }
  if (fclose(f) != 0) return PSB_E_ARG;
  return lines;
}

"""Ranking side of the oracle: full-catalog scores -> ranked ids.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows trainer.py:189-226 (Trainer.get_prod_scores: segment the catalog into
candi_batch_size chunks, score, concatenate to a [M, N] matrix on the host) and
trainer.py:136,:152 (``argsort(axis=-1)[:, ::-1]``), with ONE stated deviation:
numpy's reversed unstable argsort is not lower-id-first on ties, so
the oracle canonicalises ties to "lower id first", which is the contract the CUDA
path implements.  ``reference_rank`` keeps the reference's literal expression for
tie-free comparisons.
"""
import numpy as np

__all__ = ["reference_rank", "rank_lower_id_first", "topk_lower_id_first",
           "merge_shard_topk", "calc_metrics"]


def reference_rank(scores):
    """The literal expression at trainer.py:136 / :152."""
    return np.asarray(scores).argsort(axis=-1)[:, ::-1]


def rank_lower_id_first(scores):
    """Descending score, ties broken by ascending column id (stable)."""
    s = np.asarray(scores)
    return np.argsort(-s, axis=-1, kind="stable")


def topk_lower_id_first(scores, k, ids=None):
    """Top-k (ids, scores) per row; ids default to the column number."""
    s = np.asarray(scores)
    if ids is None:
        order = rank_lower_id_first(s)[:, :k]
        return order.astype(np.int64), np.take_along_axis(s, order, axis=-1)
    ids = np.asarray(ids)
    out_i = np.empty((s.shape[0], min(k, s.shape[1])), dtype=np.int64)
    out_s = np.empty(out_i.shape, dtype=s.dtype)
    for r in range(s.shape[0]):
        order = np.lexsort((ids[r], -s[r]))[:k]
        out_i[r], out_s[r] = ids[r][order], s[r][order]
    return out_i, out_s


def merge_shard_topk(shard_ids, shard_scores, k):
    """Merge per-shard top-k lists [G][M,k'] into a global top-k with the same
    tie rule (SURVEY.md 8(e) eval collective: all_gather then merge)."""
    ids = np.concatenate(shard_ids, axis=1)
    sc = np.concatenate(shard_scores, axis=1)
    return topk_lower_id_first(sc, k, ids)


def calc_metrics(ranked_ids, target_ids, cutoff=100):
    """MRR / P@1 as trainer.py:171-187 computes them from a ranked id list."""
    mrr, prec = 0.0, 0.0
    for row, tgt in zip(np.asarray(ranked_ids), np.asarray(target_ids)):
        hit = np.where(row == tgt)[0]
        if len(hit):
            rank = hit[0] + 1
            if cutoff < 0 or rank <= cutoff:
                mrr += 1.0 / rank
            if rank == 1:
                prec += 1
    n = len(target_ids)
    return mrr / n, prec / n

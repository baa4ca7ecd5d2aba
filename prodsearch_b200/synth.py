"""Synthetic Amazon-shaped inputs: the real Amazon review files are not
available offline, so benches and parity tests use seeded data of the same shape -- Zipf item and
word popularity, short right-padded purchase histories and queries, uniform item negatives and
count^0.75 word negatives (data/data_util.py:155-162), all SUPPLIED as tensors."""
import argparse

import numpy as np
import torch


def zipf_probs(n, s=1.0):
    p = 1.0 / np.arange(1, n + 1, dtype=np.float64) ** s
    return p / p.sum()


def word_dists(vocab_size):
    """Normalised count^0.75 over a Zipf vocabulary; the pad word (last id) has probability 0."""
    counts = zipf_probs(vocab_size - 1) * 1e7
    d = np.concatenate([counts ** 0.75, [0.0]])
    return (d / d.sum()).astype(np.float32)


def _sample(rng, probs, shape):
    cdf = np.cumsum(probs)
    cdf[-1] = 1.0
    return np.searchsorted(cdf, rng.random(shape), side="right").astype(np.int64)


def tem_batch(B, product_size, vocab_size, L=20, W=1, K=5, Wq_max=12, seed=666, permute=True):
    """One ItemPVBatch-shaped training batch (data/batch_data.py:3-37) + supplied negatives.
    Returns (batch namespace of CPU int64 tensors, neg_item_idxs [B,K], neg_word_idxs [B*W*K])."""
    rng = np.random.default_rng(seed)
    P, V = product_size, vocab_size
    item_p = zipf_probs(P)
    perm = rng.permutation(P) if permute else np.arange(P)       # popular ids spread over the table
    word_p = zipf_probs(V - 1)
    wd = word_dists(V).astype(np.float64)
    wd = wd / wd.sum()
    target = perm[_sample(rng, item_p, (B,))]
    hist_len = np.minimum(L, 1 + rng.geometric(0.15, size=B))
    hist = perm[_sample(rng, item_p, (B, L))]
    hist[np.arange(L)[None, :] >= hist_len[:, None]] = P
    qlen = rng.integers(2, Wq_max + 1, size=B)
    Wq = int(qlen.max())
    qw = _sample(rng, word_p, (B, Wq))
    qw[np.arange(Wq)[None, :] >= qlen[:, None]] = V - 1
    iw = _sample(rng, wd[:-1] / wd[:-1].sum(), (B, W))
    neg_items = rng.integers(0, P, size=(B, K)).astype(np.int64)
    neg_words = _sample(rng, wd[:-1] / wd[:-1].sum(), (B * W * K,))
    batch = argparse.Namespace(
        query_word_idxs=torch.from_numpy(qw), target_prod_idxs=torch.from_numpy(target),
        u_item_idxs=torch.from_numpy(hist), pos_iword_idxs=torch.from_numpy(iw),
        candi_prod_idxs=torch.zeros(B, 1, dtype=torch.int64), query_idxs=list(range(B)), user_idxs=list(range(B)))
    return batch, torch.from_numpy(neg_items), torch.from_numpy(neg_words)


def gather_indices(n, table_rows, seed=666, dist="zipf"):
    """Row ids for the bandwidth-regime gather / scatter benches on a big table."""
    rng = np.random.default_rng(seed)
    if dist == "uniform":
        return torch.from_numpy(rng.integers(0, table_rows, size=n).astype(np.int64))
    perm_mult = 2654435761 % table_rows                            # cheap bijection-ish spread of hot ids
    ids = _sample(rng, zipf_probs(min(table_rows, 1 << 22)), (n,))
    return torch.from_numpy((ids * perm_mult + 12345) % table_rows)

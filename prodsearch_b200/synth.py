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


def review_words_table(review_count, vocab_size, Wr=100, seed=666):
    """[R, Wr] int64 word ids of every review, right-padded with the pad word; the last review is the pad
    review (all pad words), as ProductRanker keeps it (models/ps_model.py:99-107)."""
    rng = np.random.default_rng(seed)
    R, V = review_count, vocab_size
    words = _sample(rng, zipf_probs(V - 1), (R, Wr))
    length = rng.integers(5, Wr + 1, size=R)
    words[np.arange(Wr)[None, :] >= length[:, None]] = V - 1
    words[R - 1] = V - 1
    return torch.from_numpy(words)


def rtm_batch(B, review_words, vocab_size, product_size, user_size, Ru=20, Ri=30, W=1, K=5, Wq_max=12, pvc=False,
              train_pv=True, seed=666):
    """One ProdSearchTrainBatch-shaped batch (data/batch_data.py:137-223) at BASELINE configs[2] shape:
    sequences of Ru user + Ri item reviews (right-padded with the pad review R-1, segment 3), target words
    [B,Rc,W] (the pv sliding window when ``train_pv``, the whole review otherwise), negatives' review words
    [B,K,Rc,Wr] from the review-word table.  Returns (batch namespace of CPU tensors, draws) where draws =
    dict(multinomial=[...], bernoulli=[...]) are the supplied random draws in the reference's call order."""
    rng = np.random.default_rng(seed)
    R, Wr = review_words.shape
    V, P, U = vocab_size, product_size, user_size
    Rc = Ru + Ri
    rw = review_words.numpy()

    def ridx(*shape):
        x = rng.integers(0, R - 1, size=shape).astype(np.int64)
        n_u = np.minimum(Ru, 1 + rng.geometric(0.15, size=shape[:-1]))
        n_i = np.minimum(Ri, 1 + rng.geometric(0.10, size=shape[:-1]))
        pos = np.arange(Rc)
        pad = np.where(pos < Ru, pos >= n_u[..., None], (pos - Ru) >= n_i[..., None])
        x[pad] = R - 1
        return x
    pos_r = ridx(B, Rc)
    neg_r = ridx(B, K, Rc)
    neg_r[B // 2, K - 1, :] = R - 1                      # a negative without any review: weight 0 (ps_model.py:344)
    seg = np.array([0] + [1] * Ru + [2] * Ri, dtype=np.int64)
    pos_seg = np.broadcast_to(seg, (B, Rc + 1)).copy()
    neg_seg = np.broadcast_to(seg, (B, K, Rc + 1)).copy()
    pos_seg[:, 1:][pos_r == R - 1] = 3
    neg_seg[:, :, 1:][neg_r == R - 1] = 3
    qlen = rng.integers(2, Wq_max + 1, size=B)
    Wq = int(qlen.max())
    qw = _sample(rng, zipf_probs(V - 1), (B, Wq))
    qw[np.arange(Wq)[None, :] >= qlen[:, None]] = V - 1
    if train_pv:
        pos_rw = _sample(rng, zipf_probs(V - 1), (B, Rc, W))
        pos_rw[pos_r == R - 1] = V - 1
    else:
        pos_rw = rw[pos_r]
    neg_rw = rw[neg_r]
    t = torch.from_numpy
    batch = argparse.Namespace(
        query_word_idxs=t(qw), pos_prod_ridxs=t(pos_r), pos_seg_idxs=t(pos_seg), pos_prod_rword_idxs=t(pos_rw),
        pos_prod_rword_masks=t((pos_rw != V - 1).astype(np.uint8)), neg_prod_ridxs=t(neg_r), neg_seg_idxs=t(neg_seg),
        pos_user_idxs=t(rng.integers(0, U, size=(B, Rc + 1)).astype(np.int64)),
        neg_user_idxs=t(rng.integers(0, U, size=(B, K, Rc + 1)).astype(np.int64)),
        pos_item_idxs=t(rng.integers(0, P, size=(B, Rc + 1)).astype(np.int64)),
        neg_item_idxs=t(rng.integers(0, P, size=(B, K, Rc + 1)).astype(np.int64)),
        neg_prod_rword_idxs=t(neg_rw), neg_prod_rword_masks=t((neg_rw != V - 1).astype(np.uint8)),
        pos_prod_rword_idxs_pvc=t(rw[pos_r]) if pvc else None, neg_prod_rword_idxs_pvc=t(neg_rw) if pvc else None)
    draws = dict(multinomial=[], bernoulli=[])
    if train_pv:
        wd = word_dists(V).astype(np.float64)
        draws["multinomial"].append(t(_sample(rng, wd[:-1] / wd[:-1].sum(), (B * Rc * pos_rw.shape[-1] * K,))))
    return batch, draws

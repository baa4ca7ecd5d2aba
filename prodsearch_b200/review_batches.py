"""Host-side TRAINING collate of the review-transformer (RTM) path (the caller side of
``ProductRanker.forward``, BASELINE configs[2]).

The reference assembles a training batch sample by sample from nested lists
(``ProdSearchDataLoader.prepare_train_batch`` / ``get_train_batch``, data/prod_search_dataloader.py:208-358) after a
per-epoch ``ProdSearchData.initialize_epoch`` (data/data_util.py:94-117).  ``ReviewTrainCollate`` does the same work on
the flat arrays of ``data_files.CorpusFiles`` / ``SplitFiles``: histories are slices of the CSR review sequences,
review words one ``[R + 1, review_word_limit]`` matrix indexed by whole id tensors, sliding pv windows one
reshape / transpose.  Every random decision is drawn from the same generator with the same call as in the reference
(``random.choice`` / ``random.sample``; ``numpy.random.choice`` / ``random`` / ``shuffle`` / ``permutation``), in the
reference's order, so a run seeded like main.py:172-173 produces the identical batches
(tests/golden/review_batches.npz).  Host code only: the tensors it returns are what ``ProductRanker.forward`` gathers
from on the device.

Reference behaviour kept on purpose (each visible in the golden vectors):
* ``initialize_epoch`` without ``do_subsample_mask`` compares EVERY word with the first random number (its cursor is
  never advanced, data_util.py:104-111);
* ``shuffle_words_in_reviews`` (prod_search_dataset.py:104-107) iterates over ``[B, reviews, words]`` and so permutes
  the reviews' word rows of a sample, not the words of a review -- after the masks were drawn, and in place, which
  also changes the ``*_pvc`` word tensors that alias the same array;
* a sample without item-side reviews, or whose sampled negatives all lack reviews, is dropped AFTER it consumed its
  random numbers (prod_search_dataloader.py:219-222,:255-257).
Reference behaviour NOT kept: the pv branch indexes four padded python lists with an index array
(prod_search_dataloader.py:342-345) and raises TypeError for batches larger than one; here those four fields are
indexed like their neighbours (the golden generator wraps ``util.pad`` so the reference does the same).
"""
import argparse
import random as _random

import numpy as np
import torch


class ReviewTrainBatch(argparse.Namespace):
    """Attribute names of ProdSearchTrainBatch (data/batch_data.py:137-166); ``to`` returns a new object."""

    def to(self, device):
        if device == "cpu":
            return self
        return ReviewTrainBatch(**{k: (v.to(device) if torch.is_tensor(v) else v) for k, v in vars(self).items()})


def _pad_rows(rows, pad_id, width=None):
    """util.pad (others/util.py:36-40) -> int64 [n, width]."""
    width = max(len(r) for r in rows) if width is None else width
    out = np.full((len(rows), width), pad_id, dtype=np.int64)
    for i, r in enumerate(rows):
        r = r[:width]
        out[i, :len(r)] = r
    return out


def _pad_nested(groups, pad_id):
    """util.pad_3d over dim 1 then dim 2 (others/util.py:42-61) -> int64 [n, max group size, max row width]."""
    depth = max(len(g) for g in groups)
    width = max(len(r) for g in groups for r in g)
    out = np.full((len(groups), depth, width), pad_id, dtype=np.int64)
    for i, g in enumerate(groups):
        for j, r in enumerate(g):
            out[i, j, :len(r)] = r
    return out


def _windows(x, window, pad_id):
    """slide_padded_matrices_for_pv (prod_search_dataset.py:120-135): [n, L] -> [ceil(L / window), n, window]."""
    n, L = x.shape
    segs = (L + window - 1) // window
    if segs * window != L:
        x = np.concatenate([x, np.full((n, segs * window - L), pad_id, dtype=x.dtype)], axis=1)
    return np.ascontiguousarray(x.reshape(n, segs, window).transpose(1, 0, 2))


class ReviewTrainCollate(object):
    """files: data_files.CorpusFiles; split: its "train" SplitFiles; args: the reference's flags
    (uprev_review_limit, iprev_review_limit, do_seq_review_train, do_subsample_mask, review_word_limit, neg_per_pos,
    review_encoder_name, pv_window_size, shuffle_review_words).  py_random / np_random default to the global
    generators the reference draws from."""

    def __init__(self, files, split, args, py_random=None, np_random=None):
        self.f, self.s, self.args = files, split, args
        self.py_random = py_random if py_random is not None else _random
        self.np_random = np_random if np_random is not None else np.random
        self.user_pad_idx, self.prod_pad_idx = files.user_size, files.product_size       # prod_search_dataset.py:23-24
        self.word_pad_idx, self.review_pad_idx = files.word_pad_idx, files.review_count - 1
        self.seg_pad_idx = 3
        self.u_limit, self.i_limit = int(args.uprev_review_limit), int(args.iprev_review_limit)
        self.total_limit = self.u_limit + self.i_limit
        self.word_limit = int(args.review_word_limit)
        self.in_train = files.review_in_train.astype(bool)
        self.neg_sample_products = None
        self.sub_sampling_rate = None
        self.review_words = None
        if args.do_subsample_mask:               # data_util.py:186-190: cut / pad every review, pad review last
            self.review_words = self._word_matrix(np.ones(len(files.review_word), dtype=bool))
            self.sub_sampling_rate = np.asarray(split.sub_sampling_rate)

    # ------------------------------------------------------------------ per epoch
    def _word_matrix(self, keep):
        """Kept words of every review, left-aligned, cut / padded to review_word_limit; row R is the pad review."""
        f = self.f
        R = len(f.review_length)
        out = np.full((R + 1, self.word_limit), self.word_pad_idx, dtype=np.int64)
        review_of = np.repeat(np.arange(R), f.review_length)
        kept = np.concatenate([[0], np.cumsum(keep)])                  # kept words before word i, corpus-wide
        pos = kept[:-1] - kept[f.review_word_off[review_of]]           # rank among the review's kept words
        sel = keep & (pos < self.word_limit)
        out[review_of[sel], pos[sel]] = f.review_word[sel]
        return out

    def initialize_epoch(self):
        """ProdSearchData.initialize_epoch (data_util.py:94-117): negative items per training line and, without
        do_subsample_mask, the epoch's sub-sampled review words."""
        f, s = self.f, self.s
        self.neg_sample_products = self.np_random.choice(
            f.product_size, size=(len(s.review_info), int(self.args.neg_per_pos)), replace=True, p=s.product_dists)
        if self.args.do_subsample_mask:
            return
        rand_numbers = self.np_random.random(int(f.review_length.sum()))
        rate = np.asarray(s.sub_sampling_rate)
        keep = ~(rand_numbers[0] > rate[f.review_word]) if len(rand_numbers) else np.zeros(0, dtype=bool)
        self.review_words = self._word_matrix(keep)

    # ------------------------------------------------------------------ histories
    def _train_subset(self, seq, review_idx, limit):
        """Training reviews of a sequence except review_idx, a random ``limit`` of them in sequence order
        (get_user_review_idxs / get_item_review_idxs with fix=False, prod_search_dataloader.py:135-159,:178-206)."""
        cand = seq[self.in_train[seq]]
        if review_idx is not None:
            cand = cand[cand != review_idx]
        cand = cand.tolist()
        if len(cand) > limit:
            chosen = set(self.py_random.sample(cand, limit))
            cand = [x for x in cand if x in chosen]
        return cand

    def user_reviews(self, user_idx, review_idx):
        f = self.f
        seq = f.user_seq[f.user_seq_off[user_idx]:f.user_seq_off[user_idx + 1]]
        if self.args.do_seq_review_train:
            loc = int(f.review_loc_time[review_idx, 0])
            return seq[:loc][-self.u_limit:].tolist()
        return self._train_subset(seq, review_idx, self.u_limit)

    def item_reviews(self, prod_idx, review_idx, time_stamp=None):
        f = self.f
        seq = f.item_seq[f.item_seq_off[prod_idx]:f.item_seq_off[prod_idx + 1]]
        if self.args.do_seq_review_train:
            if review_idx is None:               # reviews of the item up to the purchase time (bisect_right, :137-156)
                lo, hi = 0, len(seq)
                while lo < hi:
                    mid = (lo + hi) // 2
                    if time_stamp < f.review_loc_time[seq[mid], 2]:
                        hi = mid
                    else:
                        lo = mid + 1
                loc = lo
            else:
                loc = int(f.review_loc_time[review_idx, 1])
            if loc == 0:
                return []
            return seq[:loc][-self.i_limit:].tolist()
        return self._train_subset(seq, review_idx, self.i_limit)

    # ------------------------------------------------------------------ one batch
    def _samples(self, rows):
        """prepare_train_batch (prod_search_dataloader.py:208-271) -> per-sample python lists."""
        f, s, T = self.f, self.s, self.total_limit
        if self.neg_sample_products is None:
            raise RuntimeError("call initialize_epoch() first (negative items are drawn once per epoch)")
        out = []
        for line_id, user_idx, prod_idx, review_idx in np.asarray(rows, dtype=np.int64).reshape(-1, 4).tolist():
            queries = s.item_query[s.item_query_off[prod_idx]:s.item_query_off[prod_idx + 1]].tolist()
            query_idx = self.py_random.choice(queries)
            u_rev = self.user_reviews(user_idx, review_idx)
            i_rev = self.item_reviews(prod_idx, review_idx)
            stamp = int(f.review_loc_time[review_idx, 2]) if self.args.do_seq_review_train else None
            if len(i_rev) == 0:
                continue
            u_items = f.review_u_p[u_rev, 1].tolist() if u_rev else []

            def sequence(item, item_rev):
                users = [self.user_pad_idx] + [user_idx] * len(u_rev) + f.review_u_p[item_rev, 0].tolist()
                items = [self.prod_pad_idx] + u_items + [item] * len(item_rev)
                segs = [0] + [1] * len(u_rev) + [2] * len(item_rev)
                return (u_rev + item_rev)[:T], segs[:T + 1], users[:T + 1], items[:T + 1]
            negs = []
            for neg_i in self.neg_sample_products[line_id].tolist():
                n_rev = self.item_reviews(neg_i, None, stamp)
                if len(n_rev):
                    negs.append(sequence(neg_i, n_rev))
            if not negs:
                continue
            out.append((f.query_words[query_idx], sequence(prod_idx, i_rev), negs))
        return out

    def _word_masks(self, word_idxs):
        """get_pv_word_masks (prod_search_dataset.py:92-102)."""
        if self.sub_sampling_rate is None:
            return word_idxs != self.word_pad_idx
        rand_numbers = self.np_random.random(word_idxs.shape)
        return np.logical_and(word_idxs != self.word_pad_idx, rand_numbers < self.sub_sampling_rate[word_idxs])

    def train_batch(self, rows, prepare_pv=True, shuffle=False):
        """get_train_batch (prod_search_dataloader.py:286-358) for ``rows`` of review_info
        (line_id, user_idx, prod_idx, review_idx).  Returns None when no sample survives, ONE ReviewTrainBatch, or --
        pv / pvc encoders with prepare_pv -- the list of per-window batches."""
        samples = self._samples(rows)
        if not samples:
            return None
        if self.review_words is None:
            raise RuntimeError("call initialize_epoch() first (review words are sub-sampled once per epoch)")
        query = np.stack([q for q, _, _ in samples]).astype(np.int64)
        pos = [p for _, p, _ in samples]
        pos_ridxs = _pad_rows([p[0] for p in pos], self.review_pad_idx)
        pos_seg = _pad_rows([p[1] for p in pos], self.seg_pad_idx)
        pos_user = _pad_rows([p[2] for p in pos], self.user_pad_idx)
        pos_item = _pad_rows([p[3] for p in pos], self.prod_pad_idx)
        pos_words = self.review_words[pos_ridxs]                                   # [B, rc, word_limit]
        pos_masks = self._word_masks(pos_words)
        neg_ridxs = _pad_nested([[n[0] for n in ns] for _, _, ns in samples], self.review_pad_idx)
        neg_seg = _pad_nested([[n[1] for n in ns] for _, _, ns in samples], self.seg_pad_idx)
        neg_user = _pad_nested([[n[2] for n in ns] for _, _, ns in samples], self.user_pad_idx)
        neg_item = _pad_nested([[n[3] for n in ns] for _, _, ns in samples], self.prod_pad_idx)
        neg_words = self.review_words[neg_ridxs]                                   # [B, K', rc', word_limit]
        t = torch.from_numpy

        def u8(m):
            return t(np.ascontiguousarray(m).astype(np.uint8))
        if "pv" not in self.args.review_encoder_name or not prepare_pv:
            neg_masks = self._word_masks(neg_words)
            return ReviewTrainBatch(
                query_word_idxs=t(query), pos_prod_ridxs=t(pos_ridxs), pos_seg_idxs=t(pos_seg),
                pos_prod_rword_idxs=t(pos_words), pos_prod_rword_masks=u8(pos_masks), neg_prod_ridxs=t(neg_ridxs),
                neg_seg_idxs=t(neg_seg), pos_user_idxs=t(pos_user), neg_user_idxs=t(neg_user),
                pos_item_idxs=t(pos_item), neg_item_idxs=t(neg_item), neg_prod_rword_idxs=t(neg_words),
                neg_prod_rword_masks=u8(neg_masks), pos_prod_rword_idxs_pvc=None, neg_prod_rword_idxs_pvc=None)
        B, rc, L = pos_words.shape
        W = int(self.args.pv_window_size)
        if self.args.shuffle_review_words:
            for sample_rows in pos_words:            # [rc, L] views: permutes the word ROWS of a sample, in place
                self.np_random.shuffle(sample_rows)
        win_words = _windows(pos_words.reshape(-1, L), W, self.word_pad_idx)       # [seg, B * rc, W]
        win_masks = _windows(pos_masks.reshape(-1, L), W, False)
        seg = win_words.shape[0]
        win_words = win_words.reshape(seg * B, rc, W)
        win_masks = win_masks.reshape(seg * B, rc, W)
        pick = np.tile(np.arange(B), (seg, 1))
        if shuffle:
            perm = self.np_random.permutation(B * seg)
            pick = pick.reshape(-1)[perm].reshape(seg, B)
            win_words, win_masks = win_words[perm], win_masks[perm]
        win_words = win_words.reshape(seg, B, rc, W)
        win_masks = win_masks.reshape(seg, B, rc, W)
        out = []
        for i in range(seg):
            b = pick[i]
            out.append(ReviewTrainBatch(
                query_word_idxs=t(query[b]), pos_prod_ridxs=t(pos_ridxs[b]), pos_seg_idxs=t(pos_seg[b]),
                pos_prod_rword_idxs=t(np.ascontiguousarray(win_words[i])), pos_prod_rword_masks=u8(win_masks[i]),
                neg_prod_ridxs=t(neg_ridxs[b]), neg_seg_idxs=t(neg_seg[b]), pos_user_idxs=t(pos_user[b]),
                neg_user_idxs=t(neg_user[b]), pos_item_idxs=t(pos_item[b]), neg_item_idxs=t(neg_item[b]),
                neg_prod_rword_idxs=None, neg_prod_rword_masks=None, pos_prod_rword_idxs_pvc=t(pos_words[b]),
                neg_prod_rword_idxs_pvc=t(neg_words[b])))
        return out

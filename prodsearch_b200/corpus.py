"""Device-resident corpus index and on-device batch construction for the item-transformer (TEM) path
(SURVEY.md 8(f) N4).

The reference keeps the corpus relations as nested Python lists on ``GlobalProdSearchData`` / ``ProdSearchData``
(data/data_util.py:165-203,:10-62) and assembles every batch sample by sample in
``ItemPVDataloader.get_train_batch`` / ``get_test_batch`` (data/item_pv_dataloader.py:122-143,:31-49), calling
``get_user_review_idxs`` (:85-105) per sample.  Here the same relations are flattened ONCE into CSR arrays in
HBM (``ItemCorpus``) and ``psb_build_item_batch`` builds a whole batch in one launch; the result carries the
attribute names of ``ItemPVBatch`` (data/batch_data.py:3-37) so the models take it unchanged.
"""
import argparse
import ctypes

import numpy as np
import torch

from . import _lib


class ItemBatch(argparse.Namespace):
    """Same attribute names as the reference's ItemPVBatch; tensors already live on the device."""

    def to(self, device):
        return self


def _csr(lists, dtype=np.int32):
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    for i, l in enumerate(lists):
        off[i + 1] = off[i] + len(l)
    flat = np.empty(int(off[-1]), dtype=dtype)
    for i, l in enumerate(lists):
        flat[off[i]:off[i + 1]] = l
    return off, flat


class ItemCorpus(object):
    """CSR view of the corpus in HBM.

    user_seq      global_data.u_r_seq            list[user] -> review ids in time order
    review_u_p    global_data.review_u_p         list[review] -> (user, item)
    review_uloc   global_data.review_loc_time    list[review] -> (loc_in_user, loc_in_item, time); column 0
    train_reviews prod_data.u_reviews            per-user sets of training-split reviews
    item_queries  prod_data.product_query_idx    list[item] -> query ids
    query_words   global_data.query_words        [Q, wq] already padded with vocab_size - 1
    """

    def __init__(self, device, user_seq, review_u_p, query_words, item_queries, train_reviews=None,
                 review_uloc=None, product_size=None, vocab_size=None, item_seq=None, review_time=None):
        self._setup(device, vocab_size=vocab_size, **self.lists_to_arrays(
            user_seq, review_u_p, query_words, item_queries, train_reviews, review_uloc, product_size, item_seq,
            review_time))

    @staticmethod
    def lists_to_arrays(user_seq, review_u_p, query_words, item_queries, train_reviews=None, review_uloc=None,
                        product_size=None, item_seq=None, review_time=None):
        """Nested lists of the reference's data objects -> the keyword arguments of ``host_arrays``."""
        rup = np.asarray(review_u_p, dtype=np.int64).reshape(-1, 2)
        R = rup.shape[0]
        P = int(product_size) if product_size is not None else len(item_queries)
        in_set = np.zeros(R, dtype=np.uint8)
        if train_reviews is None:
            in_set[:] = 1
        else:
            for s in train_reviews:
                if len(s):
                    in_set[np.fromiter(s, dtype=np.int64, count=len(s))] = 1
        uloc = None
        if review_uloc is not None:
            uloc = np.asarray([x[0] if np.ndim(x) else x for x in review_uloc], dtype=np.int32)
        if review_time is None and review_uloc is not None and len(review_uloc) and np.ndim(review_uloc[0]) and \
                len(review_uloc[0]) >= 3:
            review_time = [x[2] for x in review_uloc]
        return dict(review_u_p=rup, review_uloc=uloc, review_time=review_time, review_in_set=in_set,
                    user_seq=_csr(user_seq),
                    item_seq=None if item_seq is None else _csr(list(item_seq) + [[]] * (P - len(item_seq))),
                    item_query=_csr(list(item_queries) + [[]] * (P - len(item_queries))), query_words=query_words,
                    product_size=P)

    @classmethod
    def from_arrays(cls, device, review_u_p, review_in_set, user_seq, item_query, query_words, product_size,
                    vocab_size=None, review_uloc=None, review_time=None, item_seq=None):
        """From flat arrays (``data_files.CorpusFiles``): user_seq / item_seq / item_query are (offsets, flat) pairs."""
        self = cls.__new__(cls)
        self._setup(device, review_u_p=np.asarray(review_u_p, dtype=np.int64).reshape(-1, 2), review_uloc=review_uloc,
                    review_time=review_time, review_in_set=np.asarray(review_in_set, dtype=np.uint8),
                    user_seq=user_seq, item_seq=item_seq, item_query=item_query, query_words=query_words,
                    product_size=int(product_size), vocab_size=vocab_size)
        return self

    @staticmethod
    def host_arrays(review_u_p, review_uloc, review_time, review_in_set, user_seq, item_seq, item_query, query_words,
                    product_size):
        """The flat arrays exactly as they are uploaded (name -> contiguous numpy array or None); pure host code, so
        the CPU tests can hold both constructors to the nested lists of the reference's loaders."""
        rup = np.asarray(review_u_p, dtype=np.int64).reshape(-1, 2)
        qw = np.asarray(query_words, dtype=np.int64)
        if qw.ndim != 2:
            raise ValueError("query_words must be padded to a rectangle (global_data.query_words is)")
        P = int(product_size)
        if len(item_query[0]) - 1 != P or (item_seq is not None and len(item_seq[0]) - 1 != P):
            raise ValueError("item_query / item_seq need one (possibly empty) row per product")

        def arr(a, dtype):
            return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=dtype))
        h = dict(review_user=arr(rup[:, 0], np.int32), review_item=arr(rup[:, 1], np.int32),
                 review_uloc=arr(review_uloc, np.int32), review_in_set=arr(review_in_set, np.uint8),
                 user_seq_off=arr(user_seq[0], np.int64), user_seq=arr(user_seq[1], np.int32),
                 item_query_off=arr(item_query[0], np.int64), item_query=arr(item_query[1], np.int32),
                 query_words=arr(qw, np.int64), item_seq_off=None, item_seq=None,
                 review_time=arr(review_time, np.int64))
        # review-transformer batches: every item's reviews in time order + the time stamps (i_r_seq, review_loc_time[:, 2])
        if item_seq is not None:
            h["item_seq_off"], h["item_seq"] = arr(item_seq[0], np.int64), arr(item_seq[1], np.int32)
        if h["review_in_set"].shape[0] != rup.shape[0]:
            raise ValueError("review_in_set needs one flag per review")
        return h

    def _setup(self, device, review_u_p, review_uloc, review_time, review_in_set, user_seq, item_seq, item_query,
               query_words, product_size, vocab_size):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("ItemCorpus lives in GPU memory (no CPU fallback for batch construction)")
        h = self.host_arrays(review_u_p, review_uloc, review_time, review_in_set, user_seq, item_seq, item_query,
                             query_words, product_size)
        self.device = dev
        for name, a in h.items():
            setattr(self, name, None if a is None else torch.from_numpy(a).to(dev))
        qw = h["query_words"]
        R, P, U = h["review_user"].shape[0], int(product_size), len(h["user_seq_off"]) - 1
        self.n_reviews, self.n_users, self.n_items, self.n_queries, self.wq = R, U, P, qw.shape[0], qw.shape[1]
        self.prod_pad_idx = P                                       # item_pv_dataset.py:24
        self.word_pad_idx = (int(vocab_size) if vocab_size is not None else int(qw.max()) + 1) - 1
        c = _lib.Corpus()
        self.user_pad_idx = U                                       # prod_search_dataset.py:23
        self.review_pad_idx = R                                     # review_count - 1 (:26; review_count = R + 1)
        self.seg_pad_idx = 3
        for name in ("review_user", "review_item", "review_uloc", "review_in_set", "user_seq_off", "user_seq",
                     "item_query_off", "item_query", "query_words", "item_seq_off", "item_seq", "review_time"):
            t = getattr(self, name)
            setattr(c, name, None if t is None else t.data_ptr())
        c.n_reviews, c.n_users, c.n_items, c.n_queries, c.wq, c.word_pad = R, U, P, qw.shape[0], qw.shape[1], \
            self.word_pad_idx
        self._c = c

    @classmethod
    def from_reference(cls, device, global_data, prod_data):
        """From the reference's own data objects (data/data_util.py)."""
        return cls(device, global_data.u_r_seq, global_data.review_u_p, global_data.query_words,
                   prod_data.product_query_idx, train_reviews=prod_data.u_reviews,
                   review_uloc=global_data.review_loc_time, product_size=global_data.product_size,
                   vocab_size=global_data.vocab_size, item_seq=global_data.i_r_seq)

    # ------------------------------------------------------------------ batches
    def _dev_i64(self, x):
        if x is None:
            return None
        if torch.is_tensor(x):
            return x.to(device=self.device, dtype=torch.int64).contiguous()
        return torch.as_tensor(np.asarray(x, dtype=np.int64), device=self.device)

    def build(self, review_idxs, hist_limit, mode, user_idxs=None, item_idxs=None, query_idxs=None, query_pick=None,
              seed=0, trim=True):
        """One psb_build_item_batch launch.  Returns (target_prod_idxs, query_idxs, query_word_idxs, u_item_idxs,
        hist_len); with trim=True the history is cut to the batch maximum like util.pad does (one host sync);
        trim=False keeps the fixed [B, hist_limit] width (graph-capturable, identical model output: the extra
        columns are pad items)."""
        rv = self._dev_i64(review_idxs)
        B = rv.numel()
        us, it, qi = self._dev_i64(user_idxs), self._dev_i64(item_idxs), self._dev_i64(query_idxs)
        if qi is None:
            if query_pick is None:                       # random.choice(product_query_idx[prod]) (:131)
                qp = torch.randint(0, 2 ** 31 - 1, (B,), device=self.device, dtype=torch.int32)
            elif torch.is_tensor(query_pick) and query_pick.is_cuda:
                qp = query_pick.to(torch.int32).contiguous()          # values < 2^31
            else:                                        # any uint32 words supplied from the host
                host = np.asarray(query_pick).astype(np.uint32).view(np.int32)
                qp = torch.from_numpy(np.ascontiguousarray(host)).to(self.device)
        else:
            qp = None
        L = int(hist_limit)
        target = torch.empty(B, dtype=torch.int64, device=self.device)
        q_out = torch.empty(B, dtype=torch.int64, device=self.device)
        qw = torch.empty(B, self.wq, dtype=torch.int64, device=self.device)
        hist = torch.empty(B, L, dtype=torch.int64, device=self.device)
        hlen = torch.empty(B, dtype=torch.int32, device=self.device)
        err = torch.zeros(1, dtype=torch.int32, device=self.device)
        lib = _lib.load()
        _lib.check(lib.psb_build_item_batch(
            ctypes.byref(self._c), _lib.ptr(rv), _lib.ptr(us), _lib.ptr(it), _lib.ptr(qi), _lib.ptr(qp), B, L,
            int(mode), int(seed) & 0xFFFFFFFF, self.prod_pad_idx, _lib.ptr(target), _lib.ptr(q_out), _lib.ptr(qw),
            _lib.ptr(hist), _lib.ptr(hlen), _lib.ptr(err), _lib.stream_ptr()), "psb_build_item_batch")
        if trim:
            if int(err.item()) != 0:
                raise IndexError("psb_build_item_batch: review / user / item / query id out of range")
            hist = hist[:, :max(int(hlen.max().item()), 0)].contiguous() if B else hist
        return target, q_out, qw, hist, hlen

    def train_batch(self, review_idxs, pos_iword_idxs, args, query_pick=None, seed=0, trim=True):
        """ItemPVDataloader.get_train_batch (item_pv_dataloader.py:122-143) for samples (word_idxs, review_idx)."""
        if args.do_seq_review_train:
            mode = _lib.HIST_SEQ
        else:
            mode = _lib.HIST_LAST if args.fix_train_review else _lib.HIST_RANDOM
        target, q, qw, hist, hlen = self.build(review_idxs, args.uprev_review_limit, mode, query_pick=query_pick,
                                               seed=seed, trim=trim)
        return ItemBatch(query_word_idxs=qw, target_prod_idxs=target, u_item_idxs=hist,
                         pos_iword_idxs=self._dev_i64(pos_iword_idxs), query_idxs=q,
                         user_idxs=self.review_user[self._dev_i64(review_idxs)].to(torch.int64), candi_prod_idxs=[],
                         hist_len=hlen)

    def test_batch(self, query_idxs, user_idxs, prod_idxs, review_idxs, args, candi_prod_idxs=None, trim=True):
        """ItemPVDataloader.get_test_batch (item_pv_dataloader.py:31-49) for entries
        (query_idx, user_idx, prod_idx, review_idx[, candidates]) of collect_test_samples."""
        do_seq = args.do_seq_review_test and not args.train_review_only
        mode = _lib.HIST_SEQ if do_seq else _lib.HIST_LAST
        target, q, qw, hist, hlen = self.build(review_idxs, args.uprev_review_limit, mode, user_idxs=user_idxs,
                                               item_idxs=prod_idxs, query_idxs=query_idxs, trim=trim)
        return ItemBatch(query_word_idxs=qw, target_prod_idxs=target, u_item_idxs=hist, pos_iword_idxs=[],
                         query_idxs=q, user_idxs=self._dev_i64(user_idxs),
                         candi_prod_idxs=self._dev_i64(candi_prod_idxs) if candi_prod_idxs is not None else [],
                         hist_len=hlen)


def _review_test_batch(self, query_idxs, user_idxs, prod_idxs, review_idxs, candi_prod_idxs, args, trim=True):
    """ProdSearchDataLoader.get_test_batch (data/prod_search_dataloader.py:44-109) for entries
    (query_idx, user_idx, prod_idx, review_idx) with candidate lists candi_prod_idxs [B, C] (-1 = padding, :92):
    one psb_build_review_test_batch launch instead of the python loop over every candidate.  The result carries the
    attribute names of ProdSearchTestBatch (data/batch_data.py:94-135)."""
    if self.item_seq is None:
        raise RuntimeError("this corpus was built without item_seq (global_data.i_r_seq)")
    do_seq = args.do_seq_review_test and not args.train_review_only
    mode = _lib.HIST_SEQ if do_seq else _lib.HIST_LAST
    rv, us = self._dev_i64(review_idxs), self._dev_i64(user_idxs)
    cand = self._dev_i64(candi_prod_idxs)
    B, C = cand.shape
    Lu, Li = int(args.uprev_review_limit), int(args.iprev_review_limit)
    W = Lu + Li
    dev = self.device
    ridxs = torch.empty(B, C, W, dtype=torch.int64, device=dev)
    seg = torch.empty(B, C, W + 1, dtype=torch.int64, device=dev)
    users = torch.empty(B, C, W + 1, dtype=torch.int64, device=dev)
    items = torch.empty(B, C, W + 1, dtype=torch.int64, device=dev)
    slen = torch.empty(B, C, dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(_lib.load().psb_build_review_test_batch(
        ctypes.byref(self._c), _lib.ptr(rv), _lib.ptr(us), _lib.ptr(cand), B, C, Lu, Li, mode, self.review_pad_idx,
        self.user_pad_idx, self.prod_pad_idx, self.seg_pad_idx, _lib.ptr(ridxs), _lib.ptr(seg), _lib.ptr(users),
        _lib.ptr(items), _lib.ptr(slen), _lib.ptr(err), _lib.stream_ptr()), "psb_build_review_test_batch")
    if trim:
        if int(err.item()) != 0:
            raise IndexError("psb_build_review_test_batch: review / user / candidate id out of range")
        w = int(slen.max().item()) if B * C else 0
        ridxs, seg = ridxs[:, :, :w].contiguous(), seg[:, :, :w + 1].contiguous()
        users, items = users[:, :, :w + 1].contiguous(), items[:, :, :w + 1].contiguous()
    qi = self._dev_i64(query_idxs)
    return ItemBatch(query_idxs=qi, user_idxs=us, target_prod_idxs=self._dev_i64(prod_idxs), candi_prod_idxs=cand,
                     query_word_idxs=self.query_words[qi], candi_prod_ridxs=ridxs, candi_seg_idxs=seg,
                     candi_seq_user_idxs=users, candi_seq_item_idxs=items, seq_len=slen)


ItemCorpus.review_test_batch = _review_test_batch


def subset_key(seed, sample, pos):
    """Host evaluation of the PSB_HIST_RANDOM key (the same function the kernel runs)."""
    return int(_lib.load().psb_subset_key(int(seed) & 0xFFFFFFFF, int(sample) & 0xFFFFFFFF, int(pos) & 0xFFFFFFFF))

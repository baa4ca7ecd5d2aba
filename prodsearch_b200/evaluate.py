"""Evaluation driver for the item-transformer path: fused full-catalog ranking, metrics and the TREC run file
(replaces Trainer.get_prod_scores / validate / test / calc_metrics, trainer.py:126-226).

The reference scores every (query, candidate) pair in 500-candidate segments, copies the [M, N] score matrix to
the host, argsorts it there and writes the first ``cutoff`` entries.  Here batches are built on the device
(corpus.ItemCorpus), ``model.rank_catalog`` returns the top-k lists directly, ``psb_target_rank`` finds the
target's rank inside them and ``psb_write_ranklist`` formats the run file -- [M, N] never exists.
Ranking contract: descending score, ties -> lower item id.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def target_ranks(ids, target):
    """rank [M] int32: 1-based position of target[i] in ids[i], 0 if absent (psb_target_rank)."""
    ids = ids.contiguous()
    target = target.to(device=ids.device, dtype=torch.int64).contiguous()
    m, k = ids.shape
    rank = torch.empty(m, dtype=torch.int32, device=ids.device)
    _lib.check(_lib.load().psb_target_rank(_lib.ptr(ids, torch.int64), _lib.ptr(target), m, k, _lib.ptr(rank),
                                           _lib.stream_ptr()), "psb_target_rank")
    return rank


def calc_metrics(ranks, cutoff=100):
    """MRR / P@1 with the reference's arithmetic (trainer.py:171-186): python-float accumulation in query order."""
    r = np.asarray(ranks.cpu() if torch.is_tensor(ranks) else ranks)
    mrr, prec = 0, 0
    for rank in r.tolist():
        if rank <= 0:
            continue
        if cutoff < 0 or rank <= cutoff:
            mrr += 1 / rank
        if rank == 1:
            prec += 1
    n = max(len(r), 1)
    return mrr / n, prec / n


def _c_strings(strings):
    arr = (ctypes.c_char_p * len(strings))()
    arr[:] = [s if isinstance(s, bytes) else str(s).encode() for s in strings]
    return arr


def write_ranklist(path, user_ids, user_idxs, query_idxs, product_ids, ids, scores, cutoff=100, append=False):
    """trainer.py:158-169.  ids / scores: [M, k] top-k lists (device or host); user_ids / product_ids: the
    global_data string tables (lists of str, or ctypes arrays prepared once with ``_c_strings``)."""
    ids_h = ids.detach().to("cpu", torch.int64).contiguous() if torch.is_tensor(ids) else \
        torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64))
    sc_h = scores.detach().to("cpu", torch.float32).contiguous() if torch.is_tensor(scores) else \
        torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float32))
    u_h = torch.as_tensor(np.asarray(user_idxs.cpu() if torch.is_tensor(user_idxs) else user_idxs, dtype=np.int64))
    q_h = torch.as_tensor(np.asarray(query_idxs.cpu() if torch.is_tensor(query_idxs) else query_idxs, dtype=np.int64))
    m, k = ids_h.shape
    if u_h.numel() != m or q_h.numel() != m:
        raise ValueError("one user / query index per ranked list")
    if m and (int(u_h.max()) >= len(user_ids) or int(ids_h.max()) >= len(product_ids) or int(u_h.min()) < 0):
        raise IndexError("user / product index outside the id tables")
    ua = user_ids if isinstance(user_ids, ctypes.Array) else _c_strings(user_ids)
    pa = product_ids if isinstance(product_ids, ctypes.Array) else _c_strings(product_ids)
    n = _lib.load().psb_write_ranklist(str(path).encode(), ua, u_h.data_ptr(), q_h.data_ptr(), pa, ids_h.data_ptr(),
                                       sc_h.data_ptr(), m, k, int(cutoff), 1 if append else 0)
    if n < 0:
        raise RuntimeError("psb_write_ranklist failed: %d" % n)
    return int(n)


def rank_test_set(model, corpus, entries, args, k=100, batch_size=4096, mode=None):
    """Full-catalog ranking of test entries (query_idx, user_idx, prod_idx, review_idx)
    (item_pv_dataset.py:36-68 without the per-segment candidate lists: the whole catalog is the candidate set,
    ``test_candi_size < 1``).  Returns device tensors (ids [M,k], scores [M,k], target [M], query_idx [M],
    user_idx [M]) -- the 5-tuple of Trainer.get_prod_scores (trainer.py:226) with top-k lists in place of
    the [M, N] matrices."""
    e = np.asarray(entries, dtype=np.int64).reshape(-1, 4)
    out_i, out_s = [], []
    was_training = model.training
    model.eval()
    try:
        for s in range(0, e.shape[0], batch_size):
            c = e[s:s + batch_size]
            batch = corpus.test_batch(c[:, 0], c[:, 1], c[:, 2], c[:, 3], args)
            ids, sc = model.rank_catalog(batch, k=k) if mode is None else model.rank_catalog(batch, k=k, mode=mode)
            out_i.append(ids)
            out_s.append(sc)
    finally:
        model.train(was_training)
    dev = corpus.device
    cat = (lambda xs, dt: torch.cat(xs) if xs else torch.empty(0, k, dtype=dt, device=dev))
    return (cat(out_i, torch.int64), cat(out_s, torch.float32), torch.as_tensor(e[:, 2], device=dev),
            torch.as_tensor(e[:, 0], device=dev), torch.as_tensor(e[:, 1], device=dev))


def test(model, corpus, entries, args, user_ids, product_ids, rank_path=None, cutoff=100, batch_size=4096):
    """Trainer.test (trainer.py:140-169): rank, report MRR / P@1, write the run file."""
    # cutoff < 0 = "no cutoff" (calc_metrics); the reference writes min(cutoff, candidate_size) rows (trainer.py:164)
    n_items = int(corpus.prod_pad_idx)
    k = n_items if cutoff < 0 else max(1, min(int(cutoff), n_items))
    ids, scores, target, q_idx, u_idx = rank_test_set(model, corpus, entries, args, k=k, batch_size=batch_size)
    mrr, prec = calc_metrics(target_ranks(ids, target), cutoff)
    if rank_path is not None:
        write_ranklist(rank_path, user_ids, u_idx, q_idx, product_ids, ids, scores, cutoff)
    return mrr, prec


def rank_candidates(cand_ids, cand_scores, target, pad_ids=(-1,)):
    """Rank explicit candidate lists: descending score, ties -> lower item id; entries whose id is a padding value
    rank last.  cand_ids / cand_scores [M, C] (any device), target [M].  Returns (ranked ids [M, C], ranked scores
    [M, C], rank [M] int32: 1-based position of the target's first occurrence, 0 if absent) -- what
    ``argsort(axis=-1)[:, ::-1]`` + ``np.where`` compute at trainer.py:136,:174-178, with the tie rule made definite."""
    ids = cand_ids.to(torch.int64)
    sc = cand_scores.to(torch.float32).clone()
    pad = torch.zeros_like(ids, dtype=torch.bool)
    for p in pad_ids:
        pad |= ids == int(p)
    sc[pad] = float("-inf")
    key = torch.where(pad, torch.full_like(ids, torch.iinfo(torch.int64).max), ids)
    by_id = torch.argsort(key, dim=1, stable=True)                                   # ascending id, padding last
    order = torch.gather(by_id, 1, torch.argsort(torch.gather(sc, 1, by_id), dim=1, descending=True, stable=True))
    r_ids, r_sc = torch.gather(ids, 1, order), torch.gather(sc, 1, order)
    hit = (r_ids == target.to(ids.device, torch.int64).view(-1, 1)) & ~torch.gather(pad, 1, order)
    first = torch.argmax(hit.to(torch.int8), dim=1)
    rank = torch.where(hit.any(dim=1), first + 1, torch.zeros_like(first)).to(torch.int32)
    return r_ids, r_sc, rank


def validate(model, corpus, entries, candidates, args, cutoff=100, batch_size=4096, pad_id=-1):
    """Trainer.validate (trainer.py:124-138) / the candidate-list form of Trainer.test: score explicit candidate
    lists (``data_files.SplitFiles.test_samples``: one row per (user, query) pair and per segment of
    ``candi_batch_size`` candidates, padded with ``pad_id``) with ``model.test`` and report MRR / P@1.  Segments of the
    same pair are consecutive rows and are joined before ranking, as get_prod_scores does (trainer.py:215-221).
    Returns (mrr, prec, ranked ids [pairs, C], ranked scores [pairs, C], query_idx [pairs], user_idx [pairs])."""
    e = np.asarray(entries, dtype=np.int64).reshape(-1, 4)
    cand = np.asarray(candidates, dtype=np.int64).reshape(e.shape[0], -1)
    # segments per pair: consecutive rows with the same (query, user, item, review)
    new_pair = np.ones(e.shape[0], dtype=bool)
    new_pair[1:] = (e[1:] != e[:-1]).any(axis=1)
    pair_of = np.cumsum(new_pair) - 1
    n_pairs = int(pair_of[-1]) + 1 if e.shape[0] else 0
    seg_of = np.arange(e.shape[0]) - np.flatnonzero(new_pair)[pair_of]
    seg_count = int(seg_of.max()) + 1 if e.shape[0] else 1
    item_pad = corpus.prod_pad_idx
    scores = []
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad():
            for s in range(0, e.shape[0], batch_size):
                c = e[s:s + batch_size]
                cb = np.where(cand[s:s + batch_size] == pad_id, item_pad, cand[s:s + batch_size])   # :44: item pad row
                batch = corpus.test_batch(c[:, 0], c[:, 1], c[:, 2], c[:, 3], args, candi_prod_idxs=cb)
                scores.append(model.test(batch))
    finally:
        model.train(was_training)
    dev = scores[0].device if scores else torch.device("cpu")
    sc = torch.cat(scores) if scores else torch.empty(0, cand.shape[1])
    W = cand.shape[1]
    ids_all = torch.full((n_pairs, seg_count * W), pad_id, dtype=torch.int64, device=dev)
    sc_all = torch.full((n_pairs, seg_count * W), float("-inf"), dtype=torch.float32, device=dev)
    rows = torch.as_tensor(pair_of, device=dev)
    cols = torch.as_tensor(seg_of * W, device=dev).view(-1, 1) + torch.arange(W, device=dev).view(1, -1)
    ids_all[rows.view(-1, 1), cols] = torch.as_tensor(cand, device=dev)
    sc_all[rows.view(-1, 1), cols] = sc.to(torch.float32)
    first = np.flatnonzero(new_pair)
    r_ids, r_sc, rank = rank_candidates(ids_all, sc_all, torch.as_tensor(e[first, 2], device=dev), pad_ids=(pad_id,))
    mrr, prec = calc_metrics(rank, cutoff)
    return mrr, prec, r_ids, r_sc, e[first, 0], e[first, 1]


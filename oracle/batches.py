"""Batch construction, metrics and run-file side of the oracle.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Plain-Python restatement of
  * ItemPVDataloader.get_user_review_idxs     data/item_pv_dataloader.py:85-105
  * ItemPVDataloader.get_train_batch          data/item_pv_dataloader.py:122-143
  * ItemPVDataloader.get_test_batch           data/item_pv_dataloader.py:31-49
  * util.pad                                  others/util.py:37-41
  * Trainer.test's run-file loop              trainer.py:158-169
working on the same nested lists the reference keeps (u_r_seq, review_u_p, review_loc_time, u_reviews,
product_query_idx, query_words).  Two RNG draws of the reference are inputs here: the query pick
(``random.choice`` at :131 -> ``queries[pick % len(queries)]``) and the random history subset
(``random.sample`` at :99 -> the hist_limit candidates with the smallest ``subset_key``).
Pinned by tests/golden/batches.npz, produced by the reference's own methods
(tests/golden/make_golden_batches.py); the random-subset mode cannot be pinned that way (python's RNG stream)
and is checked for its defining properties instead.
"""
import numpy as np

__all__ = ["subset_key", "user_review_idxs", "item_review_idxs", "pad", "pad_3d", "item_train_batch",
           "item_test_batch", "review_test_batch", "ranklist_lines"]

_M = 0xFFFFFFFF


def subset_key(seed, sample, pos):
    """lowbias32 of (seed, sample, position); identical to psb_subset_key."""
    h = (seed ^ ((sample * 0x9E3779B1) & _M) ^ ((pos * 0x85EBCA77) & _M)) & _M
    h ^= h >> 16
    h = (h * 0x7FEB352D) & _M
    h ^= h >> 15
    h = (h * 0x846CA68B) & _M
    h ^= h >> 16
    return h


def user_review_idxs(u_r_seq, train_set, review_loc_time, user_idx, review_idx, limit, do_seq, fix=True,
                     seed=0, sample=0):
    """item_pv_dataloader.py:85-105."""
    seq = u_r_seq[user_idx]
    if do_seq:
        loc = review_loc_time[review_idx][0]
        return list(seq[:loc][-limit:])
    cand = [(p, x) for p, x in enumerate(seq) if x in train_set[user_idx] and x != review_idx]
    if len(cand) > limit:
        if fix:
            cand = cand[-limit:]
        else:
            keyed = sorted(cand, key=lambda px: (subset_key(seed, sample, px[0]), px[0]))[:limit]
            keep = set(p for p, _ in keyed)
            cand = [(p, x) for p, x in cand if p in keep]
    return [x for _, x in cand]


def pad(data, pad_id, width=-1):
    """others/util.py:37-41."""
    if width == -1:
        width = max(len(d) for d in data) if len(data) else 0
    return [list(d[:width]) + [pad_id] * (width - len(d)) for d in data]


def item_train_batch(corpus, samples, query_pick, limit, do_seq, fix, prod_pad_idx, seed=0):
    """item_pv_dataloader.py:122-143.  samples: [(word_idxs, review_idx)]; corpus: dict of the nested lists."""
    qw, target, hist, words, users, queries = [], [], [], [], [], []
    for b, (word_idxs, review_idx) in enumerate(samples):
        user_idx, prod_idx = corpus["review_u_p"][review_idx]
        qs = corpus["product_query_idx"][prod_idx]
        query_idx = qs[int(query_pick[b]) % len(qs)]
        prev = user_review_idxs(corpus["u_r_seq"], corpus["u_reviews"], corpus["review_loc_time"], user_idx,
                                review_idx, limit, do_seq, fix, seed, b)
        qw.append(list(corpus["query_words"][query_idx]))
        target.append(prod_idx)
        hist.append([corpus["review_u_p"][x][1] for x in prev])
        words.append(list(word_idxs))
        users.append(user_idx)
        queries.append(query_idx)
    return dict(query_word_idxs=np.asarray(qw, np.int64), target_prod_idxs=np.asarray(target, np.int64),
                u_item_idxs=np.asarray(pad(hist, prod_pad_idx), np.int64).reshape(len(samples), -1),
                pos_iword_idxs=np.asarray(words, np.int64), user_idxs=np.asarray(users, np.int64),
                query_idxs=np.asarray(queries, np.int64), hist_len=np.asarray([len(h) for h in hist], np.int32))


def item_test_batch(corpus, entries, limit, do_seq, prod_pad_idx):
    """item_pv_dataloader.py:31-49.  entries: [(query_idx, user_idx, prod_idx, review_idx)]."""
    qw, hist = [], []
    for query_idx, user_idx, prod_idx, review_idx in entries:
        prev = user_review_idxs(corpus["u_r_seq"], corpus["u_reviews"], corpus["review_loc_time"], user_idx,
                                review_idx, limit, do_seq, True)
        qw.append(list(corpus["query_words"][query_idx]))
        hist.append([corpus["review_u_p"][x][1] for x in prev])
    e = np.asarray(entries, np.int64).reshape(-1, 4)
    return dict(query_word_idxs=np.asarray(qw, np.int64), target_prod_idxs=e[:, 2].copy(),
                u_item_idxs=np.asarray(pad(hist, prod_pad_idx), np.int64).reshape(len(entries), -1),
                user_idxs=e[:, 1].copy(), query_idxs=e[:, 0].copy(),
                hist_len=np.asarray([len(h) for h in hist], np.int32))


def bisect_right(review_arr, review_loc_time, timestamp):
    """data/prod_search_dataset.py:135-152."""
    lo, hi = 0, len(review_arr)
    while lo < hi:
        mid = (lo + hi) // 2
        if timestamp < review_loc_time[review_arr[mid]][2]:
            hi = mid
        else:
            lo = mid + 1
    return lo


def item_review_idxs(i_r_seq, train_set, review_loc_time, prod_idx, review_idx, limit, do_seq, review_time_stamp=None):
    """data/prod_search_dataloader.py:135-160 with fix=True."""
    seq = i_r_seq[prod_idx]
    if do_seq:
        if review_idx is None:
            loc = bisect_right(seq, review_loc_time, review_time_stamp)
        else:
            loc = review_loc_time[review_idx][1]
        if loc == 0:
            return []
        return list(seq[:loc][-limit:])
    cand = [x for x in seq if x in train_set[prod_idx] and x != review_idx]
    return cand[-limit:] if len(cand) > limit else cand


def pad_3d(data, pad_id, dim=1, width=-1):
    """others/util.py:43-62."""
    if width == -1:
        if dim == 1:
            width = max(len(d) for d in data)
        else:
            for entry in data:
                width = max(width, max(len(d) for d in entry))
    if dim == 1:
        return [list(d[:width]) + [[pad_id] * len(data[0][0])] * (width - len(d)) for d in data]
    return [[list(d[:width]) + [pad_id] * (width - len(d)) for d in entry] for entry in data]


def review_test_batch(corpus, entries, candidates, u_limit, i_limit, do_seq_review_test, train_review_only, pads):
    """data/prod_search_dataloader.py:44-109.  entries: [(query_idx, user_idx, prod_idx, review_idx)];
    candidates: per entry list of candidate items; pads: dict(review=, user=, prod=, seg=)."""
    total = u_limit + i_limit
    do_seq = do_seq_review_test and not train_review_only
    all_r, all_s, all_u, all_i = [], [], [], []
    for (query_idx, user_idx, prod_idx, review_idx), cands in zip(entries, candidates):
        u_prev = user_review_idxs(corpus["u_r_seq"], corpus["u_reviews"], corpus["review_loc_time"], user_idx,
                                  review_idx, u_limit, do_seq, True)
        ts = corpus["review_loc_time"][review_idx][2] if do_seq_review_test else None
        u_items = [corpus["review_u_p"][x][1] for x in u_prev]
        br, bs, bu, bi = [], [], [], []
        for c in cands:
            c_prev = item_review_idxs(corpus["i_r_seq"], corpus["p_reviews"], corpus["review_loc_time"], c, None,
                                      i_limit, do_seq, ts)
            c_users = [corpus["review_u_p"][x][0] for x in c_prev]
            bu.append(([pads["user"]] + [user_idx] * len(u_prev) + c_users)[:total + 1])
            bi.append(([pads["prod"]] + u_items + [c] * len(c_prev))[:total + 1])
            bs.append(([0] + [1] * len(u_prev) + [2] * len(c_prev))[:total + 1])
            br.append((u_prev + c_prev)[:total])
        all_r.append(br)
        all_s.append(bs)
        all_u.append(bu)
        all_i.append(bi)
    out = {}
    for name, data, pad_id in (("candi_prod_ridxs", all_r, pads["review"]), ("candi_seg_idxs", all_s, pads["seg"]),
                               ("candi_seq_user_idxs", all_u, pads["user"]),
                               ("candi_seq_item_idxs", all_i, pads["prod"])):
        data = pad_3d(data, pad_id, dim=1)
        data = pad_3d(data, pad_id, dim=2)
        out[name] = np.asarray(data, np.int64)
    out["candi_prod_idxs"] = np.asarray(pad([list(c) for c in candidates], -1), np.int64)
    out["query_word_idxs"] = np.asarray([corpus["query_words"][e[0]] for e in entries], np.int64)
    return out


def ranklist_lines(user_ids, user_idxs, query_idxs, product_ids, ranked_ids, ranked_scores, cutoff):
    """trainer.py:158-169 applied to already ranked (ids, scores) lists."""
    out = []
    for i in range(len(ranked_ids)):
        for rank in range(min(cutoff, len(ranked_ids[i]))):
            out.append("%s_%d Q0 %s %d %f ReviewTransformer\n" % (
                user_ids[user_idxs[i]], query_idxs[i], product_ids[ranked_ids[i][rank]], rank + 1,
                ranked_scores[i][rank]))
    return out

"""Golden vectors for batch construction and the run file, produced by RUNNING THE REFERENCE's own methods.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden_batches.py

A small seeded corpus (nested lists shaped like GlobalProdSearchData / ProdSearchData, data/data_util.py) is fed to
  * ItemPVDataloader.get_train_batch / get_test_batch (data/item_pv_dataloader.py:122-143,:31-49), called on an
    un-initialised loader object whose attributes are set by hand (no DataLoader workers, no files), with
    ``random.choice`` replaced by a queue of supplied picks, and
  * Trainer.test / Trainer.calc_metrics (trainer.py:140-186) on a Trainer whose get_prod_scores returns a supplied
    tie-free score matrix (the run-file text and MRR / P@1 are the reference's own output).
Writes tests/golden/batches.npz.  Nothing here is imported by the product.
"""
import argparse
import os
import random
import sys
import tempfile

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def make_corpus(seed=7, U=24, P=17, Q=9, V=40, wq=5):
    rng = np.random.default_rng(seed)
    n_rev = rng.integers(1, 15, size=U)                      # several users exceed the history limit of 6
    review_u_p, times = [], []
    for u in range(U):
        for _ in range(int(n_rev[u])):
            review_u_p.append([u, int(rng.integers(0, P))])
            times.append(int(rng.integers(0, 10 ** 6)))
    R = len(review_u_p)
    perm = rng.permutation(R)                                # review ids are not grouped by user
    review_u_p = [review_u_p[i] for i in perm]
    times = [times[i] for i in perm]
    u_r_seq = [[] for _ in range(U)]
    i_r_seq = [[] for _ in range(P)]
    for r in sorted(range(R), key=lambda r: (times[r], r)):
        u_r_seq[review_u_p[r][0]].append(r)
        i_r_seq[review_u_p[r][1]].append(r)
    loc = [[0, 0, times[r]] for r in range(R)]
    for seqs, col in ((u_r_seq, 0), (i_r_seq, 1)):
        for s in seqs:
            for j, r in enumerate(s):
                loc[r][col] = j
    in_train = rng.random(R) < 0.75
    u_reviews = [set() for _ in range(U)]
    for r in range(R):
        if in_train[r]:
            u_reviews[review_u_p[r][0]].add(r)
    qlen = rng.integers(1, wq + 1, size=Q)
    query_words = [[int(x) for x in rng.integers(0, V - 1, size=int(l))] + [V - 1] * (wq - int(l)) for l in qlen]
    product_query_idx = [[int(x) for x in rng.choice(Q, size=int(rng.integers(1, 4)), replace=False)]
                         for _ in range(P)]
    p_reviews = [set() for _ in range(P)]
    for r in range(R):
        if in_train[r]:
            p_reviews[review_u_p[r][1]].add(r)
    return dict(U=U, P=P, Q=Q, V=V, wq=wq, R=R, review_u_p=review_u_p, u_r_seq=u_r_seq, review_loc_time=loc,
                i_r_seq=i_r_seq, p_reviews=p_reviews,
                u_reviews=u_reviews, query_words=query_words, product_query_idx=product_query_idx,
                in_train=in_train.astype(np.uint8))


def loader(c, set_name, **flags):
    from data.item_pv_dataloader import ItemPVDataloader
    dl = ItemPVDataloader.__new__(ItemPVDataloader)
    a = dict(uprev_review_limit=6, do_seq_review_train=False, fix_train_review=True, do_seq_review_test=False,
             train_review_only=True)
    a.update(flags)
    dl.args = argparse.Namespace(**a)
    dl.global_data = argparse.Namespace(query_words=c["query_words"], review_u_p=c["review_u_p"],
                                        u_r_seq=c["u_r_seq"], review_loc_time=c["review_loc_time"])
    dl.prod_data = argparse.Namespace(u_reviews=c["u_reviews"], product_query_idx=c["product_query_idx"],
                                      set_name=set_name)
    dl.prod_pad_idx = c["P"]
    return dl


def review_loader(c, **flags):
    """ProdSearchDataLoader with hand-set attributes (no DataLoader machinery, no files)."""
    from data.prod_search_dataloader import ProdSearchDataLoader
    from data.prod_search_dataset import ProdSearchDataset
    dl = ProdSearchDataLoader.__new__(ProdSearchDataLoader)
    a = dict(uprev_review_limit=4, iprev_review_limit=5, do_seq_review_test=False, train_review_only=True)
    a.update(flags)
    dl.args = argparse.Namespace(**a)
    dl.global_data = argparse.Namespace(query_words=c["query_words"], review_u_p=c["review_u_p"],
                                        u_r_seq=c["u_r_seq"], i_r_seq=c["i_r_seq"],
                                        review_loc_time=c["review_loc_time"])
    dl.prod_data = argparse.Namespace(u_reviews=c["u_reviews"], p_reviews=c["p_reviews"], set_name="test")
    ds = ProdSearchDataset.__new__(ProdSearchDataset)
    object.__setattr__(dl, "_stub_dataset", ds)
    dl.prod_pad_idx, dl.user_pad_idx, dl.review_pad_idx, dl.seg_pad_idx = c["P"], c["U"], c["R"], 3
    dl.total_review_limit = a["uprev_review_limit"] + a["iprev_review_limit"]
    return dl, ds


def csr(lists):
    off = np.zeros(len(lists) + 1, np.int64)
    off[1:] = np.cumsum([len(l) for l in lists])
    return off, np.asarray([x for l in lists for x in l], np.int64)


def main():
    assert os.path.isdir(REF), "golden vectors can only be regenerated where /root/reference exists"
    sys.path.insert(0, REF)
    c = make_corpus()
    rng = np.random.default_rng(99)
    out = {}
    for name in ("U", "P", "Q", "V", "wq", "R"):
        out["corpus/" + name] = np.int64(c[name])
    out["corpus/review_u_p"] = np.asarray(c["review_u_p"], np.int64)
    out["corpus/review_loc_time"] = np.asarray(c["review_loc_time"], np.int64)
    out["corpus/in_train"] = c["in_train"]
    out["corpus/query_words"] = np.asarray(c["query_words"], np.int64)
    out["corpus/u_r_seq_off"], out["corpus/u_r_seq"] = csr(c["u_r_seq"])
    out["corpus/pq_off"], out["corpus/pq"] = csr(c["product_query_idx"])

    # ---- training batches: (word_idxs, review_idx) samples over training reviews
    B, W = 40, 2
    train_reviews = np.flatnonzero(c["in_train"])
    reviews = rng.choice(train_reviews, size=B, replace=True)
    words = rng.integers(0, c["V"] - 1, size=(B, W))
    picks = rng.integers(0, 2 ** 32, size=B, dtype=np.uint64).astype(np.uint32)
    out["train/review_idx"], out["train/word_idxs"], out["train/query_pick"] = reviews.astype(np.int64), \
        words.astype(np.int64), picks
    samples = [[[int(x) for x in words[b]], int(reviews[b])] for b in range(B)]
    for tag, flags in (("last", dict(fix_train_review=True)), ("seq", dict(do_seq_review_train=True))):
        queue = [int(p) for p in picks]
        orig = random.choice
        random.choice = lambda seq: seq[queue.pop(0) % len(seq)]
        try:
            b = loader(c, "train", **flags).get_train_batch(samples)
        finally:
            random.choice = orig
        assert not queue
        for k in ("query_word_idxs", "target_prod_idxs", "u_item_idxs", "pos_iword_idxs"):
            out["train_%s/%s" % (tag, k)] = getattr(b, k).numpy()

    # ---- test batches: (query_idx, user_idx, prod_idx, review_idx, candidates) over held-out reviews
    held = np.flatnonzero(c["in_train"] == 0)
    entries = []
    for r in held[:30]:
        u, p = c["review_u_p"][int(r)]
        for q in c["product_query_idx"][p][:1]:
            entries.append([int(q), int(u), int(p), int(r)])
    out["test/entries"] = np.asarray(entries, np.int64)
    for tag, flags in (("last", dict()), ("seq", dict(do_seq_review_test=True, train_review_only=False))):
        b = loader(c, "test", **flags).get_test_batch([e + [list(range(c["P"]))] for e in entries])
        for k in ("query_word_idxs", "target_prod_idxs", "u_item_idxs", "candi_prod_idxs"):
            out["test_%s/%s" % (tag, k)] = getattr(b, k).numpy()
        out["test_%s/user_idxs" % tag] = np.asarray(b.user_idxs, np.int64)
        out["test_%s/query_idxs" % tag] = np.asarray(b.query_idxs, np.int64)

    # ---- review-transformer test batches: ragged candidate lists per entry
    out["corpus/i_r_seq_off"], out["corpus/i_r_seq"] = csr(c["i_r_seq"])
    r_entries = entries[:12]
    cands = [[int(x) for x in rng.choice(c["P"], size=int(rng.integers(2, 8)), replace=False)] for _ in r_entries]
    out["rtest/entries"] = np.asarray(r_entries, np.int64)
    out["rtest/cand_off"], out["rtest/cand"] = csr(cands)
    import torch.utils.data.dataloader as tdl
    for tag, flags in (("last", dict()), ("seq", dict(do_seq_review_test=True, train_review_only=False))):
        dl, ds = review_loader(c, **flags)
        orig_prop = tdl.DataLoader.__dict__.get("dataset")
        type(dl).dataset = property(lambda self: self._stub_dataset)     # only bisect_right is used (:140)
        try:
            b = dl.get_test_batch([e + [cd] for e, cd in zip(r_entries, cands)])
        finally:
            del type(dl).dataset
        for k in ("query_word_idxs", "candi_prod_ridxs", "candi_seg_idxs", "candi_seq_user_idxs",
                  "candi_seq_item_idxs"):
            out["rtest_%s/%s" % (tag, k)] = getattr(b, k).numpy()
        out["rtest_%s/candi_prod_idxs" % tag] = np.asarray(b.candi_prod_idxs, np.int64)

    # ---- run file + metrics from Trainer.test on a supplied tie-free score matrix
    from trainer import Trainer
    M, N, cutoff = len(entries), 57, 10
    scores = rng.standard_normal((M, N)).astype(np.float32) * 3
    assert all(len(np.unique(row)) == N for row in scores)
    prod_idxs = np.tile(np.arange(N), (M, 1))
    target = rng.integers(0, N, size=M)
    user_idxs = np.asarray([e[1] for e in entries])
    query_idxs = np.asarray([e[0] for e in entries])
    user_ids = ["A%05dU" % (7 * i + 3) for i in range(c["U"])]
    product_ids = ["B00%04dX" % (13 * i + 1) for i in range(N)]
    tr = Trainer.__new__(Trainer)
    tr.ExpDataset = lambda *a, **k: None
    tr.ExpDataloader = lambda *a, **k: None
    tr.get_prod_scores = lambda *a, **k: (prod_idxs, scores, target, query_idxs, user_idxs)
    metrics = {}
    orig_calc = Trainer.calc_metrics

    def calc(self, *a, **k):
        metrics["mrr"], metrics["prec"] = orig_calc(self, *a, **k)
        return metrics["mrr"], metrics["prec"]
    tr.calc_metrics = calc.__get__(tr)
    with tempfile.TemporaryDirectory() as td:
        args = argparse.Namespace(test_candi_size=-1, valid_batch_size=24, num_workers=0, save_dir=td)
        gd = argparse.Namespace(product_size=N, user_ids=user_ids, product_ids=product_ids)
        tr.test(args, gd, None, rankfname="run.txt", cutoff=cutoff)
        text = open(os.path.join(td, "run.txt"), "rb").read()
    out["rank/scores"], out["rank/target"], out["rank/cutoff"] = scores, target.astype(np.int64), np.int64(cutoff)
    out["rank/user_idxs"], out["rank/query_idxs"] = user_idxs.astype(np.int64), query_idxs.astype(np.int64)
    out["rank/user_ids"], out["rank/product_ids"] = np.asarray(user_ids), np.asarray(product_ids)
    out["rank/text"] = np.frombuffer(text, dtype=np.uint8)
    out["rank/mrr"], out["rank/prec"] = np.float64(metrics["mrr"]), np.float64(metrics["prec"])
    np.savez_compressed(os.path.join(OUT, "batches.npz"), **out)
    print("batches ok: R=%d, %d train samples, %d test entries, run file %d bytes, MRR %.4f P@1 %.4f" %
          (c["R"], B, M, len(text), metrics["mrr"], metrics["prec"]))


if __name__ == "__main__":
    main()

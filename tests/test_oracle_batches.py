"""CPU: the batch-construction / run-file oracle against the golden vectors produced by the reference's own
ItemPVDataloader and Trainer (tests/golden/make_golden_batches.py), plus the HOST entry points of the library
(psb_write_ranklist, psb_subset_key) -- no GPU work."""
import os

import numpy as np

import oracle
from oracle import batches as ob

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "batches.npz")


def load_corpus(z):
    def uncsr(off, flat):
        return [[int(x) for x in flat[off[i]:off[i + 1]]] for i in range(len(off) - 1)]
    rup = z["corpus/review_u_p"]
    u_r_seq = uncsr(z["corpus/u_r_seq_off"], z["corpus/u_r_seq"])
    u_reviews = [set() for _ in u_r_seq]
    for r in np.flatnonzero(z["corpus/in_train"]):
        u_reviews[int(rup[r, 0])].add(int(r))
    i_r_seq = uncsr(z["corpus/i_r_seq_off"], z["corpus/i_r_seq"])
    p_reviews = [set() for _ in i_r_seq]
    for r in np.flatnonzero(z["corpus/in_train"]):
        p_reviews[int(rup[r, 1])].add(int(r))
    return dict(i_r_seq=i_r_seq, p_reviews=p_reviews,
                review_u_p=[[int(a), int(b)] for a, b in rup], u_r_seq=u_r_seq,
                review_loc_time=[[int(x) for x in row] for row in z["corpus/review_loc_time"]],
                u_reviews=u_reviews, query_words=[[int(x) for x in row] for row in z["corpus/query_words"]],
                product_query_idx=uncsr(z["corpus/pq_off"], z["corpus/pq"]))


def test_train_batches_match_reference():
    z = np.load(GOLDEN)
    c = load_corpus(z)
    samples = [(list(w), int(r)) for w, r in zip(z["train/word_idxs"], z["train/review_idx"])]
    for tag, do_seq in (("last", False), ("seq", True)):
        b = ob.item_train_batch(c, samples, z["train/query_pick"], 6, do_seq, True, int(z["corpus/P"]))
        for k in ("query_word_idxs", "target_prod_idxs", "u_item_idxs", "pos_iword_idxs"):
            assert np.array_equal(b[k], z["train_%s/%s" % (tag, k)]), (tag, k)
    assert z["train_last/u_item_idxs"].shape[1] == 6          # the limit binds for some user


def test_test_batches_match_reference():
    z = np.load(GOLDEN)
    c = load_corpus(z)
    entries = [tuple(int(x) for x in e) for e in z["test/entries"]]
    for tag, do_seq in (("last", False), ("seq", True)):
        b = ob.item_test_batch(c, entries, 6, do_seq, int(z["corpus/P"]))
        for k in ("query_word_idxs", "target_prod_idxs", "u_item_idxs", "user_idxs", "query_idxs"):
            assert np.array_equal(b[k], z["test_%s/%s" % (tag, k)]), (tag, k)


def rtest_inputs(z):
    entries = [tuple(int(x) for x in e) for e in z["rtest/entries"]]
    off, flat = z["rtest/cand_off"], z["rtest/cand"]
    cands = [[int(x) for x in flat[off[i]:off[i + 1]]] for i in range(len(entries))]
    pads = dict(review=int(z["corpus/R"]), user=int(z["corpus/U"]), prod=int(z["corpus/P"]), seg=3)
    return entries, cands, pads


def test_review_test_batches_match_reference():
    z = np.load(GOLDEN)
    c = load_corpus(z)
    entries, cands, pads = rtest_inputs(z)
    assert len(set(len(x) for x in cands)) > 1                   # ragged candidate lists: -1 / all-pad rows occur
    for tag, seq_test, tro in (("last", False, True), ("seq", True, False)):
        b = ob.review_test_batch(c, entries, cands, 4, 5, seq_test, tro, pads)
        for k in ("query_word_idxs", "candi_prod_ridxs", "candi_seg_idxs", "candi_seq_user_idxs",
                  "candi_seq_item_idxs", "candi_prod_idxs"):
            assert np.array_equal(b[k], z["rtest_%s/%s" % (tag, k)]), (tag, k)


def _ranked(z):
    ids, sc = oracle.topk_lower_id_first(z["rank/scores"], int(z["rank/cutoff"]))
    return ids, sc


def test_ranklist_text_and_metrics_match_reference(tmp_path):
    z = np.load(GOLDEN)
    ids, sc = _ranked(z)
    users, prods = [str(x) for x in z["rank/user_ids"]], [str(x) for x in z["rank/product_ids"]]
    text = "".join(ob.ranklist_lines(users, z["rank/user_idxs"], z["rank/query_idxs"], prods, ids, sc,
                                     int(z["rank/cutoff"])))
    assert text.encode() == z["rank/text"].tobytes()
    full = oracle.rank_lower_id_first(z["rank/scores"])
    mrr, prec = oracle.calc_metrics(full, z["rank/target"], cutoff=int(z["rank/cutoff"]))
    assert mrr == float(z["rank/mrr"]) and prec == float(z["rank/prec"])
    # product host entry point: the C writer produces the reference's bytes from the top-k lists
    from prodsearch_b200 import evaluate
    path = tmp_path / "run.txt"
    n = evaluate.write_ranklist(path, users, z["rank/user_idxs"], z["rank/query_idxs"], prods, ids, sc,
                                cutoff=int(z["rank/cutoff"]))
    assert n == ids.shape[0] * int(z["rank/cutoff"])
    assert path.read_bytes() == z["rank/text"].tobytes()
    # cutoff below k, append mode
    evaluate.write_ranklist(path, users, z["rank/user_idxs"][:2], z["rank/query_idxs"][:2], prods, ids[:2], sc[:2],
                            cutoff=3, append=True)
    assert path.read_bytes()[len(z["rank/text"]):] == "".join(
        ob.ranklist_lines(users, z["rank/user_idxs"][:2], z["rank/query_idxs"][:2], prods, ids[:2], sc[:2], 3)).encode()
    # ranks taken from the top-k lists give the same MRR / P@1 when k >= cutoff
    ranks = [(list(row).index(t) + 1 if t in row else 0) for row, t in zip(ids, z["rank/target"])]
    assert evaluate.calc_metrics(np.asarray(ranks), cutoff=int(z["rank/cutoff"])) == (mrr, prec)


def test_subset_key_host_matches_oracle():
    from prodsearch_b200 import corpus
    rng = np.random.default_rng(5)
    for seed, b, p in rng.integers(0, 2 ** 32, size=(200, 3), dtype=np.uint64):
        assert corpus.subset_key(seed, b, p) == ob.subset_key(int(seed), int(b), int(p))


def test_random_subset_properties():
    z = np.load(GOLDEN)
    c = load_corpus(z)
    for user, seq in enumerate(c["u_r_seq"]):
        cand = [x for x in seq if x in c["u_reviews"][user]]
        for limit in (1, 3, 6):
            a = ob.user_review_idxs(c["u_r_seq"], c["u_reviews"], c["review_loc_time"], user, -1, limit, False,
                                    fix=False, seed=11, sample=user)
            assert len(a) == min(limit, len(cand)) and set(a) <= set(cand)
            assert [x for x in cand if x in a] == a                      # sequence order kept
            assert a == ob.user_review_idxs(c["u_r_seq"], c["u_reviews"], c["review_loc_time"], user, -1, limit,
                                            False, fix=False, seed=11, sample=user)
    picks = [tuple(ob.user_review_idxs(c["u_r_seq"], c["u_reviews"], c["review_loc_time"], 0, -1, 2, False,
                                       fix=False, seed=s, sample=0)) for s in range(40)]
    if len([x for x in c["u_r_seq"][0] if x in c["u_reviews"][0]]) > 2:
        assert len(set(picks)) > 1                                        # the seed changes the subset

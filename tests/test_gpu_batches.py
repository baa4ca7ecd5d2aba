"""GPU: on-device batch construction (psb_build_item_batch), target ranks and the evaluation driver against the
golden vectors of the reference's own dataloader / trainer and against the oracle."""
import argparse
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import batches as ob
from test_oracle_batches import GOLDEN, load_corpus, rtest_inputs

pytestmark = pytest.mark.gpu


def make_corpus(c, z=None, **kw):
    from prodsearch_b200.corpus import ItemCorpus
    return ItemCorpus("cuda:0", c["u_r_seq"], c["review_u_p"], c["query_words"], c["product_query_idx"],
                      train_reviews=c["u_reviews"], review_uloc=c["review_loc_time"], item_seq=c.get("i_r_seq"), **kw)


def flags(**kw):
    a = dict(uprev_review_limit=6, do_seq_review_train=False, fix_train_review=True, do_seq_review_test=False,
             train_review_only=True)
    a.update(kw)
    return argparse.Namespace(**a)


def test_train_batches_match_reference_golden():
    z = np.load(GOLDEN)
    c = load_corpus(z)
    corpus = make_corpus(c, product_size=int(z["corpus/P"]), vocab_size=int(z["corpus/V"]))
    for tag, fl in (("last", flags()), ("seq", flags(do_seq_review_train=True))):
        b = corpus.train_batch(z["train/review_idx"], z["train/word_idxs"], fl, query_pick=z["train/query_pick"])
        for k in ("query_word_idxs", "target_prod_idxs", "u_item_idxs", "pos_iword_idxs"):
            got = getattr(b, k).cpu().numpy()
            assert got.dtype == np.int64 and np.array_equal(got, z["train_%s/%s" % (tag, k)]), (tag, k)
        # fixed-width form (graph-capturable): same ids, extra columns are pad items
        b2 = corpus.train_batch(z["train/review_idx"], z["train/word_idxs"], fl, query_pick=z["train/query_pick"],
                                trim=False)
        h = b2.u_item_idxs.cpu().numpy()
        w = z["train_%s/u_item_idxs" % tag].shape[1]
        assert h.shape[1] == 6 and np.array_equal(h[:, :w], z["train_%s/u_item_idxs" % tag])
        assert (h[:, w:] == int(z["corpus/P"])).all()


def test_test_batches_match_reference_golden():
    z = np.load(GOLDEN)
    c = load_corpus(z)
    corpus = make_corpus(c, product_size=int(z["corpus/P"]), vocab_size=int(z["corpus/V"]))
    e = z["test/entries"]
    for tag, fl in (("last", flags()), ("seq", flags(do_seq_review_test=True, train_review_only=False))):
        b = corpus.test_batch(e[:, 0], e[:, 1], e[:, 2], e[:, 3], fl)
        for k in ("query_word_idxs", "target_prod_idxs", "u_item_idxs", "user_idxs", "query_idxs"):
            assert np.array_equal(getattr(b, k).cpu().numpy(), z["test_%s/%s" % (tag, k)]), (tag, k)


def big_corpus(seed, U=300, P=500, Q=60, V=900, wq=7, max_len=150):
    rng = np.random.default_rng(seed)
    n_rev = np.minimum(max_len, rng.geometric(0.04, size=U))          # long sequences: several 32-slot chunks
    rup = []
    for u in range(U):
        rup += [[u, int(rng.integers(0, P))] for _ in range(int(n_rev[u]))]
    perm = rng.permutation(len(rup))
    rup = [rup[i] for i in perm]
    u_r_seq = [[] for _ in range(U)]
    for r in rng.permutation(len(rup)):
        u_r_seq[rup[r][0]].append(int(r))
    loc = [[0, 0, 0] for _ in rup]
    for s in u_r_seq:
        for j, r in enumerate(s):
            loc[r][0] = j
    in_train = rng.random(len(rup)) < 0.8
    u_reviews = [set() for _ in range(U)]
    for r in np.flatnonzero(in_train):
        u_reviews[rup[r][0]].add(int(r))
    qw = [[int(x) for x in rng.integers(0, V - 1, size=wq)] for _ in range(Q)]
    pq = [[int(x) for x in rng.choice(Q, size=int(rng.integers(1, 5)), replace=False)] for _ in range(P)]
    return dict(review_u_p=rup, u_r_seq=u_r_seq, review_loc_time=loc, u_reviews=u_reviews, query_words=qw,
                product_query_idx=pq), P, V


@pytest.mark.parametrize("mode", ["last", "seq", "random"])
@pytest.mark.parametrize("limit", [1, 20, 70])
def test_large_corpus_matches_oracle(mode, limit):
    c, P, V = big_corpus(3)
    corpus = make_corpus(c, product_size=P, vocab_size=V)
    rng = np.random.default_rng(17)
    B = 777
    reviews = rng.integers(0, len(c["review_u_p"]), size=B)
    words = rng.integers(0, V - 1, size=(B, 1))
    picks = rng.integers(0, 2 ** 32, size=B, dtype=np.uint64).astype(np.uint32)
    fl = flags(uprev_review_limit=limit, do_seq_review_train=mode == "seq", fix_train_review=mode != "random")
    b = corpus.train_batch(reviews, words, fl, query_pick=picks, seed=12345)
    ref = ob.item_train_batch(c, [(list(w), int(r)) for w, r in zip(words, reviews)], picks, limit, mode == "seq",
                              mode != "random", P, seed=12345)
    for k in ("query_word_idxs", "target_prod_idxs", "u_item_idxs", "user_idxs", "query_idxs", "hist_len"):
        assert np.array_equal(getattr(b, k).cpu().numpy(), ref[k]), k
    # run-to-run identical
    b2 = corpus.train_batch(reviews, words, fl, query_pick=picks, seed=12345)
    assert torch.equal(b.u_item_idxs, b2.u_item_idxs)


def test_review_test_batches_match_reference_golden():
    """psb_build_review_test_batch against ProdSearchDataLoader.get_test_batch's own output (N3)."""
    z = np.load(GOLDEN)
    c = load_corpus(z)
    corpus = make_corpus(c, product_size=int(z["corpus/P"]), vocab_size=int(z["corpus/V"]))
    entries, cands, pads = rtest_inputs(z)
    e = np.asarray(entries)
    for tag, seq_test, tro in (("last", False, True), ("seq", True, False)):
        fl = argparse.Namespace(uprev_review_limit=4, iprev_review_limit=5, do_seq_review_test=seq_test,
                                train_review_only=tro)
        cand = z["rtest_%s/candi_prod_idxs" % tag]
        b = corpus.review_test_batch(e[:, 0], e[:, 1], e[:, 2], e[:, 3], cand, fl)
        for k in ("query_word_idxs", "candi_prod_ridxs", "candi_seg_idxs", "candi_seq_user_idxs",
                  "candi_seq_item_idxs", "candi_prod_idxs"):
            assert np.array_equal(getattr(b, k).cpu().numpy(), z["rtest_%s/%s" % (tag, k)]), (tag, k)


@pytest.mark.parametrize("seq", [False, True])
def test_review_test_batches_large_match_oracle(seq):
    c, P, V = big_corpus(21, U=200, P=120, Q=30, max_len=90)
    rng = np.random.default_rng(8)
    R = len(c["review_u_p"])
    times = rng.integers(0, 5000, size=R)                      # duplicate time stamps exercise bisect_right
    c["i_r_seq"] = [[] for _ in range(P)]
    for r in sorted(range(R), key=lambda r: (times[r], r)):
        c["i_r_seq"][c["review_u_p"][r][1]].append(r)
    for r in range(R):
        c["review_loc_time"][r][2] = int(times[r])
    c["p_reviews"] = [set() for _ in range(P)]
    for u_set in c["u_reviews"]:
        for r in u_set:
            c["p_reviews"][c["review_u_p"][r][1]].add(r)
    corpus = make_corpus(c, product_size=P, vocab_size=V)
    entries = []
    for r in rng.choice(R, size=40, replace=False):
        u, p = c["review_u_p"][int(r)]
        entries.append((int(c["product_query_idx"][p][0]), u, p, int(r)))
    cands = [[int(x) for x in rng.choice(P, size=int(rng.integers(1, 60)), replace=False)] for _ in entries]
    pads = dict(review=R, user=200, prod=P, seg=3)
    fl = argparse.Namespace(uprev_review_limit=20, iprev_review_limit=30, do_seq_review_test=seq,
                            train_review_only=not seq)
    ref = ob.review_test_batch(c, entries, cands, 20, 30, seq, not seq, pads)
    e = np.asarray(entries)
    b = corpus.review_test_batch(e[:, 0], e[:, 1], e[:, 2], e[:, 3], ref["candi_prod_idxs"], fl)
    for k in ("query_word_idxs", "candi_prod_ridxs", "candi_seg_idxs", "candi_seq_user_idxs",
              "candi_seq_item_idxs"):
        assert np.array_equal(getattr(b, k).cpu().numpy(), ref[k]), k


def test_bad_ids_raise_and_empty_batch():
    c, P, V = big_corpus(4, U=20, P=30, Q=5)
    corpus = make_corpus(c, product_size=P, vocab_size=V)
    with pytest.raises(IndexError):
        corpus.train_batch([len(c["review_u_p"]) + 3], [[0]], flags(), query_pick=[0])
    b = corpus.train_batch(np.zeros(0, np.int64), np.zeros((0, 1), np.int64), flags(), query_pick=np.zeros(0))
    assert b.target_prod_idxs.numel() == 0


def test_target_ranks_and_metrics():
    from prodsearch_b200 import evaluate
    z = np.load(GOLDEN)
    ids, sc = oracle.topk_lower_id_first(z["rank/scores"], int(z["rank/cutoff"]))
    ranks = evaluate.target_ranks(torch.from_numpy(ids).cuda(), torch.from_numpy(z["rank/target"]).cuda())
    want = [(list(row).index(t) + 1 if t in row else 0) for row, t in zip(ids, z["rank/target"])]
    assert ranks.cpu().tolist() == want
    mrr, prec = evaluate.calc_metrics(ranks, cutoff=int(z["rank/cutoff"]))
    assert mrr == float(z["rank/mrr"]) and prec == float(z["rank/prec"])
    # wide lists (k > 32) and absent targets
    g = torch.Generator().manual_seed(1)
    wide = torch.stack([torch.randperm(500, generator=g)[:100] for _ in range(64)]).cuda()
    tgt = torch.randint(0, 500, (64,), generator=g).cuda()
    r = evaluate.target_ranks(wide, tgt).cpu().tolist()
    for row, t, got in zip(wide.cpu().tolist(), tgt.cpu().tolist(), r):
        assert got == (row.index(t) + 1 if t in row else 0)


def test_eval_driver_end_to_end(tmp_path):
    """corpus -> device batches -> fused full-catalog top-k -> ranks + run file, against the oracle's full
    [M, N] scoring + stable sort + the reference's run-file expression."""
    from prodsearch_b200 import evaluate
    from prodsearch_b200.item_transformer import ItemTransformerRanker
    c, P, V = big_corpus(9, U=120, P=900, Q=40, V=400, wq=6, max_len=40)
    cfg = argparse.Namespace(
        train_review_only=True, embedding_size=128, dropout=0.0, pretrain_emb_dir="", pretrain_up_emb_dir="",
        sep_prod_emb=False, model_name="item_transformer", ff_size=256, heads=8, inter_layers=1,
        query_encoder_name="fs", use_dot_prod=True, use_pos_emb=True, use_item_pos=False, sim_func="product",
        pos_weight=False, neg_per_pos=5, uprev_review_limit=20, do_seq_review_test=False, do_seq_review_train=False,
        fix_train_review=True)
    torch.manual_seed(3)
    model = ItemTransformerRanker(cfg, "cuda:0", V, P, None, word_dists=np.full(V, 1.0 / V, np.float32))
    corpus = make_corpus(c, product_size=P, vocab_size=V)
    rng = np.random.default_rng(2)
    entries = []
    for r in rng.choice(len(c["review_u_p"]), size=50, replace=False):
        u, p = c["review_u_p"][int(r)]
        entries.append((c["product_query_idx"][p][0], u, p, int(r)))
    users = ["U%04d" % i for i in range(120)]
    prods = ["P%05d" % i for i in range(P)]
    path = tmp_path / "test.ranklist"
    mrr, prec = evaluate.test(model, corpus, entries, cfg, users, prods, rank_path=path, cutoff=100, batch_size=16)
    # oracle: full score matrix on the CPU
    params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref_b = ob.item_test_batch(c, entries, 20, False, P)
    with torch.no_grad():
        _, full = oracle.tem_catalog_scores(params, cfg, torch.from_numpy(ref_b["query_word_idxs"]),
                                            torch.from_numpy(ref_b["u_item_idxs"]))
    full = full.numpy()[:, :P]
    ids, sc = oracle.topk_lower_id_first(full, 100)
    got_ids, got_sc, *_ = evaluate.rank_test_set(model, corpus, entries, cfg, k=100, batch_size=16)
    gap_ok = np.abs(np.diff(sc, axis=1)).min(axis=1) > 1e-4                # rows without near-ties: ids exact
    assert gap_ok.sum() > 25
    assert np.array_equal(got_ids.cpu().numpy()[gap_ok], ids[gap_ok])
    assert np.allclose(got_sc.cpu().numpy(), sc, rtol=1e-5, atol=1e-4)
    o_mrr, o_prec = oracle.calc_metrics(oracle.rank_lower_id_first(full), ref_b["target_prod_idxs"], cutoff=100)
    assert abs(mrr - o_mrr) < 1e-9 + (~gap_ok).sum() and prec <= 1.0
    lines = path.read_text().splitlines()
    assert len(lines) == 50 * 100
    want = ob.ranklist_lines(users, ref_b["user_idxs"], ref_b["query_idxs"], prods, ids, sc, 100)
    first = [i for i in range(50) if gap_ok[i]][0]
    for got, ref in zip(lines[first * 100:first * 100 + 100], want[first * 100:first * 100 + 100]):
        assert got.split()[:4] == ref.split()[:4]                       # qid, Q0, product, rank
        assert abs(float(got.split()[4]) - float(ref.split()[4])) <= 2e-4

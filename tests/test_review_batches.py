"""CPU: the review-transformer TRAINING collate (prodsearch_b200/review_batches.py) against the batches the reference's
own ProdSearchData.initialize_epoch + ProdSearchDataLoader.get_train_batch produced from the same files on the same
random streams (tests/golden/review_batches.npz from tests/golden/make_golden_review_batches.py): every tensor of
every batch object, bit for bit."""
import argparse
import os
import random
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_review_batches import BATCH, CASES, FIELDS  # noqa: E402  (the case table only; no reference import)

FILES = os.path.join(HERE, "golden", "files.npz")
GOLDEN = os.path.join(HERE, "golden", "review_batches.npz")


@pytest.fixture(scope="module")
def corpus(tmp_path_factory):
    from prodsearch_b200 import data_files
    z = np.load(FILES)
    root = tmp_path_factory.mktemp("corpus")
    data, inp = root / "data", root / "data" / "split"
    inp.mkdir(parents=True)
    for k in z.files:
        if k.startswith("file/"):
            _, tag, name = k.split("/")
            (data if tag == "data" else inp).joinpath(name).write_bytes(z[k].tobytes())
    files = data_files.CorpusFiles(str(data), str(inp))
    return files, files.split("train", subsampling_rate=1e-2)


def _collate(corpus, case):
    from prodsearch_b200.review_batches import ReviewTrainCollate
    files, split = corpus
    cfg = CASES[case]
    args = argparse.Namespace(**{k: v for k, v in cfg.items() if k not in ("prepare_pv", "shuffle")})
    random.seed(666)
    np.random.seed(666)
    c = ReviewTrainCollate(files, split, args)
    c.initialize_epoch()
    return c, cfg, split


@pytest.mark.parametrize("case", sorted(CASES))
def test_train_batches_match_reference(corpus, case):
    z = np.load(GOLDEN)
    c, cfg, split = _collate(corpus, case)
    assert np.array_equal(c.neg_sample_products, z[case + "/neg_sample_products"])
    assert np.array_equal(c.review_words, z[case + "/review_words"])
    rows = split.review_info
    n_checked = 0
    for bi in range(int(z[case + "/n_batches"])):
        res = c.train_batch(rows[bi * BATCH:(bi + 1) * BATCH], prepare_pv=cfg["prepare_pv"], shuffle=cfg["shuffle"])
        res = res if isinstance(res, list) else [res]
        assert len(res) == int(z["%s/b%d/count" % (case, bi)])
        for j, b in enumerate(res):
            for f in FIELDS:
                key = "%s/b%d/%d/%s" % (case, bi, j, f)
                got = getattr(b, f)
                if key not in z.files:
                    assert got is None, key
                    continue
                want = z[key]
                assert torch.is_tensor(got) and got.numpy().dtype == want.dtype, (key, got.dtype, want.dtype)
                assert np.array_equal(got.numpy(), want), key
                n_checked += 1
    assert n_checked >= 11 * int(z[case + "/n_batches"])


def test_cases_exercise_the_interesting_paths(corpus):
    """The golden cases are only worth something if the limits bind, samples get dropped / negatives go missing,
    and the pv windows need padding -- checked on the product's own bookkeeping."""
    c, cfg, split = _collate(corpus, "pvc")
    calls = {"n": 0}
    real = c.py_random.sample

    class Counting(object):
        choice = staticmethod(random.choice)

        @staticmethod
        def sample(pop, k):
            calls["n"] += 1
            return real(pop, k)
    c.py_random = Counting
    res = c.train_batch(split.review_info[:BATCH], prepare_pv=True, shuffle=False)
    assert calls["n"] > 0                                         # a history longer than its limit was sub-sampled
    assert len(res) == 3 and res[0].pos_prod_rword_idxs.shape[-1] == 2        # 6 words -> 3 windows of 2
    assert res[0].pos_prod_rword_idxs_pvc.shape[-1] == 6
    c2, _, _ = _collate(corpus, "pv")
    r2 = c2.train_batch(split.review_info[:BATCH], prepare_pv=True, shuffle=False)
    assert len(r2) == 3 and r2[0].pos_prod_rword_idxs.shape[-1] == 3          # 7 words -> padded to 9 -> 3 windows
    # without initialize_epoch there are no negatives to build sequences from
    from prodsearch_b200.review_batches import ReviewTrainCollate
    fresh = ReviewTrainCollate(corpus[0], corpus[1], c.args)
    with pytest.raises(RuntimeError, match="initialize_epoch"):
        fresh.train_batch(split.review_info[:4])


def test_batch_to_device_contract(corpus):
    c, cfg, split = _collate(corpus, "fs_seq")
    b = c.train_batch(split.review_info[:BATCH], prepare_pv=False)
    assert b.to("cpu") is b                                                  # batch_data.py:181-183
    assert b.pos_prod_rword_masks.dtype == torch.uint8 and b.neg_prod_rword_masks.dtype == torch.uint8
    assert b.pos_seg_idxs.shape[1] == b.pos_prod_ridxs.shape[1] + 1          # the query slot leads the sequence
    assert b.neg_prod_rword_idxs.shape[:3] == b.neg_prod_ridxs.shape


@pytest.mark.parametrize("enc,train_pv", [("fs", False), ("avg", False), ("pv", True), ("pvc", True)])
def test_collate_output_feeds_the_review_transformer_forward(corpus, enc, train_pv):
    """End to end on the host: files -> collate -> the oracle's restatement of ProductRanker.forward
    (ps_model.py:241-358) with the parameter shapes of the product's module: every field has the layout the forward
    pass indexes with (the product's forward runs on the GPU and is checked there against the same oracle)."""
    from oracle import ref_models as rm
    from prodsearch_b200.ps_model import ProductRanker
    from prodsearch_b200.review_batches import ReviewTrainCollate
    files, split = corpus
    args = argparse.Namespace(
        review_encoder_name=enc, do_subsample_mask=True, shuffle_review_words=True, do_seq_review_train=False,
        pv_window_size=2, review_word_limit=6, uprev_review_limit=2, iprev_review_limit=3, neg_per_pos=3,
        train_review_only=True, embedding_size=16, dropout=0.0, fix_emb=False, use_user_emb=True, use_item_emb=True,
        use_seg_emb=True, ff_size=32, heads=2, inter_layers=1, corrupt_rate=0.5, query_encoder_name="fs",
        use_pos_emb=True, model_name="review_transformer", sim_func="product", pos_weight=False)
    random.seed(1)
    np.random.seed(1)
    torch.manual_seed(0)
    c = ReviewTrainCollate(files, split, args)
    c.initialize_epoch()
    out = c.train_batch(split.review_info[:BATCH], prepare_pv=train_pv, shuffle=True)
    out = out if isinstance(out, list) else [out]
    assert len(out) == (3 if train_pv else 1)
    model = ProductRanker(args, "cpu", files.vocab_size, files.review_count, files.product_size, files.user_size,
                          c.review_words.tolist(), files.words, word_dists=split.word_dists)   # shapes only: no compute
    P = {k: v.detach().clone() for k, v in model.state_dict().items()}
    cfg = argparse.Namespace(**vars(args))
    cfg.review_pad_idx = files.review_count - 1
    for b in out:
        negw = torch.randint(0, files.vocab_size - 1, (b.pos_prod_rword_idxs.numel() * args.neg_per_pos,)) \
            if train_pv else None
        masks = None
        if enc == "pvc":
            masks = [torch.zeros(t.reshape(-1, t.shape[-1]).shape, dtype=torch.bool)
                     for t in (b.pos_prod_rword_idxs_pvc, b.neg_prod_rword_idxs_pvc)]
        loss = rm.rtm_forward(P, cfg, b, train_pv, neg_word_idxs=negw, corrupt_masks=masks, training=True)[0]
        assert torch.isfinite(loss) and float(loss) > 0

"""Golden vectors for the evaluation-sample enumeration with candidate lists, produced by the reference's own
ItemPVDataset.collect_test_samples (data/item_pv_dataset.py:36-68).

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden_samples.py

Cases on the synthetic corpus of tests/golden/files.npz: (a) the "valid" split with sampled candidates
(valid_candi_size = 5, segments of 3), (b) the test split re-ranking the candidate lists of a run file
(test_candi_size > 0, segments of 4).  The global python / numpy generators are seeded like main.py:172-173.
Stores tests/golden/samples.npz (entries, padded candidates, the run file's bytes)."""
import argparse
import os
import random
import sys
import tempfile

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def pack(data):
    entries = np.asarray([e[:4] for e in data], np.int64).reshape(-1, 4)
    width = max(len(e[4]) for e in data)
    cand = np.full((len(data), width), -1, np.int64)
    for i, e in enumerate(data):
        cand[i, :len(e[4])] = e[4]
    return entries, cand


def main():
    assert os.path.isdir(REF), "golden vectors can only be regenerated where /root/reference exists"
    sys.path.insert(0, REF)
    from data.data_util import GlobalProdSearchData, ProdSearchData
    from data.item_pv_dataset import ItemPVDataset
    z = np.load(os.path.join(OUT, "files.npz"))
    out = {}
    with tempfile.TemporaryDirectory() as root:
        data, inp = os.path.join(root, "data"), os.path.join(root, "data", "split")
        os.makedirs(inp)
        for k in z.files:
            if k.startswith("file/"):
                _, tag, name = k.split("/")
                open(os.path.join(data if tag == "data" else inp, name), "wb").write(z[k].tobytes())
        base = dict(model_name="item_transformer", do_subsample_mask=False, neg_per_pos=5, subsampling_rate=1e-3,
                    fix_emb=False, has_valid=False, prod_freq_neg_sample=False, pv_window_size=1,
                    train_review_only=True, uprev_review_limit=20)
        # (a) validation with sampled candidates
        args = argparse.Namespace(test_candi_size=-1, valid_candi_size=5, candi_batch_size=3, **base)
        g = GlobalProdSearchData(args, data, inp)
        va = ProdSearchData(args, inp, "valid", g)
        random.seed(666)
        np.random.seed(666)
        out["valid/entries"], out["valid/candidates"] = pack(ItemPVDataset(args, g, va)._data)
        # (b) candidates from a run file: every (user, query) of the test split gets 6 distinct items
        te0 = ProdSearchData(args, inp, "test", g)
        rng = np.random.default_rng(3)
        lines, seen = [], set()
        for _, u, p, _ in te0.review_info:
            for q in te0.product_query_idx[p]:
                if (u, q) in seen:
                    continue
                seen.add((u, q))
                for rank, item in enumerate(rng.choice(g.product_size, size=6, replace=False)):
                    lines.append("%s_%d Q0 %s %d %f ReviewTransformer" % (g.user_ids[u], q, g.product_ids[item], rank + 1,
                                                                          1.0 / (rank + 1)))
        run = os.path.join(inp, "test.bias_product.ranklist")
        open(run, "w").write("\n".join(lines) + "\n")
        out["run/bytes"] = np.frombuffer(open(run, "rb").read(), dtype=np.uint8)
        args_b = argparse.Namespace(test_candi_size=6, valid_candi_size=-1, candi_batch_size=4, **base)
        te = ProdSearchData(args_b, inp, "test", g)
        assert te.uq_pids is not None
        random.seed(666)
        np.random.seed(666)
        out["run/entries"], out["run/candidates"] = pack(ItemPVDataset(args_b, g, te)._data)
    np.savez_compressed(os.path.join(OUT, "samples.npz"), **out)
    print("valid: %d entries, run file: %d entries" % (len(out["valid/entries"]), len(out["run/entries"])))


if __name__ == "__main__":
    main()

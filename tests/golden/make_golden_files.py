"""Golden vectors for the dataset-file readers, produced by the reference's own loaders.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden_files.py

A tiny synthetic corpus is written in the reference's gzip text formats (data/data_util.py:165-287), loaded with
the reference's GlobalProdSearchData / ProdSearchData / ItemPVDataset.collect_test_samples, and both the file bytes
and what the reference made of them are stored in tests/golden/files.npz.  Nothing here is imported by the product."""
import argparse
import gzip
import os
import sys
import tempfile

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def write_gz(path, lines):
    with gzip.open(path, "wt") as f:
        for l in lines:
            f.write(l + "\n")


def make_files(root, seed=5, U=13, P=9, V=30, Q=7, R=60):
    rng = np.random.default_rng(seed)
    data, inp = os.path.join(root, "data"), os.path.join(root, "data", "split")
    os.makedirs(inp)
    users = rng.integers(0, U, size=R)
    prods = rng.integers(0, P, size=R)
    times = rng.integers(0, 10 ** 5, size=R)
    texts = [[int(x) for x in rng.integers(0, V, size=int(rng.integers(1, 9)))] for _ in range(R)]
    write_gz(os.path.join(data, "product.txt.gz"), ["B%05d" % (11 * i) for i in range(P)])
    write_gz(os.path.join(data, "users.txt.gz"), ["A%04dZ" % (7 * i) for i in range(U)])
    write_gz(os.path.join(data, "vocab.txt.gz"), ["w%d" % i for i in range(V)])
    write_gz(os.path.join(inp, "query.txt.gz"),
             [" ".join(str(int(x)) for x in rng.integers(0, V, size=int(rng.integers(1, 5)))) for _ in range(Q)])
    write_gz(os.path.join(data, "review_text.txt.gz"), [" ".join(map(str, t)) for t in texts])
    u_seq = [[] for _ in range(U)]
    p_seq = [[] for _ in range(P)]
    order = sorted(range(R), key=lambda r: (times[r], r))
    loc = [[0, 0, int(times[r])] for r in range(R)]
    for r in order:
        loc[r][0] = len(u_seq[users[r]])
        loc[r][1] = len(p_seq[prods[r]])
        u_seq[users[r]].append(r)
        p_seq[prods[r]].append(r)
    write_gz(os.path.join(data, "u_r_seq.txt.gz"), [" ".join(map(str, s)) for s in u_seq])
    write_gz(os.path.join(data, "p_r_seq.txt.gz"), [" ".join(map(str, s)) for s in p_seq])
    write_gz(os.path.join(data, "review_uloc_ploc_and_time.txt.gz"), [" ".join(map(str, l)) for l in loc])
    orig = rng.permutation(1000)[:R]                                   # original line ids, arbitrary
    write_gz(os.path.join(data, "review_id.txt.gz"), ["line_%d" % int(o) for o in orig])
    write_gz(os.path.join(data, "review_u_p.txt.gz"), ["%d %d" % (users[r], prods[r]) for r in range(R)])
    pq_train = [[int(x) for x in rng.choice(Q, size=int(rng.integers(1, 3)), replace=False)] for _ in range(P)]
    pq_test = [[int(x) for x in rng.choice(Q, size=int(rng.integers(1, 4)), replace=False)] for _ in range(P)]
    write_gz(os.path.join(inp, "train_query_idx.txt.gz"), [" ".join(map(str, q)) for q in pq_train])
    write_gz(os.path.join(inp, "test_query_idx.txt.gz"), [" ".join(map(str, q)) for q in pq_test])
    is_train = rng.random(R) < 0.7
    tr = [r for r in rng.permutation(R) if is_train[r]]
    te = [r for r in rng.permutation(R) if not is_train[r]]
    write_gz(os.path.join(inp, "train.txt.gz"),
             ["%d\t%d\t%s" % (users[r], prods[r], " ".join(map(str, texts[r]))) for r in tr])
    write_gz(os.path.join(inp, "train_id.txt.gz"),
             ["%d\t%d\tline_%d\t%d" % (users[r], prods[r], int(orig[r]), pq_train[prods[r]][0]) for r in tr])
    write_gz(os.path.join(inp, "test_id.txt.gz"),
             ["%d\t%d\tline_%d\t%d" % (users[r], prods[r], int(orig[r]), pq_test[prods[r]][0]) for r in te])
    return data, inp


def csr(lists):
    off = np.zeros(len(lists) + 1, np.int64)
    off[1:] = np.cumsum([len(l) for l in lists])
    return off, np.asarray([x for l in lists for x in l], np.int64)


def main():
    assert os.path.isdir(REF), "golden vectors can only be regenerated where /root/reference exists"
    sys.path.insert(0, REF)
    from data.data_util import GlobalProdSearchData, ProdSearchData
    from data.item_pv_dataset import ItemPVDataset
    out = {}
    with tempfile.TemporaryDirectory() as root:
        data, inp = make_files(root)
        for d, tag in ((data, "data"), (inp, "split")):
            for f in sorted(os.listdir(d)):
                p = os.path.join(d, f)
                if os.path.isfile(p):
                    out["file/%s/%s" % (tag, f)] = np.frombuffer(open(p, "rb").read(), dtype=np.uint8)
        args = argparse.Namespace(model_name="item_transformer", do_subsample_mask=False, neg_per_pos=5,
                                  subsampling_rate=1e-3, fix_emb=False, has_valid=False, test_candi_size=-1,
                                  prod_freq_neg_sample=False, valid_candi_size=-1, pv_window_size=1,
                                  train_review_only=True, uprev_review_limit=20, candi_batch_size=1000)
        g = GlobalProdSearchData(args, data, inp)
        tr = ProdSearchData(args, inp, "train", g)
        te = ProdSearchData(args, inp, "test", g)
        args2 = argparse.Namespace(**dict(vars(args), prod_freq_neg_sample=True))
        tr_freq = ProdSearchData(args2, inp, "train", g)
        ds = ItemPVDataset(args, g, te)
        # training samples: python / numpy global RNGs seeded like main.py:172-173 does; collect_train_samples
        # shuffles every review's word list IN PLACE, so it runs on a deep copy of the loaded corpus
        import copy
        import random
        g2 = copy.deepcopy(g)
        train_samples = {}
        for W in (1, 3):
            args_w = argparse.Namespace(**dict(vars(args), pv_window_size=W))
            g3 = copy.deepcopy(g2)
            random.seed(666)
            np.random.seed(666)
            train_samples[W] = ItemPVDataset(args_w, g3, tr)._data
    out["g/product_ids"], out["g/user_ids"] = np.asarray(g.product_ids), np.asarray(g.user_ids)
    out["g/vocab_size"], out["g/review_count"] = np.int64(g.vocab_size), np.int64(g.review_count)
    out["g/query_words"] = np.asarray(g.query_words, np.int64)
    out["g/review_length"] = np.asarray(g.review_length, np.int64)
    out["g/review_words_off"], out["g/review_words"] = csr(g.review_words)
    out["g/u_r_seq_off"], out["g/u_r_seq"] = csr(g.u_r_seq)
    out["g/i_r_seq_off"], out["g/i_r_seq"] = csr(g.i_r_seq)
    out["g/review_loc_time"] = np.asarray(g.review_loc_time, np.int64)
    out["g/review_u_p"] = np.asarray(g.review_u_p, np.int64)
    out["g/train_review_info"] = np.asarray(g.train_review_info, np.int64)
    out["g/train_query_idxs"] = np.asarray(g.train_query_idxs, np.int64)
    for tag, pd in (("train", tr), ("test", te)):
        out["%s/pq_off" % tag], out["%s/pq" % tag] = csr(pd.product_query_idx)
        out["%s/review_info" % tag] = np.asarray(pd.review_info, np.int64)
        out["%s/product_dists" % tag] = np.asarray(pd.product_dists, np.float64)
        in_u = np.zeros(len(g.review_u_p), np.uint8)
        for s in pd.u_reviews:
            for r in s:
                in_u[r] = 1
        in_p = np.zeros(len(g.review_u_p), np.uint8)
        for s in pd.p_reviews:
            for r in s:
                in_p[r] = 1
        out["%s/in_u_reviews" % tag], out["%s/in_p_reviews" % tag] = in_u, in_p
    out["train/vocab_distribute"] = np.asarray(tr.vocab_distribute, np.float64)
    out["train/sub_sampling_rate"] = np.asarray(tr.sub_sampling_rate, np.float64)
    out["train/word_dists"] = np.asarray(tr.word_dists, np.float64)
    out["train_freq/product_dists"] = np.asarray(tr_freq.product_dists, np.float64)
    entries = [e[:4] for e in ds._data]
    assert all(e[4] == list(range(g.product_size)) for e in ds._data)
    out["test/entries"] = np.asarray(entries, np.int64)
    for W, data_ in train_samples.items():
        out["train_samples_w%d/words" % W] = np.asarray([d[0] for d in data_], np.int64).reshape(len(data_), W)
        out["train_samples_w%d/review" % W] = np.asarray([d[1] for d in data_], np.int64)
    np.savez_compressed(os.path.join(OUT, "files.npz"), **out)
    print("files ok: %d files, %d reviews, %d train / %d test lines, %d test entries" % (
        sum(1 for k in out if k.startswith("file/")), len(g.review_u_p), len(tr.review_info), len(te.review_info),
        len(entries)))


if __name__ == "__main__":
    main()

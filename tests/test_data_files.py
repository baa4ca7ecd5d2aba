"""CPU: the dataset-file readers (prodsearch_b200/data_files.py) against what the reference's own loaders made of
the same files (tests/golden/files.npz from tests/golden/make_golden_files.py): every array GlobalProdSearchData /
ProdSearchData expose, the derived distributions, and the test-entry enumeration of ItemPVDataset."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "files.npz")


@pytest.fixture(scope="module")
def loaded(tmp_path_factory):
    from prodsearch_b200 import data_files
    z = np.load(GOLDEN)
    root = tmp_path_factory.mktemp("corpus")
    data, inp = root / "data", root / "data" / "split"
    inp.mkdir(parents=True)
    for k in z.files:
        if k.startswith("file/"):
            _, tag, name = k.split("/")
            (data if tag == "data" else inp).joinpath(name).write_bytes(z[k].tobytes())
    files = data_files.CorpusFiles(str(data), str(inp))
    return z, files, data_files


def test_global_arrays_match_reference(loaded):
    z, f, _ = loaded
    assert f.product_ids == [str(x) for x in z["g/product_ids"]] and f.user_ids == [str(x) for x in z["g/user_ids"]]
    assert f.vocab_size == int(z["g/vocab_size"]) and f.review_count == int(z["g/review_count"])
    assert np.array_equal(f.query_words, z["g/query_words"])
    assert np.array_equal(f.review_length, z["g/review_length"])
    for mine, ref in (((f.review_word_off, f.review_word), ("g/review_words_off", "g/review_words")),
                      ((f.user_seq_off, f.user_seq), ("g/u_r_seq_off", "g/u_r_seq")),
                      ((f.item_seq_off, f.item_seq), ("g/i_r_seq_off", "g/i_r_seq"))):
        assert np.array_equal(mine[0], z[ref[0]]) and np.array_equal(mine[1], z[ref[1]]), ref
    assert np.array_equal(f.review_loc_time, z["g/review_loc_time"])
    assert np.array_equal(f.review_u_p, z["g/review_u_p"])
    assert np.array_equal(f.train_review_info, z["g/train_review_info"])
    assert f.train_query_idxs == z["g/train_query_idxs"].tolist()


def test_splits_match_reference(loaded):
    z, f, df = loaded
    tr = f.split("train", subsampling_rate=1e-3)
    te = f.split("test", subsampling_rate=1e-3)
    for tag, s in (("train", tr), ("test", te)):
        assert np.array_equal(s.item_query_off, z[tag + "/pq_off"]) and np.array_equal(s.item_query, z[tag + "/pq"])
        assert np.array_equal(s.review_info, z[tag + "/review_info"])
        assert np.array_equal(s.product_dists, z[tag + "/product_dists"])               # bit-exact float64
        # one byte per review replaces the per-user AND the per-item sets of training reviews
        assert np.array_equal(f.review_in_train, z[tag + "/in_u_reviews"])
        assert np.array_equal(f.review_in_train, z[tag + "/in_p_reviews"])
    assert np.array_equal(tr.vocab_distribute, z["train/vocab_distribute"])
    assert np.array_equal(tr.sub_sampling_rate, z["train/sub_sampling_rate"])           # bit-exact float64
    assert np.array_equal(tr.word_dists, z["train/word_dists"])
    assert te.word_dists is None and te.sub_sampling_rate is None
    freq = f.split("train", subsampling_rate=1e-3, prod_freq_neg_sample=True)
    assert np.array_equal(freq.product_dists, z["train_freq/product_dists"])
    assert np.array_equal(f.split("train", subsampling_rate=0.0).sub_sampling_rate, np.ones(f.vocab_size))


def test_test_entries_match_reference(loaded):
    z, f, _ = loaded
    e = f.split("test").test_entries()
    assert e.dtype == np.int64 and np.array_equal(e, z["test/entries"])
    assert len(set((int(u), int(q)) for q, u, _, _ in e)) == len(e)                      # distinct (user, query)


def test_item_corpus_needs_a_gpu(loaded):
    _, f, df = loaded
    with pytest.raises(RuntimeError):
        df.item_corpus("cpu", f, f.split("test"))


def test_train_samples_match_reference_rng_stream(loaded):
    """Seeded like main.py:172-173, the sample enumeration is the reference's, word for word (python's shuffle and
    numpy's legacy generator are consumed in the same order)."""
    import random
    z, f, _ = loaded
    tr = f.split("train", subsampling_rate=1e-3)
    before = f.review_word.copy()
    for W in (1, 3):
        random.seed(666)
        np.random.seed(666)
        words, reviews = tr.train_samples(pv_window_size=W)
        assert np.array_equal(words, z["train_samples_w%d/words" % W]), W
        assert np.array_equal(reviews, z["train_samples_w%d/review" % W]), W
    assert np.array_equal(f.review_word, before)                 # no in-place shuffle of the corpus
    # private generators give the same stream without touching the globals
    words2, _ = tr.train_samples(1, py_random=random.Random(666), np_random=np.random.RandomState(666))
    assert np.array_equal(words2, z["train_samples_w1/words"])


def test_eval_samples_with_candidate_lists_match_reference(loaded, tmp_path):
    """collect_test_samples with candidate lists (tests/golden/samples.npz from make_golden_samples.py): the "valid"
    split's sampled candidates and a run file's candidate lists, segmented, on the reference's random streams."""
    import random
    z, f, df = loaded
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "samples.npz"))
    va = f.split("valid", subsampling_rate=1e-3)              # has_valid False: reads the test files (:38-40)
    random.seed(666)
    np.random.seed(666)
    entries, cand = va.test_samples(valid_candi_size=5, candi_batch_size=3)
    assert np.array_equal(entries, g["valid/entries"]) and np.array_equal(cand, g["valid/candidates"])
    assert cand.shape[1] == 3 and (cand[1::2, 2] == -1).all()          # 5 candidates -> segments of 3 + 2
    run = tmp_path / "test.bias_product.ranklist"
    run.write_bytes(g["run/bytes"].tobytes())
    uq = df.read_ranklist(str(run), f.product_ids)
    assert all(len(v) == 6 for v in uq.values())
    te = f.split("test", subsampling_rate=1e-3)
    random.seed(666)
    np.random.seed(666)
    entries, cand = te.test_samples(candi_batch_size=4, uq_pids=uq)
    assert np.array_equal(entries, g["run/entries"]) and np.array_equal(cand, g["run/candidates"])
    # the item-transformer's collate pads candidate lists with the item pad index (item_pv_dataloader.py:44)
    random.seed(666)
    uq = df.read_ranklist(str(run), f.product_ids)
    _, cand_p = te.test_samples(candi_batch_size=4, uq_pids=uq, pad_id=f.product_size)
    assert np.array_equal(cand_p == f.product_size, g["run/candidates"] == -1)
    # whole catalog: no candidate matrix, the entries of test_entries()
    e2, c2 = te.test_samples()
    assert c2 is None and np.array_equal(e2, te.test_entries())


def test_pretrained_embedding_readers_match_reference(tmp_path):
    """others/util.py readers + the word-table alignment of models/item_transformer.py:58-66
    (tests/golden/pretrain.npz from make_golden_pretrain.py): float32 bits identical."""
    from prodsearch_b200 import data_files as df
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "pretrain.npz"))
    wpath, upath = tmp_path / "word_emb.txt.gz", tmp_path / "product_emb.txt"
    wpath.write_bytes(g["word_file"].tobytes())
    upath.write_bytes(g["ui_file"].tobytes())
    index, weights = df.load_pretrain_embeddings(str(wpath))
    assert list(index.keys()) == [str(x) for x in g["index_keys"]] and list(index.values()) == g["index_vals"].tolist()
    assert weights.dtype == np.float32 and np.array_equal(weights.view(np.uint32), g["weights"].view(np.uint32))
    words = [str(x) for x in g["vocab_words"]]
    table = df.pretrained_word_table(str(wpath), words, len(words))
    assert np.array_equal(table.view(np.uint32), g["word_table"].view(np.uint32))
    assert np.array_equal(df.load_user_item_embeddings(str(upath)).view(np.uint32), g["ui"].view(np.uint32))

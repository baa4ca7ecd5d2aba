"""CPU: the host side of ``corpus.ItemCorpus`` -- the flat arrays both constructors upload (``host_arrays``) -- held to
the nested lists of the reference's loaders.  The kernels that read those arrays are checked on the GPU
(tests/test_gpu_batches.py); here the arrays are turned back into nested lists and fed to the oracle collate, which
must give the answers the reference's own dataloader gave (tests/golden/batches.npz, files.npz)."""
import os

import numpy as np
import pytest

from oracle import batches as ob
from prodsearch_b200 import data_files
from prodsearch_b200.corpus import ItemCorpus
from test_oracle_batches import GOLDEN, load_corpus, rtest_inputs

FILES = os.path.join(os.path.dirname(GOLDEN), "files.npz")


def uncsr(off, flat):
    return [[int(x) for x in flat[off[i]:off[i + 1]]] for i in range(len(off) - 1)]


def lists_from_host_arrays(h):
    """What the kernels see, back in the oracle's vocabulary (names of GlobalProdSearchData / ProdSearchData)."""
    R = h["review_user"].shape[0]
    lt = np.zeros((R, 3), dtype=np.int64)
    lt[:, 0] = h["review_uloc"]
    if h["review_time"] is not None:
        lt[:, 2] = h["review_time"]
    u_r_seq = uncsr(h["user_seq_off"], h["user_seq"])
    u_reviews = [set() for _ in u_r_seq]
    for r in np.flatnonzero(h["review_in_set"]):
        u_reviews[int(h["review_user"][r])].add(int(r))
    c = dict(review_u_p=[[int(a), int(b)] for a, b in zip(h["review_user"], h["review_item"])], u_r_seq=u_r_seq,
             review_loc_time=[[int(x) for x in row] for row in lt], u_reviews=u_reviews,
             query_words=[[int(x) for x in row] for row in h["query_words"]],
             product_query_idx=uncsr(h["item_query_off"], h["item_query"]))
    if h["item_seq"] is not None:
        c["i_r_seq"] = uncsr(h["item_seq_off"], h["item_seq"])
        c["p_reviews"] = [set() for _ in c["i_r_seq"]]
        for r in np.flatnonzero(h["review_in_set"]):
            c["p_reviews"][int(h["review_item"][r])].add(int(r))
        # loc_in_item: position of the review in its item's time-ordered sequence (review_loc_time[:, 1])
        for seq in c["i_r_seq"]:
            for pos, r in enumerate(seq):
                c["review_loc_time"][r][1] = pos
    return c


def test_dtypes_and_layout_of_uploaded_arrays():
    z = np.load(GOLDEN)
    c = load_corpus(z)
    h = ItemCorpus.host_arrays(**ItemCorpus.lists_to_arrays(
        c["u_r_seq"], c["review_u_p"], c["query_words"], c["product_query_idx"], train_reviews=c["u_reviews"],
        review_uloc=c["review_loc_time"], product_size=int(z["corpus/P"]), item_seq=c["i_r_seq"]))
    want = dict(review_user=np.int32, review_item=np.int32, review_uloc=np.int32, review_in_set=np.uint8,
                user_seq_off=np.int64, user_seq=np.int32, item_query_off=np.int64, item_query=np.int32,
                query_words=np.int64, item_seq_off=np.int64, item_seq=np.int32, review_time=np.int64)
    assert set(h) == set(want)                       # every pointer field of psb_corpus_t has its array
    for k, dt in want.items():
        assert h[k].dtype == dt and h[k].flags["C_CONTIGUOUS"], k
    assert len(h["item_query_off"]) - 1 == int(z["corpus/P"]) == len(h["item_seq_off"]) - 1
    assert np.array_equal(h["review_in_set"], z["corpus/in_train"].astype(np.uint8))
    assert np.array_equal(h["review_time"], z["corpus/review_loc_time"][:, 2])


def test_list_constructor_arrays_reproduce_reference_batches():
    z = np.load(GOLDEN)
    c = load_corpus(z)
    h = ItemCorpus.host_arrays(**ItemCorpus.lists_to_arrays(
        c["u_r_seq"], c["review_u_p"], c["query_words"], c["product_query_idx"], train_reviews=c["u_reviews"],
        review_uloc=c["review_loc_time"], product_size=int(z["corpus/P"]), item_seq=c["i_r_seq"]))
    c2 = lists_from_host_arrays(h)
    P = int(z["corpus/P"])
    samples = [(list(w), int(r)) for w, r in zip(z["train/word_idxs"], z["train/review_idx"])]
    for tag, do_seq in (("last", False), ("seq", True)):
        b = ob.item_train_batch(c2, samples, z["train/query_pick"], 6, do_seq, True, P)
        for k in ("query_word_idxs", "target_prod_idxs", "u_item_idxs", "pos_iword_idxs"):
            assert np.array_equal(b[k], z["train_%s/%s" % (tag, k)]), (tag, k)
        b = ob.item_test_batch(c2, [tuple(int(x) for x in e) for e in z["test/entries"]], 6, do_seq, P)
        for k in ("query_word_idxs", "target_prod_idxs", "u_item_idxs", "user_idxs", "query_idxs"):
            assert np.array_equal(b[k], z["test_%s/%s" % (tag, k)]), (tag, k)
    entries, cands, pads = rtest_inputs(z)
    for tag, seq_test, tro in (("last", False, True), ("seq", True, False)):
        b = ob.review_test_batch(c2, entries, cands, 4, 5, seq_test, tro, pads)
        for k in ("candi_prod_ridxs", "candi_seg_idxs", "candi_seq_user_idxs", "candi_seq_item_idxs"):
            assert np.array_equal(b[k], z["rtest_%s/%s" % (tag, k)]), (tag, k)


def _files_corpus(tmp_path):
    z = np.load(FILES)
    data, inp = tmp_path / "data", tmp_path / "data" / "split"
    inp.mkdir(parents=True)
    for k in z.files:
        if k.startswith("file/"):
            _, tag, name = k.split("/")
            (data if tag == "data" else inp).joinpath(name).write_bytes(z[k].tobytes())
    return z, data_files.CorpusFiles(str(data), str(inp))


def test_file_constructor_arrays_equal_list_constructor_arrays(tmp_path):
    """``data_files.item_corpus`` (flat arrays, no nested lists) uploads exactly what the list constructor uploads
    for the nested lists the reference's loaders made of the same files."""
    z, files = _files_corpus(tmp_path)
    split = files.split("test")
    seen = {}

    class Capture(ItemCorpus):
        def _setup(self, device, vocab_size=None, **kw):
            seen["h"] = self.host_arrays(**kw)
    data_files.item_corpus("cuda:0", files, split, cls=Capture)
    h = seen["h"]
    rup = z["g/review_u_p"]
    u_reviews = [set() for _ in range(len(z["g/user_ids"]))]
    for r in np.flatnonzero(z["test/in_u_reviews"]):
        u_reviews[int(rup[r, 0])].add(int(r))
    P = len(z["g/product_ids"])
    ref = ItemCorpus.host_arrays(**ItemCorpus.lists_to_arrays(
        uncsr(z["g/u_r_seq_off"], z["g/u_r_seq"]), rup, z["g/query_words"],
        uncsr(z["test/pq_off"], z["test/pq"]), train_reviews=u_reviews,
        review_uloc=[[int(x) for x in row] for row in z["g/review_loc_time"]], product_size=P,
        item_seq=uncsr(z["g/i_r_seq_off"], z["g/i_r_seq"])))
    assert set(h) == set(ref)
    for k in ref:
        assert h[k].dtype == ref[k].dtype and np.array_equal(h[k], ref[k]), k
    # and the oracle collate on them gives what it gives on the reference's lists
    entries = [tuple(int(x) for x in e) for e in split.test_entries()]
    c_ref = dict(review_u_p=[[int(a), int(b)] for a, b in rup], u_r_seq=uncsr(z["g/u_r_seq_off"], z["g/u_r_seq"]),
                 review_loc_time=[[int(x) for x in row] for row in z["g/review_loc_time"]], u_reviews=u_reviews,
                 query_words=[[int(x) for x in row] for row in z["g/query_words"]])
    c2 = lists_from_host_arrays(h)
    for limit, do_seq in ((3, False), (20, False), (4, True)):
        got, want = ob.item_test_batch(c2, entries, limit, do_seq, P), ob.item_test_batch(c_ref, entries, limit,
                                                                                          do_seq, P)
        for k in want:
            assert np.array_equal(got[k], want[k]), (limit, k)


def test_corpus_refuses_cpu_devices_and_ragged_inputs():
    z = np.load(GOLDEN)
    c = load_corpus(z)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ItemCorpus("cpu", c["u_r_seq"], c["review_u_p"], c["query_words"], c["product_query_idx"])
    kw = ItemCorpus.lists_to_arrays(c["u_r_seq"], c["review_u_p"], c["query_words"], c["product_query_idx"],
                                    product_size=int(z["corpus/P"]))
    bad = dict(kw, product_size=int(z["corpus/P"]) + 1)
    with pytest.raises(ValueError, match="one .* row per product"):
        ItemCorpus.host_arrays(**bad)
    with pytest.raises(ValueError, match="rectangle"):
        ItemCorpus.host_arrays(**dict(kw, query_words=np.zeros(5, dtype=np.int64)))

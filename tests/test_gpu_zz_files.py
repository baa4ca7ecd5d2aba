"""GPU: the device corpus built straight from the reference's dataset files (data_files.CorpusFiles ->
ItemCorpus.from_arrays -> psb_build_item_batch) against the oracle collate on the nested lists the reference's
loaders produce from the same files (tests/golden/files.npz).  Added after round 1's GPU budget was spent: its host
half is held on CPU by tests/test_corpus_host.py; the file sorts last among the GPU tests on purpose."""
import os

import numpy as np
import pytest

from oracle import batches as ob
from test_gpu_batches import flags
from test_oracle_batches import GOLDEN

pytestmark = pytest.mark.gpu


def test_corpus_from_dataset_files_matches_oracle(tmp_path):
    """gzip text files (the reference's formats) -> flat arrays -> device corpus -> test batches, against the oracle
    collate on the nested lists the reference's loaders produce from the same files (tests/golden/files.npz)."""
    from prodsearch_b200 import data_files
    z = np.load(os.path.join(os.path.dirname(GOLDEN), "files.npz"))
    data, inp = tmp_path / "data", tmp_path / "data" / "split"
    inp.mkdir(parents=True)
    for k in z.files:
        if k.startswith("file/"):
            _, tag, name = k.split("/")
            (data if tag == "data" else inp).joinpath(name).write_bytes(z[k].tobytes())
    files = data_files.CorpusFiles(str(data), str(inp))
    split = files.split("test")
    corpus = data_files.item_corpus("cuda:0", files, split)
    entries = split.test_entries()

    def uncsr(off, flat):
        return [[int(x) for x in flat[off[i]:off[i + 1]]] for i in range(len(off) - 1)]
    rup = z["g/review_u_p"]
    u_reviews = [set() for _ in range(len(z["g/user_ids"]))]
    for r in np.flatnonzero(z["test/in_u_reviews"]):
        u_reviews[int(rup[r, 0])].add(int(r))
    c = dict(review_u_p=[[int(a), int(b)] for a, b in rup], u_r_seq=uncsr(z["g/u_r_seq_off"], z["g/u_r_seq"]),
             review_loc_time=[[int(x) for x in row] for row in z["g/review_loc_time"]], u_reviews=u_reviews,
             query_words=[[int(x) for x in row] for row in z["g/query_words"]])
    P = len(z["g/product_ids"])
    for limit, seq_test, tro in ((3, False, True), (20, False, True), (4, True, False)):
        fl = flags(uprev_review_limit=limit, do_seq_review_test=seq_test, train_review_only=tro)
        b = corpus.test_batch(entries[:, 0], entries[:, 1], entries[:, 2], entries[:, 3], fl)
        ref = ob.item_test_batch(c, [tuple(int(x) for x in e) for e in entries], limit, seq_test and not tro, P)
        for k in ("query_word_idxs", "target_prod_idxs", "u_item_idxs", "user_idxs", "query_idxs"):
            assert np.array_equal(getattr(b, k).cpu().numpy(), ref[k]), (limit, k)

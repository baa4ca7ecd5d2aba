"""Golden vectors for the pretrained-embedding file readers, produced by the reference's own others/util.py.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden_pretrain.py

Writes a small word-embedding file (gzip) and a user/item embedding file (text) in the reference's formats, reads
them with load_pretrain_embeddings / load_user_item_embeddings and the table construction of
models/item_transformer.py:58-66, and stores file bytes + results in tests/golden/pretrain.npz."""
import gzip
import os
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    assert os.path.isdir(REF), "golden vectors can only be regenerated where /root/reference exists"
    sys.path.insert(0, REF)
    from others.util import load_pretrain_embeddings, load_user_item_embeddings
    rng = np.random.default_rng(11)
    V, d = 12, 5
    words = ["w%d" % i for i in range(V)]
    order = rng.permutation(V)                      # the file lists the words in another order than the vocabulary
    out = {}
    with tempfile.TemporaryDirectory() as root:
        wpath, upath = os.path.join(root, "word_emb.txt.gz"), os.path.join(root, "product_emb.txt")
        with gzip.open(wpath, "wt") as f:
            f.write("%d\n%d\n" % (V + 1, d))
            for i in order:
                f.write("%s\t%s \n" % (words[i], " ".join("%.9g" % x for x in rng.normal(size=d) * 10.0 ** int(rng.integers(-3, 3)))))
            f.write("</s>\t%s\n" % " ".join("%.9g" % x for x in rng.normal(size=d)))
        with open(upath, "w") as f:
            f.write("7\n%d\n" % d)
            for _ in range(7):
                f.write(" ".join("%.9g" % x for x in rng.normal(size=d)) + "\n")
        out["word_file"] = np.frombuffer(open(wpath, "rb").read(), dtype=np.uint8)
        out["ui_file"] = np.frombuffer(open(upath, "rb").read(), dtype=np.uint8)
        index, weights = load_pretrain_embeddings(wpath)
        out["index_keys"] = np.asarray(list(index.keys()))
        out["index_vals"] = np.asarray(list(index.values()), np.int64)
        out["weights"] = torch.FloatTensor(weights).numpy()
        word_pad_idx = V                                            # vocab_size - 1 with vocab_size = len(words) + 1
        idx = torch.tensor([0] + [index[x] for x in words[1:]] + [word_pad_idx])
        out["word_table"] = torch.FloatTensor(weights)[idx].numpy()          # item_transformer.py:62-65
        out["vocab_words"] = np.asarray(words)
        out["ui"] = torch.FloatTensor(load_user_item_embeddings(upath)).numpy()
    np.savez_compressed(os.path.join(OUT, "pretrain.npz"), **out)
    print("pretrain ok:", out["weights"].shape, out["word_table"].shape, out["ui"].shape)


if __name__ == "__main__":
    main()

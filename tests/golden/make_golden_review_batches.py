"""Golden vectors for the review-transformer TRAINING collate, produced by the reference's own loaders.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden_review_batches.py

The synthetic corpus of tests/golden/files.npz (file bytes) is loaded with the reference's GlobalProdSearchData /
ProdSearchData (model_name review_transformer), ``initialize_epoch`` (data/data_util.py:94-117) and
``ProdSearchDataLoader.get_train_batch`` (data/prod_search_dataloader.py:208-358) are run on the global python / numpy
generators seeded like main.py:172-173, and every tensor of the resulting ProdSearchTrainBatch objects is stored in
tests/golden/review_batches.npz.  Nothing here is imported by the product.

Shim (reference untouched): the pv branch of get_train_batch indexes four padded PYTHON LISTS with an index array
(prod_search_dataloader.py:342-345: pos/neg user and item ids are not part of the ``map(np.asarray, ...)`` two lines
above), which raises TypeError for any batch larger than one.  ``others.util.pad`` / ``pad_3d`` are wrapped here to
return a list subclass that accepts an index array the way the neighbouring ndarray fields do -- the evident intent."""
import argparse
import os
import random
import sys
import tempfile

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # limits that bind (random.sample), sub-sampling masks, in-place review shuffling, sliding windows, batch shuffle
    "pvc": dict(review_encoder_name="pvc", do_subsample_mask=True, prepare_pv=True, shuffle=True,
                shuffle_review_words=True, do_seq_review_train=False, pv_window_size=2, review_word_limit=6,
                uprev_review_limit=2, iprev_review_limit=3, neg_per_pos=3),
    # sequential history (loc_in_user / bisect on the time stamp), words sub-sampled once per epoch, no pv batches
    "fs_seq": dict(review_encoder_name="fs", do_subsample_mask=False, prepare_pv=False, shuffle=False,
                   shuffle_review_words=False, do_seq_review_train=True, pv_window_size=1, review_word_limit=5,
                   uprev_review_limit=3, iprev_review_limit=2, neg_per_pos=4),
    # pv windows that do not divide the word limit, nothing shuffled
    "pv": dict(review_encoder_name="pv", do_subsample_mask=True, prepare_pv=True, shuffle=False,
               shuffle_review_words=False, do_seq_review_train=False, pv_window_size=3, review_word_limit=7,
               uprev_review_limit=4, iprev_review_limit=4, neg_per_pos=2),
}
FIELDS = ("query_word_idxs", "pos_prod_ridxs", "pos_seg_idxs", "pos_prod_rword_idxs", "pos_prod_rword_masks",
          "neg_prod_ridxs", "neg_seg_idxs", "pos_user_idxs", "neg_user_idxs", "pos_item_idxs", "neg_item_idxs",
          "neg_prod_rword_idxs", "neg_prod_rword_masks", "pos_prod_rword_idxs_pvc", "neg_prod_rword_idxs_pvc")
BATCH = 16


class ArrList(list):
    """A padded list that can also be indexed by an index array (see the module docstring)."""

    def __getitem__(self, i):
        if isinstance(i, np.ndarray):
            return np.asarray(list(self))[i]
        return list.__getitem__(self, i)


def main():
    assert os.path.isdir(REF), "golden vectors can only be regenerated where /root/reference exists"
    sys.path.insert(0, REF)
    from data.data_util import GlobalProdSearchData, ProdSearchData
    from data.prod_search_dataloader import ProdSearchDataLoader
    from data.prod_search_dataset import ProdSearchDataset
    import others.util as util
    pad0, pad3 = util.pad, util.pad_3d
    util.pad = lambda *a, **k: ArrList(pad0(*a, **k))
    util.pad_3d = lambda *a, **k: ArrList(pad3(*a, **k))
    z = np.load(os.path.join(OUT, "files.npz"))
    out = {}
    with tempfile.TemporaryDirectory() as root:
        data, inp = os.path.join(root, "data"), os.path.join(root, "data", "split")
        os.makedirs(inp)
        for k in z.files:
            if k.startswith("file/"):
                _, tag, name = k.split("/")
                open(os.path.join(data if tag == "data" else inp, name), "wb").write(z[k].tobytes())
        for case, cfg in CASES.items():
            args = argparse.Namespace(model_name="review_transformer", subsampling_rate=1e-2, fix_emb=False,
                                      has_valid=False, test_candi_size=-1, prod_freq_neg_sample=False,
                                      valid_candi_size=-1, train_review_only=True, candi_batch_size=1000,
                                      corrupt_rate=0.9, do_seq_review_test=False,
                                      **{k: v for k, v in cfg.items() if k not in ("prepare_pv", "shuffle")})
            g = GlobalProdSearchData(args, data, inp)
            tr = ProdSearchData(args, inp, "train", g)
            random.seed(666)
            np.random.seed(666)
            tr.initialize_epoch()
            ds = ProdSearchDataset(args, g, tr)
            loader = ProdSearchDataLoader(args, ds, prepare_pv=cfg["prepare_pv"], batch_size=BATCH,
                                          shuffle=cfg["shuffle"])
            out["%s/neg_sample_products" % case] = np.asarray(tr.neg_sample_products, np.int64)
            words = g.review_words if args.do_subsample_mask else g.padded_review_words
            out["%s/review_words" % case] = np.asarray(words, np.int64)
            rows = list(tr.review_info)
            n_made = 0
            for b0 in range(0, len(rows), BATCH):
                res = loader.get_train_batch(rows[b0:b0 + BATCH])
                res = res if isinstance(res, list) else [res]
                out["%s/b%d/count" % (case, b0 // BATCH)] = np.int64(len(res))
                for j, b in enumerate(res):
                    for f in FIELDS:
                        v = getattr(b, f)
                        if v is not None:
                            out["%s/b%d/%d/%s" % (case, b0 // BATCH, j, f)] = v.numpy()
                    n_made += 1
            out["%s/n_batches" % case] = np.int64((len(rows) + BATCH - 1) // BATCH)
            print("%s: %d train rows -> %d batch objects" % (case, len(rows), n_made))
    np.savez_compressed(os.path.join(OUT, "review_batches.npz"), **out)


if __name__ == "__main__":
    main()

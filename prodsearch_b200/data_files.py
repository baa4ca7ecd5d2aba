"""Readers for the reference's dataset files ("I/O formats").

``GlobalProdSearchData`` / ``ProdSearchData`` (data/data_util.py:10-62,:165-287) parse the gzip text files of the
Amazon product-search corpora into nested Python lists.  ``CorpusFiles`` / ``SplitFiles`` read the SAME files into
the flat arrays the device-side batch construction works on (``corpus.ItemCorpus``): CSR offsets + int32 payloads
instead of lists of lists, a byte flag per review instead of per-user / per-item sets.  File formats, field
meanings and the derived quantities (sub-sampling rates, count^0.75 negative-sampling distribution, the test
entry enumeration) follow the reference line by line; host-side I/O only, no model code.
"""
import gzip
import os

import numpy as np


def _lines(path):
    with gzip.open(path, "rt") as f:
        for line in f:
            yield line


def read_lines(path):
    """data_util.py:264-269 (GlobalProdSearchData.read_lines)."""
    return [line.strip() for line in _lines(path)]


def read_csr(path, strict=False):
    """Lines of space-separated ints -> (offsets int64 [n+1], flat int32).  strict=False skips empty fields like
    read_arr_from_lines (data_util.py:251-262); strict=True is read_words_in_lines (:272-281)."""
    off, flat = [0], []
    for line in _lines(path):
        parts = line.strip().split(" ") if strict else [x for x in line.strip().split(" ") if len(x) >= 1]
        flat.extend(int(x) for x in parts)
        off.append(len(flat))
    return np.asarray(off, dtype=np.int64), np.asarray(flat, dtype=np.int32)


def _uncsr(off, flat):
    return [flat[off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]


def read_review_id(path, line_review_id_map):
    """data_util.py:221-236: rows (line_no, user_idx, product_idx, review_idx) + the trailing query ids."""
    info, queries = [], []
    for line_no, line in enumerate(_lines(path)):
        arr = line.strip().split("\t")
        review_id = line_review_id_map[int(arr[2].split("_")[-1])]
        info.append((line_no, int(arr[0]), int(arr[1]), review_id))
        if arr[-1].isdigit():
            queries.append(int(arr[-1]))
    return np.asarray(info, dtype=np.int64).reshape(-1, 4), queries


class CorpusFiles(object):
    """GlobalProdSearchData (data_util.py:165-203) as flat arrays."""

    def __init__(self, data_path, input_train_dir, model_name="item_transformer"):
        self.data_path, self.input_train_dir = data_path, input_train_dir
        self.product_ids = read_lines(os.path.join(data_path, "product.txt.gz"))
        self.user_ids = read_lines(os.path.join(data_path, "users.txt.gz"))
        self.words = read_lines(os.path.join(data_path, "vocab.txt.gz"))
        self.product_size, self.user_size = len(self.product_ids), len(self.user_ids)
        self.vocab_size = len(self.words) + 1
        self.word_pad_idx = self.vocab_size - 1
        q_off, q_flat = read_csr(os.path.join(input_train_dir, "query.txt.gz"), strict=True)
        wq = int(np.diff(q_off).max()) if len(q_off) > 1 else 0
        self.query_words = np.full((len(q_off) - 1, wq), self.word_pad_idx, dtype=np.int64)     # util.pad
        for i in range(len(q_off) - 1):
            self.query_words[i, :q_off[i + 1] - q_off[i]] = q_flat[q_off[i]:q_off[i + 1]]
        self.review_word_off, self.review_word = read_csr(os.path.join(data_path, "review_text.txt.gz"), strict=True)
        self.review_length = np.diff(self.review_word_off)
        self.review_count = len(self.review_length) + 1
        self.user_seq_off, self.user_seq = read_csr(os.path.join(data_path, "u_r_seq.txt.gz"))
        self.item_seq_off, self.item_seq = read_csr(os.path.join(data_path, "p_r_seq.txt.gz"))
        lt_off, lt = read_csr(os.path.join(data_path, "review_uloc_ploc_and_time.txt.gz"))
        self.review_loc_time = lt.astype(np.int64).reshape(-1, 3)          # (loc_in_user, loc_in_item, time)
        self.line_review_id_map = {}
        for idx, line in enumerate(_lines(os.path.join(data_path, "review_id.txt.gz"))):
            self.line_review_id_map[int(line.strip().split("_")[-1])] = idx
        self.train_review_info, self.train_query_idxs = read_review_id(
            os.path.join(input_train_dir, "train_id.txt.gz"), self.line_review_id_map)
        up_off, up = read_csr(os.path.join(data_path, "review_u_p.txt.gz"))
        self.review_u_p = up.astype(np.int64).reshape(-1, 2)
        # membership of the training split (ProdSearchData.get_u_i_reviews_set, data_util.py:86-92): one flag per
        # review serves both the per-user and the per-item sets (a review has one user and one item)
        self.review_in_train = np.zeros(self.review_u_p.shape[0], dtype=np.uint8)
        self.review_in_train[self.train_review_info[:, 3]] = 1

    def split(self, set_name, subsampling_rate=1e-5, has_valid=False, fix_emb=False, prod_freq_neg_sample=False):
        return SplitFiles(self, set_name, subsampling_rate, has_valid, fix_emb, prod_freq_neg_sample)


class SplitFiles(object):
    """ProdSearchData (data_util.py:10-62) for one of train / valid / test."""

    def __init__(self, corpus, set_name, subsampling_rate=1e-5, has_valid=False, fix_emb=False,
                 prod_freq_neg_sample=False):
        self.corpus, self.set_name = corpus, set_name
        d = corpus.input_train_dir
        self.word_dists = self.sub_sampling_rate = self.vocab_distribute = None
        if fix_emb:
            subsampling_rate = 0
        if set_name == "train":
            self.vocab_distribute = np.zeros(corpus.vocab_size)
            for line in _lines(os.path.join(d, "train.txt.gz")):            # read_reviews, :125-136
                for w in line.strip().split("\t")[2].split(" "):
                    self.vocab_distribute[int(w)] += 1
            self.sub_sampling_rate = sub_sampling(self.vocab_distribute, subsampling_rate)
            self.word_dists = neg_distributes(self.vocab_distribute)
            self.item_query_off, self.item_query = read_csr(os.path.join(d, "train_query_idx.txt.gz"))
            self.review_info, self.review_query_idx = corpus.train_review_info, corpus.train_query_idxs
        else:
            read_name = set_name if has_valid else "test"                   # :38-40
            self.item_query_off, self.item_query = read_csr(os.path.join(d, "test_query_idx.txt.gz"))
            self.review_info, self.review_query_idx = read_review_id(
                os.path.join(d, "%s_id.txt.gz" % read_name), corpus.line_review_id_map)
        if prod_freq_neg_sample:
            dist = np.zeros(corpus.product_size)
            np.add.at(dist, corpus.train_review_info[:, 2], 1)              # collect_product_distribute, :119-123
        else:
            dist = np.ones(corpus.product_size)
        self.product_dists = neg_distributes(dist)

    @property
    def product_query_idx(self):
        return _uncsr(self.item_query_off, self.item_query)

    def train_samples(self, pv_window_size=1, py_random=None, np_random=None):
        """ItemPVDataset.collect_train_samples (item_pv_dataset.py:71-93): the (word window, review) samples of one
        epoch -> (words int64 [n, pv_window_size], review_idx int64 [n]).  The reference draws from the GLOBAL python and
        numpy generators (one ``np.random.random`` call for all words, one ``random.shuffle`` per review, in file
        order); the same calls are made here on ``py_random`` / ``np_random`` (default: those globals), so a run
        seeded like main.py:172-173 enumerates the identical samples.  Quirk kept: a sub-sampled (skipped) word does
        not advance the random-number cursor (:84-85).  The corpus' review words are NOT shuffled in place."""
        import random as _random
        py_random = py_random if py_random is not None else _random
        np_random = np_random if np_random is not None else np.random
        c = self.corpus
        rand_numbers = np_random.random(int(c.review_length.sum()))
        rate = self.sub_sampling_rate
        W = int(pv_window_size)
        words, reviews, cur, entry_id = [], [], [], 0
        off, flat = c.review_word_off, c.review_word
        for _, _, _, review_idx in self.review_info.tolist():
            ws = flat[off[review_idx]:off[review_idx + 1]].tolist()
            py_random.shuffle(ws)
            for w in ws:
                if rand_numbers[entry_id] > rate[w]:
                    continue
                cur.append(w)
                if len(cur) == W:
                    words.append(cur)
                    reviews.append(review_idx)
                    cur = []
                entry_id += 1
        if len(cur) > 0:
            words.append(cur + [c.word_pad_idx] * (W - len(cur)))
            reviews.append(review_idx)
        return np.asarray(words, dtype=np.int64).reshape(len(words), W), np.asarray(reviews, dtype=np.int64)

    def test_entries(self):
        """ItemPVDataset.collect_test_samples (item_pv_dataset.py:36-68) when the whole catalog is the candidate set
        (``test_candi_size < 1``, no ranklist file): the distinct (user, query) pairs in file order, one query per
        query of the purchased item -> int64 [n, 4] rows (query_idx, user_idx, prod_idx, review_idx)."""
        seen, out = set(), []
        off, flat = self.item_query_off, self.item_query
        for _, user_idx, prod_idx, review_idx in self.review_info.tolist():
            for query_idx in flat[off[prod_idx]:off[prod_idx + 1]].tolist():
                if (user_idx, query_idx) in seen:
                    continue
                seen.add((user_idx, query_idx))
                out.append((query_idx, user_idx, prod_idx, review_idx))
        return np.asarray(out, dtype=np.int64).reshape(-1, 4)


    def test_samples(self, valid_candi_size=-1, candi_batch_size=1000, uq_pids=None, py_random=None, np_random=None,
                     pad_id=-1):
        """collect_test_samples in full (item_pv_dataset.py:36-68 = prod_search_dataset.py:43-84): one entry per
        distinct (user, query) pair and per segment of ``candi_batch_size`` candidates.  Candidates are
        * ``uq_pids[(user id string, query_idx)]`` (``read_ranklist``) shuffled with ``random.shuffle``, or
        * for the "valid" split with ``valid_candi_size > 1``: ``valid_candi_size - 1`` items drawn without
          replacement by ``numpy.random.choice(p=product_dists)`` plus the purchased item (which may therefore
          appear twice), shuffled, or
        * the whole catalog (then ``candidates`` is None and ``test_entries`` is the cheaper call).
        The generators are the reference's (global ``random`` / ``numpy.random`` unless given), called in its order.
        Returns (entries int64 [n, 4] of (query_idx, user_idx, prod_idx, review_idx), candidates int64
        [n, width] padded with ``pad_id`` (-1 in the review-transformer's test collate, prod_search_dataloader.py:92;
        the item-transformer's pads with the item pad index, item_pv_dataloader.py:44) -- or None)."""
        import random as _random
        py_random = py_random if py_random is not None else _random
        np_random = np_random if np_random is not None else np.random
        c = self.corpus
        whole = uq_pids is None and not (self.set_name == "valid" and valid_candi_size > 1)
        if whole and c.product_size <= candi_batch_size:
            return self.test_entries(), None
        seen, entries, cands = set(), [], []
        off, flat = self.item_query_off, self.item_query
        for _, user_idx, prod_idx, review_idx in self.review_info.tolist():
            for query_idx in flat[off[prod_idx]:off[prod_idx + 1]].tolist():
                if (user_idx, query_idx) in seen:
                    continue
                seen.add((user_idx, query_idx))
                if uq_pids is not None:
                    items = uq_pids[(c.user_ids[user_idx], query_idx)]
                    py_random.shuffle(items)                      # in place, like the reference
                elif whole:
                    items = list(range(c.product_size))
                else:
                    items = np_random.choice(c.product_size, size=valid_candi_size - 1, replace=False,
                                             p=self.product_dists).tolist()
                    items.append(prod_idx)
                    py_random.shuffle(items)
                for i in range(int((len(items) - 1) / candi_batch_size) + 1):
                    entries.append((query_idx, user_idx, prod_idx, review_idx))
                    cands.append(items[i * candi_batch_size:(i + 1) * candi_batch_size])
        width = max((len(x) for x in cands), default=0)
        cand = np.full((len(cands), width), pad_id, dtype=np.int64)
        for i, x in enumerate(cands):
            cand[i, :len(x)] = x
        return np.asarray(entries, dtype=np.int64).reshape(-1, 4), cand

def read_ranklist(path, product_ids):
    """ProdSearchData.read_ranklist (data_util.py:64-72): a TREC run file (``<user>_<query> Q0 <asin> <rank> ...``) ->
    {(user id string, query idx): [product idx, ...]} in file order -- the candidate lists of ``test_candi_size > 0``."""
    index = {x: i for i, x in enumerate(product_ids)}
    out = {}
    with open(path, "r") as f:
        for line in f:
            arr = line.strip().split(" ")
            uid, qid = arr[0].split("_")
            out.setdefault((uid, int(qid)), []).append(index[arr[2]])
    return out


def load_pretrain_embeddings(path):
    """others/util.py:4-20: gzip text, line 1 = count, line 2 = size, then ``<key>\t<v0> <v1> ...`` per row ->
    ({key: row}, float32 [rows, size]).  The reference parses to python floats and builds a FloatTensor: the same
    double -> float32 rounding as here."""
    index, rows = {}, []
    with gzip.open(path, "rt") as f:
        f.readline()
        f.readline()
        for line_no, line in enumerate(f):
            arr = line.strip(" ").split("\t")
            index[arr[0]] = line_no
            rows.append([float(x) for x in arr[1].split()])
    return index, np.asarray(rows, dtype=np.float64).astype(np.float32)


def load_user_item_embeddings(path):
    """others/util.py:22-34: plain text, count / size header, one space-separated row per line -> float32 [rows, size]."""
    rows = []
    with open(path, "r") as f:
        f.readline()
        f.readline()
        for line in f:
            rows.append([float(x) for x in line.strip().split(" ")])
    return np.asarray(rows, dtype=np.float64).astype(np.float32)


def pretrained_word_table(path, vocab_words, word_pad_idx):
    """The word table the reference builds from a pretrained file (models/item_transformer.py:58-66,
    models/PVC.py:24-28): row i = the file's vector of ``vocab_words[i]``, except row 0 (the file's first row) and
    the last row (the file's row ``word_pad_idx``) -- float32 [len(vocab_words) + 1, size], ready for
    ``model.word_embeddings.weight.data.copy_`` / ``load_state_dict``.  The reference wraps it in
    ``nn.Embedding.from_pretrained`` (frozen): set ``requires_grad = False`` on the table to train like it."""
    index, weights = load_pretrain_embeddings(path)
    rows = [0] + [index[w] for w in vocab_words[1:]] + [word_pad_idx]
    return weights[np.asarray(rows, dtype=np.int64)]


def sub_sampling(vocab_distribute, subsample_threshold):
    """data_util.py:138-153."""
    vd = np.asarray(vocab_distribute, dtype=np.float64)
    rate = np.ones(len(vd))
    if subsample_threshold == 0.0:
        return rate
    threshold = sum(vd.tolist()) * subsample_threshold          # python float sum, as on the reference's list
    for i in range(len(vd)):
        if vd[i] == 0:
            rate[i] = 0
            continue
        rate[i] = min(1.0, (np.sqrt(float(vd[i]) / threshold) + 1) * threshold / float(vd[i]))
    return rate


def neg_distributes(weights, distortion=0.75):
    """data_util.py:155-162."""
    weights = np.asarray(weights)
    wf = weights / weights.sum()
    wf = np.power(wf, distortion)
    return wf / wf.sum()


def item_corpus(device, files, split, cls=None):
    """``corpus.ItemCorpus`` (CSR arrays in HBM) straight from the files: no nested lists in between."""
    if cls is None:
        from .corpus import ItemCorpus as cls
    return cls.from_arrays(
        device, review_u_p=files.review_u_p, review_uloc=files.review_loc_time[:, 0],
        review_time=files.review_loc_time[:, 2], review_in_set=files.review_in_train,
        user_seq=(files.user_seq_off, files.user_seq), item_seq=(files.item_seq_off, files.item_seq),
        item_query=(split.item_query_off, split.item_query), query_words=files.query_words,
        product_size=files.product_size, vocab_size=files.vocab_size)

"""CPU: the candidate-list evaluation glue of prodsearch_b200/evaluate.py (``rank_candidates`` / ``validate``): the
regrouping of candidate segments and the ranking rule, against the oracle's restatement of Trainer.get_prod_scores +
calc_metrics.  The batch construction and the scoring run on the GPU in the product (and are tested there); here
they are stand-ins that return fixed scores, so only the glue is exercised."""
import argparse

import numpy as np
import torch

import oracle
from prodsearch_b200 import evaluate


def test_rank_candidates_ties_padding_and_missing_target():
    ids = torch.tensor([[7, 3, 9, -1, 5], [4, 4, 2, 8, -1], [1, 2, 3, 4, 6]])
    sc = torch.tensor([[1.0, 2.0, 2.0, 99.0, 2.0], [0.5, 0.5, 0.5, 0.1, 7.0], [3.0, 1.0, 2.0, 0.0, -1.0]])
    r_ids, r_sc, rank = evaluate.rank_candidates(ids, sc, torch.tensor([9, 4, 5]))
    assert r_ids.tolist() == [[3, 5, 9, 7, -1], [2, 4, 4, 8, -1], [1, 3, 2, 4, 6]]     # ties -> lower id; padding last
    assert rank.tolist() == [3, 2, 0]                                                    # first occurrence; absent -> 0
    assert r_sc[0, :4].tolist() == [2.0, 2.0, 2.0, 1.0] and r_sc[0, 4] == float("-inf")


class _Corpus(object):
    prod_pad_idx = 50

    def test_batch(self, q, u, p, r, args, candi_prod_idxs=None):
        assert (candi_prod_idxs != -1).all()                       # padding arrives as the item pad index
        return argparse.Namespace(candi_prod_idxs=torch.as_tensor(candi_prod_idxs), query_idxs=q, user_idxs=u)


class _Model(torch.nn.Module):
    def __init__(self, table):
        super().__init__()
        self.table = table

    def test(self, batch):
        return self.table[batch.candi_prod_idxs]                    # a fixed score per item; pad item scores 0


def test_validate_joins_segments_and_matches_oracle_metrics():
    rng = np.random.default_rng(0)
    P, n_pairs, C, W = 50, 23, 7, 3
    table = torch.cat([torch.tensor(np.round(rng.normal(size=P), 1), dtype=torch.float32), torch.zeros(1)])
    entries, cands, full = [], [], []
    for i in range(n_pairs):
        items = rng.choice(P, size=C, replace=False).tolist()
        tgt = items[int(rng.integers(0, C))] if i % 5 else int((set(range(P)) - set(items)).pop())   # some absent
        full.append(items)
        for s in range(0, C, W):
            entries.append((i % 4, i, tgt, 100 + i))
            seg = items[s:s + W]
            cands.append(seg + [-1] * (W - len(seg)))
    mrr, prec, r_ids, r_sc, q_idx, u_idx = evaluate.validate(_Model(table), _Corpus(), entries, cands,
                                                             argparse.Namespace(), cutoff=4, batch_size=10)
    full = np.asarray(full)
    sc = table.numpy()[full]
    ref_ids, ref_sc = oracle.topk_lower_id_first(sc, C, full)
    assert np.array_equal(r_ids[:, :C].numpy(), ref_ids) and np.array_equal(r_sc[:, :C].numpy(), ref_sc)
    assert (r_ids[:, C:] == -1).all()                                                # the segments' padding ranks last
    tgt = np.asarray([e[2] for e in entries[::3]])
    ref_mrr, ref_prec = oracle.calc_metrics(ref_ids, tgt, cutoff=4)
    assert mrr == ref_mrr and prec == ref_prec
    assert u_idx.tolist() == list(range(n_pairs)) and q_idx.tolist() == [i % 4 for i in range(n_pairs)]

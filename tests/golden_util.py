"""Loader for tests/golden/*.npz (written by tests/golden/make_golden.py)."""
import argparse
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# defaults of make_golden.base_args (reference flag defaults, small sizes)
DEFAULTS = dict(
    train_review_only=True, embedding_size=64, dropout=0.0, pretrain_emb_dir="",
    pretrain_up_emb_dir="", sep_prod_emb=False, model_name="item_transformer", ff_size=64,
    heads=8, inter_layers=1, query_encoder_name="fs", use_dot_prod=True, use_pos_emb=True,
    use_item_pos=False, sim_func="product", pos_weight=False, neg_per_pos=3,
    review_encoder_name="pv", fix_emb=False, do_subsample_mask=False, review_word_limit=6,
    use_user_emb=False, use_item_emb=False, use_seg_emb=True, corrupt_rate=0.5)


def _parse(v, like):
    if isinstance(like, bool):
        return v == "True"
    return type(like)(v)


class Golden(object):
    def __init__(self, name, **extra_cfg):
        z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        self.z = z
        cfg = dict(DEFAULTS)
        cfg.update(extra_cfg)
        if "cfg/keys" in z.files:
            for k, v in zip(z["cfg/keys"].tolist(), z["cfg/vals"].tolist()):
                cfg[k] = _parse(v, cfg[k])
        self.cfg = argparse.Namespace(**cfg)
        self.params, self.grads, self.inputs, self.outputs, self.draws = {}, {}, {}, {}, {}
        for k in z.files:
            head, _, tail = k.partition("/")
            dst = {"param": self.params, "grad": self.grads, "in": self.inputs,
                   "out": self.outputs, "draw": self.draws}.get(head)
            if dst is not None:
                dst[tail] = torch.from_numpy(z[k])

    def leaf_params(self, pe_builder=None):
        """Parameters as autograd leaves; the sinusoid buffer is rebuilt to full length."""
        P = {}
        for k, v in self.params.items():
            if k.endswith("pos_emb.pe"):
                full = pe_builder(5000, v.shape[-1])
                assert torch.equal(full[:, :v.shape[1]], v), "sinusoid table mismatch"
                P[k] = full
            elif v.dtype.is_floating_point:
                P[k] = v.clone().requires_grad_(True)
            else:
                P[k] = v.clone()
        return P

    def batch(self):
        return argparse.Namespace(**{k: v for k, v in self.inputs.items()})

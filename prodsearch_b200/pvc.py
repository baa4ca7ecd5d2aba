"""ParagraphVectorCorruption review encoder (reference models/PVC.py) on the fused kernels."""
import torch
import torch.nn as nn

from . import functional as F_


class ParagraphVectorCorruption(nn.Module):
    """Same constructor as models/PVC.py:16.  The [N, Wr, d] gather the reference materialises
    twice per step (PVC.py:76-79) is replaced by fused gather + mean kernels; the corruption
    mask enters as a per-token scale (0 or 1/(1-rate)) and -- reproducing the reference's
    ``.data`` edit (PVC.py:53) -- is NOT seen by the backward pass."""

    def __init__(self, word_embeddings, word_dists, corrupt_rate, dropout=0.0, pretrain_emb_path=None,
                 vocab_words=None, fix_emb=False, word_sink=None):
        super().__init__()
        if pretrain_emb_path is not None:
            raise NotImplementedError("pretrained context embeddings: build the table with "
                                      "data_files.pretrained_word_table and copy it into the word table")
        self.word_embeddings = word_embeddings
        self.context_embeddings = word_embeddings           # PVC.py:30
        self.word_dists = word_dists
        self._embedding_size = word_embeddings.weight.size(-1)
        self.word_pad_idx = word_embeddings.weight.size(0) - 1
        self.corrupt_rate = corrupt_rate
        self.train_corrupt_rate = corrupt_rate
        self.dropout_ = dropout
        self.word_sink = word_sink or F_.RowGradSink(word_embeddings.weight, self.word_pad_idx)
        self.injected_negatives = None
        self.injected_corruption = []                        # queue of [N,Wr] 0/1 masks (1 = dropped)

    @property
    def embedding_size(self):
        return self._embedding_size

    def set_to_evaluation_mode(self):
        self.corrupt_rate = 0.

    def set_to_train_mode(self):
        self.corrupt_rate = self.train_corrupt_rate

    def _token_scale(self, idx, rate):
        """apply_token_dropout (PVC.py:46-54) as a multiplier: dropped -> 0, kept -> 1/(1-rate)."""
        if self.injected_corruption:
            dropped = self.injected_corruption.pop(0).to(idx.device).float()
        else:
            dropped = torch.bernoulli(torch.full(idx.shape, rate, dtype=torch.float32, device=idx.device))
        return ((1.0 - dropped) * (1.0 / (1.0 - rate))).contiguous()

    def get_para_vector(self, prod_rword_idxs_pvc):
        """PVC.py:56-61."""
        scale = self._token_scale(prod_rword_idxs_pvc, self.corrupt_rate) if self.corrupt_rate > 0. else None
        return F_.meanpool(self.context_embeddings.weight, prod_rword_idxs_pvc, self.word_sink,
                           pad_idx=self.word_pad_idx, tok_scale=scale)

    def forward(self, review_word_idxs, review_word_mask, prod_rword_idxs_pvc, n_negs):
        """PVC.py:69-95 -> (uncorrupted review_emb [N,d], loss [N,1]).  ``review_word_idxs``: the target-word
        indices [N,W] (fused form) or the reference's dense ``review_word_emb`` [N,W,d] (PVC.py:69)."""
        dense = review_word_idxs.is_floating_point()
        n, w = review_word_idxs.shape[:2]
        table = self.context_embeddings.weight
        review_emb = F_.meanpool(table, prod_rword_idxs_pvc, self.word_sink, pad_idx=self.word_pad_idx)
        # the reference draws the mask even at rate 0 (PVC.py:78); rate 0 keeps every token
        scale = self._token_scale(prod_rword_idxs_pvc, self.corrupt_rate)
        corr = F_.meanpool(table, prod_rword_idxs_pvc, self.word_sink, pad_idx=self.word_pad_idx, tok_scale=scale)
        if self.injected_negatives is not None:
            neg = self.injected_negatives
        else:
            neg = torch.multinomial(self.word_dists, n * w * n_negs, replacement=True)
        mask = review_word_mask.to(torch.uint8).contiguous()
        if dense:
            loss = F_.ns_loss_dense_pos(corr, review_word_idxs, self.word_embeddings.weight, neg.view(n, w, n_negs),
                                        self.word_sink, mask=mask)
        else:
            loss = F_.ns_loss(corr, self.word_embeddings.weight, review_word_idxs, neg.view(n, w, n_negs),
                              self.word_sink, mask=mask)
        return review_emb, loss.unsqueeze(-1)

    def initialize_parameters(self, logger=None):
        pass

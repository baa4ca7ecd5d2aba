"""ParagraphVector review encoder (reference models/PV.py) on the fused NS-loss kernel."""
import torch
import torch.nn as nn

from . import functional as F_


class ParagraphVector(nn.Module):
    """Same constructor as models/PV.py:16.  ``forward`` accepts both call forms: the reference's
    ``(review_ids, review_word_emb [N,W,d], review_word_mask, n_negs)`` (PV.py:50; the caller has gathered
    the target-word rows, their gradient returns through autograd) and the fused form with the target-word
    INDICES [N,W] in the same argument position, which skips the [N,W,d] round trip (what the model classes
    of this package pass)."""

    def __init__(self, word_embeddings, word_dists, review_count, dropout=0.0, pretrain_emb_path=None,
                 fix_emb=False, word_sink=None):
        super().__init__()
        if pretrain_emb_path is not None:
            raise NotImplementedError("pretrained review embeddings: read them with data_files.load_pretrain_embeddings "
                                      "and copy them into review_embeddings.weight")
        self.word_embeddings = word_embeddings
        self.fix_emb = fix_emb
        self.dropout_ = 0 if fix_emb else dropout
        self.word_dists = word_dists
        self._embedding_size = word_embeddings.weight.size(-1)
        self.review_count = review_count
        self.review_pad_idx = review_count - 1
        self.review_embeddings = nn.Embedding(review_count, self._embedding_size, padding_idx=self.review_pad_idx)
        if fix_emb:
            self.review_embeddings.weight.requires_grad = False
        self.word_sink = word_sink or F_.RowGradSink(word_embeddings.weight, word_embeddings.weight.size(0) - 1)
        self.review_sink = F_.RowGradSink(self.review_embeddings.weight, self.review_pad_idx)
        self.injected_negatives = None

    @property
    def embedding_size(self):
        return self._embedding_size

    def get_para_vector(self, review_ids):
        """PV.py:46-48."""
        return F_.gather_rows(self.review_embeddings.weight, review_ids, self.review_sink)

    def forward(self, review_ids, review_word_idxs, review_word_mask, n_negs):
        """PV.py:50-80 -> (review_emb [N,d], loss [N,1]).  ``review_word_idxs``: int64 indices [N,W], or the
        reference's dense ``review_word_emb`` [N,W,d]."""
        dense = review_word_idxs.is_floating_point()
        n, w = review_word_idxs.shape[:2]
        review_emb = self.get_para_vector(review_ids)
        keep = F_._dropout_keep(review_emb.shape, self.dropout_, self.training, review_emb.device)
        if keep is not None:
            review_emb = review_emb * keep
        if self.injected_negatives is not None:
            neg = self.injected_negatives
        else:
            neg = torch.multinomial(self.word_dists, n * w * n_negs, replacement=True)
        mask = review_word_mask.to(torch.uint8).contiguous()
        if dense:
            loss = F_.ns_loss_dense_pos(review_emb, review_word_idxs, self.word_embeddings.weight,
                                        neg.view(n, w, n_negs), self.word_sink, mask=mask)
        else:
            loss = F_.ns_loss(review_emb, self.word_embeddings.weight, review_word_idxs, neg.view(n, w, n_negs),
                              self.word_sink, mask=mask)
        return review_emb, loss.unsqueeze(-1)

    def initialize_parameters(self, logger=None):
        nn.init.normal_(self.review_embeddings.weight)

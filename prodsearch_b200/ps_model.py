"""RTM: ``ProductRanker`` (reference models/ps_model.py) on the sm_100a hot path.

Same constructor / forward / test / get_review_embeddings / clear_review_embbeddings surface and
state_dict keys as the reference (ps_model.py:53-371).  The gather-heavy parts run on the fused
kernels: query encoder (:257-258), review-word gathers and PV / PVC objectives (:261-300), fs / avg
review encoders (:301-305), review / segment / user / item row gathers (:281-334,:214), and all
embedding-table gradients (deterministic sort + segmented reduce).  The transformer encoder with
its ``wo`` head stays on cuBLAS/ATen ("next" row N1), as does the final
[B, 1+K] BCE (:351-356).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as F_
from . import ops
from .pv import ParagraphVector
from .pvc import ParagraphVectorCorruption
from .text_encoder import AVGEncoder, FSEncoder
from .transformer import TransformerEncoder


def pad_reviews(data, pad_id, width):
    """others/util.py:36-40 ``pad`` (truncate / right-pad every review to ``width`` words)."""
    return [list(d[:width]) + [pad_id] * (width - len(d)) for d in data]


class ProductRanker(F_.LazyFlushMixin, nn.Module):
    def __init__(self, args, device, vocab_size, review_count, product_size, user_size, review_words,
                 vocab_words, word_dists=None, grad_mode="dense"):
        super().__init__()
        self.args = args
        self.device = device
        self.train_review_only = args.train_review_only
        self.embedding_size = d = args.embedding_size
        self.vocab_words = vocab_words
        self.word_dists = None
        if word_dists is not None:
            self.word_dists = torch.as_tensor(word_dists, dtype=torch.float32, device=device)
        self.prod_pad_idx = product_size
        self.user_pad_idx = user_size
        self.word_pad_idx = vocab_size - 1
        self.seg_pad_idx = 3
        self.review_pad_idx = review_count - 1
        self.emb_dropout = args.dropout
        self.review_encoder_name = args.review_encoder_name
        self.fix_emb = args.fix_emb
        self.grad_mode = grad_mode
        if not args.do_subsample_mask:
            review_words = pad_reviews(review_words, self.word_pad_idx, args.review_word_limit)
        self.review_words = torch.as_tensor(review_words, dtype=torch.int64, device=device)
        self.dropout_layer = nn.Dropout(p=args.dropout)
        if args.use_user_emb:
            self.user_emb = nn.Embedding(user_size + 1, d, padding_idx=self.user_pad_idx)
        if args.use_item_emb:
            self.product_emb = nn.Embedding(product_size + 1, d, padding_idx=self.prod_pad_idx)
        self.word_embeddings = nn.Embedding(vocab_size, d, padding_idx=self.word_pad_idx)
        if self.fix_emb and args.review_encoder_name == "pvc":
            self.review_encoder_name = "pv"                       # ps_model.py:125-128
        self.transformer_encoder = TransformerEncoder(d, args.ff_size, args.heads, args.dropout, args.inter_layers)
        self.word_sink = F_.RowGradSink(self.word_embeddings.weight, self.word_pad_idx, None, grad_mode)
        if self.review_encoder_name == "pv":
            self.review_encoder = ParagraphVector(self.word_embeddings, self.word_dists, review_count,
                                                  self.emb_dropout, None, fix_emb=self.fix_emb,
                                                  word_sink=self.word_sink)
        elif self.review_encoder_name == "pvc":
            self.review_encoder = ParagraphVectorCorruption(self.word_embeddings, self.word_dists,
                                                            args.corrupt_rate, self.emb_dropout, None,
                                                            self.vocab_words, fix_emb=self.fix_emb,
                                                            word_sink=self.word_sink)
        elif self.review_encoder_name == "fs":
            self.review_encoder = FSEncoder(d, self.emb_dropout)
        else:
            self.review_encoder = AVGEncoder(d, self.emb_dropout)
        if args.query_encoder_name == "fs":
            self.query_encoder = FSEncoder(d, self.emb_dropout)
        else:
            self.query_encoder = AVGEncoder(d, self.emb_dropout)
        self.seg_embeddings = nn.Embedding(4, d, padding_idx=self.seg_pad_idx)
        self.review_embeddings = None
        self.initialize_parameters()
        self.to(device)
        self._make_sinks()
        if self.fix_emb:
            self.get_review_embeddings()

    def _make_sinks(self):
        m = self.grad_mode
        self.word_sink.weight = self.word_embeddings.weight
        if hasattr(self.review_encoder, "review_sink"):
            self.review_encoder.review_sink = F_.RowGradSink(self.review_encoder.review_embeddings.weight,
                                                             self.review_pad_idx, None, m)
        self.seg_sink = F_.RowGradSink(self.seg_embeddings.weight, self.seg_pad_idx, None, m)
        self.user_sink = F_.RowGradSink(self.user_emb.weight, self.user_pad_idx, None, m) \
            if self.args.use_user_emb else None
        self.item_sink = F_.RowGradSink(self.product_emb.weight, self.prod_pad_idx, None, m) \
            if self.args.use_item_emb else None

    def load_cp(self, pt, strict=True):
        self.load_state_dict(pt["model"], strict=strict)

    def initialize_parameters(self, logger=None):
        """ps_model.py:360-370."""
        nn.init.normal_(self.word_embeddings.weight)
        nn.init.normal_(self.seg_embeddings.weight)
        self.review_encoder.initialize_parameters(logger)
        self.query_encoder.initialize_parameters(logger)
        self.transformer_encoder.initialize_parameters(logger)

    # ---- review table for evaluation (ps_model.py:177-203) --------------------------------
    def clear_review_embbeddings(self):
        if not self.fix_emb:
            self.review_embeddings = None

    def get_review_embeddings(self, batch_size=128):
        """[R, d] table of review vectors; the reference fills it 128 rows at a time
        (ps_model.py:194-203), here one fused gather + mean (+fs) launch covers all reviews."""
        if self.review_embeddings is not None:
            return
        self.flush_lazy_rows()
        if self.review_encoder_name == "pv":
            # a plain attribute, not a registered alias of the parameter (state_dict stays clean)
            object.__setattr__(self, "review_embeddings", self.review_encoder.review_embeddings.weight)
            return
        with torch.no_grad():
            R = self.review_pad_idx
            table = torch.zeros(R + 1, self.embedding_size, device=self.review_words.device)
            words = self.review_words[:R].contiguous()
            w = self.word_embeddings.weight
            if self.review_encoder_name == "fs":
                enc = self.review_encoder
                vec, _, _ = ops.gather_meanpool(w, words, pad_idx=self.word_pad_idx, fs_weight=enc.f_W.weight,
                                                fs_bias=enc.f_W.bias)
            else:                                                 # pvc (rate 0 in eval) and avg: plain mean
                vec, _, _ = ops.gather_meanpool(w, words, pad_idx=self.word_pad_idx)
            table[:R] = vec
            object.__setattr__(self, "review_embeddings", table)

    # ---- sequence assembly (ps_model.py:316-334 / :221-232) -------------------------------
    def _sequence(self, query_emb, review_emb, seg_idxs, user_idxs, item_idxs):
        seq = torch.cat((query_emb, review_emb), dim=-2)
        if self.args.use_seg_emb:
            seq = seq + F_.gather_rows(self.seg_embeddings.weight, seg_idxs, self.seg_sink)
        if self.args.use_item_emb:
            seq = seq + F_.gather_rows(self.product_emb.weight, item_idxs, self.item_sink)
        if self.args.use_user_emb:
            seq = seq + F_.gather_rows(self.user_emb.weight, user_idxs, self.user_sink)
        return seq

    def test(self, batch_data):
        """ps_model.py:205-239."""
        self.flush_lazy_rows()
        with torch.no_grad():
            cand = batch_data.candi_prod_ridxs
            B, C, Rc = cand.shape
            q = self.query_encoder.encode_indices(self.word_embeddings.weight, batch_data.query_word_idxs,
                                                  self.word_sink, pad_idx=self.word_pad_idx)
            rev = ops.gather_rows(self.review_embeddings, cand)
            mask = torch.cat([torch.ones(B, C, 1, dtype=torch.bool, device=cand.device),
                              cand.ne(self.review_pad_idx)], dim=2)
            seq = self._sequence(q.unsqueeze(1).expand(-1, C, -1).unsqueeze(2), rev, batch_data.candi_seg_idxs,
                                 batch_data.candi_seq_user_idxs, batch_data.candi_seq_item_idxs)
            scores = self.transformer_encoder(seq.reshape(B * C, Rc + 1, -1), mask.reshape(B * C, Rc + 1),
                                              use_pos=self.args.use_pos_emb)
            return scores.view(B, C)

    def forward(self, batch_data, train_pv=True):
        """ps_model.py:241-358."""
        b = batch_data
        K = self.args.neg_per_pos
        name = self.review_encoder_name
        w_table = self.word_embeddings.weight
        B, Rp, Wp = b.pos_prod_rword_idxs.shape
        _, Kn, Rn = b.neg_prod_ridxs.shape
        q_emb = self.query_encoder.encode_indices(w_table, b.query_word_idxs, self.word_sink,
                                                  pad_idx=self.word_pad_idx)
        pos_words = b.pos_prod_rword_idxs.reshape(-1, Wp)
        pos_masks = b.pos_prod_rword_masks.reshape(-1, Wp)
        pv_loss = None
        enc = self.review_encoder
        if "pv" in name:
            if train_pv:
                if name == "pv":
                    pos_rev, pos_loss = enc(b.pos_prod_ridxs.reshape(-1), pos_words, pos_masks, K)
                else:
                    pvc_idx = b.pos_prod_rword_idxs_pvc.reshape(-1, b.pos_prod_rword_idxs_pvc.size(-1))
                    pos_rev, pos_loss = enc(pos_words, pos_masks, pvc_idx, K)
                n_valid = b.pos_prod_ridxs.ne(self.review_pad_idx).float().sum()
                pv_loss = pos_loss.sum() / n_valid
            else:
                if self.fix_emb:
                    pos_rev = ops.gather_rows(self.review_embeddings, b.pos_prod_ridxs)
                elif name == "pv":
                    pos_rev = enc.get_para_vector(b.pos_prod_ridxs)
                else:
                    pos_rev = enc.get_para_vector(pos_words)
            if self.fix_emb:
                neg_rev = ops.gather_rows(self.review_embeddings, b.neg_prod_ridxs)
            elif name == "pv":
                neg_rev = enc.get_para_vector(b.neg_prod_ridxs)
            else:
                neg_idx = b.neg_prod_rword_idxs_pvc if train_pv else b.neg_prod_rword_idxs
                neg_rev = enc.get_para_vector(neg_idx.reshape(-1, neg_idx.size(-1)))
            pos_rev = self.dropout_layer(pos_rev)
            neg_rev = self.dropout_layer(neg_rev)
        else:
            Wn = b.neg_prod_rword_idxs.size(-1)
            pos_rev = enc.encode_indices(w_table, pos_words, self.word_sink, mask=pos_masks)
            neg_rev = enc.encode_indices(w_table, b.neg_prod_rword_idxs.reshape(-1, Wn), self.word_sink,
                                         mask=b.neg_prod_rword_masks.reshape(-1, Wn))
        pos_rev = pos_rev.reshape(B, Rp, -1)
        neg_rev = neg_rev.reshape(B, Kn, Rn, -1)
        dev = pos_rev.device
        pos_mask = torch.cat([torch.ones(B, 1, dtype=torch.bool, device=dev),
                              b.pos_prod_ridxs.ne(self.review_pad_idx)], dim=1)
        neg_ridx_mask = b.neg_prod_ridxs.ne(self.review_pad_idx)
        neg_mask = torch.cat([torch.ones(B, Kn, 1, dtype=torch.bool, device=dev), neg_ridx_mask], dim=2)
        pos_seq = self._sequence(q_emb.unsqueeze(1), pos_rev, b.pos_seg_idxs, b.pos_user_idxs, b.pos_item_idxs)
        neg_seq = self._sequence(q_emb.unsqueeze(1).expand(-1, Kn, -1).unsqueeze(2), neg_rev, b.neg_seg_idxs,
                                 b.neg_user_idxs, b.neg_item_idxs)
        pos_scores = self.transformer_encoder(pos_seq, pos_mask, use_pos=self.args.use_pos_emb)
        neg_scores = self.transformer_encoder(neg_seq.reshape(B * Kn, Rn + 1, -1),
                                              neg_mask.reshape(B * Kn, Rn + 1),
                                              use_pos=self.args.use_pos_emb).view(B, Kn)
        w_pos = float(K) if self.args.pos_weight else 1.0
        weight = torch.cat([torch.full((B, 1), w_pos, device=dev), neg_ridx_mask.sum(-1).ne(0).float()], dim=-1)
        scores = torch.cat([pos_scores.unsqueeze(-1), neg_scores], dim=-1)
        target = torch.cat([torch.ones(B, 1, device=dev), torch.zeros(B, Kn, device=dev)], dim=-1)
        ps_loss = F.binary_cross_entropy_with_logits(scores, target, weight=weight, reduction="none")
        ps_loss = ps_loss.sum(-1).mean()
        return ps_loss + pv_loss if pv_loss is not None else ps_loss

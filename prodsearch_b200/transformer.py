"""Transformer encoder used by TEM / RTM between the gathers and the loss.

Same parameter names as the reference (models/transformer.py, models/neural.py) so a reference
checkpoint loads unchanged: ``pos_emb.pe``, ``transformer_inter.{i}.self_attn.linear_{keys,values,
query}``, ``.final_linear``, ``.feed_forward.{w_1,w_2,layer_norm}``, ``.layer_norm``, ``layer_norm``,
``wo``.  The encoder is row N1 of SURVEY.md 8(f): the last layer + final LayerNorm at the one position
the models read run in the fused sm_100a encoder (``encode_position`` -> functional.SeqEncoderFn ->
psb_encoder_fwd / _bwd; 3xTF32 products on tcgen05 by default, fp32 FFMA kernels as the fallback,
DESIGN.md section 4).  Earlier layers (inter_layers > 1) and the full-sequence ``forward`` the
reference also exposes run the reference's arithmetic through cuBLAS / ATen on the device.
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn


def sinusoid_table(max_len, dim):
    """PositionalEncoding buffer (models/transformer.py:10-18)."""
    pe = torch.zeros(max_len, dim)
    position = torch.arange(0, max_len).unsqueeze(1).float()
    div_term = torch.exp(torch.arange(0, dim, 2, dtype=torch.float) * -(math.log(10000.0) / dim))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


class PositionalEncoding(nn.Module):
    def __init__(self, dropout, dim, max_len=5000):
        super().__init__()
        self.register_buffer("pe", sinusoid_table(max_len, dim))
        self.dim = dim


def gelu(x):
    """tanh-approximated gelu (models/neural.py:7-8)."""
    return 0.5 * x * (1 + torch.tanh(math.sqrt(2 / math.pi) * (x + 0.044715 * x * x * x)))


class MultiHeadedAttention(nn.Module):
    """models/neural.py:36-231 (self-attention path, no layer cache)."""

    def __init__(self, head_count, model_dim, dropout=0.1):
        super().__init__()
        assert model_dim % head_count == 0
        self.dim_per_head = model_dim // head_count
        self.head_count = head_count
        self.linear_keys = nn.Linear(model_dim, model_dim)
        self.linear_values = nn.Linear(model_dim, model_dim)
        self.linear_query = nn.Linear(model_dim, model_dim)
        self.dropout = nn.Dropout(dropout)
        self.final_linear = nn.Linear(model_dim, model_dim)

    def forward(self, key, value, query, mask=None):
        B, H, dh = key.size(0), self.head_count, self.dim_per_head

        def shape(x):
            return x.view(B, -1, H, dh).transpose(1, 2)

        k = shape(self.linear_keys(key))
        v = shape(self.linear_values(value))
        q = shape(self.linear_query(query)) / math.sqrt(dh)
        scores = torch.matmul(q, k.transpose(2, 3))
        if mask is not None:
            scores = scores.masked_fill(mask.unsqueeze(1).expand_as(scores), -1e18)
        attn = self.dropout(torch.softmax(scores, dim=-1))
        ctx = torch.matmul(attn, v).transpose(1, 2).contiguous().view(B, -1, H * dh)
        return self.final_linear(ctx)


class PositionwiseFeedForward(nn.Module):
    """models/neural.py:11-33."""

    def __init__(self, d_model, d_ff, dropout=0.1):
        super().__init__()
        self.w_1 = nn.Linear(d_model, d_ff)
        self.w_2 = nn.Linear(d_ff, d_model)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)
        self.dropout_1 = nn.Dropout(dropout)
        self.dropout_2 = nn.Dropout(dropout)

    def forward(self, x):
        inter = self.dropout_1(gelu(self.w_1(self.layer_norm(x))))
        return self.dropout_2(self.w_2(inter)) + x


class TransformerEncoderLayer(nn.Module):
    """models/transformer.py:37-57 (layer 0 skips the pre-attention LayerNorm)."""

    def __init__(self, d_model, heads, d_ff, dropout):
        super().__init__()
        self.self_attn = MultiHeadedAttention(heads, d_model, dropout=dropout)
        self.feed_forward = PositionwiseFeedForward(d_model, d_ff, dropout)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)
        self.dropout = nn.Dropout(dropout)

    def forward(self, i, inputs, pad_mask):
        h = self.layer_norm(inputs) if i != 0 else inputs
        ctx = self.self_attn(h, h, h, mask=pad_mask.unsqueeze(1))
        return self.feed_forward(self.dropout(ctx) + inputs)


class TransformerEncoder(nn.Module):
    """models/transformer.py:59-119."""

    def __init__(self, d_model, d_ff, heads, dropout, num_inter_layers=0):
        super().__init__()
        self.d_model = d_model
        self.num_inter_layers = num_inter_layers
        self.pos_emb = PositionalEncoding(dropout, d_model)
        self.transformer_inter = nn.ModuleList(
            [TransformerEncoderLayer(d_model, heads, d_ff, dropout) for _ in range(num_inter_layers)])
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)
        self.wo = nn.Linear(d_model, 1, bias=True)

    def encode(self, input_vecs, mask, use_pos=True):
        """input_vecs [S,T,d]; mask [S,T] true/1 at real tokens (transformer.py:71-88)."""
        valid = mask.bool()
        x = input_vecs * valid.unsqueeze(-1).to(input_vecs.dtype)
        if use_pos:
            x = x + self.pos_emb.pe[:, :input_vecs.size(1)]
        pad = ~valid
        for i in range(self.num_inter_layers):
            x = self.transformer_inter[i](i, x, pad)
        return self.layer_norm(x)

    def forward(self, input_vecs, mask, use_pos=True, out_pos=0):
        x = self.encode_position(dense=input_vecs, mask=mask, use_pos=use_pos, out_pos=out_pos)
        return self.wo(x).squeeze(-1)

    # ---- the hot path: only top_vecs[:, out_pos, :] is ever read (item_transformer.py:479,:487,
    #      ps_model.py:336-339), so the last layer runs for that one position in fused sm_100a kernels
    def _last_layer_params(self):
        L = self.transformer_inter[-1]
        a, f = L.self_attn, L.feed_forward
        return OrderedDict([
            ("wq", a.linear_query.weight), ("bq", a.linear_query.bias), ("wk", a.linear_keys.weight),
            ("bk", a.linear_keys.bias), ("wv", a.linear_values.weight), ("bv", a.linear_values.bias),
            ("wo", a.final_linear.weight), ("bo", a.final_linear.bias), ("ln_attn_g", L.layer_norm.weight),
            ("ln_attn_b", L.layer_norm.bias), ("ln_ff_g", f.layer_norm.weight), ("ln_ff_b", f.layer_norm.bias),
            ("w1", f.w_1.weight), ("b1", f.w_1.bias), ("w2", f.w_2.weight), ("b2", f.w_2.bias),
            ("ln_out_g", self.layer_norm.weight), ("ln_out_b", self.layer_norm.bias)])

    def encode_position(self, first=None, table=None, idx=None, sink=None, pad_idx=-1, dense=None, mask=None,
                        use_pos=True, out_pos=0, copies=1, first_ready=None):
        """encode(...)[:, out_pos, :] for ``copies`` independent dropout draws of every sequence ->
        [S*copies, d] (copy-minor).  Tokens: ``first`` [S,d] + rows ``table[idx]`` (idx [S,T-1], invalid =
        ``pad_idx``) or ``dense`` [S,T,d] with ``mask`` [S,T].  Layers before the last (inter_layers > 1)
        need every position and stay on cuBLAS/ATen with the reference's arithmetic; the last layer + the
        final LayerNorm are the fused kernels."""
        from . import functional as F_
        nl = self.num_inter_layers
        p_drop = self.transformer_inter[0].dropout.p if (nl > 0 and self.training) else 0.0
        if nl == 0:
            raise NotImplementedError("inter_layers == 0 (LayerNorm only) is not built")
        if nl > 1:
            if first_ready is not None:       # the ATen layers below read ``first`` on the current stream
                torch.cuda.current_stream(first.device).wait_event(first_ready)
                first_ready = None
            if dense is None:
                rows = F_.gather_rows(table, idx, sink)
                dense = torch.cat([first.unsqueeze(1), rows], dim=1)
                mask = torch.cat([torch.ones_like(idx[:, :1], dtype=torch.bool), idx.ne(pad_idx)], dim=1)
                first = table = idx = sink = None
            valid = mask.bool() if mask is not None else torch.ones(dense.shape[:2], dtype=torch.bool,
                                                                    device=dense.device)
            if copies > 1:
                dense = dense.repeat_interleave(copies, dim=0)
                valid = valid.repeat_interleave(copies, dim=0)
                copies = 1
            x = dense * valid.unsqueeze(-1).to(dense.dtype)
            if use_pos:
                x = x + self.pos_emb.pe[:, :dense.size(1)]
            for i in range(nl - 1):
                x = self.transformer_inter[i](i, x, ~valid)
            # raw layer input: masked rows keep their activations, the mask only removes keys
            T = x.shape[1]
            seed = self._next_seed(x.device) if p_drop > 0 else None
            opts = dict(heads=self.transformer_inter[-1].self_attn.head_count, copies=1, out_pos=out_pos % T,
                        pre_ln=True, eps=1e-6, p_drop=p_drop, seed=seed, raw_input=True)
            return F_.seq_encoder(self._last_layer_params(), opts, dense=x.contiguous(), mask=valid, pe=None)
        T = (1 + idx.shape[1]) if first is not None else dense.shape[1]
        dev = first.device if first is not None else dense.device
        seed = self._next_seed(dev) if p_drop > 0 else None
        opts = dict(heads=self.transformer_inter[-1].self_attn.head_count, copies=copies, out_pos=out_pos % T,
                    pre_ln=False, eps=1e-6, p_drop=p_drop, seed=seed, raw_input=False, first_ready=first_ready)
        pe = self.pos_emb.pe[0, :T] if use_pos else None
        if dense is not None:
            dense = dense.contiguous()
        return F_.seq_encoder(self._last_layer_params(), opts, first=first, table=table, idx=idx, sink=sink,
                              pad_idx=pad_idx, dense=dense, mask=mask, pe=pe)

    def _next_seed(self, device):
        s = getattr(self, "_drop_seed", None)
        if s is None or s.device != torch.empty(0, device=device).device:
            s = self._drop_seed = torch.zeros(1, dtype=torch.int64, device=device)
        s.random_()          # device generator: a fresh key every call, CUDA-graph safe
        return s

    def initialize_parameters(self, logger=None):
        """xavier-normal matrices, zero biases, N(0,1) for LayerNorm gains -- the rule at
        transformer.py:99-119 (its ``else`` branch hits the 1-D LayerNorm weights)."""
        for name, p in self.named_parameters():
            if "weight" in name and p.dim() > 1:
                nn.init.xavier_normal_(p)
            elif "bias" in name:
                nn.init.constant_(p, 0)
            else:
                nn.init.normal_(p)

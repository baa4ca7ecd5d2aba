"""Transformer encoder used by TEM / RTM between the gathers and the loss.

Same parameter names as the reference (models/transformer.py, models/neural.py) so a reference
checkpoint loads unchanged: ``pos_emb.pe``, ``transformer_inter.{i}.self_attn.linear_{keys,values,
query}``, ``.final_linear``, ``.feed_forward.{w_1,w_2,layer_norm}``, ``.layer_norm``, ``layer_norm``,
``wo``.  The encoder is NOT one of the hot-path subsystems (1)-(4) of the north star (SURVEY.md 2.1
C5/C6: "stays torch"); it is row N1 of the "next" list.  This implementation runs the reference's
arithmetic through cuBLAS/ATen on the device and is the piece a fused sm_100a encoder replaces.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


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
        x = self.encode(input_vecs, mask, use_pos)
        return self.wo(x[:, out_pos, :]).squeeze(-1)

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

"""Functional fp32 CPU restatement of the reference's hot-path modules.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference file:line it follows (paths relative to kepingbi/ProdSearch).  Parameters
are passed as a flat dict keyed by the reference's own ``state_dict`` names, so a
reference checkpoint loads without renaming.  RNG never runs inside the oracle:
sampled negatives and the PVC corruption mask are *supplied* by the caller in the
order the reference draws them.

Pinned by tests/golden/*.npz (outputs of the reference's own modules).
"""
import math

import torch
import torch.nn.functional as F

__all__ = [
    "masked_mean", "avg_encoder", "fs_encoder", "query_encoder",
    "bce_with_logits", "gelu_tanh", "multi_head_attention", "encoder_layer",
    "encoder_encode", "encoder_scores", "ns_word_loss", "item_to_words",
    "tem_encode_queries", "tem_forward", "tem_test_scores", "tem_catalog_scores",
    "pv_forward", "pvc_para_vector", "pvc_forward", "rtm_forward", "rtm_test_scores",
    "rtm_review_embeddings", "sinusoid_table",
]


# --------------------------------------------------------------------------
# text encoders  (models/text_encoder.py)
# --------------------------------------------------------------------------
def masked_mean(x, mask):
    """get_vector_mean, models/text_encoder.py:6-16.

    x [N,W,d] fp32, mask [N,W] (bool or uint8).  Sum of masked rows divided by the
    integer count of valid positions, the count clamped to >= 1."""
    m = mask.to(x.dtype).unsqueeze(-1)
    total = (x * m).sum(1)
    count = mask.sum(-1)
    count = torch.where(count == 0, torch.ones_like(count), count).unsqueeze(-1)
    return total / count.to(x.dtype)


def avg_encoder(x, mask, p=0.0, training=False, keep_mul=None):
    """AVGEncoder.forward, models/text_encoder.py:75-82 (mean -> nn.Dropout).  ``keep_mul`` [N,d] supplies the
    dropout multipliers (0 or 1/(1-p)) instead of drawing them (parity runs under dropout)."""
    mean = masked_mean(x, mask)
    return mean * keep_mul if keep_mul is not None else F.dropout(mean, p=p, training=training)


def fs_encoder(x, mask, weight, bias, p=0.0, training=False, keep_mul=None):
    """FSEncoder.forward, models/text_encoder.py:32-40 (mean -> dropout -> tanh(W x + b))."""
    mean = masked_mean(x, mask)
    mean = mean * keep_mul if keep_mul is not None else F.dropout(mean, p=p, training=training)
    return torch.tanh(F.linear(mean, weight, bias))


def query_encoder(P, cfg, word_emb, mask, training=False, keep_mul=None):
    """Dispatch used at models/item_transformer.py:80-83,:450 / models/ps_model.py:153-156,:258."""
    if cfg.query_encoder_name == "fs":
        return fs_encoder(word_emb, mask, P["query_encoder.f_W.weight"],
                          P["query_encoder.f_W.bias"], cfg.dropout, training, keep_mul)
    return avg_encoder(word_emb, mask, cfg.dropout, training, keep_mul)


# --------------------------------------------------------------------------
# loss
# --------------------------------------------------------------------------
class _BCEWithLogits(torch.autograd.Function):
    """Value max(x,0) - x*t + log1p(exp(-|x|)); gradient the analytic (sigmoid(x) - t)
    that ATen's binary_cross_entropy_with_logits_backward uses.  Written as an explicit
    Function because differentiating the stable value formula through clamp/abs gives the
    wrong sub-gradient at exactly x == 0 (which happens: a fully corrupted PVC review
    scores 0 against every word)."""

    @staticmethod
    def forward(ctx, x, t):
        ctx.save_for_backward(x, t)
        return torch.clamp(x, min=0) - x * t + torch.log1p(torch.exp(-x.abs()))

    @staticmethod
    def backward(ctx, g):
        x, t = ctx.saved_tensors
        return g * (torch.sigmoid(x) - t), None


def bce_with_logits(x, t, w=None):
    """binary_cross_entropy_with_logits(reduction='none'), the form SURVEY.md 8(a) A4
    quotes: w * [max(x,0) - x*t + log1p(exp(-|x|))]; call sites
    models/item_transformer.py:280,:510-514, models/PV.py:64, models/PVC.py:90,
    models/ps_model.py:351-354."""
    loss = _BCEWithLogits.apply(x, t)
    return loss if w is None else loss * w


# --------------------------------------------------------------------------
# transformer encoder  (models/transformer.py, models/neural.py)
# --------------------------------------------------------------------------
def sinusoid_table(max_len, dim):
    """PositionalEncoding.__init__, models/transformer.py:10-18."""
    pe = torch.zeros(max_len, dim)
    pos = torch.arange(0, max_len).unsqueeze(1).float()
    div = torch.exp(torch.arange(0, dim, 2, dtype=torch.float) * -(math.log(10000.0) / dim))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0)


def gelu_tanh(x):
    """gelu, models/neural.py:7-8 (tanh approximation)."""
    return 0.5 * x * (1 + torch.tanh(math.sqrt(2 / math.pi) * (x + 0.044715 * torch.pow(x, 3))))


def _drop(x, p, training, mul=None):
    """nn.Dropout; ``mul`` supplies the keep/scale multipliers instead of drawing them (parity runs)."""
    return x * mul if mul is not None else F.dropout(x, p=p, training=training)


def multi_head_attention(P, pre, key, value, query, pad_mask, heads, p=0.0, training=False, attn_mul=None):
    """MultiHeadedAttention.forward, models/neural.py:98-231 (layer_cache=None path).

    pad_mask: bool [B,Tq or 1,Tk], True where the key is padding (filled with -1e18
    at :214).  query is scaled by 1/sqrt(dh) before the product (:206)."""
    B, Tk, D = key.shape
    dh = D // heads

    def split(x):
        return x.view(B, -1, heads, dh).transpose(1, 2)

    k = split(F.linear(key, P[pre + "linear_keys.weight"], P[pre + "linear_keys.bias"]))
    v = split(F.linear(value, P[pre + "linear_values.weight"], P[pre + "linear_values.bias"]))
    q = split(F.linear(query, P[pre + "linear_query.weight"], P[pre + "linear_query.bias"]))
    q = q / math.sqrt(dh)
    scores = torch.matmul(q, k.transpose(2, 3))
    if pad_mask is not None:
        scores = scores.masked_fill(pad_mask.unsqueeze(1).expand_as(scores), -1e18)
    attn = _drop(torch.softmax(scores, dim=-1), p, training, attn_mul)
    ctx = torch.matmul(attn, v).transpose(1, 2).contiguous().view(B, -1, heads * dh)
    return F.linear(ctx, P[pre + "final_linear.weight"], P[pre + "final_linear.bias"])


def encoder_layer(P, pre, i, x, pad_mask, heads, p=0.0, training=False, muls=None):
    """TransformerEncoderLayer.forward, models/transformer.py:47-57 and
    PositionwiseFeedForward.forward, models/neural.py:30-33.  Layer 0 skips the
    pre-attention LayerNorm (:48-51).  ``muls`` = dict(attn [S,H,T,T], ctx [S,T,d], inner [S,T,ff],
    out [S,T,d]) replaces the four dropout draws by supplied multipliers."""
    d = x.shape[-1]
    m = muls or {}
    h = x if i == 0 else F.layer_norm(x, (d,), P[pre + "layer_norm.weight"],
                                      P[pre + "layer_norm.bias"], 1e-6)
    ctx = multi_head_attention(P, pre + "self_attn.", h, h, h, pad_mask.unsqueeze(1),
                               heads, p, training, m.get("attn"))
    out = _drop(ctx, p, training, m.get("ctx")) + x
    ff = pre + "feed_forward."
    n = F.layer_norm(out, (d,), P[ff + "layer_norm.weight"], P[ff + "layer_norm.bias"], 1e-6)
    inter = _drop(gelu_tanh(F.linear(n, P[ff + "w_1.weight"], P[ff + "w_1.bias"])), p, training, m.get("inner"))
    y = _drop(F.linear(inter, P[ff + "w_2.weight"], P[ff + "w_2.bias"]), p, training, m.get("out"))
    return y + out


def encoder_encode(P, cfg, x, valid_mask, use_pos=True, pre="transformer_encoder.",
                   training=False, last_layer_muls=None):
    """TransformerEncoder.encode, models/transformer.py:71-88.

    x [S,T,d]; valid_mask [S,T] (1 = real token).  Input rows are zeroed where
    invalid (:76), the sinusoid table is added un-scaled (:78-80), each layer gets
    the inverted mask (:83-84), final LayerNorm eps=1e-6 (:86)."""
    T, d = x.shape[1], x.shape[2]
    valid = valid_mask.bool()
    h = x * valid.unsqueeze(-1).to(x.dtype)
    if use_pos:
        h = h + P[pre + "pos_emb.pe"][:, :T]
    for i in range(cfg.inter_layers):
        h = encoder_layer(P, "%stransformer_inter.%d." % (pre, i), i, h, ~valid, cfg.heads,
                          cfg.dropout, training,
                          last_layer_muls if i == cfg.inter_layers - 1 else None)
    return F.layer_norm(h, (d,), P[pre + "layer_norm.weight"], P[pre + "layer_norm.bias"], 1e-6)


def encoder_scores(P, cfg, x, valid_mask, use_pos=True, out_pos=0,
                   pre="transformer_encoder.", training=False):
    """TransformerEncoder.forward, models/transformer.py:90-97 (wo: d -> 1 head)."""
    top = encoder_encode(P, cfg, x, valid_mask, use_pos, pre, training)[:, out_pos, :]
    return F.linear(top, P[pre + "wo.weight"], P[pre + "wo.bias"]).squeeze(-1)


# --------------------------------------------------------------------------
# negative-sampling word objective (shared by item_to_words / PV / PVC)
# --------------------------------------------------------------------------
def ns_word_loss(anchor, word_table, pos_word_idxs, neg_word_idxs, word_mask, n_negs,
                 word_bias=None):
    """The block repeated at models/item_transformer.py:263-281, models/PV.py:57-65,
    models/PVC.py:83-91.

    anchor [N,d]; pos_word_idxs [N,W]; neg_word_idxs flat [N*W*K] in the order
    torch.multinomial returned them (viewed [N, W*K] then [N,W,K]); word_mask [N,W].
    Returns per-anchor loss [N,1] = masked mean over W of sum_j BCE(score_j, t_j)."""
    N, W = pos_word_idxs.shape
    pos_emb = word_table[pos_word_idxs]                                  # [N,W,d]
    neg_emb = word_table[neg_word_idxs.view(N, -1)]                      # [N,W*K,d]
    out_pos = torch.bmm(pos_emb, anchor.unsqueeze(2))                    # [N,W,1]
    out_neg = torch.bmm(neg_emb, anchor.unsqueeze(2)).view(N, W, -1)     # [N,W,K]
    if word_bias is not None:
        out_pos = out_pos + word_bias[pos_word_idxs.reshape(-1)].view(N, W, 1)
        out_neg = out_neg + word_bias[neg_word_idxs].view(N, W, -1)
    scores = torch.cat((out_pos, out_neg), dim=-1)
    target = torch.cat((torch.ones_like(out_pos), torch.zeros_like(out_neg)), dim=-1)
    loss = bce_with_logits(scores, target).sum(-1)                       # [N,W]
    return masked_mean(loss.unsqueeze(-1), word_mask)                    # [N,1]


def item_to_words(P, cfg, target_prod_idxs, target_word_idxs, neg_word_idxs, n_negs):
    """ItemTransformerRanker.item_to_words, models/item_transformer.py:260-283."""
    V = P["word_embeddings.weight"].shape[0]
    anchor = P["product_emb.weight"][target_prod_idxs]
    loss = ns_word_loss(anchor, P["word_embeddings.weight"], target_word_idxs, neg_word_idxs,
                        target_word_idxs.ne(V - 1), n_negs, P["word_bias"])
    return loss.mean()


# --------------------------------------------------------------------------
# TEM  (models/item_transformer.py)
# --------------------------------------------------------------------------
def tem_encode_queries(P, cfg, query_word_idxs, u_item_idxs, training=False, copies=1, q_emb=None, muls=None):
    """Shared front half of forward_dotproduct (:449-484) and test_dotproduct
    (:118-140): query encoder + history gather + transformer encode -> [B*copies,d].

    ``copies`` replays the reference's expansion of the same sequence for every
    negative / candidate (:473-476, :130-133).  ``q_emb``: the query vector computed once by the caller
    (the reference encodes the query once, :450, and feeds it to both encoder calls); ``muls``: supplied
    dropout multipliers of the (last) encoder layer for the B*copies sequences (encoder_layer)."""
    V = P["word_embeddings.weight"].shape[0]
    Pp1 = P["product_emb.weight"].shape[0]
    B, L = u_item_idxs.shape
    if q_emb is None:
        q_emb = query_encoder(P, cfg, P["word_embeddings.weight"][query_word_idxs],
                              query_word_idxs.ne(V - 1), training)
    hist_table = P["hist_product_emb.weight"] if cfg.sep_prod_emb else P["product_emb.weight"]
    u_emb = hist_table[u_item_idxs]
    seq = torch.cat([q_emb.unsqueeze(1), u_emb], dim=1)                   # [B,1+L,d]
    mask = torch.cat([torch.ones(B, 1, dtype=torch.bool), u_item_idxs.ne(Pp1 - 1)], dim=1)
    if copies > 1:
        seq = seq.unsqueeze(1).expand(-1, copies, -1, -1).reshape(B * copies, 1 + L, -1)
        mask = mask.unsqueeze(1).expand(-1, copies, -1).reshape(B * copies, 1 + L)
    out_pos = -1 if cfg.use_item_pos else 0
    top = encoder_encode(P, cfg, seq, mask, cfg.use_pos_emb, training=training, last_layer_muls=muls)
    return top[:, out_pos, :]


def tem_forward(P, cfg, query_word_idxs, target_prod_idxs, u_item_idxs, pos_iword_idxs,
                neg_item_idxs, neg_word_idxs, training=False, drop=None):
    """ItemTransformerRanker.forward_dotproduct, models/item_transformer.py:440-520.

    neg_item_idxs [B,K] replaces the multinomial at :447, neg_word_idxs [B*W*K] the
    one at :268.  ``drop`` (parity under dropout > 0): dict(query_keep [B,d], enc_pos, enc_neg) -- the
    multipliers of the query encoder's dropout (drawn ONCE, :450) and of the two encoder calls (:479, :482-486;
    dicts as encoder_layer takes them, for B and B*K sequences).  Returns (loss, ps_loss, item_loss)."""
    B = target_prod_idxs.shape[0]
    K = cfg.neg_per_pos
    E = P["product_emb.weight"]
    if drop is not None:
        V = P["word_embeddings.weight"].shape[0]
        q_emb = query_encoder(P, cfg, P["word_embeddings.weight"][query_word_idxs], query_word_idxs.ne(V - 1),
                              True, drop["query_keep"])
        pos_out = tem_encode_queries(P, cfg, query_word_idxs, u_item_idxs, True, q_emb=q_emb, muls=drop["enc_pos"])
        neg_out = tem_encode_queries(P, cfg, query_word_idxs, u_item_idxs, True, copies=K, q_emb=q_emb,
                                     muls=drop["enc_neg"])
    else:
        pos_out = tem_encode_queries(P, cfg, query_word_idxs, u_item_idxs, training)          # [B,d]
        neg_out = tem_encode_queries(P, cfg, query_word_idxs, u_item_idxs, training, copies=K)  # [B*K,d]
    pos_scores = torch.bmm(pos_out.unsqueeze(1), E[target_prod_idxs].unsqueeze(2)).view(B)
    neg_scores = torch.bmm(neg_out.unsqueeze(1), E[neg_item_idxs].view(B * K, -1).unsqueeze(2)).view(B, K)
    if cfg.sim_func == "bias_product":
        pos_scores = pos_scores + P["product_bias"][target_prod_idxs]
        neg_scores = neg_scores + P["product_bias"][neg_item_idxs.reshape(-1)].view(B, K)
    w = torch.ones(B, 1 + K)
    if cfg.pos_weight:
        w[:, 0] = K
    scores = torch.cat([pos_scores.unsqueeze(-1), neg_scores], dim=-1)
    target = torch.cat([torch.ones(B, 1), torch.zeros(B, K)], dim=-1)
    ps_loss = bce_with_logits(scores, target, w).sum(-1).mean()
    item_loss = item_to_words(P, cfg, target_prod_idxs, pos_iword_idxs, neg_word_idxs, K)
    return ps_loss + item_loss, ps_loss, item_loss


def tem_test_scores(P, cfg, query_word_idxs, u_item_idxs, candi_prod_idxs):
    """ItemTransformerRanker.test_dotproduct, models/item_transformer.py:111-146, with
    the reference's per-candidate re-encode kept (copies=candi_k)."""
    B, C = candi_prod_idxs.shape
    out = tem_encode_queries(P, cfg, query_word_idxs, u_item_idxs, False, copies=C)  # [B*C,d]
    cand = P["product_emb.weight"][candi_prod_idxs].view(B * C, -1)
    scores = torch.bmm(out.unsqueeze(1), cand.unsqueeze(2)).view(B, C)
    if cfg.sim_func == "bias_product":
        scores = scores + P["product_bias"][candi_prod_idxs.reshape(-1)].view(B, C)
    return scores


def tem_catalog_scores(P, cfg, query_word_idxs, u_item_idxs, n_items=None):
    """Single-GEMM restatement of full-catalog scoring: the encoder
    output does not depend on the candidate, so S = Qout . E^T (+bias).  Must agree
    with tem_test_scores; checked in tests/test_oracle_golden.py."""
    q = tem_encode_queries(P, cfg, query_word_idxs, u_item_idxs, False)
    E = P["product_emb.weight"]
    n = E.shape[0] - 1 if n_items is None else n_items
    scores = q @ E[:n].t()
    if cfg.sim_func == "bias_product":
        scores = scores + P["product_bias"][:n]
    return q, scores


# --------------------------------------------------------------------------
# PV / PVC  (models/PV.py, models/PVC.py)
# --------------------------------------------------------------------------
def pv_forward(review_table, word_table, review_ids, pos_word_idxs, word_mask,
               neg_word_idxs, n_negs, p=0.0, training=False):
    """ParagraphVector.forward, models/PV.py:50-80.  Returns (review_emb [N,d], loss [N,1])."""
    review_emb = F.dropout(review_table[review_ids], p=p, training=training)
    loss = ns_word_loss(review_emb, word_table, pos_word_idxs, neg_word_idxs, word_mask, n_negs)
    return review_emb, loss


def pvc_para_vector(context_table, word_idxs, pad_idx, corrupt_mask=None, corrupt_rate=0.0):
    """ParagraphVectorCorruption.get_para_vector, models/PVC.py:56-61 (+ apply_token_dropout
    :46-54).  corrupt_mask [N,Wr] (1 = token dropped) replaces the bernoulli draw at :51.
    The corruption edits ``.data`` in place, so autograd sees an ordinary masked mean of
    the *un-corrupted* gather; reproduced with a detached delta."""
    emb = context_table[word_idxs]
    if corrupt_rate > 0.0 and corrupt_mask is not None:
        keep = (~corrupt_mask.bool()).to(emb.dtype).unsqueeze(-1)
        corrupted = (emb.detach() * keep) * (1.0 / (1 - corrupt_rate))
        emb = corrupted + (emb - emb.detach())  # value = corrupted exactly, d/d emb = 1
    return masked_mean(emb, word_idxs.ne(pad_idx))


def pvc_forward(context_table, word_table, pos_word_idxs, word_mask, rword_idxs_pvc,
                neg_word_idxs, n_negs, corrupt_mask, corrupt_rate):
    """ParagraphVectorCorruption.forward, models/PVC.py:69-95.  Returns the UNcorrupted
    review_emb (:77,:95) and the loss scored against the corrupted mean (:79,:85-86)."""
    pad_idx = word_table.shape[0] - 1
    review_emb = masked_mean(context_table[rword_idxs_pvc], rword_idxs_pvc.ne(pad_idx))
    # apply_token_dropout is called unconditionally in forward (:78), rate 0 keeps all
    corr = pvc_para_vector(context_table, rword_idxs_pvc, pad_idx,
                           corrupt_mask if corrupt_rate > 0 else None, corrupt_rate)
    loss = ns_word_loss(corr, word_table, pos_word_idxs, neg_word_idxs, word_mask, n_negs)
    return review_emb, loss


# --------------------------------------------------------------------------
# RTM  (models/ps_model.py)
# --------------------------------------------------------------------------
def rtm_review_embeddings(P, cfg, review_words):
    """ProductRanker.get_review_embeddings, models/ps_model.py:184-203 (pvc eval mode has
    corrupt_rate 0; fs/avg run the review encoder in eval mode).  Last row stays zero."""
    V = P["word_embeddings.weight"].shape[0]
    if cfg.review_encoder_name == "pv":
        return P["review_encoder.review_embeddings.weight"]
    R = review_words.shape[0]
    table = torch.zeros(R, cfg.embedding_size)
    words = review_words[: R - 1]
    emb = P["word_embeddings.weight"][words]
    mask = words.ne(V - 1)
    if cfg.review_encoder_name == "pvc":
        vec = masked_mean(emb, mask)
    elif cfg.review_encoder_name == "fs":
        vec = fs_encoder(emb, mask, P["review_encoder.f_W.weight"], P["review_encoder.f_W.bias"])
    else:
        vec = avg_encoder(emb, mask)
    table[: R - 1] = vec
    return table


def _rtm_sequence(P, cfg, query_emb, review_emb, seg_idxs, user_idxs, item_idxs):
    """Sequence assembly shared by forward (:316-334) and test (:221-232)."""
    seq = torch.cat((query_emb, review_emb), dim=-2)
    if cfg.use_seg_emb:
        seq = seq + P["seg_embeddings.weight"][seg_idxs]
    if cfg.use_item_emb:
        seq = seq + P["product_emb.weight"][item_idxs]
    if cfg.use_user_emb:
        seq = seq + P["user_emb.weight"][user_idxs]
    return seq


def rtm_forward(P, cfg, batch, train_pv, neg_word_idxs=None, corrupt_masks=None, training=False):
    """ProductRanker.forward, models/ps_model.py:241-358.

    ``batch`` is any object with the ProdSearchTrainBatch fields (data/batch_data.py:137).
    neg_word_idxs: the multinomial draw inside PV/PVC forward (only when train_pv).
    corrupt_masks: list of bernoulli masks in call order (pvc: pos first when train_pv
    [PVC.py:78], or pos get_para_vector [ps_model.py:286]; then neg [ps_model.py:297])."""
    V = P["word_embeddings.weight"].shape[0]
    E = P["word_embeddings.weight"]
    name = cfg.review_encoder_name
    K = cfg.neg_per_pos
    B, Rp, Wp = batch.pos_prod_rword_idxs.shape
    _, Kn, Rn = batch.neg_prod_ridxs.shape
    R = None
    q_emb = query_encoder(P, cfg, E[batch.query_word_idxs], batch.query_word_idxs.ne(V - 1), training)
    pos_words = batch.pos_prod_rword_idxs.view(-1, Wp)
    pos_masks = batch.pos_prod_rword_masks.view(-1, Wp)
    pv_loss = None
    masks = list(corrupt_masks or [])
    if "pv" in name:
        if name == "pv":
            R = P["review_encoder.review_embeddings.weight"]
            review_pad = R.shape[0] - 1
        else:
            review_pad = cfg.review_pad_idx
        if train_pv:
            if name == "pv":
                pos_rev, pos_loss = pv_forward(R, E, batch.pos_prod_ridxs.view(-1), pos_words, pos_masks,
                                               neg_word_idxs, K, cfg.dropout, training)
            else:
                pvc_idx = batch.pos_prod_rword_idxs_pvc.view(-1, batch.pos_prod_rword_idxs_pvc.size(-1))
                pos_rev, pos_loss = pvc_forward(E, E, pos_words, pos_masks, pvc_idx, neg_word_idxs, K,
                                                masks.pop(0) if masks else None, cfg.corrupt_rate)
            n_valid = batch.pos_prod_ridxs.ne(review_pad).float().sum(-1)
            pv_loss = pos_loss.sum() / n_valid.sum()
        else:
            if name == "pv":
                pos_rev = R[batch.pos_prod_ridxs]
            else:
                pos_rev = pvc_para_vector(E, pos_words, V - 1, masks.pop(0) if masks else None,
                                          cfg.corrupt_rate)
        if name == "pv":
            neg_rev = R[batch.neg_prod_ridxs]
        else:
            neg_idx = batch.neg_prod_rword_idxs_pvc if train_pv else batch.neg_prod_rword_idxs
            neg_rev = pvc_para_vector(E, neg_idx.view(-1, neg_idx.size(-1)), V - 1,
                                      masks.pop(0) if masks else None, cfg.corrupt_rate)
        pos_rev = F.dropout(pos_rev, p=cfg.dropout, training=training)
        neg_rev = F.dropout(neg_rev, p=cfg.dropout, training=training)
    else:
        review_pad = cfg.review_pad_idx
        Wn = batch.neg_prod_rword_idxs.size(-1)
        neg_words = batch.neg_prod_rword_idxs.view(-1, Wn)
        neg_masks = batch.neg_prod_rword_masks.view(-1, Wn)
        if name == "fs":
            w, b = P["review_encoder.f_W.weight"], P["review_encoder.f_W.bias"]
            pos_rev = fs_encoder(E[pos_words], pos_masks, w, b, cfg.dropout, training)
            neg_rev = fs_encoder(E[neg_words], neg_masks, w, b, cfg.dropout, training)
        else:
            pos_rev = avg_encoder(E[pos_words], pos_masks, cfg.dropout, training)
            neg_rev = avg_encoder(E[neg_words], neg_masks, cfg.dropout, training)
    pos_rev = pos_rev.view(B, Rp, -1)
    neg_rev = neg_rev.view(B, Kn, Rn, -1)

    pos_mask = torch.cat([torch.ones(B, 1, dtype=torch.bool), batch.pos_prod_ridxs.ne(review_pad)], dim=1)
    neg_ridx_mask = batch.neg_prod_ridxs.ne(review_pad)
    neg_mask = torch.cat([torch.ones(B, Kn, 1, dtype=torch.bool), neg_ridx_mask], dim=2)
    pos_seq = _rtm_sequence(P, cfg, q_emb.unsqueeze(1), pos_rev, batch.pos_seg_idxs,
                            batch.pos_user_idxs, batch.pos_item_idxs)
    neg_seq = _rtm_sequence(P, cfg, q_emb.unsqueeze(1).expand(-1, Kn, -1).unsqueeze(2), neg_rev,
                            batch.neg_seg_idxs, batch.neg_user_idxs, batch.neg_item_idxs)
    pos_scores = encoder_scores(P, cfg, pos_seq, pos_mask, cfg.use_pos_emb, training=training)
    neg_scores = encoder_scores(P, cfg, neg_seq.reshape(B * Kn, Rn + 1, -1),
                                neg_mask.reshape(B * Kn, Rn + 1), cfg.use_pos_emb,
                                training=training).view(B, Kn)
    w_pos = float(K) if cfg.pos_weight else 1.0
    w = torch.cat([torch.full((B, 1), w_pos), neg_ridx_mask.sum(-1).ne(0).float()], dim=-1)
    scores = torch.cat([pos_scores.unsqueeze(-1), neg_scores], dim=-1)
    target = torch.cat([torch.ones(B, 1), torch.zeros(B, Kn)], dim=-1)
    ps_loss = bce_with_logits(scores, target, w).sum(-1).mean()
    return (ps_loss + pv_loss if pv_loss is not None else ps_loss), ps_loss, pv_loss


def rtm_test_scores(P, cfg, review_table, batch):
    """ProductRanker.test, models/ps_model.py:205-239 (batch: ProdSearchTestBatch fields)."""
    V = P["word_embeddings.weight"].shape[0]
    review_pad = review_table.shape[0] - 1
    B, C, Rc = batch.candi_prod_ridxs.shape
    q_emb = query_encoder(P, cfg, P["word_embeddings.weight"][batch.query_word_idxs],
                          batch.query_word_idxs.ne(V - 1), False)
    rev = review_table[batch.candi_prod_ridxs]
    mask = torch.cat([torch.ones(B, C, 1, dtype=torch.bool),
                      batch.candi_prod_ridxs.ne(review_pad)], dim=2)
    seq = _rtm_sequence(P, cfg, q_emb.unsqueeze(1).expand(-1, C, -1).unsqueeze(2), rev,
                        batch.candi_seg_idxs, batch.candi_seq_user_idxs, batch.candi_seq_item_idxs)
    scores = encoder_scores(P, cfg, seq.reshape(B * C, Rc + 1, -1), mask.reshape(B * C, Rc + 1),
                            cfg.use_pos_emb)
    return scores.view(B, C)

"""Thin Python mirrors of the C-ABI entry points (include/psb.h), on torch CUDA tensors.

No autograd here and no fallback: these are the operator-level calls; the nn.Module
surface (autograd.Functions, gradient sink) is in ``functional.py`` / the model files.
"""
import ctypes

import torch

from . import _lib
from ._lib import Contrib, check, load, ptr, stream_ptr

f32, i64, i32, u8 = torch.float32, torch.int64, torch.int32, torch.uint8


def _idx(t):
    if t.dtype != i64:
        t = t.long()
    return t.contiguous()


def _on(stream, dev, *inputs):
    """Context that enqueues on ``stream`` (a side stream the CALLER forked from and will join into the current
    one) or, when None, on the current stream.  Output buffers are always allocated before entering it, i.e. on
    the current stream, so the caching allocator's stream bookkeeping stays valid.  ``inputs``: tensors the side
    kernel reads; temporaries among them (a converted index / mask copy) would otherwise be freed -- and their
    memory reused by the current stream -- as soon as the Python call returns, so the last few calls' inputs stay
    referenced on the stream object (the caller joins the stream long before the list wraps around)."""
    if stream is None:
        return torch.cuda.stream(torch.cuda.current_stream(dev))
    keep = getattr(stream, "_psb_keep", None)
    if keep is None:
        keep = stream._psb_keep = []
    # detached aliases: a tensor with a grad_fn would keep its autograd graph -- and the AccumulateGrad nodes of the
    # parameters behind it, bound to the stream of THAT pass -- alive into the next step (breaks a later graph capture)
    keep.append(tuple(t.detach() for t in inputs if t is not None))
    del keep[:-8]
    return torch.cuda.stream(stream)


def gather_rows(table, idx, err_flag=None, stream=None):
    """out[..., :] = table[idx[...], :]   (psb_gather_rows; aten::embedding forward)."""
    idx = _idx(idx)
    n, d = idx.numel(), table.shape[1]
    out = torch.empty(idx.shape + (d,), dtype=f32, device=table.device)
    with _on(stream, table.device, table, idx, out):
        check(load().psb_gather_rows(ptr(table, f32), table.shape[0], d, ptr(idx), n, ptr(out), ptr(err_flag),
                                     stream_ptr()), "psb_gather_rows")
    return out


def gather_meanpool(table, idx, pad_idx=-1, mask=None, tok_scale=None, keep_scale=None, fs_weight=None,
                    fs_bias=None, want_mean=False, want_inv_count=False, stream=None):
    """Fused gather + masked mean (+dropout multiplier, +fs projection); psb_gather_meanpool_fwd.
    idx [n, w].  Returns (out [n,d], mean or None, inv_count or None).
    stream: enqueue the kernel on this side stream (forked from the current one; the outputs are allocated on the
    current stream).  The CALLER orders later readers after it (an event recorded on ``stream``)."""
    idx = _idx(idx)
    n, w = idx.shape
    d = table.shape[1]
    dev = table.device
    out = torch.empty((n, d), dtype=f32, device=dev)
    need_mean = want_mean or fs_weight is not None
    mean = torch.empty((n, d), dtype=f32, device=dev) if need_mean else None
    inv = torch.empty((n,), dtype=f32, device=dev) if want_inv_count else None
    if mask is not None and mask.dtype != u8:
        mask = mask.to(u8)
    mask_c = mask.contiguous() if mask is not None else None
    if stream is not None:
        stream.wait_stream(torch.cuda.current_stream(dev))
    with _on(stream, dev, table, idx, mask_c, tok_scale, keep_scale, fs_weight, fs_bias, mean, out, inv):
        check(load().psb_gather_meanpool_fwd(
            ptr(table, f32), table.shape[0], d, ptr(idx), n, w, int(pad_idx), ptr(mask_c), ptr(tok_scale, f32),
            ptr(keep_scale, f32), ptr(fs_weight, f32), ptr(fs_bias, f32), ptr(mean), ptr(out), ptr(inv),
            stream_ptr()), "psb_gather_meanpool_fwd")
    return out, mean, inv


def fs_bwd(grad_out, out, mean, keep_scale, fs_weight):
    n, d = out.shape
    dev = out.device
    gw = torch.empty((d, d), dtype=f32, device=dev)
    gb = torch.empty((d,), dtype=f32, device=dev)
    gm = torch.empty((n, d), dtype=f32, device=dev)
    check(load().psb_fs_bwd(ptr(grad_out.contiguous(), f32), ptr(out, f32), ptr(mean, f32), ptr(keep_scale, f32),
                            ptr(fs_weight, f32), n, d, ptr(gw), ptr(gb), ptr(gm), stream_ptr()), "psb_fs_bwd")
    return gw, gb, gm


def token_weights(idx, pad_idx=-1, mask=None):
    """valid / max(#valid, 1) per token, the backward weights of a masked mean."""
    idx = _idx(idx)
    n, w = idx.shape
    tw = torch.empty((n, w), dtype=f32, device=idx.device)
    if mask is not None and mask.dtype != u8:
        mask = mask.to(u8)
    check(load().psb_meanpool_token_weights(ptr(idx), n, w, int(pad_idx),
                                            ptr(mask.contiguous() if mask is not None else None), ptr(tw),
                                            stream_ptr()), "psb_meanpool_token_weights")
    return tw


def ns_loss(anchor_a, table, pos_idx, neg_idx, anchor_b=None, bias=None, mask=None, pad_idx=-1,
            neg_weight=None, pos_weight=1.0, stream=None):
    """Fused negative-sampling loss forward + analytic score gradient (psb_ns_loss_fwd).
    pos_idx [n,w], neg_idx [n,w,k].  Returns loss [n], coef_pos [n,w], coef_neg [n,w,k],
    grad_anchor_a [n,d], grad_anchor_b [n*k,d] or None."""
    pos_idx = _idx(pos_idx)
    neg_idx = _idx(neg_idx)
    n, w = pos_idx.shape
    k = neg_idx.numel() // max(n * w, 1)
    d = table.shape[1]
    dev = table.device
    loss = torch.empty((n,), dtype=f32, device=dev)
    cp = torch.empty((n, w), dtype=f32, device=dev)
    cn = torch.empty((n, w, k), dtype=f32, device=dev)
    ga = torch.empty((n, d), dtype=f32, device=dev)
    gb = torch.empty((n * k, d), dtype=f32, device=dev) if anchor_b is not None else None
    if mask is not None and mask.dtype != u8:
        mask = mask.to(u8)
    mask_c = mask.contiguous() if mask is not None else None
    with _on(stream, dev, anchor_a, anchor_b, table, bias, pos_idx, neg_idx, mask_c, neg_weight, loss, cp, cn, ga, gb):
        check(load().psb_ns_loss_fwd(
            ptr(anchor_a, f32), ptr(anchor_b, f32), ptr(table, f32), table.shape[0], d, ptr(bias, f32), ptr(pos_idx),
            ptr(neg_idx), ptr(mask_c), int(pad_idx), ptr(neg_weight, f32), float(pos_weight), n, w, k, ptr(loss),
            ptr(cp), ptr(cn), ptr(ga), ptr(gb), stream_ptr()), "psb_ns_loss_fwd")
    return loss, cp, cn, ga, gb


def tem_loss(enc_out, table, pos_idx, neg_idx, bias=None, pos_weight=1.0, grad_scale=1.0):
    """TEM score + loss tail on the encoder's output block enc_out [n, 1 + k, d] (psb_tem_loss_fwd).  Returns
    (loss_rows [n], coef_pos [n], coef_neg [n, k], grad_enc_out [n, 1 + k, d]); coefficients and gradient are
    multiplied by grad_scale."""
    pos_idx, neg_idx = _idx(pos_idx), _idx(neg_idx)
    n = pos_idx.numel()
    k = neg_idx.numel() // max(n, 1)
    d = table.shape[1]
    dev = table.device
    rows = torch.empty((n,), dtype=f32, device=dev)
    cp = torch.empty((n,), dtype=f32, device=dev)
    cn = torch.empty((n, k), dtype=f32, device=dev)
    g = torch.empty((n, 1 + k, d), dtype=f32, device=dev)
    check(load().psb_tem_loss_fwd(ptr(enc_out, f32), ptr(table, f32), table.shape[0], d, ptr(bias, f32), ptr(pos_idx),
                                  ptr(neg_idx), float(pos_weight), n, k, float(grad_scale), ptr(rows), ptr(cp), ptr(cn),
                                  ptr(g), stream_ptr()), "psb_tem_loss_fwd")
    return rows, cp, cn, g


def tem_loss_finish(ps_rows, il_rows, acc_ps=None, acc_il=None):
    """0-dim loss = mean(ps_rows) + mean(il_rows); the running sums are advanced in the same launch."""
    out = torch.empty((), dtype=f32, device=ps_rows.device)
    check(load().psb_tem_loss_finish(ptr(ps_rows, f32), ptr(il_rows, f32), ps_rows.numel(),
                                     il_rows.numel() if il_rows is not None else 0, ptr(out), ptr(acc_ps, f32),
                                     ptr(acc_il, f32), stream_ptr()), "psb_tem_loss_finish")
    return out


def score_rows(anchor, table, idx, bias=None):
    """scores[i,c] = <anchor[i], table[idx[i,c]]> (+bias) for an explicit candidate list (psb_score_rows)."""
    idx = _idx(idx)
    n, c = idx.shape
    out = torch.empty((n, c), dtype=f32, device=table.device)
    check(load().psb_score_rows(ptr(anchor, f32), ptr(table, f32), table.shape[0], table.shape[1], ptr(bias, f32),
                                ptr(idx), n, c, ptr(out), stream_ptr()), "psb_score_rows")
    return out


def make_contrib(idx, src, src_row=None, src_div=1, scale=None, scale2=None, scale2_div=1, to_bias=False):
    """One gradient contribution (psb_contrib_t).  Returns (struct, keepalive tensors)."""
    idx = _idx(idx).reshape(-1)
    keep = [idx, src, src_row, scale, scale2]
    c = Contrib(ptr(idx), idx.numel(), ptr(src, f32), ptr(src_row, i64), int(src_div), ptr(scale, f32),
                ptr(scale2, f32), int(scale2_div), 1 if to_bias else 0, 0)
    return c, keep


def scatter_reduce(contribs, table_rows, d, drop_idx=-1, dense_grad=None, dense_bias_grad=None,
                   want_rows=True, want_bias=False, device=None, out_uniq=None, out_nu=None, out_red=None):
    """Deterministic sort + segmented-reduce embedding backward (psb_scatter_reduce_rows).
    contribs: list of (Contrib, keepalive).  Returns (unique_rows [n_total] int32, reduced or None,
    reduced_bias or None, n_unique device int32 [1])."""
    n_total = sum(int(c.n) for c, _ in contribs)
    if device is None:
        device = contribs[0][1][1].device
    arr = (Contrib * len(contribs))(*[c for c, _ in contribs])
    ws_bytes = int(load().psb_scatter_reduce_workspace_bytes(n_total, table_rows))
    ws = torch.empty((ws_bytes,), dtype=u8, device=device)
    cap = max(n_total, 1)
    # out_uniq / out_nu: caller-owned persistent buffers (a CUDA-graph replay must find last step's rows there)
    uniq = out_uniq if out_uniq is not None and out_uniq.numel() >= cap else torch.empty((cap,), dtype=i32, device=device)
    if out_red is not None and want_rows:     # caller-owned (e.g. peer-visible staging) row buffer
        if out_red.shape[0] < cap or out_red.shape[1] != d:
            raise RuntimeError("scatter_reduce: out_red too small")
        red = out_red
    else:
        red = torch.empty((cap, d), dtype=f32, device=device) if want_rows else None
    redb = torch.empty((cap,), dtype=f32, device=device) if want_bias else None
    nu = out_nu.zero_() if out_nu is not None else torch.zeros((1,), dtype=i32, device=device)
    if n_total > 0:
        check(load().psb_scatter_reduce_rows(arr, len(contribs), table_rows, d, int(drop_idx), ptr(ws), ws_bytes,
                                             ptr(uniq), ptr(red), ptr(redb), ptr(nu), ptr(dense_grad, f32),
                                             ptr(dense_bias_grad, f32), stream_ptr()), "psb_scatter_reduce_rows")
    return uniq, red, redb, nu


def scatter_sort(idx_list, table_rows, drop_idx, ws, uniq, nu):
    """Sort phase of the embedding backward (psb_scatter_sort_rows): only the index lists are read.  ``ws`` /
    ``uniq`` / ``nu``: caller-owned workspace (psb_scatter_reduce_workspace_bytes), unique-row list and count."""
    arr = (Contrib * len(idx_list))(*[Contrib(ptr(t), t.numel(), None, None, 1, None, None, 1, 0, 0) for t in idx_list])
    check(load().psb_scatter_sort_rows(arr, len(idx_list), table_rows, int(drop_idx), ptr(ws), ws.numel(), ptr(uniq),
                                       ptr(nu), stream_ptr()), "psb_scatter_sort_rows")


def scatter_reduce_sorted(contribs, table_rows, d, drop_idx, ws, uniq, nu, dense_grad=None, dense_bias_grad=None,
                          want_rows=False, want_bias=False, out_red=None):
    """Reduce phase over a workspace psb_scatter_sort_rows has filled for the same index lists in the same order.
    out_red: caller-owned row buffer (e.g. the peer-visible staging list) instead of a fresh one."""
    n_total = sum(int(c.n) for c, _ in contribs)
    arr = (Contrib * len(contribs))(*[c for c, _ in contribs])
    dev = uniq.device
    cap = max(n_total, 1)
    if out_red is not None and want_rows:
        if out_red.shape[0] < cap or out_red.shape[1] != d:
            raise RuntimeError("scatter_reduce_sorted: out_red too small")
        red = out_red
    else:
        red = torch.empty((cap, d), dtype=f32, device=dev) if want_rows else None
    redb = torch.empty((cap,), dtype=f32, device=dev) if want_bias else None
    check(load().psb_scatter_reduce_sorted(arr, len(contribs), table_rows, d, int(drop_idx), ptr(ws), ws.numel(),
                                           ptr(uniq), ptr(red), ptr(redb), ptr(nu), ptr(dense_grad, f32),
                                           ptr(dense_bias_grad, f32), stream_ptr()), "psb_scatter_reduce_sorted")
    return red, redb


def scatter_workspace_bytes(n_total, table_rows):
    return int(load().psb_scatter_reduce_workspace_bytes(n_total, table_rows))


def zero_rows(rows, n_rows, d, dense=None, dense_bias=None):
    check(load().psb_zero_rows(ptr(rows, i32), ptr(n_rows, i32), rows.numel(), d, ptr(dense, f32),
                               ptr(dense_bias, f32), stream_ptr()), "psb_zero_rows")


def table_max_row_sqnorm(table, n_rows=None):
    """Device scalar max_r |table[r]|^2 (psb_table_max_row_sqnorm); cache it while the table is static."""
    out = torch.empty((1,), dtype=f32, device=table.device)
    n_rows = table.shape[0] if n_rows is None else int(n_rows)
    check(load().psb_table_max_row_sqnorm(ptr(table, f32), n_rows, table.shape[1], ptr(out), stream_ptr()),
          "psb_table_max_row_sqnorm")
    return out


class PreparedCatalog(object):
    """fp16 copy of a static evaluation table + its measured quantisation statistics (psb_catalog_prepare_f16)."""
    __slots__ = ("half", "stats", "n_items", "fits")


def catalog_prepare_f16(table, n_items=None):
    """One pass over table[0..n_items): fp16 copy, max |e|^2, max |e - half(e)|^2, overflow flag.  ``fits`` (read
    once on the host) says whether the fp16 shortlist may be used; callers fall back to TOPK_TC otherwise."""
    n_items = table.shape[0] if n_items is None else int(n_items)
    d = table.shape[1]
    p = PreparedCatalog()
    p.half = torch.empty((n_items, d), dtype=torch.float16, device=table.device)
    p.stats = torch.zeros((4,), dtype=f32, device=table.device)
    p.n_items = n_items
    check(load().psb_catalog_prepare_f16(ptr(table, f32), n_items, d, p.half.data_ptr(), ptr(p.stats), stream_ptr()),
          "psb_catalog_prepare_f16")
    p.fits = float(p.stats[2].item()) == 0.0
    return p


def catalog_topk(queries, table, k, n_items=None, bias=None, id_base=0, id_stride=1, mode=_lib.TOPK_EXACT,
                 max_row_sqnorm=None, prepared=None):
    """Top-k items per query over the whole table with fused selection (psb_catalog_topk).
    Returns (ids [m,k] int64, scores [m,k] fp32), descending score / ascending id.
    mode TOPK_TC16 needs ``prepared`` (catalog_prepare_f16 of the same table)."""
    m, d = queries.shape
    n_items = table.shape[0] if n_items is None else int(n_items)
    dev = queries.device
    if mode == _lib.TOPK_TC16:
        if prepared is None or prepared.n_items != n_items:
            raise RuntimeError("TOPK_TC16 needs catalog_prepare_f16(table, n_items) of the same table")
        ws_bytes = int(load().psb_catalog_topk_workspace_bytes(m, n_items, d, k, mode))
        if ws_bytes < 0:
            check(ws_bytes, "psb_catalog_topk_workspace_bytes")
        ws = torch.empty((ws_bytes,), dtype=u8, device=dev)
        ids = torch.empty((m, k), dtype=i64, device=dev)
        sc = torch.empty((m, k), dtype=f32, device=dev)
        check(load().psb_catalog_topk_f16(ptr(queries, f32), m, ptr(table, f32), prepared.half.data_ptr(),
                                          ptr(prepared.stats), n_items, d, ptr(bias, f32), k, int(id_base),
                                          int(id_stride), ptr(ws), ws_bytes, ptr(ids), ptr(sc), stream_ptr()),
              "psb_catalog_topk_f16")
        return ids, sc
    ws_bytes = int(load().psb_catalog_topk_workspace_bytes(m, n_items, d, k, mode))
    if ws_bytes < 0:
        check(ws_bytes, "psb_catalog_topk_workspace_bytes")
    ws = torch.empty((ws_bytes,), dtype=u8, device=dev)
    ids = torch.empty((m, k), dtype=i64, device=dev)
    sc = torch.empty((m, k), dtype=f32, device=dev)
    check(load().psb_catalog_topk(ptr(queries, f32), m, ptr(table, f32), n_items, d, ptr(bias, f32), k,
                                  int(id_base), int(id_stride), mode, ptr(max_row_sqnorm, f32), ptr(ws), ws_bytes,
                                  ptr(ids), ptr(sc),
                                  stream_ptr()), "psb_catalog_topk")
    return ids, sc


def topk_merge(ids, scores):
    """Merge per-shard lists [g, m, k] (all_gather layout) into the global top-k [m, k]."""
    g, m, k = ids.shape
    out_i = torch.empty((m, k), dtype=i64, device=ids.device)
    out_s = torch.empty((m, k), dtype=f32, device=ids.device)
    check(load().psb_topk_merge(ptr(ids, i64), ptr(scores, f32), g, m, k, ptr(out_i), ptr(out_s), stream_ptr()),
          "psb_topk_merge")
    return out_i, out_s


# ---- fused single-position sequence encoder (psb_encoder_fwd / psb_encoder_bwd) ---------------
class EncoderCall(object):
    """Host-side handle of one encoder forward: the cfg / params structs, the saved-state buffer
    and the tensors that must stay alive until the backward call."""
    __slots__ = ("cfg", "params", "saved", "keep", "S", "T", "d", "copies", "tem", "names")


def encoder_fwd(params, heads, first=None, table=None, idx=None, pad_idx=-1, dense=None, mask=None, pe=None,
                copies=1, out_pos=0, pre_ln=False, eps=1e-6, p_drop=0.0, seed=None, raw_input=False,
                first_ready=None):
    """One TransformerEncoderLayer + final LayerNorm evaluated at ONE output position (include/psb.h N1).
    params: dict name -> fp32 CUDA tensor with the psb_encoder_params_t member names.  Tokens: ``first`` [S,d]
    + ``table`` / ``idx`` [S,T-1] (TEM), or ``dense`` [S,T,d] (+ ``mask`` [S,T] uint8/bool, 1 = real).
    ``seed``: int64 CUDA tensor [1] (required when p_drop > 0).  ``first_ready``: torch.cuda.Event recorded on the
    side stream that produces ``first`` (psb_encoder_cfg_t.first_ready).  Returns (out [S*copies, d], EncoderCall)."""
    tem = first is not None
    if tem:
        S, d = first.shape
        idx = _idx(idx)
        T = 1 + idx.shape[1]
        dev = first.device
    else:
        S, T, d = dense.shape
        dev = dense.device
        if mask is not None and mask.dtype != u8:
            mask = mask.to(u8)
        if mask is not None:
            mask = mask.contiguous()
    if out_pos < 0:
        out_pos += T
    ff = params["w1"].shape[0]
    if pe is not None:
        pe = pe.reshape(-1, d)[:T].contiguous()
    P = _lib.EncoderParams()
    keep = [first, table, idx, dense, mask, pe, seed]
    for n in _lib._ENC_NAMES:
        t = params.get(n)
        if t is not None:
            t = t.detach()
            keep.append(t)
        setattr(P, n, ptr(t, f32))
    cfg = _lib.EncoderCfg(S, T, d, int(heads), ff, int(copies), int(out_pos), 1 if pre_ln else 0, 1 if raw_input else 0, float(eps),
                          float(p_drop), ptr(seed, i64), ptr(first, f32), ptr(table, f32),
                          table.shape[0] if table is not None else 0, ptr(idx), int(pad_idx), ptr(dense, f32),
                          ptr(mask), ptr(pe, f32), first_ready.cuda_event if first_ready is not None else None)
    lib = load()
    sb = int(lib.psb_encoder_saved_bytes(ctypes.byref(cfg)))
    if sb < 0:
        check(sb, "psb_encoder_saved_bytes")
    wb = int(lib.psb_encoder_workspace_bytes(ctypes.byref(cfg), 0))
    saved = torch.empty((sb,), dtype=u8, device=dev)
    ws = torch.empty((wb,), dtype=u8, device=dev)
    out = torch.empty((S * copies, d), dtype=f32, device=dev)
    check(lib.psb_encoder_fwd(ctypes.byref(cfg), ctypes.byref(P), ptr(saved), sb, ptr(ws), wb, ptr(out),
                              stream_ptr()), "psb_encoder_fwd")
    cfg.first_ready = None                       # forward-only: the backward call must not see a stale event
    call = EncoderCall()
    call.cfg, call.params, call.saved, call.keep = cfg, P, saved, keep
    call.S, call.T, call.d, call.copies, call.tem = S, T, d, copies, tem
    call.names = [n for n in _lib._ENC_NAMES if params.get(n) is not None]
    return out, call


def encoder_bwd(call, grad_out, shapes, wgrad_event=None):
    """Backward of encoder_fwd.  ``shapes``: dict name -> shape of every parameter to differentiate.
    Returns (grad_first, grad_rest, grad_dense, grads dict, workspace).  ``wgrad_event`` (torch.cuda.Event that has
    been recorded at least once): the weight gradients are produced on the library's side stream and the event is
    recorded behind them -- the caller must make every reader of ``grads`` wait for it and keep ``workspace`` alive
    until then (psb_encoder_cfg_t.wgrad_done)."""
    S, T, d = call.S, call.T, call.d
    dev = grad_out.device
    g_first = g_rest = g_dense = None
    if call.tem:
        g_first = torch.empty((S, d), dtype=f32, device=dev)
        g_rest = torch.empty((S, T - 1, d), dtype=f32, device=dev)
    else:
        g_dense = torch.empty((S, T, d), dtype=f32, device=dev)
    G = _lib.EncoderParams()
    grads = {}
    for n in _lib._ENC_NAMES:
        t = None
        if n in shapes:
            t = grads[n] = torch.empty(shapes[n], dtype=f32, device=dev)
        setattr(G, n, ptr(t, f32))
    lib = load()
    wb = int(lib.psb_encoder_workspace_bytes(ctypes.byref(call.cfg), 1))
    ws = torch.empty((wb,), dtype=u8, device=dev)
    call.cfg.wgrad_done = wgrad_event.cuda_event if wgrad_event is not None else None
    check(lib.psb_encoder_bwd(ctypes.byref(call.cfg), ctypes.byref(call.params), ptr(call.saved),
                              call.saved.numel(), ptr(ws), wb, ptr(grad_out.contiguous(), f32), ptr(g_first),
                              ptr(g_rest), ptr(g_dense), ctypes.byref(G), stream_ptr()), "psb_encoder_bwd")
    call.cfg.wgrad_done = None
    return g_first, g_rest, g_dense, grads, ws


# ---- optional per-op device timing (bench.py roofline leg) ----------------------------------
PROFILE = None   # dict name -> list of (start_event, end_event) when enabled


def _profiled(name, fn):
    def wrapper(*a, **k):
        if PROFILE is None:
            return fn(*a, **k)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = fn(*a, **k)
        e.record()
        PROFILE.setdefault(name, []).append((s, e))
        return out
    wrapper.__name__, wrapper.__doc__ = fn.__name__, fn.__doc__
    return wrapper


for _n in ("catalog_prepare_f16", "table_max_row_sqnorm", "gather_rows", "gather_meanpool", "fs_bwd", "token_weights", "ns_loss", "tem_loss", "tem_loss_finish", "score_rows",
           "scatter_reduce", "scatter_sort", "scatter_reduce_sorted", "zero_rows", "catalog_topk", "topk_merge", "encoder_fwd", "encoder_bwd"):
    globals()[_n] = _profiled(_n, globals()[_n])

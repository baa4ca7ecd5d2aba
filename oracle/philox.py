"""Philox4x32-10 (Salmon et al., SC'11) in numpy: the counter-based generator the fused encoder
kernels use for dropout (prodsearch_b200/csrc/encoder_common.cuh).  TEST INFRASTRUCTURE ONLY: lets the
CPU oracle apply exactly the multipliers the device applied."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised: counters c0..c3 (uint64 arrays holding 32-bit values), key words k0, k1 (ints)."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3)]
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & MASK, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def dropout_multipliers(seed, stream, elems, p):
    """Multiplier (0 or 1/(1-p), fp32) of every element index in ``elems`` (int array) of dropout
    stream ``stream``: element e uses word e % 4 of philox(counter = (e // 4, stream), key = seed)."""
    elems = np.asarray(elems, dtype=np.uint64)
    if p <= 0:
        return np.ones(elems.shape, dtype=np.float32)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    c = elems >> np.uint64(2)
    out = philox4x32_10(c & MASK, c >> np.uint64(32), np.full(c.shape, stream, np.uint64), np.zeros(c.shape, np.uint64),
                        seed & 0xFFFFFFFF, seed >> 32)
    word = np.choose((elems & np.uint64(3)).astype(np.int64), out)
    thr = max(1, int(np.float64(np.float32(p)) * 16777216.0 + 0.5))
    keep = (word >> np.uint64(8)) >= np.uint64(thr)
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(keep, scale, np.float32(0.0)).astype(np.float32)


def encoder_dropout_muls(seed, p, S, C, T, H, d, ff, out_pos):
    """The four multiplier tensors of one fused-encoder call, laid out for oracle.encoder_layer on
    S*C replicated sequences: ones everywhere except the query position ``out_pos``."""
    rows = np.arange(S * C, dtype=np.int64)
    attn = np.ones((S * C, H, T, T), np.float32)
    e = (rows[:, None, None] * H + np.arange(H)[None, :, None]) * T + np.arange(T)[None, None, :]
    attn[:, :, out_pos, :] = dropout_multipliers(seed, 1, e, p)
    ctx = np.ones((S * C, T, d), np.float32)
    ctx[:, out_pos, :] = dropout_multipliers(seed, 2, rows[:, None] * d + np.arange(d)[None, :], p)
    inner = np.ones((S * C, T, ff), np.float32)
    inner[:, out_pos, :] = dropout_multipliers(seed, 3, rows[:, None] * ff + np.arange(ff)[None, :], p)
    out = np.ones((S * C, T, d), np.float32)
    out[:, out_pos, :] = dropout_multipliers(seed, 4, rows[:, None] * d + np.arange(d)[None, :], p)
    return dict(attn=attn, ctx=ctx, inner=inner, out=out)

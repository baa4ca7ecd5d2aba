"""CPU study behind DESIGN.md section 8 (encoder tails on tcgen05): does the 3xTF32 split keep a 128 -> 512 -> 128 FFN
inside the parity bars (1e-5 relative on outputs, 1e-4 on gradients) that plain TF32 misses?

Model of tcgen05.mma kind::tf32 used here (pessimistic where the hardware is undocumented): operands are fp32 words
whose low 13 mantissa bits are IGNORED (truncation, not rounding); the 8 products of one K = 8 instruction are summed
exactly and added to the fp32 accumulator with truncation toward zero (no round-to-nearest).  The split:
    a = a_hi + a_lo,  a_hi = a with the low 13 bits cleared (what the hardware sees when it is handed a),
    a_lo = a - a_hi   (exact in fp32; the hardware truncates it once more to 11 significant bits)
    a.b ~= a_hi.b_hi + a_lo.b_hi + a_hi.b_lo          (three MMAs into the same accumulator)
Run:  python tools/tf32x3_numerics.py        (numpy only, ~10 s)"""
import json

import numpy as np


def trunc13(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def trunc_to_f32(x64):
    """fp64 -> fp32 toward zero."""
    y = x64.astype(np.float32)
    over = np.abs(y.astype(np.float64)) > np.abs(x64)
    y[over] = np.nextafter(y[over], np.float32(0))
    return y


ACCUMULATE = "trunc"       # "trunc" (toward zero, pessimistic) | "rn" (round to nearest even)


def mma_chain(terms):
    """terms: list of (A [M,K], B [N,K]) fp32 operand pairs accumulated into one fp32 accumulator, K = 8 per step."""
    m, n = terms[0][0].shape[0], terms[0][1].shape[0]
    acc = np.zeros((m, n), np.float32)
    k = terms[0][0].shape[1]
    for k0 in range(0, k, 8):
        for a, b in terms:
            part = trunc13(a[:, k0:k0 + 8]).astype(np.float64) @ trunc13(b[:, k0:k0 + 8]).astype(np.float64).T
            tot = acc.astype(np.float64) + part
            acc = trunc_to_f32(tot) if ACCUMULATE == "trunc" else tot.astype(np.float32)
    return acc


def gemm_modes(a, b):
    a_lo, b_lo = a - trunc13(a), b - trunc13(b)
    return {
        "fp32_fma_order": np.add.reduce([np.outer(a[:, j], b[:, j]).astype(np.float32) for j in range(a.shape[1])],
                                        dtype=np.float32) if a.shape[0] * b.shape[0] <= 1 << 16 else (a @ b.T),
        "tf32_x1": mma_chain([(a, b)]),
        "tf32_x3": mma_chain([(a, b), (a_lo, b), (a, b_lo)]),
        # the two correction products in an accumulator of their own (its truncation errors are relative to a sum
        # 2^-11 times smaller), added to the main one in fp32 by the epilogue
        "tf32_x3_two_accumulators": mma_chain([(a, b)]) + mma_chain([(a_lo, b), (a, b_lo)]),
    }


def rel(x, ref, scale):
    return float(np.max(np.abs(x.astype(np.float64) - ref) / scale))


def main():
    rng = np.random.default_rng(0)
    rows, d, ff = 2304, 128, 512
    x = rng.standard_normal((rows, d)).astype(np.float32)                     # LayerNorm output: unit variance
    w1 = (rng.standard_normal((ff, d)) * np.sqrt(2.0 / (d + ff))).astype(np.float32)
    w2 = (rng.standard_normal((d, ff)) * np.sqrt(2.0 / (d + ff))).astype(np.float32)
    gelu = lambda v: 0.5 * v * (1 + np.tanh(np.sqrt(2 / np.pi) * (v + 0.044715 * v ** 3)))
    out = {}
    h_ref = x.astype(np.float64) @ w1.astype(np.float64).T
    y_ref = gelu(h_ref) @ w2.astype(np.float64).T
    gw_ref = (x.astype(np.float64).T @ y_ref).T                               # a K = 2304 weight-gradient-shaped product
    for mode in ("fp32_fma_order", "tf32_x1", "tf32_x3", "tf32_x3_two_accumulators"):
        h = gemm_modes(x, w1)[mode]
        y = gemm_modes(gelu(h.astype(np.float64)).astype(np.float32), w2)[mode]
        gw = gemm_modes(np.ascontiguousarray(y_ref.astype(np.float32).T), np.ascontiguousarray(x.T))[mode]
        out[mode] = {
            # the tests' metric: |x - ref| <= tol * max|ref| (outputs), tol * sum|terms| floor (gradients)
            "ffn_up_rel_to_max": rel(h, h_ref, np.abs(h_ref).max()),
            "ffn_out_rel_to_max": rel(y, y_ref, np.abs(y_ref).max()),
            "wgrad_K2304_rel_to_max": rel(gw, gw_ref, np.abs(gw_ref).max()),
        }
    out["bars"] = {"outputs": 1e-5, "gradients": 1e-4}
    return out


if __name__ == "__main__":
    res = {}
    for ACCUMULATE in ("trunc", "rn"):
        res["accumulate_" + ACCUMULATE] = main()
    print(json.dumps(res, indent=1))

"""First GPU run of the 3xTF32 tcgen05 GEMM (csrc/gemm3_tf32.cu, psb_debug_gemm3_tf32): accuracy against an fp64
product and time per launch, in a subprocess under a timeout (a hand-off bug in a tcgen05 pipeline hangs).

    timeout 300 python profiles/check_gemm3.py

Prints one JSON line per shape: max |out - ref| / max |ref| for the kernel and, next to it, for an fp32 cuBLAS product
(TF32 off) and a plain-TF32 one -- the kernel has to sit with the fp32 number (~1e-6), far from the TF32 one (~1e-3) --
and the microseconds per launch (CUDA events, 20 launches).  Shapes: the encoder's projections at batch 384
(q: 384 x 128 -> 128; K|V: ~3.5k active tokens x 128 -> 256), a ragged row count, K = 256 and K = 512 (ring reuse)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import ctypes, json, sys, torch
sys.path.insert(0, %r)
from prodsearch_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
torch.backends.cuda.matmul.allow_tf32 = False
ok = True
for m, k, j, with_bias in ((384, 128, 128, True), (3500, 128, 256, True), (1000, 256, 128, False), (130, 512, 64, True), (8064, 128, 256, True)):
    a = torch.randn(m, k, device="cuda")
    bt = torch.randn(j, k, device="cuda") * 0.1
    bias = torch.randn(j, device="cuda") if with_bias else None
    out = torch.full((m, j), float("nan"), device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    call = lambda: lib.psb_debug_gemm3_tf32(a.data_ptr(), k, m, k, bt.data_ptr(), j, bias.data_ptr() if with_bias else None,
                                            out.data_ptr(), j, s)
    st = call()
    torch.cuda.synchronize()
    ref = a.double() @ bt.double().t() + (bias.double() if with_bias else 0.0)
    scale = ref.abs().max().item()
    err = (out.double() - ref).abs().max().item() / scale
    f32 = a @ bt.t() + (bias if with_bias else 0.0)
    torch.backends.cuda.matmul.allow_tf32 = True
    tf32 = a @ bt.t() + (bias if with_bias else 0.0)
    torch.backends.cuda.matmul.allow_tf32 = False
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): call()
    e1.record(); torch.cuda.synchronize()
    good = st == 0 and err < 5e-6
    ok = ok and good
    print(json.dumps({"m": m, "k": k, "j": j, "bias": with_bias, "status": st, "rel_err_3xtf32": err,
                      "rel_err_cublas_fp32": (f32.double() - ref).abs().max().item() / scale,
                      "rel_err_cublas_tf32": (tf32.double() - ref).abs().max().item() / scale,
                      "us_per_launch": round(e0.elapsed_time(e1) / 20 * 1e3, 2), "ok": good}), flush=True)
# which accumulate rounding does tcgen05 kind::tf32 use?  The kernel's two-accumulator result against the two CPU models
# of tools/tf32x3_numerics.py (exact sum of the 8 products of a K = 8 step, then fp32 accumulate with truncation /
# round-to-nearest): the closer model -- ideally bit-equal -- decides whether one accumulator would do.
import numpy as np
sys.path.insert(0, %r)
import tf32x3_numerics as tn
m, k, j = 128, 128, 64
a = torch.randn(m, k, device="cuda"); bt = torch.randn(j, k, device="cuda")
out = torch.zeros(m, j, device="cuda")
st = lib.psb_debug_gemm3_tf32(a.data_ptr(), k, m, k, bt.data_ptr(), j, None, out.data_ptr(), j, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
an, bn, got = a.cpu().numpy(), bt.cpu().numpy(), out.cpu().numpy()
a_lo, b_lo = an - tn.trunc13(an), bn - tn.trunc13(bn)
for model in ("trunc", "rn"):
    tn.ACCUMULATE = model
    pred = tn.mma_chain([(an, bn)]) + tn.mma_chain([(a_lo, bn), (an, b_lo)])
    print(json.dumps({"probe": "accumulate model " + model, "status": st, "bit_equal_fraction": float((pred == got).mean()),
                      "max_abs_diff": float(np.abs(pred.astype(np.float64) - got).max())}), flush=True)
print("VERDICT:", "3xTF32 GEMM within 5e-6 of fp64 on every shape" if ok else "FAILED -- keep PSB_ENC_TC unset")
sys.exit(0 if ok else 1)
''' % (ROOT, os.path.join(ROOT, "tools"))

if __name__ == "__main__":
    try:
        r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, timeout=240)
    except subprocess.TimeoutExpired as ex:
        print(json.dumps({"error": "timeout (hang?)", "partial": (ex.stdout or b"")[-600:].decode("utf8", "replace")}))
        sys.exit(1)
    print(r.stdout.strip())
    if r.returncode != 0:
        print(r.stderr[-800:])
    sys.exit(r.returncode)

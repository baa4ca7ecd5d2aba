"""Timing of full-catalog top-100 (G5) for both modes; prints one JSON line per (M, N, mode)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prodsearch_b200 import _lib, ops  # noqa: E402


def timed(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


d = 128
sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1000000"])]
for n in sizes:
    table = torch.empty(n + 1, d, device="cuda").normal_()
    norm = ops.table_max_row_sqnorm(table, n)
    prep = ops.catalog_prepare_f16(table, n)
    for m in (24, 128, 384, 1024, 4096):
        q = torch.randn(m, d, device="cuda")
        for mode, name in ((_lib.TOPK_TC16, "tcgen05_f16"), (_lib.TOPK_TC, "tcgen05_tf32"), (_lib.TOPK_EXACT, "exact_fp32")):
            if mode == _lib.TOPK_EXACT and (n > 2_000_000 or m > 384):
                continue
            sec = timed(lambda: ops.catalog_topk(q, table, 100, n_items=n, mode=mode, max_row_sqnorm=norm, prepared=prep if mode == _lib.TOPK_TC16 else None))
            _lib.profile_enable(True)          # per-kernel split of the same call (CUDA events per launch)
            for _ in range(3):
                ops.catalog_topk(q, table, 100, n_items=n, mode=mode, max_row_sqnorm=norm, prepared=prep if mode == _lib.TOPK_TC16 else None)
            split = {k_: round(v[1] / 3 * 1e3, 1) for k_, v in _lib.profile_dump().items()}
            _lib.profile_enable(False)
            print(json.dumps({"n_items": n, "m": m, "mode": name, "ms": round(sec * 1e3, 4), "kernel_us": split,
                              "queries_per_s": round(m / sec, 1), "tflops": round(2.0 * m * n * d / sec / 1e12, 2),
                              "table_GBps": round(n * d * 4 / sec / 1e9, 1)}))
    del table

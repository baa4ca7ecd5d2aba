"""Two full-catalog top-100 calls (1M x 128 table, M from argv, default 384; mode from argv: f16|tf32) for ncu
captures; prints no bench value."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prodsearch_b200 import _lib, ops  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 384
mode = _lib.TOPK_TC if (len(sys.argv) > 2 and sys.argv[2] == "tf32") else _lib.TOPK_TC16
n, d = int(os.environ.get("PSB_N", "1000000")), 128
table = torch.empty(n + 1, d, device="cuda").normal_()
norm = ops.table_max_row_sqnorm(table, n)
prep = ops.catalog_prepare_f16(table, n)
q = torch.randn(m, d, device="cuda")
for _ in range(2):
    ops.catalog_topk(q, table, 100, n_items=n, mode=mode, max_row_sqnorm=norm, prepared=prep if mode == _lib.TOPK_TC16 else None)
torch.cuda.synchronize()

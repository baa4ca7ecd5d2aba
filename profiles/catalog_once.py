"""Two full-catalog top-100 calls (1M x 128 table, M from argv, default 384) for ncu captures; prints no bench value."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prodsearch_b200 import _lib, ops  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 384
n, d = 1_000_000, 128
table = torch.empty(n + 1, d, device="cuda").normal_()
norm = ops.table_max_row_sqnorm(table, n)
q = torch.randn(m, d, device="cuda")
for _ in range(2):
    ops.catalog_topk(q, table, 100, n_items=n, mode=_lib.TOPK_TC, max_row_sqnorm=norm)
torch.cuda.synchronize()

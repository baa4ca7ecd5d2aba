"""Runs each hot-path kernel once (after one warm-up) in the bandwidth regime on a 16M x 128 table.
Used under ncu (profiles/capture.sh); prints nothing that counts as a bench value."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prodsearch_b200 import _lib, ops, synth  # noqa: E402

rows, d = 16_000_000, 128
table = torch.empty(rows + 1, d, device="cuda").normal_()
n = 4_000_000
idx = synth.gather_indices(n, rows, seed=1, dist="uniform").cuda()
idz = synth.gather_indices(n, rows, seed=2, dist="zipf").cuda()
src = torch.randn(n, d, device="cuda")
na, k = 500_000, 5
anchor = torch.randn(na, d, device="cuda")
pos = idx[:na].view(na, 1).contiguous()
neg = idx[na:na + na * k].view(na, 1, k).contiguous()
idx2 = idx[:4_000_000].view(400_000, 10).contiguous()
q = torch.randn(384, d, device="cuda")
modes = [_lib.TOPK_EXACT] + ([_lib.TOPK_TC] if "--tc" in sys.argv else [])
for rep in range(2):
    ops.gather_rows(table, idx)
    ops.gather_meanpool(table, idx2, pad_idx=rows)
    ops.ns_loss(anchor, table, pos, neg)
    ops.scatter_reduce([ops.make_contrib(idx, src)], rows + 1, d, drop_idx=rows)
    ops.scatter_reduce([ops.make_contrib(idz, src)], rows + 1, d, drop_idx=rows)
    for m in modes:
        ops.catalog_topk(q, table, 100, n_items=1_000_000, mode=m)
torch.cuda.synchronize()

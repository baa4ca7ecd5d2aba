import os, sys, torch
sys.path.insert(0, os.getcwd())
from prodsearch_b200 import _lib, ops
n, d = 1000000, 128
table = torch.empty(n + 1, d, device="cuda").normal_()
q = torch.randn(24, d, device="cuda")
for _ in range(3):
    ops.catalog_topk(q, table, 100, n_items=n, mode=_lib.TOPK_TC)
torch.cuda.synchronize()

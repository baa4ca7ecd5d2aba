"""A/B timing of kernel variants selected by environment knobs (each knob is read once per process, so every
variant runs in its own subprocess).  Prints one JSON line per variant; run under gpurun:
    python profiles/ab_variants.py
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys, torch
sys.path.insert(0, %r)
from prodsearch_b200 import ops, synth
what = sys.argv[1]
rows, d = 16_000_000, 128
table = torch.empty(rows + 1, d, device="cuda").normal_()
n = 4_000_000
idx = synth.gather_indices(n, rows, seed=1, dist="uniform").cuda()
def timed(fn, iters=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
if what == "g3":
    na, k = 500_000, 5
    anchor = torch.randn(na, d, device="cuda")
    pos = idx[:na].view(na, 1).contiguous(); neg = idx[na:na + na * k].view(na, 1, k).contiguous()
    ms = timed(lambda: ops.ns_loss(anchor, table, pos, neg))
    nbytes = na * ((1 + k) * (d * 4 + 8) + 2 * d * 4)
else:
    src = torch.randn(n, d, device="cuda")
    ids = idx if what == "g2" else synth.gather_indices(n, rows, seed=2, dist="zipf").cuda()
    c = [ops.make_contrib(ids, src)]
    out = ops.scatter_reduce(c, rows + 1, d, drop_idx=rows)
    nu = int(out[3].item())
    ms = timed(lambda: ops.scatter_reduce(c, rows + 1, d, drop_idx=rows))
    nbytes = n * (d * 4 + 8) + nu * d * 4
print(json.dumps({"what": what, "env": {k: v for k, v in os.environ.items() if k.startswith("PSB_")}, "ms": round(ms, 4),
                  "GBps": round(nbytes / ms / 1e6, 1), "frac_of_6548": round(nbytes / ms / 1e6 / 6548.5, 3)}))
''' % ROOT
for what, envs in (("g3", [{"PSB_NS_W1": "0"}, {"PSB_NS_W1": "1"}, {"PSB_NS_W1": "2"}]),
                   ("g2", [{"PSB_RADIX_MATCH": "match"}, {"PSB_RADIX_MATCH": "ballot"}]),
                   ("g2z", [{"PSB_RADIX_MATCH": "ballot"}])):
    for e in envs:
        env = dict(os.environ)
        env.update(e)
        r = subprocess.run([sys.executable, "-c", CHILD, what], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-800:])

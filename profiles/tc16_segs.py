"""fp16 shortlist at 1M / 2M items: main-pass segment schedules (PSB_TC16_SCHED="r0:r": first segment r0 x the pilot, then r x steps), each in its own
subprocess (the knob is read once); lists compared by digest."""
import hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import hashlib, json, os, sys, torch
sys.path.insert(0, %r)
from prodsearch_b200 import _lib, ops
n = int(sys.argv[1])
g = torch.Generator(device="cuda").manual_seed(1)
table = torch.empty(n + 1, 128, device="cuda").normal_(generator=g)
prep = ops.catalog_prepare_f16(table, n)
for m in (24, 384, 4096):
    q = torch.randn(m, 128, device="cuda", generator=g)
    f = lambda: ops.catalog_topk(q, table, 100, n_items=n, mode=_lib.TOPK_TC16, prepared=prep)
    for _ in range(2): ids, sc = f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): f()
    b.record(); torch.cuda.synchronize()
    h = hashlib.sha256(ids.cpu().numpy().tobytes() + sc.cpu().numpy().tobytes()).hexdigest()[:12]
    print(json.dumps({"n": n, "m": m, "sched": os.environ.get("PSB_TC16_SCHED", "auto"), "ms": round(a.elapsed_time(b) / 5, 4), "sha": h}), flush=True)
''' % ROOT
for n in (1_000_000, 16_000_000):
    for sched in ("10:5", "auto", "3:3", "4:4", "10:0"):
        env = dict(os.environ)
        if sched != "auto":
            env["PSB_TC16_SCHED"] = sched
        r = subprocess.run([sys.executable, "-c", CHILD, str(n)], env=env, capture_output=True, text=True, timeout=200)
        print(r.stdout.strip() or r.stderr[-300:])

"""G5 fp16 shortlist: validate and time the opt-in variants of the main-pass kernel (csrc/catalog_tc.cu:
tc16_score_v2_kernel<.., VAR>, PSB_TC16_EPI = 2 | 3 | 4, and the PSB_TC16_MT tile-residency knob) against the default v1
kernel.  The knobs are read once per process, so each variant runs in its own subprocess, under a hard timeout (a
hand-off bug in a tcgen05 pipeline shows up as a hang, not as a wrong answer).  Run under gpurun on ONE GPU:

    timeout 900 python profiles/check_tc16_v2.py                       # v1 against v2: one JSON line per (M, N, variant) + a verdict
    timeout 900 python profiles/check_tc16_v2.py --variants 3,4,3/2,4/2   # "epi" or "epi/max tiles per CTA", each against v1
    timeout 900 python profiles/check_tc16_v2.py --stats               # + the cycle split per item tile (who waits for whom)
    timeout 200 python profiles/check_tc16_v2.py --tmem                # TMEM read bytes / clock / SM (psb_debug_tmem_read_bw)

Every variant must return bit-identical ids and scores (every mode rescores exactly in fp32); a variant is only worth
making the default if its ms is lower.  Round-1 results: profiles/r01Z_tc16_v2_stats.jsonl (v2: identical, 4.5-6 %
faster), r01Z_tc16_v3_stats.jsonl (v3: identical, 13-19 % faster).  Variant 4 and PSB_TC16_MT were written after the
round's GPU budget was spent: compiled and SASS-checked only."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import hashlib, json, os, sys, torch
sys.path.insert(0, %r)
from prodsearch_b200 import _lib, ops
n, d = int(sys.argv[1]), 128
torch.manual_seed(0)
table = torch.empty(n + 1, d, device="cuda").normal_()
prep = ops.catalog_prepare_f16(table, n)
for m in ((384, 4096) if os.environ.get("PSB_CHECK_QUICK") == "1" else (24, 128, 384, 1024, 4096)):
    q = torch.randn(m, d, device="cuda")
    max_mt = int(os.environ.get("PSB_TC16_MT", "4"))
    m_tiles = (m + 127) // 128
    groups = (m_tiles + max_mt - 1) // max_mt
    MT = (m_tiles + groups - 1) // groups          # query tiles resident per CTA (plan16_for)
    TN = 128 if MT <= 2 else 64                    # items per tcgen05.mma = TN; a 128 x TN x 16 f16 MMA takes TN / 2 cycles
    f = lambda: ops.catalog_topk(q, table, 100, n_items=n, mode=_lib.TOPK_TC16, prepared=prep)
    ids, sc = f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): f()
    b.record(); torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(3): f()
    split = {k_: round(v[1] / 3 * 1e3, 1) for k_, v in _lib.profile_dump().items()}
    _lib.profile_enable(False)
    stats = None
    if os.environ.get("PSB_TC16_STATS") == "1" and os.environ.get("PSB_TC16_EPI") in ("2", "3", "4"):
        import ctypes
        buf = (ctypes.c_uint64 * 8)()
        _lib.load().psb_debug_tc16_stats(buf, 1)                      # reset, then one call on its own
        f(); _lib.load().psb_debug_tc16_stats(buf, 1)
        c = [int(x) for x in buf]
        tiles, ctas = max(c[6], 1), max(c[7], 1)
        stats = {"cycles_per_tile": {"issuer_total": round(c[0] / tiles, 1), "issuer_wait_accumulator": round(c[1] / tiles, 1),
                                     "issuer_wait_items": round(c[2] / tiles, 1), "issuer_issue": round((c[0] - c[1] - c[2]) / tiles, 1),
                                     "epilogue_total": round(c[3] / tiles, 1), "epilogue_wait_scores": round(c[4] / tiles, 1),
                                     "producer_wait_stage": round(c[5] / tiles, 1)},
                 "tiles": c[6], "cta_launches": c[7], "mma_floor_cycles_per_tile": 8 * MT * TN // 2, "MT": MT, "TN": TN}
    h = hashlib.sha256(ids.cpu().numpy().tobytes() + sc.cpu().numpy().tobytes()).hexdigest()[:16]
    print(json.dumps({"n_items": n, "m": m, "epi": os.environ.get("PSB_TC16_EPI", "1"), "max_mt": max_mt, "ms": round(a.elapsed_time(b) / 5, 4),
                      "kernel_us": split, "tflops": round(2.0 * m * n * d / (a.elapsed_time(b) / 5) / 1e9, 1), "sha": h, "stats": stats}), flush=True)
''' % ROOT


def run(n, epi, max_mt=4):
    env = dict(os.environ, PSB_TC16_EPI=str(epi), PSB_TC16_MT=str(max_mt))
    quick = "--quick" in sys.argv
    if quick:
        env["PSB_CHECK_QUICK"] = "1"
    if epi in (2, 3, 4) and "--stats" in sys.argv:
        env["PSB_TC16_STATS"] = "1"     # the instrumented kernel (a few clock reads per tile): not a timing run
    try:
        r = subprocess.run([sys.executable, "-c", CHILD, str(n)], env=env, capture_output=True, text=True,
                           timeout=25 if quick else 240)
    except subprocess.TimeoutExpired as ex:
        print(json.dumps({"n_items": n, "epi": epi, "max_mt": max_mt, "error": "timeout (hang?)",
                          "partial": (ex.stdout or b"")[-400:].decode("utf8", "replace")}))
        return {}
    if r.returncode != 0:
        print(json.dumps({"n_items": n, "epi": epi, "max_mt": max_mt, "error": r.stderr[-600:]}))
    out = {}
    for line in r.stdout.splitlines():
        if line.startswith("{"):
            print(line)
            j = json.loads(line)
            out[j["m"]] = j
    return out


def parse_variants(text):
    """"3,4,3/2" -> [(3, 4), (4, 4), (3, 2)]: (PSB_TC16_EPI, PSB_TC16_MT)."""
    out = []
    for item in text.split(","):
        epi, _, mt = item.strip().partition("/")
        out.append((int(epi), int(mt) if mt else 4))
    return out


TMEM_CHILD = r'''
import ctypes, json, sys
sys.path.insert(0, %r)
import torch
from prodsearch_b200 import _lib
torch.zeros(1, device="cuda")
out = (ctypes.c_double * 2)()
for warps in (4, 8, 12, 16):
    _lib.check(_lib.load().psb_debug_tmem_read_bw(warps, 20000, out), "psb_debug_tmem_read_bw")
    print(json.dumps({"tmem_read": "tcgen05.ld.32x32b.x32 + wait::ld", "warps": warps, "bytes_per_clk_one_cta": round(out[0], 1),
                      "bytes_per_clk_all_sms": round(out[1], 1), "needed_for_f16_peak_at_K128": 128.0}))
''' % ROOT


if __name__ == "__main__":
    if "--tmem" in sys.argv:
        r = subprocess.run([sys.executable, "-c", TMEM_CHILD], capture_output=True, text=True, timeout=120)
        print(r.stdout.strip() or r.stderr[-600:])
        sys.exit(r.returncode)
    if "--only" in sys.argv:       # one variant alone (its sha is compared by hand with an earlier run: same seed, same data)
        (epi, mt), = parse_variants(sys.argv[sys.argv.index("--only") + 1])
        got = run(1_000_000, epi, mt)
        sys.exit(0 if got else 1)
    variants = parse_variants(sys.argv[sys.argv.index("--variants") + 1]) if "--variants" in sys.argv else [(2, 4)]
    ok = True
    for n in ((1_000_000,) if "--quick" in sys.argv else (1_000_000, 16_000_000 if "--big" in sys.argv else 250_000)):
        v1 = run(n, 1)
        ok = ok and bool(v1)
        for epi, mt in variants:
            vx = run(n, epi, mt)
            for m in sorted(v1):
                same = m in vx and vx[m]["sha"] == v1[m]["sha"]
                ok = ok and same
                print(json.dumps({"n_items": n, "m": m, "variant": "%d/%d" % (epi, mt), "identical": same, "v1_ms": v1[m]["ms"],
                                  "ms": vx.get(m, {}).get("ms"),
                                  "speedup": round(v1[m]["ms"] / vx[m]["ms"], 3) if m in vx else None}))
    print("VERDICT:", "every variant returns v1's lists bit for bit" if ok
          else "a variant DIFFERS or failed -- keep PSB_TC16_EPI / PSB_TC16_MT unset")
    sys.exit(0 if ok else 1)

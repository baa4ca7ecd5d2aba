#!/bin/bash
# PSB_ENC_TC=L (3: fused forward tail, 4: + fused backward tail): stage diffs against the FFMA kernels, then bench lines
# with and without it
R=${1:-r02G}
L=${2:-3}
mkdir -p gpurun_out
timeout 200 python profiles/diff_enc_tc.py 3 > gpurun_out/${R}_diff_enc_tc3.jsonl 2>&1; echo "fwd diff exit $?"
timeout 200 python profiles/diff_enc_bwd_tc.py $L > gpurun_out/${R}_diff_enc_bwd_tc.jsonl 2>&1; echo "bwd diff exit $?"
PSB_ENC_TC=$L timeout 300 python bench.py --no-extra --no-cpu > gpurun_out/${R}_bench_tc$L.json 2> gpurun_out/${R}_bench_tc$L.err; echo "bench tc$L exit $?"
PSB_ENC_TC=0 timeout 300 python bench.py --no-extra --no-cpu > gpurun_out/${R}_bench_tc0.json 2> gpurun_out/${R}_bench_tc0.err; echo "bench tc0 exit $?"
python - <<PY
import json
for t in ("tc$L","tc0"):
    try:
        j=json.loads(open("gpurun_out/${R}_bench_%s.json"%t).read().strip().splitlines()[-1])
        print(t, j["value"], j["ms_per_step"], j.get("roofline",{}).get("kernel"), j.get("roofline",{}).get("frac"))
        print("  ", {n: round(v["avg_launch_us"],1) for n,v in j["kernels"].items() if "tail" in n or "gemm" in n or "wgrad" in n})
    except Exception as e:
        print(t, "failed", e)
PY

#!/bin/bash
# PSB_ENC_TC=3 (fused cluster tail): stage diff against the FFMA kernels, then bench lines with and without it
R=${1:-r02G}
mkdir -p gpurun_out
timeout 200 python profiles/diff_enc_tc.py 3 > gpurun_out/${R}_diff_enc_tc3.jsonl 2>&1; echo "diff exit $?"
tail -14 gpurun_out/${R}_diff_enc_tc3.jsonl
PSB_ENC_TC=3 timeout 300 python bench.py --no-extra --no-cpu > gpurun_out/${R}_bench_tc3.json 2> gpurun_out/${R}_bench_tc3.err; echo "bench tc3 exit $?"
timeout 300 python bench.py --no-extra --no-cpu > gpurun_out/${R}_bench_tc0.json 2> gpurun_out/${R}_bench_tc0.err; echo "bench tc0 exit $?"
python - <<PY
import json
for t in ("tc3","tc0"):
    try:
        j=json.loads(open("gpurun_out/${R}_bench_%s.json"%t).read().strip().splitlines()[-1])
        print(t, j["value"], j["ms_per_step"], j.get("roofline",{}).get("kernel"), j.get("roofline",{}).get("frac"))
    except Exception as e:
        print(t, "failed", e)
PY

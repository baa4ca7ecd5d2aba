#!/bin/bash
# ncu evidence for the tensor-core encoder tails (default build, one GPU):  bash profiles/capture_r02_tc.sh r02T
R=${1:-r02T}
mkdir -p gpurun_out
# 1. launch list of the default bench command (graph replays): per-kernel durations, cold cache / serialised -> shares only
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-extra --no-cpu > gpurun_out/${R}_bench_under_ncu.log 2>&1
# 2. full sections of the tail kernels inside the step
ncu --set full --clock-control none --import-source on -k regex:'tail_fused_tc_kernel|tail_bwd_fused_tc_kernel|tail_attn_bwd_kernel|tail_ctx_kernel|wgrad_kernel' \
    -s 12 -c 6 -o gpurun_out/${R}_tails_full python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/${R}_tails_under_ncu.log 2>&1
ncu -i gpurun_out/${R}_tails_full.ncu-rep --page raw --csv > gpurun_out/${R}_tails_full_raw.csv 2>/dev/null
rm -f gpurun_out/${R}_tails_full.ncu-rep
wc -l gpurun_out/${R}_launches_bench.csv gpurun_out/${R}_tails_full_raw.csv

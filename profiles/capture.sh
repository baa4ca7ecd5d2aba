#!/bin/bash
# ncu captures for one round (run under gpurun from the repo root):  bash profiles/capture.sh r01
# 1) launch list (per-launch device time) of the bench command and of the bandwidth-regime script
# 2) --set full of the dominant hand-written kernels (one launch each)
R=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/${R}_bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${R}_launches_bw.csv python profiles/bw_regime.py ${2} > gpurun_out/${R}_bw_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ns_loss_kernel|seg_reduce_kernel|radix_scatter_kernel|gather_rows_kernel|tc_score' \
    -s 8 -c 8 -o gpurun_out/${R}_full python profiles/bw_regime.py ${2} > gpurun_out/${R}_full.log 2>&1
ls -la gpurun_out | tail -8

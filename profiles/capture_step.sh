#!/bin/bash
# ncu captures of the batch-384 TEM step (run under gpurun from the repo root):  bash profiles/capture_step.sh r01c
# 1) launch list of the eager bench command (per-launch device time: compare SHARES only)
# 2) --set full of the dominant hand-written kernels of the step (one launch each)
R=${1:-r01c}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --eager > gpurun_out/${R}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'tail_fwd_kernel|tail_bwd_kernel|count_sort_segments_kernel|wgrad_kernel|rows_gemm_kernel|fs_bwd_kernel|meanpool_kernel' \
    -s 40 -c 14 -o gpurun_out/${R}_step_full python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --eager > gpurun_out/${R}_step_full.log 2>&1
ls -la gpurun_out | tail -6

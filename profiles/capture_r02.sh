#!/bin/bash
# ncu evidence of round 2 (one GPU; run under gpurun from the repo root):  bash profiles/capture_r02.sh r02
#  1) launch list of the default bench command (graph replays: per-node device time; compare SHARES only)
#  2) --set full of the two dominant kernels of the step (eager launch of the same step)
#  3) --set full of the fp16 shortlist main pass at 16M items / 4096 queries (the last main-pass segment of the 2nd call)
#  4) --set full of the row-sparse optimizer kernels and the lazily-updated-row catch-up in the 16M-row training step
R=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-extra --no-cpu > gpurun_out/${R}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'tail_fwd_kernel|tail_bwd_kernel|wgrad_kernel|rows_gemm_kernel|ns_loss_w1_kernel|seg_reduce_kernel' \
    -s 30 -c 9 -o gpurun_out/${R}_step_full python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --eager > gpurun_out/${R}_step_full.log 2>&1
ncu -i gpurun_out/${R}_step_full.ncu-rep --page raw --csv > gpurun_out/${R}_step_full_raw.csv 2>/dev/null
PSB_N=16000000 ncu --set full --clock-control none --import-source on -k regex:'tc16_score' -s 7 -c 1 \
    -o gpurun_out/${R}_catalog16_m4096_full python profiles/catalog_once.py 4096 > gpurun_out/${R}_catalog16_full.log 2>&1
ncu -i gpurun_out/${R}_catalog16_m4096_full.ncu-rep --page raw --csv > gpurun_out/${R}_catalog16_m4096_full_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:'adam_rows|sqnorm_rows' -s 6 -c 6 \
    -o gpurun_out/${R}_sparse_adam_full python profiles/run_train16.py > gpurun_out/${R}_sparse_adam_full.log 2>&1
ncu -i gpurun_out/${R}_sparse_adam_full.ncu-rep --page raw --csv > gpurun_out/${R}_sparse_adam_full_raw.csv 2>/dev/null
rm -f gpurun_out/${R}_sparse_adam_full.ncu-rep gpurun_out/${R}_step_full.ncu-rep
ls -la gpurun_out | grep ${R}_ | tail -12

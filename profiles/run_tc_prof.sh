ncu --set full --clock-control none --import-source on -k regex:'tc_score_kernel' -s 3 -c 1 -o gpurun_out/r01_tc python profiles/bench_catalog.py 1000000 > gpurun_out/r01_tc.log 2>&1
ncu -i gpurun_out/r01_tc.ncu-rep --page raw --csv > gpurun_out/r01_tc_raw.csv 2>/dev/null
ncu -i gpurun_out/r01_tc.ncu-rep --page source --csv > gpurun_out/r01_tc_source.csv 2>/dev/null
ls -la gpurun_out | tail -5

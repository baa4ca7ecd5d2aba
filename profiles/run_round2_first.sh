#!/bin/bash
# First gpurun call of round 2 (one GPU): everything written after round 1's GPU budget was spent gets its first run.
#   gpurun --timeout 1500 -- 'bash profiles/run_round2_first.sh r02a'
# 1. GPU parity suite (includes the Adam kernel's per-CTA coefficient change and the corpus-from-files test);
# 2. TMEM read-bandwidth probe, then the v2 epilogue of the fp16 shortlist kernel against v1 (bit-identical lists
#    required), then v2's cycle split per role -- each in subprocesses under timeouts (a tcgen05 hand-off bug hangs);
# 3. the default bench line (now with extra.bandwidth_regime.G5_catalog_topk_16M);
# 4. one ncu --set full capture of the v2 kernel at M = 4096 if (2) passed.
R=${1:-r02a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${R}_pytest.log
tail -5 gpurun_out/${R}_pytest.log
timeout 200 python profiles/check_tc16_v2.py --tmem > gpurun_out/${R}_tmem.jsonl 2>&1; cat gpurun_out/${R}_tmem.jsonl
timeout 900 python profiles/check_tc16_v2.py > gpurun_out/${R}_tc16_v2.jsonl 2>&1; V2=$?; tail -12 gpurun_out/${R}_tc16_v2.jsonl
if [ $V2 -eq 0 ]; then
  timeout 600 python profiles/check_tc16_v2.py --stats > gpurun_out/${R}_tc16_v2_stats.jsonl 2>&1
  grep '"epi": "2"' gpurun_out/${R}_tc16_v2_stats.jsonl | cut -c1-700
fi
timeout 700 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench exit $?"
tail -c 400 gpurun_out/${R}_bench.err
if [ $V2 -eq 0 ]; then
  PSB_TC16_EPI=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'tc16_score_v2_kernel' -s 2 -c 1 \
      -o gpurun_out/${R}_tc16v2_full python profiles/catalog_once.py 4096 > gpurun_out/${R}_tc16v2_full.log 2>&1
fi
ls -la gpurun_out | tail -8

#!/bin/bash
# First gpurun call of round 2 (one GPU): everything written after round 1's GPU budget was spent gets its first run.
#   gpurun --timeout 1700 -- 'bash profiles/run_round2_first.sh r02a'
# 1. GPU parity suite with the DEFAULT kernels (includes the corpus-from-files test);
# 2. the fp16 shortlist tests again with PSB_TC16_EPI=3 (measured in round 1: identical lists, 13-19 % faster) -- if they
#    pass, 3 becomes the default (tc16_epilogue_variant() in csrc/catalog_tc.cu);
# 3. variants 3, 4 and the PSB_TC16_MT=2 plan (128 items per MMA, 2 query tiles per CTA) against v1 at 1M items, lists
#    must be bit-identical; then the cycle split of each -- all in subprocesses under timeouts (a tcgen05 hand-off bug hangs);
# 3b. the 3xTF32 GEMM standalone (accuracy against fp64, us per launch), then the encoder tests and a bench line with PSB_ENC_TC=1;
# 4. the default bench line (extra.bandwidth_regime.G5_catalog_topk_16M gets its first run here);
# 5. one ncu --set full capture of the best variant's main pass at M = 4096.
R=${1:-r02a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${R}_pytest.log
tail -5 gpurun_out/${R}_pytest.log
PSB_TC16_EPI=3 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_full_size.py -m gpu -q -k "f16 or catalog or rank" \
    > gpurun_out/${R}_pytest_epi3.log 2>&1; echo "pytest (EPI=3) exit $?" >> gpurun_out/${R}_pytest_epi3.log
tail -3 gpurun_out/${R}_pytest_epi3.log
timeout 900 python profiles/check_tc16_v2.py --quick --variants 3,4,3/2,4/2 > gpurun_out/${R}_tc16_variants.jsonl 2>&1; VX=$?
grep -v '"kernel_us"' gpurun_out/${R}_tc16_variants.jsonl | tail -12
timeout 600 python profiles/check_tc16_v2.py --quick --stats --variants 3,4,3/2,4/2 > gpurun_out/${R}_tc16_variants_stats.jsonl 2>&1
grep '"stats": {' gpurun_out/${R}_tc16_variants_stats.jsonl | cut -c1-900
# 3b. the 3xTF32 tcgen05 GEMM (csrc/gemm3_tf32.cu) on its own, then the encoder / model / train-step tests with it switched in
timeout 300 python profiles/check_gemm3.py > gpurun_out/${R}_gemm3.jsonl 2>&1; G3=$?; cat gpurun_out/${R}_gemm3.jsonl
if [ $G3 -eq 0 ]; then
  timeout 400 python profiles/diff_enc_tc.py > gpurun_out/${R}_diff_enc_tc.jsonl 2>&1; grep -v '"ok": true\|"identical": true' gpurun_out/${R}_diff_enc_tc.jsonl | tail -8
  PSB_ENC_TC=1 timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_models.py tests/test_gpu_train_step.py -m gpu -q \
      > gpurun_out/${R}_pytest_enc_tc.log 2>&1; echo "pytest (PSB_ENC_TC=1) exit $?" >> gpurun_out/${R}_pytest_enc_tc.log
  tail -3 gpurun_out/${R}_pytest_enc_tc.log
  PSB_ENC_TC=1 timeout 400 python bench.py --no-extra --no-cpu > gpurun_out/${R}_bench_enc_tc.json 2> gpurun_out/${R}_bench_enc_tc.err
  PSB_ENC_TC=2 timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_models.py tests/test_gpu_train_step.py -m gpu -q \
      > gpurun_out/${R}_pytest_enc_tc2.log 2>&1; echo "pytest (PSB_ENC_TC=2) exit $?" >> gpurun_out/${R}_pytest_enc_tc2.log
  tail -3 gpurun_out/${R}_pytest_enc_tc2.log
  PSB_ENC_TC=2 timeout 400 python bench.py --no-extra --no-cpu > gpurun_out/${R}_bench_enc_tc2.json 2> gpurun_out/${R}_bench_enc_tc2.err
fi
timeout 700 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench exit $?"
tail -c 400 gpurun_out/${R}_bench.err
if [ $VX -eq 0 ]; then
  PSB_TC16_EPI=4 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'tc16_score_v2_kernel' -s 2 -c 1 \
      -o gpurun_out/${R}_tc16v4_full python profiles/catalog_once.py 4096 > gpurun_out/${R}_tc16v4_full.log 2>&1
fi
ls -la gpurun_out | tail -8

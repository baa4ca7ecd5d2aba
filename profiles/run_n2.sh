#!/bin/bash
# Two-GPU validation (gpurun --gpus 2): sharded parity check under torchrun, then the 2-GPU bench line.
R=${1:-r01j}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tests/multi_gpu_check.py > gpurun_out/${R}_multi_gpu_check.log 2>&1; echo "check exit $?"
tail -12 gpurun_out/${R}_multi_gpu_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/${R}_bench_n2.json 2> gpurun_out/${R}_bench_n2.err; echo "bench exit $?"
tail -c 1500 gpurun_out/${R}_bench_n2.json | head -c 1500; tail -5 gpurun_out/${R}_bench_n2.err

#!/bin/bash
# Two-GPU validation (gpurun --gpus $N): sharded parity check under torchrun, then the 2-GPU bench line.
R=${1:-r02}; N=${2:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/multi_gpu_check.py > gpurun_out/${R}_multi_gpu_check.log 2>&1; echo "check exit $?"
tail -12 gpurun_out/${R}_multi_gpu_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/${R}_bench_n$N.json 2> gpurun_out/${R}_bench_n$N.err; echo "bench exit $?"
tail -c 1500 gpurun_out/${R}_bench_n$N.json | head -c 1500; tail -5 gpurun_out/${R}_bench_n$N.err

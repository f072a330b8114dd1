#!/bin/bash
# driver-style runs of bench.py at HEAD: default flags on 1 GPU, torchrun on 2 GPUs (both arms)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/r02_bench_head_n1.log 2>&1; echo "n1 rc=$?"; tail -4 gpurun_out/r02_bench_head_n1.log | cut -c1-400
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_head_n2.log 2>&1; echo "n2 rc=$?"; grep '^{"metric"' gpurun_out/r02_bench_head_n2.log | cut -c1-300; tail -3 gpurun_out/r02_bench_head_n2.log | cut -c1-200
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/r02_bench_head_n2_ref.log 2>&1; echo "n2 ref rc=$?"; grep '^{' gpurun_out/r02_bench_head_n2_ref.log | cut -c1-300; tail -3 gpurun_out/r02_bench_head_n2_ref.log | cut -c1-200

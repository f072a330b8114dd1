#!/bin/bash
# 8-GPU call: the bench at N=8 (fused peer update, then NCCL), the peer check, and N=1 on the same box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-diffpool --no-genconv > gpurun_out/r02_bench_n8_peer.log 2>&1; echo "bench n8 peer rc=$?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r02_bench_n8_peer.log | head -2
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --nccl-update --no-cpu-baseline --no-diffpool --no-genconv > gpurun_out/r02_bench_n8_nccl.log 2>&1; echo "bench n8 nccl rc=$?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r02_bench_n8_nccl.log | head -2
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tools/peer_check.py > gpurun_out/r02_peer_check_n8.log 2>&1; echo "peer_check rc=$?"; tail -2 gpurun_out/r02_peer_check_n8.log | cut -c1-500
timeout 120 python bench.py --no-cpu-baseline --no-diffpool --no-genconv > gpurun_out/r02_bench_n1_same_box.log 2>&1; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r02_bench_n1_same_box.log | head -2

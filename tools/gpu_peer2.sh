#!/bin/bash
# 2-GPU call: fused peer update vs NCCL (flat buffers + trainer, eager + graph), then the bench at N=2 both ways
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_peer.py -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest_peer.log 2>&1; echo "pytest peer rc=$?"; tail -3 gpurun_out/r02_pytest_peer.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/peer_check.py > gpurun_out/r02_peer_check_n2.log 2>&1; echo "peer_check rc=$?"; tail -3 gpurun_out/r02_peer_check_n2.log | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2_peer.log 2>&1; echo "bench n2 peer rc=$?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r02_bench_n2_peer.log | head -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --nccl-update > gpurun_out/r02_bench_n2_nccl.log 2>&1; echo "bench n2 nccl rc=$?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r02_bench_n2_nccl.log | head -2
python bench.py --no-cpu-baseline --no-diffpool --no-genconv > gpurun_out/r02_bench_n1_ref.log 2>&1; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r02_bench_n1_ref.log | head -2

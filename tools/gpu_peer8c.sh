#!/bin/bash
# end of round 2: 8 GPUs of one box, fused NVLink update vs the NCCL path, and one GPU of the same box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-diffpool --no-genconv"
run8() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 $B $2 2>/dev/null | grep '^{"metric"' | tail -1; }
run8 29541 "" > gpurun_out/r02_bench_n8_peer_final.json; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r02_bench_n8_peer_final.json | head -2 | tr "\n" " "; echo " <- fused update"
run8 29542 "--nccl-update" > gpurun_out/r02_bench_n8_nccl_final.json; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r02_bench_n8_nccl_final.json | head -2 | tr "\n" " "; echo " <- NCCL update"
python bench.py --no-cpu-baseline --no-diffpool --no-genconv 2>/dev/null | grep '^{"metric"' | tail -1 > gpurun_out/r02_bench_n1_same_box_final.json; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r02_bench_n1_same_box_final.json | head -2 | tr "\n" " "; echo " <- 1 GPU same box"

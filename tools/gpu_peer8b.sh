#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-diffpool --no-genconv"
run8() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 $B 2>/dev/null | grep '^{"metric"' | tail -1; }
run8 29541 > gpurun_out/r02_bench_n8_peer_1chunk.json; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r02_bench_n8_peer_1chunk.json | head -2 | tr "\n" " "; echo " <- one chunk"
MLG_PEER_CHUNKS=2 run8 29542 | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | head -2 | tr "\n" " "; echo " <- two chunks"
run8 29543 | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | head -2 | tr "\n" " "; echo " <- one chunk again"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/peer_check.py 2>/dev/null | tail -1 | cut -c1-400
python bench.py --no-cpu-baseline --no-diffpool --no-genconv 2>/dev/null | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | head -2 | tr "\n" " "; echo " <- 1 GPU same box"

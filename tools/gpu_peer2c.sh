#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-diffpool --no-genconv"
run2() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 $B 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; }
echo "default (2 chunks, 24 early blocks): $(run2 29531) $(run2 29532)"
echo "one chunk: $(MLG_PEER_ONE_CHUNK=1 run2 29533) $(MLG_PEER_ONE_CHUNK=1 run2 29534)"
echo "early blocks 8: $(MLG_PEER_EARLY_BLOCKS=8 run2 29535) $(MLG_PEER_EARLY_BLOCKS=8 run2 29536)"
echo "early blocks 64: $(MLG_PEER_EARLY_BLOCKS=64 run2 29537)"
# two independent single-GPU runs side by side: GPU-to-GPU variance + host contention without any exchange
(CUDA_VISIBLE_DEVICES=1 python bench.py --no-cpu-baseline --no-diffpool --no-genconv 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1 | sed 's/^/gpu1 alone-in-parallel /') &
CUDA_VISIBLE_DEVICES=0 python bench.py --no-cpu-baseline --no-diffpool --no-genconv 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1 | sed 's/^/gpu0 alone-in-parallel /'
wait
CUDA_VISIBLE_DEVICES=1 python bench.py --no-cpu-baseline --no-diffpool --no-genconv 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1 | sed 's/^/gpu1 alone /'
CUDA_VISIBLE_DEVICES=0 python bench.py --no-cpu-baseline --no-diffpool --no-genconv 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1 | sed 's/^/gpu0 alone /'

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/peer_check.py > gpurun_out/r02_peer_check_n2.log 2>&1; echo "peer_check rc=$?"; tail -2 gpurun_out/r02_peer_check_n2.log | cut -c1-900

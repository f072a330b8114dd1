#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "layernorm or deepergcn or genconv or gemm or linear_bf16" 2>&1 | tail -5
timeout 600 python tools/bench_components.py --only deepergcn > gpurun_out/r02_comp_deepergcn.log 2>&1; cut -c1-330 gpurun_out/r02_comp_deepergcn.log

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "genconv or gen_aggr or deepergcn or affine" 2>&1 | tail -6
timeout 600 python tools/bench_components.py --only genconv > gpurun_out/r02_comp_genconv.log 2>&1; cat gpurun_out/r02_comp_genconv.log | cut -c1-400
timeout 600 python tools/bench_components.py --only deepergcn > gpurun_out/r02_comp_deepergcn.log 2>&1; cat gpurun_out/r02_comp_deepergcn.log | cut -c1-400

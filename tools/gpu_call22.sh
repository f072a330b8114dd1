#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "genconv or gen_ or deepergcn or affine or dynconv or pathway" 2>&1 | tail -8
timeout 600 python tools/bench_components.py --only genconv > gpurun_out/r02_comp_genconv.log 2>&1; cut -c1-400 gpurun_out/r02_comp_genconv.log
timeout 600 python tools/bench_components.py --only deepergcn > gpurun_out/r02_comp_deepergcn.log 2>&1; cut -c1-400 gpurun_out/r02_comp_deepergcn.log
MLG_GEN_BWD_DETERMINISTIC=1 timeout 600 python tools/bench_components.py --only deepergcn 2>&1 | cut -c1-400

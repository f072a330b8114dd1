#!/bin/bash
# diffpool-focused check first (fast), then the whole GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_diffpool.py tests/test_gpu_vae.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
python tools/bench_components.py --only diffpool > gpurun_out/r02_comp_diffpool.log 2>&1; tail -8 gpurun_out/r02_comp_diffpool.log
bash tools/gpu_full.sh

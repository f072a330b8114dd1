#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x -p no:cacheprovider -k "bf16 or diffpool" 2>&1 | tail -15
timeout 300 python tools/bench_components.py --only diffpool > gpurun_out/r02_comp_diffpool_sk.log 2>&1; cut -c1-600 gpurun_out/r02_comp_diffpool_sk.log
MLG_GEMM_NO_STREAMK=1 timeout 300 python tools/bench_components.py --only diffpool > gpurun_out/r02_comp_diffpool_nosk.log 2>&1; cut -c1-600 gpurun_out/r02_comp_diffpool_nosk.log

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/diag_b40.py > gpurun_out/r02_diag_b40.log 2>&1; echo "diag rc=$?"; tail -32 gpurun_out/r02_diag_b40.log | cut -c1-600
python tools/bench_components.py --only diffpool > gpurun_out/r02_diffpool_comp.log 2>&1; head -1 gpurun_out/r02_diffpool_comp.log | cut -c1-300
python bench.py --no-cpu-baseline --no-diffpool --no-genconv > gpurun_out/r02_bench4.log 2>&1; echo "bench rc=$?"; tail -c 400 gpurun_out/r02_bench4.log

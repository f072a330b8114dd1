#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "knn" 2>&1 | tail -15
timeout 300 python tools/bench_components.py --only knn > gpurun_out/r02_comp_knn.log 2>&1; tail -3 gpurun_out/r02_comp_knn.log
F="--no-cpu-baseline --no-diffpool --no-genconv --no-strong --steps 50 --warmup 10"
python bench.py $F 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rb4', d['ms_per_step'], d['value'])"
MLG_POOL_RB=2 python bench.py $F 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rb2', d['ms_per_step'], d['value'])"

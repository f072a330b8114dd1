#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
F="--no-cpu-baseline --no-diffpool --no-genconv --no-strong --steps 50 --warmup 10"
for i in 1 2; do
python bench.py $F 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rb8', d['ms_per_step'], d['value'])"
MLG_POOL_RB4=1 python bench.py $F 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rb4', d['ms_per_step'], d['value'])"
done
python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "pathwayconv or pool" 2>&1 | tail -4

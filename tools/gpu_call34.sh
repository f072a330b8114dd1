#!/bin/bash
# stacked weight folding (one launch instead of fold + permute copy + zero fill + bias copy + transpose copy + split): whole suite + step time
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest_fold.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)" gpurun_out/r02_pytest_fold.log | head; tail -3 gpurun_out/r02_pytest_fold.log
for i in 1 2; do
python bench.py --no-cpu-baseline --no-diffpool --no-genconv --no-strong > gpurun_out/r02_ab_fold_on$i.log 2>&1; echo "on rc=$?"
done
python - <<'PY'
import json
for n in ("on1","on2"):
    d=json.loads(open(f"gpurun_out/r02_ab_fold_{n}.log").read().strip().splitlines()[-1])
    print(n, d["ms_per_step"], d["value"], d["e2e"]["value"], d["loss"], d["gpu_launches"])
PY

#!/bin/bash
# coalesced xty_tc / wcolsum reductions: tests, step time, ncu launch list of the step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "xty or wcolsum or sage or multilevel or fullsize or trainer or wide" > gpurun_out/r02_pytest_reduce.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_reduce.log
for i in 1 2; do
python bench.py --no-cpu-baseline --no-diffpool --no-genconv --no-strong > gpurun_out/r02_ab_reduce_on$i.log 2>&1; echo "on rc=$?"
done
python - <<'PY'
import json
for n in ("on1","on2"):
    d=json.loads(open(f"gpurun_out/r02_ab_reduce_{n}.log").read().strip().splitlines()[-1])
    print(n, d["ms_per_step"], d["value"], d["e2e"]["value"], d["loss"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r02_launches_step6.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-diffpool --no-genconv --no-strong > gpurun_out/r02_ncu_launches6.log 2>&1; echo "ncu rc=$?"

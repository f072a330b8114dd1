#!/bin/bash
# staging -> static copies as one multi-tensor launch: trainer tests + the e2e leg
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "trainer or prefetch or graph or keyless or topology or loader or e2e" > gpurun_out/r02_pytest_e2e.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_e2e.log
for i in 1 2; do
python bench.py --no-cpu-baseline --no-diffpool --no-genconv --no-strong > gpurun_out/r02_ab_e2e_on$i.log 2>&1; echo "on rc=$?"
done
python - <<'PY'
import json
for n in ("on1","on2"):
    d=json.loads(open(f"gpurun_out/r02_ab_e2e_{n}.log").read().strip().splitlines()[-1])
    print(n, d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_full_upload"]["value"], d["loss"])
PY

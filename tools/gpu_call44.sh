#!/bin/bash
# node-major gradient between the pool backward and the transform-first layer's backward aggregation: tests + A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "multilevel or fullsize or trainer or wide or sage or activation or pool or graph" > gpurun_out/r02_pytest_h1nm.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)" gpurun_out/r02_pytest_h1nm.log | head; tail -2 gpurun_out/r02_pytest_h1nm.log
for i in 1 2; do
MLG_H1_NODE_MAJOR=0 python bench.py --no-cpu-baseline --no-diffpool --no-genconv --no-strong > gpurun_out/r02_ab_h1nm_off$i.log 2>&1; echo "off rc=$?"
python bench.py --no-cpu-baseline --no-diffpool --no-genconv --no-strong > gpurun_out/r02_ab_h1nm_on$i.log 2>&1; echo "on rc=$?"
done
python - <<'PY'
import json
for n in ("off1","on1","off2","on2"):
    d=json.loads(open(f"gpurun_out/r02_ab_h1nm_{n}.log").read().strip().splitlines()[-1])
    ak=d["roofline"]["all_kernels"]
    print(n, d["ms_per_step"], d["value"], d["loss"], {k:v["ms_per_step"] for k,v in ak.items() if "aggr" in k or "gather" in k or "pool_bwd" in k})
PY

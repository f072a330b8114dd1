#!/bin/bash
# (historical: the environment switch this A/B used existed only in the commit it was run on; results: profiles/r02_ab_end_of_round.jsonl)
# programmatic dependent launch: the whole GPU suite with it on (default), then A/B of the gbm step and the DeeperGCN step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest_pdl.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_pdl.log
for i in 1 2; do
MLG_PDL=0 python bench.py --no-cpu-baseline --no-diffpool --no-genconv --no-strong > gpurun_out/r02_ab_pdl_off$i.log 2>&1; echo "off rc=$?"
python bench.py --no-cpu-baseline --no-diffpool --no-genconv --no-strong > gpurun_out/r02_ab_pdl_on$i.log 2>&1; echo "on rc=$?"
done
python - <<'PY'
import json
for n in ("off1","on1","off2","on2"):
    try:
        d=json.loads(open(f"gpurun_out/r02_ab_pdl_{n}.log").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], d["e2e"]["value"], d["loss"])
    except Exception as e:
        print(n, "failed", e)
PY
MLG_PDL=0 timeout 600 python tools/bench_components.py --only deepergcn > gpurun_out/r02_comp_deepergcn_pdl_off.log 2>&1; cut -c1-330 gpurun_out/r02_comp_deepergcn_pdl_off.log
timeout 600 python tools/bench_components.py --only deepergcn > gpurun_out/r02_comp_deepergcn_pdl_on.log 2>&1; cut -c1-330 gpurun_out/r02_comp_deepergcn_pdl_on.log

#!/bin/bash
# (historical: the environment switch this A/B used existed only in the commit it was run on; results: profiles/r02_ab_end_of_round.jsonl)
# packed-pair first-layer forward (MLG_R1F_PACKED_OFF=1 selects the scalar pass): tests + A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "rank1 or factored or multilevel or fullsize or trainer or wide or sage or activation" > gpurun_out/r02_pytest_r1f.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)" gpurun_out/r02_pytest_r1f.log | head; tail -2 gpurun_out/r02_pytest_r1f.log
for i in 1 2; do
MLG_R1F_PACKED_OFF=1 python bench.py --no-cpu-baseline --no-diffpool --no-genconv --no-strong > gpurun_out/r02_ab_r1f_off$i.log 2>&1; echo "off rc=$?"
python bench.py --no-cpu-baseline --no-diffpool --no-genconv --no-strong > gpurun_out/r02_ab_r1f_on$i.log 2>&1; echo "on rc=$?"
done
python - <<'PY'
import json
for n in ("off1","on1","off2","on2"):
    d=json.loads(open(f"gpurun_out/r02_ab_r1f_{n}.log").read().strip().splitlines()[-1])
    ak=d["roofline"]["all_kernels"]
    print(n, d["ms_per_step"], d["value"], d["loss"], {k:v["ms_per_step"] for k,v in ak.items() if "rank1" in k})
PY

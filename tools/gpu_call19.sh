#!/bin/bash
# (historical: the environment switch this A/B used existed only in the commit it was run on; results: profiles/r02_ab_end_of_round.jsonl)
# A/B of the C == 32 128-bit-lane pool kernels (MLG_POOL_C32_OFF=1 selects the previous kernels) + the tests that cover them
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_wide_batch.py -m gpu -q -p no:cacheprovider -k "pool or multilevel or fullsize or wide or trainer" > gpurun_out/r02_pytest_pool.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_pool.log
for i in 1 2; do
MLG_POOL_C32_OFF=1 python bench.py --no-cpu-baseline --no-diffpool --no-genconv --no-strong > gpurun_out/r02_ab_pool_off$i.log 2>&1; echo "off rc=$?"
python bench.py --no-cpu-baseline --no-diffpool --no-genconv --no-strong > gpurun_out/r02_ab_pool_on$i.log 2>&1; echo "on rc=$?"
done
python - <<'PY'
import json
for n in ("off1","on1","off2","on2"):
    d=json.loads(open(f"gpurun_out/r02_ab_pool_{n}.log").read().strip().splitlines()[-1])
    ak=d["roofline"]["all_kernels"]
    print(n, d["ms_per_step"], d["value"], {k:v["ms_per_step"] for k,v in ak.items() if "pool" in k})
PY

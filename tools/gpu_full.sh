#!/bin/bash
# full validation: the whole GPU suite (no -x), smoke(), the bench line and the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_full.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02_pytest_full.log | head -30
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
python bench.py > gpurun_out/r02_bench_full.log 2>&1; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_full.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_full_ref.log 2>&1; echo "bench ref rc=$?"; tail -c 400 gpurun_out/r02_bench_full_ref.log

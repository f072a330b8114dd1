#!/bin/bash
# end of round 2 (after the node-major gradient): whole GPU suite, smoke(), bench line + reference arm, launch list, ncu --set full
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export_rep() {
  ncu -i "$1.ncu-rep" --page raw --csv > "gpurun_out/$2_raw.csv" 2>/dev/null
  ncu -i "$1.ncu-rep" --page source --csv 2>/dev/null | gzip -9 > "gpurun_out/$2_source.csv.gz"
  ls -la "gpurun_out/$2_raw.csv" "gpurun_out/$2_source.csv.gz"
}
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_final.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/r02_pytest_final.log | head
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke_final.log
python bench.py > gpurun_out/r02_bench_final.log 2>&1; echo "bench rc=$?"; tail -c 200 gpurun_out/r02_bench_final.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_final_ref.log 2>&1; echo "bench ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r02_launches_step8.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-diffpool --no-genconv --no-strong > gpurun_out/r02_ncu_launches8.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on \
    -k regex:'gather_nm|sage_rank1_fwd_rows|sage_rank1_bwd_rows' \
    -s 8 -c 8 -o /tmp/r02_step_nm python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-diffpool --no-genconv --no-strong > gpurun_out/r02_ncu_full_nm.log 2>&1
echo "ncu full rc=$?"; export_rep /tmp/r02_step_nm r02_step_nm

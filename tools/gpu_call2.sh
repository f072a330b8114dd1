#!/bin/bash
# round 2, GPU call 2: full GPU suite, B=40 diagnostic, bench, launch list, ncu of the new head / DiffPool kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export_rep() {
  ncu -i "$1.ncu-rep" --page raw --csv > "gpurun_out/$2_raw.csv" 2>/dev/null
  ncu -i "$1.ncu-rep" --page source --csv 2>/dev/null | gzip -9 > "gpurun_out/$2_source.csv.gz"
  ls -la "gpurun_out/$2_raw.csv" "gpurun_out/$2_source.csv.gz"
}
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest2.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02_pytest2.log | head -30
python tools/diag_b40.py > gpurun_out/r02_diag_b40.log 2>&1; echo "diag rc=$?"; cat gpurun_out/r02_diag_b40.log | tail -12
python bench.py > gpurun_out/r02_bench2.log 2>&1; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench2.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench2_ref.log 2>&1; echo "bench ref rc=$?"; tail -c 600 gpurun_out/r02_bench2_ref.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-diffpool --no-genconv > gpurun_out/r02_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_step2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-diffpool --no-genconv > gpurun_out/r02_ncu_launches2.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on \
    -k regex:'head_|sage_rank1_bwd_rows|pool_fwd_vec' \
    -s 30 -c 9 -o /tmp/r02_head python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-diffpool --no-genconv > gpurun_out/r02_ncu_head.log 2>&1
echo "ncu head rc=$?"; export_rep /tmp/r02_head r02_head
python tools/bench_components.py --only diffpool > gpurun_out/r02_diffpool_comp.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'diffpool_' -s 2 -c 2 \
    -o /tmp/r02_diffpool python tools/bench_components.py --only diffpool > gpurun_out/r02_ncu_diffpool.log 2>&1
echo "ncu diffpool rc=$?"; export_rep /tmp/r02_diffpool r02_diffpool
cat gpurun_out/r02_diffpool_comp.log | tail -8
du -sh gpurun_out

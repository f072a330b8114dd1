#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export_rep() {
  ncu -i "$1.ncu-rep" --page raw --csv > "gpurun_out/$2_raw.csv" 2>/dev/null
  ncu -i "$1.ncu-rep" --page source --csv 2>/dev/null | gzip -9 > "gpurun_out/$2_source.csv.gz"
  ls -la "gpurun_out/$2_raw.csv" "gpurun_out/$2_source.csv.gz"
}
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest5.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02_pytest5.log | head -30
python bench.py > gpurun_out/r02_bench5.log 2>&1; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench5.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-diffpool --no-genconv > gpurun_out/r02_plain5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_step5.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-diffpool --no-genconv > gpurun_out/r02_ncu_launches5.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on \
    -k regex:'sage_rank1_fwd_rows|head_mlp|pool_fwd_vec|pool_bwd_fused2|head_conv_pool' \
    -s 30 -c 8 -o /tmp/r02_c5 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-diffpool --no-genconv > gpurun_out/r02_ncu_c5.log 2>&1
echo "ncu rc=$?"; export_rep /tmp/r02_c5 r02_c5
du -sh gpurun_out

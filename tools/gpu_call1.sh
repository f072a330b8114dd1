#!/bin/bash
# round 2, GPU call: full GPU suite (no -x), bench line, launch list, ncu --set full of the step kernels.
# .ncu-rep files stay on the box (gpurun_out/ is capped at 64 MiB): only CSV exports come back.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export_rep() {  # $1 = report path (without extension), $2 = name under gpurun_out
  ncu -i "$1.ncu-rep" --page raw --csv > "gpurun_out/$2_raw.csv" 2>/dev/null
  ncu -i "$1.ncu-rep" --page source --csv 2>/dev/null | gzip -9 > "gpurun_out/$2_source.csv.gz"
  ls -la "gpurun_out/$2_raw.csv" "gpurun_out/$2_source.csv.gz"
}
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
tail -5 gpurun_out/r02_pytest1.log
python bench.py > gpurun_out/r02_bench1.log 2>&1; echo "bench rc=$?"
MLG_R1_SELF_MASK=0 python bench.py --no-cpu-baseline --no-genconv > gpurun_out/r02_bench1_leakypass.log 2>&1; echo "bench leaky-pass rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_step.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on \
    -k regex:'gather_sum_rep|sage_rank1_fwd|sage_rank1_bwd_rows|pool_bwd_fused2|pool_fwd_kernel|gemm_tf32x3_kernel|xty_tc_kernel' \
    -s 30 -c 14 -o /tmp/r02_step_kernels python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full rc=$?"; export_rep /tmp/r02_step_kernels r02_step_kernels
python tools/microbench.py genconv --bwd > gpurun_out/r02_genconv_micro.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gen_fwd_ring|gen_bwd_kernel' -s 8 -c 2 \
    -o /tmp/r02_genconv python tools/microbench.py genconv --bwd > gpurun_out/r02_ncu_genconv.log 2>&1
echo "ncu genconv rc=$?"; export_rep /tmp/r02_genconv r02_genconv
python tools/microbench.py gemm > gpurun_out/r02_gemm_micro.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16' -s 2 -c 2 \
    -o /tmp/r02_gemm_bf16 python tools/microbench.py gemm > gpurun_out/r02_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"; export_rep /tmp/r02_gemm_bf16 r02_gemm_bf16
du -sh gpurun_out

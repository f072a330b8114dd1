#!/bin/bash
# round 2, GPU call 1: full GPU suite (no -x), bench line, rank-1 backward probe, launch list, ncu --set full of the step kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
tail -5 gpurun_out/r02_pytest1.log
python bench.py > gpurun_out/r02_bench1.log 2>&1; echo "bench rc=$?"
python tools/rank1_bwd_probe.py > gpurun_out/r02_rank1_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r02_rank1_probe.log
# launch list of 2 graph-replayed steps (skip the eager warm-up launches is not possible by count: keep everything, filter here)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_step.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
# full captures: one of each hot kernel of the step, taken from the third (eager) warm-up step
ncu --set full --clock-control none --import-source on \
    -k regex:'gather_sum_rep|sage_rank1_fwd|sage_rank1_bwd_rows|pool_bwd_fused2|pool_fwd_kernel|gemm_tf32x3_kernel|xty_tc_kernel' \
    -s 30 -c 14 -o gpurun_out/r02_step_kernels python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full rc=$?"
# GENConv aggregation fwd (ring kernel) + bwd at cfg4
python tools/microbench.py genconv --bwd > gpurun_out/r02_genconv_micro.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gen_fwd_ring|gen_bwd_kernel' -s 8 -c 2 \
    -o gpurun_out/r02_genconv python tools/microbench.py genconv --bwd > gpurun_out/r02_ncu_genconv.log 2>&1
echo "ncu genconv rc=$?"
# DiffPool contractions on tcgen05 (tensor pipe utilisation)
python tools/microbench.py gemm > gpurun_out/r02_gemm_micro.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16' -s 2 -c 2 \
    -o gpurun_out/r02_gemm_bf16 python tools/microbench.py gemm > gpurun_out/r02_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
ls -la gpurun_out | tail -20

#!/bin/bash
# r02 evidence: ncu --set full of the stream-K GEMM (tensor pipe), the ring backward kernels (plain: microbench, affine: DeeperGCN)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export_rep() {
  ncu -i "$1.ncu-rep" --page raw --csv > "gpurun_out/$2_raw.csv" 2>/dev/null
  ncu -i "$1.ncu-rep" --page source --csv 2>/dev/null | gzip -9 > "gpurun_out/$2_source.csv.gz"
  ls -la "gpurun_out/$2_raw.csv" "gpurun_out/$2_source.csv.gz"
}
timeout 300 python tools/microbench.py gemm > gpurun_out/r02_gemm_micro_sk.log 2>&1; cat gpurun_out/r02_gemm_micro_sk.log | cut -c1-250
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16_streamk' -s 2 -c 4 \
    -o /tmp/r02_gemm_sk python tools/microbench.py gemm > gpurun_out/r02_ncu_gemm_sk.log 2>&1
echo "ncu gemm rc=$?"; export_rep /tmp/r02_gemm_sk r02_gemm_sk
timeout 300 python tools/microbench.py genconv --bwd > gpurun_out/r02_genconv_micro2.log 2>&1; cat gpurun_out/r02_genconv_micro2.log | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gen_bwd_ring' -s 4 -c 1 \
    -o /tmp/r02_genbwd python tools/microbench.py genconv --bwd > gpurun_out/r02_ncu_genbwd.log 2>&1
echo "ncu genbwd rc=$?"; export_rep /tmp/r02_genbwd r02_genbwd
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gen_bwd_ring|gen_fwd_kernel' -s 8 -c 2 \
    -o /tmp/r02_genaff python tools/bench_components.py --only deepergcn --quick > gpurun_out/r02_ncu_genaff.log 2>&1
echo "ncu genaff rc=$?"; export_rep /tmp/r02_genaff r02_genaff

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'gemm_tf32x3' -s 500 -c 1 -o /tmp/r02_knn python tools/knn_diag.py 100000 > gpurun_out/r02_ncu_knn.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02_knn.ncu-rep --page raw --csv > gpurun_out/r02_knn_raw.csv 2>/dev/null
ncu -i /tmp/r02_knn.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r02_knn_source.csv.gz
ls -la gpurun_out/r02_knn_raw.csv gpurun_out/r02_knn_source.csv.gz

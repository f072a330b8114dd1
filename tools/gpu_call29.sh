#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gen_bwd_ring' -s 4 -c 1 \
    -o /tmp/r02_genbwd2 python tools/microbench.py genconv --bwd > gpurun_out/r02_ncu_genbwd2.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02_genbwd2.ncu-rep --page raw --csv > gpurun_out/r02_genbwd2_raw.csv 2>/dev/null
ncu -i /tmp/r02_genbwd2.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r02_genbwd2_source.csv.gz

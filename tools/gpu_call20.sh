#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"head_conv_pool" -s 4 -c 2 -o gpurun_out/r02_head4 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-diffpool --no-genconv > gpurun_out/r02_head4_full.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r02_head4.ncu-rep --page raw --csv > gpurun_out/r02_head4_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_head4.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02_head4_source.csv.gz
rm -f gpurun_out/r02_head4.ncu-rep
ls -la gpurun_out/r02_head4*

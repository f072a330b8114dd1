#!/bin/bash
# source-level ncu capture of the fused DiffPool kernels (after the same command ran plain)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/bench_components.py --only diffpool --quick > gpurun_out/r02_comp_diffpool.log 2>&1; head -1 gpurun_out/r02_comp_diffpool.log
ncu --set full --clock-control none --import-source on -k regex:'diffpool_(fwd|bwd)_kernel' -s 2 -c 2 -o /tmp/r02_dp \
    python tools/bench_components.py --only diffpool --quick > gpurun_out/r02_ncu_dp.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02_dp.ncu-rep --page raw --csv > gpurun_out/r02_dp2_raw.csv 2>/dev/null
ncu -i /tmp/r02_dp.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r02_dp2_source.csv.gz
ls -la gpurun_out/r02_dp2_raw.csv gpurun_out/r02_dp2_source.csv.gz

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_vae.py -m gpu -q -p no:cacheprovider -k "diffpool or vae or decoder" 2>&1 | tail -8
python tools/bench_components.py --only diffpool --quick > gpurun_out/r02_comp_diffpool.log 2>&1; head -1 gpurun_out/r02_comp_diffpool.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'diffpool' -c 6 --csv --log-file gpurun_out/r02_dp_launches.csv \
    python tools/bench_components.py --only diffpool --quick > /dev/null 2>&1
grep diffpool gpurun_out/r02_dp_launches.csv | awk -F'","' '{print $5, $NF}' | tail -3
bash tools/gpu_full.sh

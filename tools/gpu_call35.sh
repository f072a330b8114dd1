#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/trace_small_ops.py > gpurun_out/r02_small_ops.log 2>&1; echo "rc=$?"; tail -60 gpurun_out/r02_small_ops.log | cut -c1-260

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "diffpool or gemm or vae or DiffPool" 2>&1 | tail -8
timeout 300 python tools/probes/diffpool_large_profile.py --train 2>&1 | grep -E "forward|sgemm|streamk|cast_bf16" | cut -c1-160

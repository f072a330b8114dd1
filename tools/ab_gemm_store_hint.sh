#!/bin/bash
# A/B of the L2 cache hint on the 3xTF32 GEMM's output tiles (csrc/gemm_tf32x3.cu, MLG_TF32X3_STORE_HINT: 0 none, 1 evict_first,
# 2 evict_last) for DESIGN.md section 6 item 1.  `build` (no GPU needed) compiles the variants next to the shipped library;
# `run` (on the GPU box) times the by-row backward kernel behind the GEMM and the training step with each of them.
set -e
cd "$(dirname "$0")/.."
PKG=multilevel-gnn_b200
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
case "$1" in
build)
  python $PKG/build.py
  for h in 1 2; do
    $NVCC $FLAGS -DMLG_TF32X3_STORE_HINT=$h -c $PKG/csrc/gemm_tf32x3.cu -o /tmp/gemm_tf32x3_hint$h.o
    objs=$(ls $PKG/csrc/_obj/*.o | grep -v '/gemm_tf32x3.o')
    $NVCC -shared -o $PKG/lib/libmlg_b200_hint$h.so $objs /tmp/gemm_tf32x3_hint$h.o -gencode arch=compute_100a,code=sm_100a
    echo "built $PKG/lib/libmlg_b200_hint$h.so"
  done ;;
run)
  for v in default hint1 hint2; do
    if [ $v = default ]; then unset MLG_B200_LIB; else export MLG_B200_LIB=$PWD/$PKG/lib/libmlg_b200_$v.so; fi
    echo "== $v"
    python tools/rank1_bwd_probe.py --reps 5 | grep -E '"producer": "(tensor-core GEMM|GEMM -> aten)' | grep '"mask": "bits"'
    for sm in 0 1; do
      MLG_R1_SELF_MASK=$sm python bench.py --no-cpu-baseline --no-genconv --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['all_kernels']
print('  self_mask=$sm step ms', d['ms_per_step'], 'rank1_bwd', k['sage_rank1_bwd']['ms_per_step'])"
    done
  done ;;
*) echo "usage: $0 build|run"; exit 2 ;;
esac

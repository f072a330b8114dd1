# A/B of the layer-2 formulation / layer-1 activation backward (functional.TRANSFORM_FIRST, RANK1_SELF_MASK)
for v in "0 1" "1 0"; do
  set -- $v
  MLG_TRANSFORM_FIRST=$1 MLG_R1_SELF_MASK=$2 timeout 30 python bench.py --no-cpu-baseline --no-genconv --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['all_kernels']
print('transform_first=$1 self_mask=$2 :', d['ms_per_step'], 'rank1_bwd', k['sage_rank1_bwd']['ms_per_step'], 'gathers', k['gather_sum_rep_kernel (SAGE mean aggregation fwd+bwd)']['ms_per_step'])" | tee -a gpurun_out/ab_layer2.log
done

# A/B of the row visiting order per kernel family (see functional.ORDER_*): prints step ms and the per-kernel timers
for v in "0 1 1" "0 0 1" "0 1 0" "0 0 0"; do
  set -- $v
  MLG_ORDER_R1B=$1 MLG_ORDER_R1F=$2 MLG_ORDER_TF=$3 timeout 100 python bench.py --no-cpu-baseline --no-genconv --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['all_kernels']
print('r1b/r1f/tf sorted = $1 $2 $3 :', d['ms_per_step'], 'rank1_bwd', k['sage_rank1_bwd']['ms_per_step'], 'rank1_fwd', k['sage_rank1_fwd']['ms_per_step'], 'gathers', k['gather_sum_rep_kernel (SAGE mean aggregation fwd+bwd)']['ms_per_step'])"
done

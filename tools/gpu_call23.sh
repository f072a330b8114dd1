#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_deepergcn4_launches.csv python tools/bench_components.py --only deepergcn --quick > gpurun_out/r02_deepergcn4_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02_deepergcn4_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=i;break
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
recs=[(r[ki],float(r[vi].replace(',',''))) for r in rows[hdr+1:] if len(r)>vi]
print(len(recs),'launches')
# affine leg = first half (warm 1 + reps 2 => 3 steps), take launches between knn and the materialised leg: approximate by taking all and grouping
d=collections.defaultdict(lambda:[0,0.0])
# find step boundaries: use the first third of affine leg
names=[n for n,_ in recs]
first_mat=next((i for i,n in enumerate(names) if 'gen_bwd_ring_kernel<1, false>' in n or 'gen_bwd_ring_kernel<(int)1, (bool)0>' in n), len(recs))
aff=recs[:first_mat]
for n,v in aff:
    k=n[:100]; d[k][0]+=1; d[k][1]+=v/1e3
tot=sum(v[1] for v in d.values())
print('affine leg total us', round(tot), 'launches', len(aff))
for k,v in sorted(d.items(), key=lambda kv:-kv[1][1])[:28]:
    print('%9.1f us %5d x  %5.1f%%  %s'%(v[1],v[0],100*v[1]/tot,k))
PY

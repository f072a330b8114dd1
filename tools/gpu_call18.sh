#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"head_|pool_|sage_rank1|gather_sum" -c 200 --csv --log-file gpurun_out/r02_head4_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-diffpool --no-genconv > gpurun_out/r02_head4_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_head4_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=i;break
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
recs=[(r[ki],float(r[vi].replace(',',''))) for r in rows[hdr+1:] if len(r)>vi]
import collections
d=collections.OrderedDict()
for n,v in recs[-60:]:
    d.setdefault(n[:90],[]).append(round(v/1e3,1))
for k,v in d.items(): print(v,k)
PY

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/knn_diag.py 100000; python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "knn" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_knn_launches.csv python tools/knn_diag.py 20000 > /dev/null 2>&1
python - <<'PY'
import csv,collections,re
rows=[r for r in csv.reader(open('gpurun_out/r02_knn_launches.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=re.sub(r'\(.*','',r[4])[-50:]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[-1])
for k,(c,x) in agg.items(): print('%9.1f us x%d %s'%(x/1e3,c,k))
PY

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "genconv or gen_ or deepergcn or affine" 2>&1 | tail -3
timeout 300 python tools/microbench.py genconv --bwd 2>&1 | grep -o '"bwd_ms_est": [0-9.]*'
timeout 600 python tools/bench_components.py --only deepergcn 2>&1 | grep -o '"fwd_bwd_ms": [0-9.]*'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'gen_bwd_ring' -c 10 --csv --log-file gpurun_out/r02_genbwd_pol.csv python tools/bench_components.py --only deepergcn --quick > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_genbwd_pol.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=i;break
h=rows[hdr]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit'); ii=h.index('ID')
d={}
for r in rows[hdr+1:]:
    if len(r)>vi: d.setdefault((r[ii],r[ki][:60]),{})[r[mi]]=(r[vi],r[ui])
for k,v in list(d.items())[:10]: print(k, v)
PY

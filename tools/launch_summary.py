"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the kernels of the LAST training step
(delimited by adam_kernel launches), aggregated by name, plus the ordered list with --order."""
import csv, re, sys, collections
path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
rows = [(d['Kernel Name'], float(d['Metric Value'].replace(',', '')) / 1e3) for d in csv.DictReader(lines)]
idx = [i for i, (n, v) in enumerate(rows) if 'adam_kernel' in n]
a, b = idx[-2] + 1, idx[-1] + 1
st = rows[a:b]
tot = sum(v for n, v in st)
short = lambda n: re.sub(r'void |<unnamed>::|at::native::|native::', '', n)
if '--order' in sys.argv:
    for i, (n, v) in enumerate(st):
        print('%3d %7.1f  %s' % (i, v, short(n)[:120]))
agg = collections.OrderedDict()
for n, v in st:
    k = re.sub(r'\(.*', '', short(n))[:80]
    c = agg.setdefault(k, [0, 0.0]); c[0] += 1; c[1] += v
small = [(n, v) for n, v in st if v < 12]
print('kernels %d  total %.1f us  (<12us: %d kernels, %.1f us)' % (len(st), tot, len(small), sum(v for n, v in small)))
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 24]:
    print('%7.1f us %3d %5.1f%%  %s' % (v, c, 100 * v / tot, n))

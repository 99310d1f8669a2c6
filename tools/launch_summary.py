"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python tools/launch_summary.py launches.csv [--seq] (--seq prints the launches in order instead)"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}
if '--seq' in sys.argv:
    for r in rows[1:]:
        print("%10.2f us  %s" % (float(r[vi].replace(',', '')) * scale.get(r[ui], 1e-3), r[ki][:90]))
    sys.exit(0)
gi = h.index('Grid Size') if 'Grid Size' in h else None
d = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    k = r[ki][:60] + ((" grid " + r[gi]) if gi is not None and '--grid' in sys.argv else "")
    d[k][0] += 1
    d[k][1] += float(r[vi].replace(',', '')) * scale.get(r[ui], 1e-3)
tot = sum(v[1] for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
    print("%-82s %5d %12.1f us %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))

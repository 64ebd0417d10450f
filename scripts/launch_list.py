#!/usr/bin/env python
"""Print the per-launch durations of the LAST refactorize+solve step in an ncu launch list CSV."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; data = rows[hi + 1:]
kn = h.index('Kernel Name'); mv = h.index('Metric Value'); gs = h.index('Grid Size'); bs = h.index('Block Size')
idxs = [i for i, r in enumerate(data) if 'k_rowscale' in r[kn]]
start = idxs[-1]
tot = 0; agg = {}
for r in data[start:]:
    name = r[kn].split('::')[-1].split('(')[0]
    t = float(r[mv].replace(',', '')) / 1000.0
    tot += t
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
    if len(sys.argv) < 3:
        print("%-42s %-16s %-12s %9.2f" % (name, r[gs], r[bs], t))
print("total us %.1f" % tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-42s n=%4d  %9.1f us" % (k, v[0], v[1]))

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time per kernel."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    a = agg[r[ki][:70]]; a[0] += 1; a[1] += v; tot += v
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]): print(f"{t:12.1f} us {100*t/tot:5.1f}%  x{c:4d}  {k}")

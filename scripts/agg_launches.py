"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    n = re.sub(r'\(.*', '', r[ki])[:70]
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    if r[ui] in ('us', 'usecond'): v *= 1e3
    if r[ui] in ('ms', 'msecond'): v *= 1e6
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot/1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{v/1e3:10.1f} us {100*v/tot:5.1f}% {c:5d}  {n}")

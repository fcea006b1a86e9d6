"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per (kernel, grid)."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value'); gi = H.index('Grid Size')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    n = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('b200::', '')[:60] + ' ' + r[gi]
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    agg[n][0] += 1; agg[n][1] += v
tot = sum(v[1] for v in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{t/1e3:10.1f} us {c:5d} {t/c/1e3:8.1f} us/launch {100*t/tot:5.1f}% {n}")
print(f"total {tot/1e3:.1f} us over {sum(v[0] for v in agg.values())} launches")

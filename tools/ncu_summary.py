"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value'); ui = H.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else v * 1000 if r[ui] == 'ms' else v
    a = agg.setdefault(r[ki][:70], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f'# {sys.argv[1]}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised per-launch times)')
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{n:70s} n={a[0]:4d} total_us={a[1]:10.1f} avg_us={a[1] / a[0]:8.1f} share={a[1] / tot:.3f}')

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals, and the launches
of the last complete search step (between the last two prepare_queries_kernel launches)."""
import csv, collections, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
seq = []
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('vdbk::', '')
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(unit, 1e-3)
    seq.append((name, v))
agg = collections.OrderedDict()
for n, v in seq:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{len(seq)} launches, {tot/1e3:.2f} ms")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"  {k[:64]:64s} n={n:5d} mean={t/n:9.1f} us share={t/tot:.3f}")
idx = [i for i, (n, _) in enumerate(seq) if n.startswith('prepare_queries')]
if len(idx) > 1:
    which = int(sys.argv[2]) if len(sys.argv) > 2 else -2
    a, b = idx[which], idx[which + 1] if which + 1 < 0 or which + 1 < len(idx) else len(seq)
    print("one step:")
    s = 0
    for n, v in seq[a:b]:
        print(f"  {n[:64]:64s} {v:9.1f} us"); s += v
    print(f"  sum {s:.1f} us")

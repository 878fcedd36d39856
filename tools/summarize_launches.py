# coding=utf-8
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]; kn, mv, mn = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
  if len(r) <= mv or r[mn] != 'gpu__time_duration.sum':
    continue
  name = r[kn].split('(')[0].replace('void ', '').replace('mvb::', '')
  a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[mv].replace(',', '')) / 1e3
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
  print("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k[:70], c, t, t / c, 100 * t / tot))
print("total us %.1f" % tot)

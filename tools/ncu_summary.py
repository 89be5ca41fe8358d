"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value')
agg = collections.OrderedDict(); tot = 0
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    name = re.sub(r'\(.*', '', r[kn]); name = re.sub(r'^void ', '', name)
    v = float(r[mv].replace(',', '')); a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print('total %.2f ms over %d launches' % (tot / 1e6, sum(c for c, _ in agg.values())))
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print('%-72s n=%4d  %8.2f ms  %5.1f%%  avg %.3f ms' % (k[:72], c, t / 1e6, 100 * t / tot, t / 1e6 / c))

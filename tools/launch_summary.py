"""Aggregates an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list by (kernel, grid).
usage: python tools/launch_summary.py launches.csv [top]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]; hdr = rows[hi]
kn, mn, mv, idc, gs = [hdr.index(x) for x in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID', 'Grid Size')]
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    d = per.setdefault(r[idc], {'name': re.sub(r'\(.*', '', r[kn]).replace('void ', ''), 'grid': r[gs]})
    d[r[mn]] = float(r[mv].replace(',', ''))
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault((d['name'], d['grid']), [0, 0, 0, 0]); a[0] += 1; a[1] += d.get('gpu__time_duration.sum', 0) / 1e6
    a[2] += d.get('dram__bytes_read.sum', 0) / 1e9; a[3] += d.get('dram__bytes_write.sum', 0) / 1e9
tot = sum(a[1] for a in agg.values())
print("total %.2f ms over %d launches" % (tot, len(per)))
for k, (n, t, rd, wr) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print('%-45s %-18s n=%3d tot %7.2f ms (%4.1f%%) avg %.3f ms rd %.3f wr %.3f GB/launch %.2f TB/s' % (k[0][:45], k[1], n, t, 100 * t / tot, t / n, rd / n, wr / n, (rd + wr) / t if t else 0))

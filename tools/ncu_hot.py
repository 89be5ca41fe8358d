"""Summary of an `ncu --page raw --csv` + `--page source --csv` pair of one kernel launch: key metrics, stall reasons,
executed instructions by opcode, instructions with the most stall samples.  Usage: python tools/ncu_hot.py raw.csv source.csv [units]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
d = dict(zip(rows[0], rows[2]))
units = float(sys.argv[3]) if len(sys.argv) > 3 else 0
for k in ('gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'):
    print(k, d.get(k))
st = [(k, float(v)) for k, v in d.items() if 'issue_stalled' in k and 'ratio' in k and 'not_issued' not in k]
for k, v in sorted(st, key=lambda x: -x[1])[:7]:
    print('   ', k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), round(v, 2))
rows = list(csv.reader(open(sys.argv[2])))
ex = collections.Counter(); st = collections.Counter(); tot = 0; top = []
for r in rows[2:]:
    if len(r) < 6:
        continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[1])
    if not m:
        continue
    op = m.group(2).split('.')[0]
    e = int(r[5] or 0); s = int(r[2] or 0)
    ex[op] += e; st[op] += s; tot += e
    top.append((s, r[1].strip(), e))
print('total executed', tot, ('per unit %.1f' % (tot / units)) if units else '')
for op, e in ex.most_common(16):
    print('%-10s exec %10d (%.1f%%) stall samples %6d' % (op, e, 100 * e / tot, st[op]))
print('top stall instrs:')
for s, i, e in sorted(top, reverse=True)[:16]:
    print(s, e, i)

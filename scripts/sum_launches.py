#!/usr/bin/env python
"""Generic ncu launch-list aggregator: per kernel name -> launches, total gpu__time_duration (and DRAM
bytes when the csv has them).  usage: sum_launches.py <csv> [--skip-init PATTERN ...]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = rows[0]
ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
per = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    per.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = v
agg = collections.OrderedDict()
for (i, n), m in per.items():
    short = n.split('(')[0].replace('void ', '')
    if len(short) > 90:
        short = short[:87] + '...'
    a = agg.setdefault(short, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += m.get('gpu__time_duration.sum', 0.0)
    a[2] += m.get('dram__bytes_read.sum', 0.0) + m.get('dram__bytes_write.sum', 0.0)
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | time [us] | share | DRAM [MB] |\n|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1]/1e3:.1f} | {100*a[1]/max(tot,1e-9):.1f} % | {a[2]/1e6:.1f} |")
print(f"| **total** | {sum(a[0] for a in agg.values())} | {tot/1e3:.1f} | 100 % | {sum(a[2] for a in agg.values())/1e6:.1f} |")

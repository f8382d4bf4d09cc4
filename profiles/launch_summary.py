#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (read on the CPU box).
Usage: python profiles/launch_summary.py gpurun_out/<tag>/launches.csv > profiles/rNN_bench_launches_summary.txt"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[12] == 'gpu__time_duration.sum']
tot = OrderedDict()
for r in rows:
    name = re.sub(r'^void ', '', r[4])
    name = re.sub(r'\(.*$', '', name).replace('cledb::', '')
    ns = float(r[14].replace(',', ''))
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += ns
allns = sum(v[1] for v in tot.values())
print('# ncu --metrics gpu__time_duration.sum --clock-control none over the bench command (cold-cache, serialised launches: shares, not absolutes)')
for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print('%-70s n=%4d total %9.3f ms avg %9.1f us share %5.1f%%' % (name[:70], n, ns / 1e6, ns / n / 1e3, 100 * ns / allns))

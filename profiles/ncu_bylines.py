#!/usr/bin/env python
"""Executed-instruction and stall-sample share per CUDA source line of an .ncu-rep (needs -lineinfo and --import-source on).
Usage: python profiles/ncu_bylines.py gpurun_out/x.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
for i, r in enumerate(rows):
    if r and r[0] == 'Line No':
        h, start = r, i + 1
        break
ii, si = h.index('Instructions Executed'), h.index('# Samples')


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


data = [r for r in rows[start:] if len(r) == len(h) and r[0].strip().isdigit()]
tot = sum(num(r[ii]) for r in data)
ts = sum(num(r[si]) for r in data)
print('# %s: %.4e warp instructions, %d stall samples' % (rep.split('/')[-1], tot, ts))
data.sort(key=lambda r: -num(r[ii]))
for r in data[:top]:
    print('%5s inst %6.2f%% samp %6.2f%%  %s' % (r[0], num(r[ii]) / tot * 100, num(r[si]) / ts * 100, r[1].strip()[:120]))

# optional: share per named line range, e.g.  name:lo-hi,lo-hi  as extra arguments
for spec in sys.argv[3:]:
    name, rngs = spec.split(':')
    rr = [tuple(int(x) for x in r.split('-')) for r in rngs.split(',')]
    sel = [r for r in data if any(lo <= int(r[0]) <= hi for lo, hi in rr)]
    print('range %-14s inst %6.2f%% samp %6.2f%%' % (name, sum(num(r[ii]) for r in sel) / tot * 100, sum(num(r[si]) for r in sel) / ts * 100))

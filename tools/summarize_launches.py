#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total
time and share.  usage: summarize_launches.py launches.csv [title]"""
import csv
import collections
import re
import sys


def short(name):
    name = re.sub(r'\(anonymous namespace\)::|<unnamed>::|void ', '', name)
    m = re.match(r'([\w:]+)(<[^(]*>)?', name)
    base = m.group(1) if m else name
    targs = (m.group(2) or '') if m else ''
    targs = targs.replace('__nv_bfloat16', 'bf16')
    if base.startswith('at::') or base.startswith('torch') or 'elementwise' in base:
        return 'torch:' + base[:60]
    return (base + targs)[:70]


def main():
    path = sys.argv[1]
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, mi, vi = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
    ui = hdr.index('Metric Unit')
    for r in rd:
        if len(r) <= vi or r[mi] != 'gpu__time_duration.sum':
            continue
        v = float(r[vi].replace(',', ''))
        u = r[ui]
        v_ms = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v
        rows.append((short(r[ki]), v_ms))
    agg = collections.OrderedDict()
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(v for _, v in rows)
    print('# %s' % (sys.argv[2] if len(sys.argv) > 2 else path))
    print('# total %.3f ms over %d launches (per-launch times are cold-cache and serialised: read the SHARES)' % (total, len(rows)))
    print('%-72s %8s %10s %7s %9s' % ('kernel', 'launches', 'total_ms', 'share', 'avg_us'))
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-72s %8d %10.3f %6.1f%% %9.1f' % (k, n, v, 100 * v / total, 1e3 * v / n))


if __name__ == '__main__':
    main()

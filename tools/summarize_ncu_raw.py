#!/usr/bin/env python
"""Per-kernel one-liners from `ncu -i X.ncu-rep --page raw --csv`: duration, DRAM bytes, DRAM / SM
throughput %, registers, occupancy, top stall.  usage: summarize_ncu_raw.py raw.csv [--json out.json]"""
import csv
import json
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]

    def col(name):
        for i, h in enumerate(hdr):
            if h == name:
                return i
        return None
    names = {
        'dur_us': 'gpu__time_duration.sum', 'rd_mb': 'dram__bytes_read.sum', 'wr_mb': 'dram__bytes_write.sum',
        'dram_pct': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm_pct': 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'regs': 'launch__registers_per_thread',
        'occ_pct': 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l2_hit': 'lts__t_sector_hit_rate.pct', 'ipc': 'sm__inst_executed.avg.per_cycle_elapsed',
        'issue_active': 'sm__inst_issued.avg.pct_of_peak_sustained_active'}
    idx = {k: col(v) for k, v in names.items()}
    ki, gi, bi = col('Kernel Name'), col('Grid Size'), col('Block Size')
    out = []
    print('%-44s %-16s %5s %9s %9s %9s %7s %6s %5s %6s' % ('kernel', 'grid', 'blk', 'dur_us', 'rd_MB', 'wr_MB', 'dram%', 'sm%', 'regs', 'occ%'))
    for r in rows[2:]:
        name = re.sub(r'\(anonymous namespace\)::|<unnamed>::|void ', '', r[ki])
        name = re.sub(r'\(.*', '', name).replace('__nv_bfloat16', 'bf16')
        vals = {}
        for k, i in idx.items():
            try:
                vals[k] = float(r[i].replace(',', '')) if i is not None and r[i] not in ('', 'n/a') else None
            except ValueError:
                vals[k] = None
        for k, key in (('rd_mb', 'rd_mb'), ('wr_mb', 'wr_mb')):
            i = idx[k]
            if i is not None and vals[k] is not None:
                u = units[i].lower()
                vals[k] *= {'byte': 1e-6, 'kbyte': 1e-3, 'mbyte': 1.0, 'gbyte': 1e3}.get(u, 1.0)
        i = idx['dur_us']
        if vals['dur_us'] is not None:
            u = units[i].lower()
            vals['dur_us'] *= {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3}.get(u, 1.0)
        f = lambda v, p='%9.1f': (p % v) if v is not None else ' ' * 8 + '-'
        print('%-44s %-16s %5s %s %s %s %s %s %5s %s' % (
            name[:44], r[gi].replace(' ', ''), r[bi].split(',')[0].strip('('), f(vals['dur_us']), f(vals['rd_mb']), f(vals['wr_mb']),
            f(vals['dram_pct'], '%7.1f'), f(vals['sm_pct'], '%6.1f'), int(vals['regs'] or 0), f(vals['occ_pct'], '%6.1f')))
        out.append(dict(kernel=name, grid=r[gi], block=r[bi], **vals))
    if '--json' in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index('--json') + 1], 'w'), indent=1)


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Per-kernel one-liners from `ncu -i X.ncu-rep --page raw --csv`: duration, DRAM bytes, DRAM / SM throughput %,
tensor-pipe utilisation % (sm__pipe_tensor*cycles_active, of peak sustained elapsed: the figure the north_star asks for
on the GEMMs), registers, occupancy, the two largest warp-stall reasons.
usage: summarize_ncu_raw.py raw.csv [--json out.json]"""
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
    tensor_cols = [i for i, h in enumerate(hdr) if ('pipe_tensor_cycles_active' in h and h.endswith('pct_of_peak_sustained_elapsed'))
                   or h == 'sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed']
    stall_cols = [(i, re.sub(r'.*issue_stalled_', '', h).replace('_per_issue_active.ratio', ''))
                  for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')
                  and 'selected' not in h]
    ki, gi, bi = col('Kernel Name'), col('Grid Size'), col('Block Size')
    out = []
    print('%-44s %-16s %5s %9s %9s %9s %7s %6s %7s %5s %6s  %s' % ('kernel', 'grid', 'blk', 'dur_us', 'rd_MB', 'wr_MB', 'dram%', 'sm%',
                                                                   'tensor%', 'regs', 'occ%', 'top stalls'))
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
        def num(i):
            try:
                return float(r[i].replace(',', ''))
            except (ValueError, IndexError):
                return None
        tvals = [v for v in (num(i) for i in tensor_cols) if v is not None]
        vals['tensor_pct'] = max(tvals) if tvals else None
        stalls = sorted(((num(i) or 0.0, nm) for i, nm in stall_cols), reverse=True)[:2]
        vals['top_stalls'] = ['%s %.1f' % (nm, v) for v, nm in stalls if v > 0]
        print('%-44s %-16s %5s %s %s %s %s %s %s %5s %s  %s' % (
            name[:44], r[gi].replace(' ', ''), r[bi].split(',')[0].strip('('), f(vals['dur_us']), f(vals['rd_mb']), f(vals['wr_mb']),
            f(vals['dram_pct'], '%7.1f'), f(vals['sm_pct'], '%6.1f'), f(vals['tensor_pct'], '%7.2f'), int(vals['regs'] or 0),
            f(vals['occ_pct'], '%6.1f'), ', '.join(vals['top_stalls'])))
        out.append(dict(kernel=name, grid=r[gi], block=r[bi], **vals))
    if '--json' in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index('--json') + 1], 'w'), indent=1)


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Join an `ncu --page raw --csv` export of tools/run_kernels.py --once with its op log:
per op of bench.py's kernel table -> measured duration, DRAM bytes (dram__bytes_read.sum +
dram__bytes_write.sum), algorithmic bytes.  Writes profiles/kernel_traffic.json (read by bench.py for
`roofline.traffic`) and prints a table.

usage: kernel_traffic.py prof_raw.csv run_kernels_ops.json [out.json] [--merge]
(--merge: keep the entries of out.json for ops that are not in this op log -- partial captures of tools/run_kernels.py --only.)
An op whose profiled kernels do not carry the name its label promises (a slipped join: the op log counts launches that the
ncu -k filter did not capture) is reported and dropped instead of being recorded under the wrong kernel."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_KERNEL = re.compile(r'dw_|pw_tc|wgrad_tc|bn_|upsample|stem|_simt|pack_weights|pool_|bilinear|resize_|cm_|adamw|'
                        r'ce_fwd|count_valid|ce_finalize|add_kernel|relu_bwd|cast_from|scale_inplace|im2col|col2im|permute_w|select_|ohem_')


EXPECT = [('bn_bwd_reduce', 'bn_bwd_reduce'), ('bn_bwd_apply', 'bn_bwd_apply'), ('bn_bwd_onepass', 'bn_bwd_onepass'), ('bn_apply', 'bn_apply'),
          ('pwconv_wgrad', 'wgrad_tc'), ('pwconv_', 'pw_tc'), ('dwconv_', 'dw_'), ('upsample_', 'upsample_'), ('stem', 'stem')]


def plausible(op, kernel_names):
    for prefix, token in EXPECT:
        if op.startswith(prefix):
            return any(token in k for k in kernel_names)
    return True


def main():
    argv = [a for a in sys.argv[1:] if a != '--merge']
    merge = '--merge' in sys.argv
    raw, ops = argv[0], json.load(open(argv[1]))
    out = argv[2] if len(argv) > 2 else os.path.join(ROOT, 'profiles', 'kernel_traffic.json')
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        i = col[name]
        v = float(r[i].replace(',', '')) if r[i] not in ('', 'n/a') else 0.0
        u = units[i].lower()
        scale = {'byte': 1.0, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0,
                 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3}.get(u, 1.0)
        return v * scale
    kernels = [r for r in rows[2:] if LIB_KERNEL.search(r[col['Kernel Name']])]
    need = sum(o['launches'] for o in ops)
    if need != len(kernels):
        print('warning: %d profiled library kernels vs %d launches in the op log' % (len(kernels), need), file=sys.stderr)
    result, k = {}, 0
    if merge and os.path.exists(out):
        result = json.load(open(out))
    print('%-44s %3s %9s %10s %10s %8s' % ('op', 'krn', 'dur_us', 'dram_MB', 'algo_MB', 'ratio'))
    for o in ops:
        ks = kernels[k:k + o['launches']]
        k += o['launches']
        dur = sum(val(r, 'gpu__time_duration.sum') for r in ks)
        dram = sum(val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum') for r in ks)
        if not ks:          # beyond the end of a capture that was cut short: leave the existing entry alone
            continue
        names = [re.sub(r'\(.*', '', re.sub(r'void |<unnamed>::', '', r[col['Kernel Name']])) for r in ks]
        if not plausible(o['op'], names):
            print('dropped %s: profiled kernels %s do not match the op (slipped join)' % (o['op'], names), file=sys.stderr)
            result.pop(o['op'], None)
            continue
        result[o['op']] = {'dram_bytes': dram, 'duration_us': dur, 'algorithmic_bytes': o['algorithmic_bytes'], 'kernels': names}
        print('%-44s %3d %9.1f %10.1f %10.1f %8.2f' % (o['op'][:44], len(ks), dur, dram / 1e6, o['algorithmic_bytes'] / 1e6,
                                                       dram / max(o['algorithmic_bytes'], 1)))
    with open(out, 'w') as f:
        json.dump(result, f, indent=1)


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Launch every op of bench.py's kernel table (the hot kernels of the Fast-SCNN training step at the
benchmark shapes, 12 x 768 x 768, bf16) ONCE after a warm-up launch and an L2 flush, so that ncu can
capture them, and record which library kernels each op launched:

    python tools/run_kernels.py                      # plain run (must exit 0 before profiling)
    ncu --set full --clock-control none --import-source on -k regex:'dw_|pw_tc|wgrad_tc|bn_|upsample|stem' \
        -o /tmp/prof python tools/run_kernels.py --once
    ncu -i /tmp/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv
    python tools/kernel_traffic.py gpurun_out/prof_raw.csv gpurun_out/run_kernels_ops.json   # -> profiles/kernel_traffic.json

Writes gpurun_out/run_kernels_ops.json = [{"op": label, "launches": n, "algorithmic_bytes": b}, ...] in launch order.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from torch_semantic_segmentation_b200 import _lib  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    once = '--once' in sys.argv          # under ncu: no warm-up launch, keeps the report small
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ops_log = []

    only = tuple(sys.argv[sys.argv.index('--only') + 1].split(',')) if '--only' in sys.argv else None   # op-name prefix filter(s)

    def runner(name, count, nbytes, fn):
        if only is not None and not name.startswith(only):
            return
        if not once:
            fn()
        torch.cuda.synchronize()
        flush.zero_()
        before = _lib.launch_count()
        fn()
        torch.cuda.synchronize()
        ops_log.append({'op': name, 'launches': _lib.launch_count() - before, 'launches_per_step': count,
                        'algorithmic_bytes': nbytes})

    bench.kernel_table(dev, runner=runner)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'run_kernels_ops.json'), 'w') as f:
        json.dump(ops_log, f, indent=1)
    print('ok: %d ops' % len(ops_log))


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Per-op device times of bench.py's kernel table (each op alone, L2 flushed, CUDA-graph replays between events):
    python tools/time_ops.py [--only PREFIX[,PREFIX...]]
One line per op: name, launches per step, us, GB/s of algorithmic bytes.  For A/B runs of a kernel variant (env gates)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    only = sys.argv[sys.argv.index('--only') + 1].split(',') if '--only' in sys.argv else None
    dev = torch.device('cuda:0')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    total = [0.0]

    def runner(name, count, nbytes, fn):
        if only is not None and not any(name.startswith(p) for p in only):
            return
        ms = bench.time_kernel(fn, 10, flush)
        total[0] += ms * count
        print('%-44s x%d %8.1f us %8.0f GB/s' % (name, count, ms * 1e3, nbytes / ms / 1e6))

    bench.kernel_table(dev, runner=runner)
    print('sum over the step: %.1f us' % (total[0] * 1e3))


if __name__ == '__main__':
    main()

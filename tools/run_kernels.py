#!/usr/bin/env python
"""Launch each hot kernel of the Fast-SCNN training step once (after a warm-up launch) at the
benchmark shapes (12 x 768 x 768, bf16), so that `ncu --set full -k regex:...` can capture them:

    python tools/run_kernels.py                      # plain run (must exit 0 before profiling)
    ncu --set full --clock-control none --import-source on -k regex:'dw_|pw_tc|wgrad_tc|bn_|upsample|stem' \
        -s 60 -c 60 -o gpurun_out/prof python tools/run_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_semantic_segmentation_b200 import ops  # noqa: E402

N, CROP, CLASSES = 12, 768, 19
dev = torch.device('cuda:0')
bf = torch.bfloat16


def act(C, div):
    return ops.empty_nhwc(N, C, CROP // div, CROP // div, bf, dev).normal_()


def kernels():
    small = ops.empty_nhwc(N, CLASSES, CROP // 8, CROP // 8, bf, dev, pitch=32).normal_()
    target = torch.randint(0, CLASSES, (N, CROP, CROP), device=dev)
    target[torch.rand(N, CROP, CROP, device=dev) < 0.1] = 255
    yield lambda: ops.upsample_logits_fwd(small, CROP, CROP)
    yield lambda: ops.upsample_ce_forward(small, target, CROP, CROP, 255, True)
    x = torch.randn(N, 3, CROP, CROP, device=dev)
    w = torch.randn(32, 3, 3, 3, device=dev)
    st = torch.zeros(64, dtype=torch.float64, device=dev)
    y2 = act(32, 2)
    sc = torch.ones(32, device=dev)
    yield lambda: ops.stem_fwd(x, w, bf, stats=st)
    yield lambda: ops.stem_wgrad(x, y2, torch.zeros_like(w))
    yield lambda: ops.bn_apply(y2, sc, sc, relu=True)
    yield lambda: ops.bn_backward(y2, None, y2, sc, sc, sc, True, beta=sc)
    for C, div, s, d in [(32, 2, 2, 1), (128, 8, 1, 1), (128, 8, 1, 4), (384, 8, 2, 1), (384, 16, 1, 1)]:
        xi = act(C, div)
        wd = torch.randn(C, 1, 3, 3, device=dev)
        sd = torch.zeros(2 * C, dtype=torch.float64, device=dev)
        yo = ops.dwconv_fwd(xi, wd, s, d)
        yield lambda: ops.dwconv_fwd(xi, wd, s, d, stats=sd)
        yield lambda: ops.dwconv_dgrad(yo, wd, xi.shape[2], xi.shape[3], s, d)
        yield lambda: ops.dwconv_wgrad(xi, yo, torch.zeros_like(wd), s, d)
    for K, Nc, div in [(32, 48, 4), (64, 384, 8), (128, 128, 8), (384, 64, 16), (576, 96, 32)]:
        xi = act(K, div)
        wp = torch.randn(Nc, K, 1, 1, device=dev) * 0.05
        packed = ops.pack_weights_bf16(wp)
        yo = act(Nc, div)
        sp = torch.zeros(2 * Nc, dtype=torch.float64, device=dev)
        yield lambda: ops.pwconv_fwd(xi, wp, stats=sp, wp=packed[0], impl=1)
        yield lambda: ops.pwconv_dgrad(yo, wp, wpT=packed[1], impl=1)
        yield lambda: ops.pwconv_wgrad(xi, yo, torch.zeros_like(wp), impl=1)
    f = act(128, 8)
    yield lambda: ops.bilinear_bwd(f, CROP // 32, CROP // 32)


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    once = '--once' in sys.argv   # under ncu: a single launch per kernel keeps the report small
    for fn in kernels():
        if not once:
            fn()                  # warm-up launch
        torch.cuda.synchronize()
        flush.zero_()
        fn()                      # the launch to look at
        torch.cuda.synchronize()
    print('ok')


if __name__ == '__main__':
    main()

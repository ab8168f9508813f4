#!/usr/bin/env python
"""The kernels that are still behind gates, timed alone at the benchmark shapes
(12 x 768 x 768, bf16) next to the kernels they replace: CUDA-graph timing with an L2 flush like bench.py's kernel
table, one JSON line per op on stdout.  Every op runs under its own try/except so that one faulty kernel does
not hide the others; run it under `timeout` on the GPU box:

    timeout 600 python tools/experimental_kernels.py > gpurun_out/experimental_kernels.jsonl
    ncu --set full --clock-control none --import-source on -k regex:'pw_tc_bwd|dw_bwd_fused|stem_tc|bn_finalize_apply|ppm_|augment|dropout' \
        -o gpurun_out/exp python tools/experimental_kernels.py --once
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from torch_semantic_segmentation_b200 import _lib, ops  # noqa: E402
from torch_semantic_segmentation_b200.functional import _BnLink  # noqa: E402


def main():
    dry = '--dry' in sys.argv          # CPU dry run on the emulated ABI (tests/fake_backend.py): checks every call signature
    dev = torch.device('cpu' if dry else 'cuda:0')
    once = '--once' in sys.argv or dry
    bf = torch.bfloat16
    N, CROP = (2, 64) if dry else (bench.BATCH, bench.CROP)
    if dry:
        from tests.fake_backend import FakeBackend
        _lib.set_backend(FakeBackend())
    flush = None if dry else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    FH, FW, FCROP = (64, 128, (32, 48)) if dry else (1024, 2048, (512, 768))
    act = lambda C, div: ops.empty_nhwc(N, C, CROP // div, CROP // div, bf, dev).normal_()
    vec = lambda C: torch.rand(C, device=dev) + 0.5
    px = lambda div: N * (CROP // div) ** 2

    def run(name, nbytes, fn):
        row = {'op': name, 'algorithmic_bytes': nbytes}
        try:
            if once:
                fn()
                if not dry:
                    torch.cuda.synchronize()
            else:
                ms = bench.time_kernel(fn, 10, flush)
                row.update(ms=ms, gbs=nbytes / ms / 1e6)
        except Exception as e:      # noqa: BLE001 -- report and go on with the next op
            row['error'] = '%s: %s' % (type(e).__name__, str(e)[:200])
        print(json.dumps(row), flush=True)

    class FakeBN:
        _tss_dirty = False

    def link(C, div):
        scratch = torch.zeros(3 * C, dtype=torch.float64, device=dev)
        return _BnLink(act(C, div), vec(C) - 1, vec(C), vec(C), vec(C) - 1, True, scratch, FakeBN(), C)

    # ---- stem: SIMT vs tensor cores
    x = torch.randn(N, 3, CROP, CROP, device=dev)
    w = torch.randn(32, 3, 3, 3, device=dev) / 5
    st = torch.zeros(64, dtype=torch.float64, device=dev)
    dy2 = act(32, 2)
    b_stem = 4 * 3 * px(1) + 2 * 32 * px(2)
    run('stem_fwd (simt)', b_stem, lambda: ops.stem_fwd(x, w, bf, stats=st))
    run('stem_fwd_tc', b_stem, lambda: ops.stem_fwd_tc(x, w, stats=st))
    run('stem_wgrad (simt)', b_stem, lambda: ops.stem_wgrad(x, dy2, torch.zeros_like(w)))
    run('stem_wgrad_tc', b_stem, lambda: ops.stem_wgrad_tc(x, dy2, torch.zeros_like(w)))

    # ---- BatchNorm finalize + apply: two launches vs one
    for C, div in ((32, 2), (384, 8), (384, 16)):
        y = act(C, div)
        bn = torch.nn.BatchNorm2d(C).to(dev)
        scratch = torch.zeros(3 * C, dtype=torch.float64, device=dev)

        def two(y=y, bn=bn, scratch=scratch, C=C, div=div):
            sc, sh, _, _ = ops.bn_finalize(scratch, px(div), bn, 0.1, 1e-5, clear_n=3 * C, C=C)
            ops.bn_apply(y, sc, sh, relu=True)
        run('bn_finalize + bn_apply %dch@1/%d' % (C, div), 4 * C * px(div), two)
        run('bn_finalize_apply %dch@1/%d' % (C, div), 4 * C * px(div),
            lambda y=y, bn=bn, scratch=scratch, C=C, div=div: ops.bn_finalize_apply(scratch, px(div), bn, 0.1, 1e-5, y, relu=True,
                                                                                 clear_n=3 * C, C=C))

    # ---- pointwise backward: apply + dgrad(+reduction) vs the fused kernel
    for K, Nc, div in ((64, 384, 8), (64, 384, 16), (96, 576, 32), (128, 768, 32)):
        dz, y = act(Nc, div), act(Nc, div)
        mean, rstd, gamma, beta, sums = vec(Nc) - 1, vec(Nc), vec(Nc), vec(Nc) - 1, torch.randn(2 * Nc, device=dev)
        wpT = (torch.randn(K, Nc, device=dev) / Nc ** 0.5).to(bf)
        lk = link(K, div)
        nb = 2 * px(div) * (3 * Nc + 2 * K) + 2 * K * Nc        # read g, y; write dy; read yp; write g_in; weights

        def two(dz=dz, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, wpT=wpT, lk=lk):
            dy, _ = ops.bn_backward(dz, None, y, mean, rstd, gamma, False, beta=beta, sums=sums, prereduced=True)
            ops.pwconv_dgrad_bnred(dy, wpT, lk)
        run('bn_bwd_apply + pw dgrad_bnred %d<-%d@1/%d' % (K, Nc, div), nb + 2 * Nc * px(div), two)
        run('pwconv_bwd_fused %d<-%d@1/%d' % (K, Nc, div), nb,
            lambda dz=dz, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, wpT=wpT, lk=lk:
            ops.pwconv_bwd_fused(dz, y, mean, rstd, gamma, beta, sums, False, wpT, link=lk))

    # ---- depthwise backward: apply + dgrad_bnred vs the fused kernel
    for C, div in ((384, 16), (576, 32), (768, 32), (128, 8)):
        dz, y = act(C, div), act(C, div)
        mean, rstd, gamma, beta, sums = vec(C) - 1, vec(C), vec(C), vec(C) - 1, torch.randn(2 * C, device=dev)
        w = torch.randn(C, 1, 3, 3, device=dev) / 3
        lk = link(C, div)

        def two(dz=dz, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, w=w, lk=lk):
            dy, _ = ops.bn_backward(dz, None, y, mean, rstd, gamma, False, beta=beta, sums=sums, prereduced=True)
            ops.dwconv_dgrad_bnred(dy, w, lk)
        run('bn_bwd_apply + dw dgrad_bnred %dch@1/%d' % (C, div), 2 * C * px(div) * 6, two)
        run('dwconv_bwd_fused %dch@1/%d' % (C, div), 2 * C * px(div) * 5,
            lambda dz=dz, y=y, mean=mean, rstd=rstd, gamma=gamma, beta=beta, sums=sums, w=w, lk=lk:
            ops.dwconv_bwd_fused(dz, y, w, mean, rstd, gamma, beta, sums, False, lk))

    # ---- stride-2 depthwise dgrad with the fused reduction (producer = stem / first expand conv)
    for C, div in ((32, 2), (384, 8)):
        dy = act(C, div * 2)
        w = torch.randn(C, 1, 3, 3, device=dev) / 3
        lk = link(C, div)
        run('dw dgrad s2 + bn_bwd_reduce %dch@1/%d' % (C, div), 2 * C * (px(div) * 4 + px(div * 2)),
            lambda dy=dy, w=w, lk=lk, C=C, div=div: ops.bn_backward_reduce(ops.dwconv_dgrad(dy, w, CROP // div, CROP // div, 2, 1), lk.y, lk.mean,
                                                                          lk.rstd, lk.gamma, lk.beta, True, lk.sums))
        run('dwconv_dgrad_s2_bnred %dch@1/%d' % (C, div), 2 * C * (px(div) * 2 + px(div * 2)),
            lambda dy=dy, w=w, lk=lk: ops.dwconv_dgrad_s2_bnred(dy, w, lk))

    # ---- dropout, input pipeline, confusion matrix
    z = act(128, 8)
    ops.rng_state(dev, seed=1)
    run('dropout_fwd (own) 128ch@1/8', 4 * 128 * px(8), lambda: ops.dropout_fwd(z, 0.1))
    run('dropout (aten) 128ch@1/8', 4 * 128 * px(8), lambda: torch.nn.functional.dropout(z, 0.1, True))
    from torch_semantic_segmentation_b200.data import DeviceTransform
    frames = torch.randint(0, 256, (N, FH, FW, 3), dtype=torch.uint8, device=dev)
    ids = torch.randint(0, 35, (N, FH, FW), dtype=torch.uint8, device=dev)
    t = DeviceTransform(crop=FCROP, seed=0)
    geom = t.draw_geometry(N, FH, FW).to(dev)
    run('augment_batch %d x %dx%d -> %dx%d' % (N, FH, FW, FCROP[0], FCROP[1]), N * FCROP[0] * FCROP[1] * 20, lambda: t.apply(frames, ids, geom))
    pred = torch.randint(0, 19, (8 * FH * FW,), device=dev)
    lab = torch.randint(0, 19, (8 * FH * FW,), device=dev)
    cm = torch.zeros(19, 19, dtype=torch.int64, device=dev)
    for v in ('0', '1'):
        os.environ['TSS_CM_VARIANT'] = v
        run('confusion_from_labels variant %s (8 random maps)' % v, 16 * pred.numel(),
            lambda: _lib.call('tss_confusion_from_labels', pred=pred, target=lab, n=pred.numel(), C=19, cm=cm))
    os.environ.pop('TSS_CM_VARIANT', None)


if __name__ == '__main__':
    main()

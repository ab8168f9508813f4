"""Where does a CTA's time go?  Runs single kernels of the training step at the shapes of the low-resolution
bottlenecks (where the step is latency-bound, profiles/r2_step_timeline.txt) from the INSTRUMENTED library
(`make trace` -> libtss_b200_trace.so: TSS_MARK(slot) stores %globaltimer per CTA) inside a small CUDA graph
[BatchNorm apply -> kernel -> BatchNorm apply], as in the step, and prints per mark the median / min / max time since the
first CTA of the kernel started.

    make trace && python tools/trace_kernels.py [case ...] [--res 32|16|8]
"""
import argparse
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ['TSS_LIB'] = os.path.join(ROOT, 'torch_semantic_segmentation_b200', 'libtss_b200_trace.so')
sys.path.insert(0, ROOT)

import ctypes  # noqa: E402

import torch  # noqa: E402

SLOTS, MAX_CTAS = 16, 65536

MARKS = {
    'pw': {0: 'entry', 1: 'setup done (before griddepcontrol.wait)', 2: 'predecessor complete', 3: 'producer: last TMA issued',
           4: 'MMA: first stage landed', 5: 'MMA: accumulator committed', 10: 'epilogue: row of yp requested', 6: 'epilogue: accumulator ready',
           7: 'epilogue: tile stored', 8: 'CTA joined', 9: 'statistics atomics issued', 11: 'epilogue: first tcgen05.ld complete',
           12: 'epilogue: first 16 columns reduced', 13: 'epilogue: first 16 columns stored'},
    'dw': {0: 'entry', 1: 'setup done (before griddepcontrol.wait)', 2: 'predecessor complete', 3: 'weights requested', 4: 'tile 0 landed',
           5: 'tile 0 computed', 6: 'tile 0 stored', 7: 'tile 1 landed', 8: 'tile 1 computed', 9: 'tile 1 stored', 10: 'tile 2 landed',
           11: 'tile 2 computed', 12: 'tile 2 stored', 13: 'all tiles done', 14: 'statistics atomics issued'},
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('cases', nargs='*')
    ap.add_argument('--res', type=int, default=32)
    ap.add_argument('--batch', type=int, default=12)
    ap.add_argument('--crop', type=int, default=768)
    args = ap.parse_args()
    from torch_semantic_segmentation_b200 import _lib, ops
    lib = _lib.backend().lib
    lib.tss_trace_set.argtypes = [ctypes.c_void_p]
    dev = torch.device('cuda:0')
    N, H = args.batch, args.crop // args.res
    cin, cout = {32: (96, 96), 16: (64, 64), 8: (64, 64)}[args.res]
    ce = 6 * cin
    M = N * H * H
    bf = torch.bfloat16

    def act(c):
        return ops.as_nhwc(torch.randn(N, c, H, H, device=dev).to(bf).contiguous(memory_format=torch.channels_last))

    def weights(co, ci):
        w = torch.randn(co, ci, 1, 1, device=dev) * ci ** -0.5
        return (w,) + tuple(ops.pack_weights_bf16(w))

    def scratch(c):
        return torch.zeros(3 * c, dtype=torch.float64, device=dev)

    x_in, x_e, dy_out, dy_e = act(cin), act(ce), act(cout), act(ce)
    w1, w1p, w1t = weights(ce, cin)
    w2, w2p, w2t = weights(cout, ce)
    wd = torch.randn(ce, 1, 3, 3, device=dev) / 3
    s_e, s_o = scratch(ce), scratch(cout)
    affine = torch.rand(4, ce, device=dev) + 0.5
    link = types.SimpleNamespace(y=x_e, mean=affine[0], rstd=affine[1], gamma=affine[2], beta=affine[3], relu=True,
                                 sums=torch.zeros(2 * ce, dtype=torch.float32, device=dev))
    cases = {
        'pw1': ('pw', lambda: ops.pwconv_fwd(x_in, w1, stats=s_e, wp=w1p, impl=1)),
        'pw2': ('pw', lambda: ops.pwconv_fwd(x_e, w2, stats=s_o, wp=w2p, impl=1)),
        'pw2_dgrad_bnred': ('pw', lambda: ops.pwconv_dgrad_bnred(dy_out, w2t, link)),
        'pw1_dgrad': ('pw', lambda: ops.pwconv_dgrad(dy_e, w1, wpT=w1t, impl=1)),
        'dw': ('dw', lambda: ops.dwconv_fwd(x_e, wd, 1, 1, stats=s_e)),
        'dw_bnin': ('dw', lambda: ops.dwconv_fwd_bnin(x_e, affine[0], affine[1], True, wd, 1, s_e)),
        'dw_dgrad_bnred': ('dw', lambda: ops.dwconv_dgrad_bnred(dy_e, wd, link)),
    }
    names = args.cases or list(cases)
    trace = torch.zeros(MAX_CTAS * SLOTS, dtype=torch.int64, device=dev)
    big = act(ce)
    lib.tss_trace_set(ctypes.c_void_p(trace.data_ptr()))
    print('# %d x %d x %d maps (1/%d of %d^2), %d -> %d -> %d channels, M = %d rows' % (N, H, H, args.res, args.crop, cin, ce, cout, M))
    for name in names:
        family, fn = cases[name]
        for _ in range(3):                           # (the first launch of a translation unit binds its trace pointer with a
            ops.bn_apply(big, affine[0], affine[1], relu=True)       # synchronous copy: not something to do under capture)
            fn()
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            ops.bn_apply(big, affine[0], affine[1], relu=True)
            fn()
            ops.bn_apply(big, affine[0], affine[1], relu=True)
        rows = []
        for _ in range(5):
            trace.zero_()
            torch.cuda.synchronize()
            graph.replay()
            torch.cuda.synchronize()
            t = trace.view(MAX_CTAS, SLOTS).cpu()
            used = t[:, 0] > 0
            rows.append(t[used])
        t = rows[-1]
        n = t.shape[0]
        t0 = int(t[:, 0].min())
        ends = t[:, :SLOTS - 1].max()
        sms = len(set(t[:, SLOTS - 1].tolist()))
        print('\n== %s: %d CTAs on %d SMs, kernel span %.2f us (first entry -> last mark; 5 replays: %s)' % (
            name, n, sms, (int(ends) - t0) / 1e3,
            ' '.join('%.1f' % ((int(r[:, :SLOTS - 1].max()) - int(r[:, 0].min())) / 1e3) for r in rows)))
        print('   %-44s %8s %8s %8s   %s' % ('mark', 'median', 'min', 'max', 'median step from previous mark of the same CTA'))
        order = sorted(MARKS[family], key=lambda s: float(((t[:, s][t[:, s] > 0]).double().median()) if (t[:, s] > 0).any() else 1e30))
        prev = None
        for s in order:
            col = t[:, s]
            ok = col > 0
            if not ok.any():
                continue
            rel = (col[ok] - t0).double() / 1e3
            step = ''
            if prev is not None:
                both = ok & (t[:, prev] > 0)
                if both.any():
                    step = '%+.2f' % float(((col[both] - t[:, prev][both]).double() / 1e3).median())
            print('   %-44s %8.2f %8.2f %8.2f   %s' % (MARKS[family][s], float(rel.median()), float(rel.min()), float(rel.max()), step))
            prev = s
    lib.tss_trace_set(None)


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Benchmark of the hot path: Fast-SCNN training, bf16, 12 crops of 768x768 per GPU, softmax-CE
with ignore_index=255, data-parallel over N B200s (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` = images/s of the whole job with the batch resident in
HBM; `e2e` = the same metric through the public engine API (create_segmentation_trainer) with the
batch in pinned HOST memory (H2D copy + loss read-back inside the timed region); `roofline` =
the dominant kernel of the step timed alone with CUDA events; `cpu_baseline` = the CPU oracle
(a restatement of the reference's own stock-PyTorch path) on this box's host cores.
`--impl reference` times that CPU path alone, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import types
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'fastscnn_train_images_per_sec'
UNIT = 'img/s'
BATCH, CROP, CLASSES = 12, 768, 19
CPU_SAMPLE_BATCH = 2
WORKLOAD = ('Fast-SCNN training step (fwd + CE ignore_index=255 + bwd + grad all-reduce + AdamW), '
            '19 classes, 12 crops of 768x768 per GPU, random-init weights')


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of a CUDA graph')
    ap.add_argument('--kernels', action='store_true', help='also dump the per-kernel table to stderr')
    ap.add_argument('--step-only', action='store_true',
                    help='A/B aid: print the device-timed step alone (no end-to-end leg, no kernel table) and leave')
    ap.add_argument('--headline-only', action='store_true',
                    help='skip the 1024x2048 training leg, the 500-map mIoU leg and the bs1 inference leg (A/B runs)')
    ap.add_argument('--e2e-fp32', action='store_true',
                    help='end-to-end leg fed with fp32 crops + int64 labels (what the reference DataLoader yields after its '
                         'CPU transforms: 141 MB host->device per step) instead of decoded uint8 frames + uint8 label ids '
                         'normalised / mapped on the device by data.DeviceTransform (28 MB per step, the default)')
    ap.add_argument('--e2e-sync-loss', action='store_true',
                    help='end-to-end leg with loss.item() inside every step (engine.py:39) instead of the lazy read-back')
    return ap.parse_args()


def synthetic_batch(n, size, seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = torch.randn(n, 3, size, size, generator=g, device=device)
    y = torch.randint(0, CLASSES, (n, size, size), generator=g, device=device)
    y[torch.rand(n, size, size, generator=g, device=device) < 0.1] = 255
    return x, y


# ------------------------------------------------------------------ CPU arm -----------------
def cpu_train_throughput(steps, warmup, batch=CPU_SAMPLE_BATCH):
    """The reference's path (stock PyTorch fp32 on the host cores) restated by the oracle:
    forward (train mode) + CE(ignore 255) + backward + AdamW on `batch` crops of 768x768."""
    from oracle.init_state import init_state
    from oracle.train_step import AdamW, split_state, train_step
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = split_state(init_state('fastscnn', 0))
    opt = AdamW(sd, lr=1e-3, weight_decay=1e-5)
    x, y = synthetic_batch(batch, CROP, 1234, 'cpu')
    for _ in range(warmup):
        train_step('fastscnn', sd, opt, x, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        train_step('fastscnn', sd, opt, x, y)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, cores


def run_reference(args, rank):
    if rank != 0:
        return
    # bounded: a CPU step on the 2-crop sample takes 0.15-0.3 s; more than 40 steps add nothing but wall time.  The line
    # says what was asked for and what ran (config.requested / steps / warmup) instead of clamping silently.
    steps, warmup = max(1, min(args.steps, 40)), max(1, min(args.warmup, 5))
    value, spt, cores = cpu_train_throughput(steps, warmup)
    sample = ('%d steps of %d crops %dx%d (of the %d-crop step), fp32, %d torch threads'
              % (steps, CPU_SAMPLE_BATCH, CROP, CROP, BATCH, cores))
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': spt * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'global_batch': max(args.gpus, 1) * BATCH, 'parallelism': 'dp%d' % max(args.gpus, 1),
                   'cpu_sample': 'each step = %d of the %d crops of one rank, stock PyTorch fp32 on the host cores' % (CPU_SAMPLE_BATCH, BATCH),
                   'requested': {'steps': args.steps, 'warmup': args.warmup},
                   'ran': {'steps': steps, 'warmup': warmup, 'note': 'steps capped at 40, warm-up at 5 (CPU arm; img/s is per image, the sample size does not enter it)'}},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


# ------------------------------------------------------------------ clocks ------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------ roofline ----------------
def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs'
    except Exception:
        return 6650.0, 'fallback 6.65 TB/s (B200_PROFILING.md)'


def time_kernel(fn, iters=20, flush=None):
    """Average DEVICE time of `fn()` in ms.  `fn` (one or a few launches + their allocations) is
    captured into a CUDA graph together with an L2 flush (a 256 MB write) in front of it and the
    graph is replayed `iters` times between two CUDA events; the time of a flush-only graph,
    measured the same way, is subtracted.  Host launch overhead is therefore outside the number."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()

    def graph_ms(body):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            g.replay()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / iters

    if flush is None:
        return graph_ms(fn)
    both = graph_ms(lambda: (flush.zero_(), fn()))
    only = graph_ms(lambda: flush.zero_())
    return max(both - only, 1e-6)


def kernel_family(op):
    """The CUDA kernel function an op of the table runs (ops that launch the same kernel share a family)."""
    if op.startswith('pwconv_wgrad'):
        return 'wgrad_tc_kernel (pointwise wgrad, tcgen05)'
    if op.startswith('pwconv_dgrad_bnred'):
        return 'pw_tc_bnred_kernel (pointwise dgrad + BatchNorm-backward reduction, tcgen05)'
    if op.startswith('dwconv_dgrad_bnred') and ' s2 ' in op:
        return 'dw_dgrad_s2_bnred_kernel (depthwise stride-2 dgrad + BatchNorm-backward reduction)'
    if op.startswith('dwconv_dgrad_bnred'):
        return 'dw_dgrad_bnred_persistent_kernel (depthwise stride-1 dgrad + BatchNorm-backward reduction)'
    if op.startswith('pwconv_'):
        return 'pw_tc_kernel (pointwise fwd + dgrad, tcgen05)'
    if op.startswith('dwconv_wgrad'):
        return 'dw_wgrad_tma_kernel (depthwise wgrad)'
    if op.startswith('dwconv_dgrad') and ' s2 ' in op:
        return 'dw_dgrad_s2_quad_kernel (depthwise stride-2 dgrad)'
    if op.startswith('dwconv_'):
        return 'dw_tma_kernel (depthwise fwd + stride-1 dgrad)'
    if op.startswith('bn_bwd_reduce'):
        return 'bn_bwd_reduce_kernel (BatchNorm backward, stand-alone reduction)'
    if op.startswith('bn_bwd_onepass'):
        return 'bn_bwd_onepass_kernel (BatchNorm backward, both passes in one launch)'
    if op.startswith('bn_bwd_apply'):
        return 'bn_bwd_apply_kernel (BatchNorm backward apply)'
    if op.startswith('bn_apply'):
        return 'bn_apply_kernel (BatchNorm forward apply)'
    return op


def kernel_table(device, runner=None):
    """Per-launch time and algorithmic bytes (SURVEY.md section 8d formulas) of the kernels of one
    training step at the TRN shapes, each timed alone.  Returns rows sorted by share of the step.
    `runner(name, launches_per_step, algorithmic_bytes, fn)` replaces the timing (tools/run_kernels.py
    launches every op once under ncu with it)."""
    from torch_semantic_segmentation_b200 import ops
    bf = torch.bfloat16
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    N = BATCH
    rows = []

    def act(C, div):
        return ops.empty_nhwc(N, C, CROP // div, CROP // div, bf, device).normal_()

    def add(name, count, nbytes, fn):
        if runner is not None:
            runner(name, count, nbytes, fn)
            return
        ms = time_kernel(fn, 10, flush)
        rows.append({'kernel': name, 'launches_per_step': count, 'ms': ms, 'bytes': nbytes,
                     'gbs': nbytes / ms / 1e6, 'share_ms': ms * count})

    px = lambda div: N * (CROP // div) ** 2

    def link(y, C):
        """The BatchNorm in front of a fused dgrad: its raw input, batch statistics, affine, ReLU, reduction target."""
        a = torch.rand(4, C, device=device) + 0.5
        return types.SimpleNamespace(y=y, mean=a[0] - 1.0, rstd=a[1], gamma=a[2], beta=a[3] - 1.0, relu=True,
                                     sums=torch.zeros(2 * C, dtype=torch.float32, device=device))
    # final x8 up-sampling of the class scores + fused CE + its backward (the 269 MB logits tensor)
    small = ops.empty_nhwc(N, CLASSES, CROP // 8, CROP // 8, bf, device, pitch=32).normal_()
    target = torch.randint(0, CLASSES, (N, CROP, CROP), device=device)
    add('upsample_logits_fwd', 1, 2 * CLASSES * (px(8) + px(1)), lambda: ops.upsample_logits_fwd(small, CROP, CROP))
    # fused head (x8 interpolation + CE + gradient w.r.t. the 1/8 scores): reads the int64 labels and the
    # small scores, writes the small fp32 gradient accumulator + its bf16 copy
    add('upsample_ce (fused head, 3 launches)', 1, 8 * px(1) + (2 * 32 + 4 * 24 * 2 + 2 * 24) * px(8),
        lambda: ops.upsample_ce_forward(small, target, CROP, CROP, 255, True))
    # stem
    x = torch.randn(N, 3, CROP, CROP, device=device)
    w = torch.randn(32, 3, 3, 3, device=device)
    st = torch.zeros(64, dtype=torch.float64, device=device)
    y2 = act(32, 2)
    add('stem3x3s2_fwd (tcgen05)', 1, 4 * 3 * px(1) + 2 * 32 * px(2), lambda: ops.stem_fwd_tc(x, w, stats=st))
    add('stem3x3s2_patches', 1, 4 * 3 * px(1) + 2 * 32 * px(2), lambda: ops.stem_patches(x))
    patches = ops.stem_patches(x)
    add('stem3x3s2_wgrad_from_patches', 1, 2 * 2 * 32 * px(2), lambda: ops.stem_wgrad_from_patches(patches, y2, torch.zeros_like(w)))
    # BatchNorm passes: forward apply (reads y, writes z), backward reduce (reads dz, y), backward apply (reads dz, y, writes
    # dy; the ReLU mask is recomputed from y) -- one row per kernel and shape (channels, level, layers of that shape per
    # step; the reduce of 26 of the 44 layers is folded into the consumer's dgrad epilogue and does not appear here)
    for C, div, cnt, cnt_red in [(32, 2, 1, 0), (32, 4, 1, 0), (48, 4, 1, 1), (48, 8, 1, 0), (64, 8, 1, 1), (128, 8, 7, 4), (384, 8, 1, 0),
                                 (384, 16, 6, 0), (64, 16, 3, 3), (384, 32, 1, 0), (576, 32, 6, 0), (96, 32, 3, 3), (768, 32, 4, 0),
                                 (128, 32, 4, 4)]:
        yb, gb = act(C, div), act(C, div)
        scb = torch.ones(C, device=device)
        sums = torch.zeros(2 * C, device=device)
        add('bn_apply %dch@1/%d' % (C, div), cnt, 2 * 2 * C * px(div), lambda: ops.bn_apply(yb, scb, scb, relu=True))
        # layers that run their own reduction: one launch for both passes where dz + y stay in L2 (gate BN_BWD_ONEPASS);
        # the row is timed either way (0 launches per step when the gate is off: tools/time_ops.py A/B)
        onepass = bool(cnt_red) and ops.BN_BWD_ONEPASS and 2 * 2 * C * px(div) <= ops._ONEPASS_MAX_BYTES
        if cnt_red:
            add('bn_bwd_reduce %dch@1/%d' % (C, div), 0 if onepass else cnt_red, 2 * 2 * C * px(div),
                lambda: ops.bn_backward_reduce(gb, yb, scb, scb, scb, scb, True, sums))

            def both_passes():
                keep, ops.BN_BWD_ONEPASS = ops.BN_BWD_ONEPASS, True
                try:
                    ops.bn_backward(gb, None, yb, scb, scb, scb, True, beta=scb, sums=sums)
                finally:
                    ops.BN_BWD_ONEPASS = keep
            if 2 * 2 * C * px(div) <= ops._ONEPASS_MAX_BYTES:
                add('bn_bwd_onepass %dch@1/%d' % (C, div), cnt_red if onepass else 0, 2 * 3 * C * px(div), both_passes)
        add('bn_bwd_apply %dch@1/%d' % (C, div), cnt - (cnt_red if onepass else 0), 2 * 3 * C * px(div),
            lambda: ops.bn_backward(gb, None, yb, scb, scb, scb, False, beta=scb, sums=sums, prereduced=True))
    # depthwise: every shape of the network (channels, input level, stride, dilation, layers of that shape)
    for C, div, s, d, cnt in [(32, 2, 2, 1, 1), (48, 4, 2, 1, 1), (384, 8, 2, 1, 1), (384, 16, 1, 1, 2), (384, 16, 2, 1, 1),
                              (576, 32, 1, 1, 3), (768, 32, 1, 1, 2), (128, 8, 1, 4, 1), (128, 8, 1, 1, 2)]:
        xi = act(C, div)
        wd = torch.randn(C, 1, 3, 3, device=device)
        sd = torch.zeros(2 * C, dtype=torch.float64, device=device)
        yo = act(C, div * s)             # a gradient of the output's shape
        io = 2 * C * (px(div) + px(div * s))
        add('dwconv_fwd C%d s%d d%d @1/%d' % (C, s, d, div), cnt, io, lambda: ops.dwconv_fwd(xi, wd, s, d, stats=sd))
        if d == 1:
            # what the step launches for these layers: the dgrad with the BatchNorm-backward reduction of the layer in front
            # fused in (reads the gradient and that layer's raw output, writes the masked gradient)
            lk = link(xi, C)
            fused = ops.dwconv_dgrad_s2_bnred if s == 2 else ops.dwconv_dgrad_bnred
            add('dwconv_dgrad_bnred C%d s%d d%d @1/%d' % (C, s, d, div), cnt, io + 2 * C * px(div), lambda: fused(yo, wd, lk))
        else:
            add('dwconv_dgrad C%d s%d d%d @1/%d' % (C, s, d, div), cnt, io, lambda: ops.dwconv_dgrad(yo, wd, xi.shape[2], xi.shape[3], s, d))
        add('dwconv_wgrad C%d s%d d%d @1/%d' % (C, s, d, div), cnt, io, lambda: ops.dwconv_wgrad(xi, yo, torch.zeros_like(wd), s, d))
    # pointwise GEMMs: the tcgen05/TMEM/TMA kernels the bf16 model runs (impl 1), every shape of the network
    # (K -> N, level, layers of that shape)
    for K, Nc, div, cnt in [(32, 48, 4, 1), (48, 64, 8, 1), (64, 384, 8, 1), (384, 64, 16, 3), (64, 384, 16, 3), (384, 96, 32, 1),
                            (96, 576, 32, 3), (576, 96, 32, 2), (576, 128, 32, 1), (128, 768, 32, 2), (768, 128, 32, 2),
                            (256, 128, 32, 1), (128, 128, 8, 3), (64, 128, 8, 1)]:
        xi = act(K, div)
        wp = torch.randn(Nc, K, 1, 1, device=device) * 0.05
        yo = act(Nc, div)
        sp = torch.zeros(2 * Nc, dtype=torch.float64, device=device)
        pk = ops.pack_weights_bf16(wp)
        dwp = torch.zeros_like(wp)
        io = 2 * px(div) * (K + Nc)
        add('pwconv_fwd %d->%d @1/%d' % (K, Nc, div), cnt, io + 2 * K * Nc, lambda: ops.pwconv_fwd(xi, wp, stats=sp, wp=pk[0], impl=1))
        if (K, Nc, div) in FUSED_PW_DGRAD:
            # the conv reads a depthwise conv's BatchNorm+ReLU output: its dgrad carries that BatchNorm's backward reduction
            lk = link(xi, K)
            add('pwconv_dgrad_bnred %d->%d @1/%d' % (K, Nc, div), cnt, io + 2 * px(div) * K + 2 * K * Nc, lambda: ops.pwconv_dgrad_bnred(yo, pk[1], lk))
        else:
            add('pwconv_dgrad %d->%d @1/%d' % (K, Nc, div), cnt, io + 2 * K * Nc, lambda: ops.pwconv_dgrad(yo, wp, wpT=pk[1], impl=1))
        add('pwconv_wgrad %d->%d @1/%d' % (K, Nc, div), cnt, io + 4 * K * Nc, lambda: ops.pwconv_wgrad(xi, yo, dwp, impl=1))
    rows.sort(key=lambda r: -r['share_ms'])
    return rows


# pointwise convs (K -> Nc @ level) whose dgrad runs pw_tc_bnred_kernel in the step (14 launches: the project convs of the
# bottlenecks, the pointwise halves of the separable convs)
FUSED_PW_DGRAD = {(32, 48, 4), (48, 64, 8), (384, 64, 16), (384, 96, 32), (576, 96, 32), (576, 128, 32), (768, 128, 32), (128, 128, 8)}


# ------------------------------------------------------------------ the metric's other legs --
STEP_ALGORITHMIC_BYTES = 3 * 1849e6 + 595e6      # 3 x the fused-ideal forward traffic + the CE head, per 12 x 768 x 768 step
FWD_ALGORITHMIC_BYTES_1024x2048 = 549e6          # fused-ideal inference forward per 1024 x 2048 image


def train_1024x2048(world, rank, device, steps, batch=8):
    """BASELINE.json `metric`: Fast-SCNN training at the full 1024 x 2048 resolution (per-GPU batch `batch`, stated in the
    result; 8 x 1024 x 2048 = 16.8 Mpx per step against 7.1 Mpx for configs[1]'s 12 crops -- the batch BASELINE.json's
    configs[2] uses for ContextNet at this resolution), same step as the headline."""
    import torch.distributed as dist
    from torch_semantic_segmentation_b200.distributed import GradientAllReducer, broadcast_parameters
    from torch_semantic_segmentation_b200.engine import GraphedTrainStep
    from torch_semantic_segmentation_b200.functional import enable_deferred_logits
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    torch.manual_seed(0)
    model = fastscnn(3, CLASSES).to(device).set_compute_dtype(torch.bfloat16)
    broadcast_parameters(model)
    opt = FlatAdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    GradientAllReducer(opt, num_buckets=4).install()
    loss_fn = CrossEntropyLoss(ignore_index=255)
    g = torch.Generator(device=device).manual_seed(4321 + rank)
    x = torch.randn(batch, 3, 1024, 2048, generator=g, device=device)
    y = torch.randint(0, CLASSES, (batch, 1024, 2048), generator=g, device=device)
    y[torch.rand(batch, 1024, 2048, generator=g, device=device) < 0.1] = 255
    model.train()
    enable_deferred_logits(model, loss_fn)
    step = GraphedTrainStep(model, opt, loss_fn, x, y)
    for _ in range(3):
        step.graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        step.graph.replay()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    per = float(ms) / steps
    out = {'value': world * batch / (per / 1e3), 'unit': 'img/s', 'ms_per_step': per, 'per_gpu_batch': batch,
           'resolution': [1024, 2048], 'steps': steps, 'loss': float(step.loss), 'dtype': 'bf16'}
    del step, model, opt
    torch.cuda.empty_cache()
    return out


def miou_500_exact(world, rank, device, n_maps=500):
    """BASELINE.json configs[3]: confusion matrix over 500 synthetic 1024 x 2048 prediction / label maps, sharded over the
    ranks without padding, ONE int64 all-reduce; every rank's shard is recounted on the host by the oracle
    (oracle/confusion.py: numpy bincount, the definition) and compared bit for bit, and so is the mIoU (float64)."""
    import numpy as np
    import torch.distributed as dist
    from oracle import confusion as o_cm
    from torch_semantic_segmentation_b200.distributed import shard_range
    from torch_semantic_segmentation_b200.metrics import ConfusionMatrix, metrics_from_cm
    lo, hi = shard_range(n_maps, world, rank)

    def pair(i):
        g = torch.Generator(device=device).manual_seed(4321 + i)
        p = torch.randint(0, CLASSES, (1024, 2048), generator=g, device=device)
        l = torch.randint(0, CLASSES, (1024, 2048), generator=g, device=device)
        l[torch.rand(1024, 2048, generator=g, device=device) < 0.1] = 255
        return p, l
    pairs = [pair(i) for i in range(lo, hi)]
    batches = [(torch.stack([p for p, _ in pairs[i:i + 16]]), torch.stack([l for _, l in pairs[i:i + 16]]))
               for i in range(0, len(pairs), 16)]
    cm = ConfusionMatrix(CLASSES, device=device)
    cm.update(batches[0])
    cm.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for b in batches:
        cm.update(b)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    local = cm.compute(sync=False).clone()
    total = cm.compute()                       # one int64 all-reduce (SUM)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    want = np.zeros((CLASSES, CLASSES), dtype=np.int64)
    for p, l in pairs:                          # the oracle on the host, on the very maps the GPU counted
        want += o_cm.confusion_matrix(p.cpu().numpy(), l.cpu().numpy(), CLASSES)
    ok = torch.tensor([int(np.array_equal(local.cpu().numpy(), want))], device=device)
    want_total = torch.from_numpy(want).to(device)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        dist.all_reduce(want_total)
    miou = float(metrics_from_cm(total)['miou'])
    miou_oracle = float(o_cm.metrics(want_total.cpu().numpy())['miou'])
    px = n_maps * 1024 * 2048
    return {'maps': n_maps, 'ms': float(ms), 'maps_per_s': n_maps / (float(ms) / 1e3), 'gbs_per_gpu': 16.0 * px / world / float(ms) / 1e6,
            'hbm_frac': 16.0 * px / world / float(ms) / 1e6 / measured_peak()[0],
            'confusion_matrix_bit_exact_vs_oracle': bool(int(ok)) and bool(torch.equal(total.cpu(), want_total.cpu())),
            'miou': miou, 'miou_oracle': miou_oracle, 'miou_bit_exact': miou == miou_oracle,
            'pixels_counted': int(total.sum()), 'shards': world}


def infer_bs1(device, iterations=200, warmup=20):
    """BASELINE.json `metric`: bs1 inference FPS at 1 x 3 x 1024 x 2048 (eval mode, bf16, CUDA-graphed forward).
    `fps` follows the reference's benchmark_model (utils/benchmark.py:6-29: per-iteration host time, fps = 1 / mean) with
    a synchronize inside every iteration (the reference's loop would time asynchronous launches on a GPU);
    `fps_device` = back-to-back replays between two CUDA events."""
    import numpy as np
    from torch_semantic_segmentation_b200 import _lib
    from torch_semantic_segmentation_b200.models import fastscnn
    torch.manual_seed(0)
    model = fastscnn(3, CLASSES).to(device).set_compute_dtype(torch.bfloat16).eval()
    x = torch.randn(1, 3, 1024, 2048, device=device)
    with torch.no_grad():
        for _ in range(3):
            model(x)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(graph):
            out = model(x)
        kernels = _lib.launch_count() - before
    for _ in range(warmup):
        graph.replay()
    torch.cuda.synchronize()
    record = np.zeros(iterations)
    for it in range(iterations):
        t0 = time.perf_counter()
        graph.replay()
        torch.cuda.synchronize()
        record[it] = time.perf_counter() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iterations):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / iterations
    peak = measured_peak()[0]
    res = {'fps': float(1.0 / record.mean()), 'mean_ms': float(record.mean() * 1e3), 'min_ms': float(record.min() * 1e3),
           'fps_device': 1e3 / dev_ms, 'device_ms': dev_ms, 'iterations': iterations, 'kernels_per_forward': kernels,
           'hbm_ideal_frac': (FWD_ALGORITHMIC_BYTES_1024x2048 / (peak * 1e9)) / (dev_ms / 1e3),
           'out_shape': list(out.shape), 'dtype': 'bf16'}
    del graph, out, model
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------ our arm -----------------
def main():
    args = parse()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product has no CPU path')
    import torch.distributed as dist
    from torch_semantic_segmentation_b200 import _lib
    from torch_semantic_segmentation_b200.distributed import GradientAllReducer, broadcast_parameters
    from torch_semantic_segmentation_b200.engine import GraphedTrainStep, create_segmentation_trainer
    from torch_semantic_segmentation_b200.functional import enable_deferred_logits as Fn_enable_deferred, unit_loss_grad
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    from torch_semantic_segmentation_b200.optim import FlatAdamW

    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', init_method='env://', device_id=device)
    torch.manual_seed(0)
    model = fastscnn(3, CLASSES).to(device).set_compute_dtype(torch.bfloat16)
    broadcast_parameters(model)
    opt = FlatAdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    reducer = GradientAllReducer(opt, num_buckets=4).install()
    loss_fn = CrossEntropyLoss(ignore_index=255)
    x, y = synthetic_batch(BATCH, CROP, 1234 + rank, device)
    model.train()

    use_graph = not args.no_graph
    graphed = None

    def eager_step():
        opt.zero_grad()
        out = model(x)
        loss = loss_fn(out, y)
        with unit_loss_grad(loss):
            loss.backward()
        opt.step()
        return loss.detach()

    Fn_enable_deferred(model, loss_fn)                # gated (TSS_DEFER_LOGITS=1): no full-resolution logits in training
    if use_graph:
        graphed = GraphedTrainStep(model, opt, loss_fn, x, y)      # static inputs = the resident batch

        def step():
            graphed.graph.replay()
            return graphed.loss
    else:
        step = eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = _lib.launch_count() - launches0
    if use_graph:
        launches = graphed.kernels_per_step * args.steps
    ms_total = float(ms)
    value = world * BATCH * args.steps / (ms_total / 1e3)

    if args.step_only:
        if rank == 0:
            print(json.dumps({'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                              'ms_per_step': ms_total / args.steps, 'gpu_launches': launches, 'step_only': True,
                              'gates': sorted('%s=%s' % (k, v) for k, v in os.environ.items() if k.startswith('TSS_'))}))
        sampler.stop() if rank == 0 else None
        sys.stdout.flush()
        if world > 1:
            os._exit(0)
        return

    # ---- end to end: the public engine API, batch in pinned host memory -------------------
    e2e_input = 'fp32 crops + int64 labels (what the reference DataLoader yields)'
    if not args.e2e_fp32:
        # the same crops as decoded uint8 frames + Cityscapes label ids whose train ids are exactly y
        from torch_semantic_segmentation_b200.data import DeviceTransform, TRAIN_MAPPING
        inverse = torch.zeros(256, dtype=torch.uint8)
        for label_id, train_id in enumerate(TRAIN_MAPPING.tolist()):
            if train_id != 255:
                inverse[train_id] = label_id
        gen = torch.Generator().manual_seed(99 + rank)
        xh = torch.randint(0, 256, (BATCH, CROP, CROP, 3), dtype=torch.uint8, generator=gen).pin_memory()
        yh = inverse[y.cpu().clamp(max=255)].pin_memory()               # 255 (ignore) -> id 0 ('unlabeled' -> 255)
        transform = DeviceTransform(crop=None, scale_limit=None, flip_p=0.0)
        h2d_bytes = xh.numel() + yh.numel()
        e2e_input = 'uint8 frames + uint8 label ids, normalise + label table on the device'
    else:
        xh, yh = x.cpu().pin_memory(), y.cpu().pin_memory()
        transform = None
        h2d_bytes = xh.numel() * 4 + yh.numel() * 8
    trainer = create_segmentation_trainer(model, opt, loss_fn, device, use_f16=True, logging=False,
                                          cuda_graph=use_graph, transform=transform, lazy_loss=not args.e2e_sync_loss)
    trainer.run([(xh, yh)] * 2)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    trainer.run([(xh, yh)] * args.steps)
    t1.record()
    barrier()
    ems = torch.tensor([t0.elapsed_time(t1)], device=device)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = world * BATCH * args.steps / (float(ems) / 1e3)

    # ---- the other legs of BASELINE.json's metric: every rank takes part (data-parallel / sharded) ----------------
    del trainer, graphed
    torch.cuda.empty_cache()
    extra = {}
    if not args.headline_only:
        extra['train_1024x2048'] = train_1024x2048(world, rank, device, max(5, min(args.steps, 20)))
        extra['miou_500_exact'] = miou_500_exact(world, rank, device)

    def finish():
        # All collectives of this run are behind us (the last one is the MAX of the e2e time) and the
        # result line is printed: multi-rank processes leave without the NCCL teardown handshake, which
        # can block for minutes when the ranks arrive far apart (rank 0 still runs its single-rank
        # roofline / CPU legs) -- a stuck teardown must never outlive the measurement.
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            # the other ranks stay alive (CPU-side wait on the rendezvous store) until rank 0 is done, so
            # that no NCCL peer disappears under a rank that is still working
            import datetime
            try:
                store = dist.distributed_c10d._get_default_store()
                if rank == 0:
                    store.set('tss_bench_done', '1')
                    time.sleep(0.5)
                else:
                    store.wait(['tss_bench_done'], datetime.timedelta(seconds=240))
            except Exception:
                pass
            os._exit(0)

    if rank != 0:
        finish()
        return

    # ---- roofline of the dominant kernel (timed alone, L2 flushed) -------------------------
    peak, peak_src = measured_peak()
    rows = kernel_table(device)
    if args.kernels:
        for r in rows:
            print('%-44s x%d  %8.3f ms  %8.1f MB  %7.0f GB/s  share %7.3f ms' % (
                r['kernel'], r['launches_per_step'], r['ms'], r['bytes'] / 1e6, r['gbs'], r['share_ms']), file=sys.stderr)
    # `roofline` = the single (kernel, shape) row with the largest share of the step (time per launch x launches of that
    # shape per step), timed alone in this run: algorithmic bytes of ONE launch / its device time.  `kernel_families`
    # = the same table summed per kernel function (what the ncu launch lists under profiles/ rank).
    fams = {}
    for r in rows:
        if r['launches_per_step'] == 0:      # timed for the A/B table only (a gated-off alternative): not part of the step
            continue
        f = fams.setdefault(kernel_family(r['kernel']), {'launches': 0, 'ms': 0.0, 'bytes': 0.0})
        f['launches'] += r['launches_per_step']
        f['ms'] += r['ms'] * r['launches_per_step']
        f['bytes'] += r['bytes'] * r['launches_per_step']
    families = [{'kernel': k, 'launches_per_step': v['launches'], 'ms_per_step': v['ms'], 'gbs': v['bytes'] / v['ms'] / 1e6,
                 'frac': v['bytes'] / v['ms'] / 1e6 / peak} for k, v in sorted(fams.items(), key=lambda kv: -kv[1]['ms'])]
    # the dominant kernel FUNCTION of the step (largest summed time over its launches), and of its launches the shape with
    # the largest share: `roofline` describes that one launch; the family's average over all its shapes goes along
    dom = families[0]
    top = max((r for r in rows if kernel_family(r['kernel']) == dom['kernel']), key=lambda r: r['share_ms'])
    traffic = None        # dram__bytes_read.sum + dram__bytes_write.sum of that launch from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, 'profiles', 'kernel_traffic.json')) as f:
            tr = json.load(f)
        if top['kernel'] in tr:
            traffic = tr[top['kernel']]['dram_bytes']
    except Exception:
        traffic = None
    roofline = {'bound': 'hbm', 'kernel': '%s: %s' % (dom['kernel'], top['kernel']), 'achieved': top['gbs'], 'peak': peak,
                'unit': 'GB/s', 'frac': top['gbs'] / peak, 'traffic': traffic, 'traffic_source': 'profiles/kernel_traffic.json (ncu --set full)',
                'peak_source': peak_src, 'launch_ms': top['ms'], 'algorithmic_bytes': top['bytes'],
                'launches_per_step': top['launches_per_step'], 'share_of_step_ms': top['share_ms'],
                'family': {'launches_per_step': dom['launches_per_step'], 'ms_per_step': dom['ms_per_step'], 'gbs': dom['gbs'], 'frac': dom['frac']},
                'largest_single_launch': {'kernel': rows[0]['kernel'], 'ms': rows[0]['ms'], 'gbs': rows[0]['gbs'], 'frac': rows[0]['gbs'] / peak,
                                          'note': 'the fused head is bound by its softmax arithmetic (19 ex2 + ~130 FP32 ops per output pixel), not by HBM'
                                          if rows[0]['kernel'].startswith('upsample_ce') else ''},
                'timing': 'this run: the launch alone in a CUDA graph behind a 256 MB L2 flush, 10 replays between CUDA events, flush-only graph subtracted'}

    if not args.headline_only:
        extra['infer_bs1'] = infer_bs1(device)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, spt, cores = cpu_train_throughput(3, 1)
        cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': '3 steps of %d crops %dx%d (of the %d-crop step), fp32 oracle, %d torch threads'
                         % (CPU_SAMPLE_BATCH, CROP, CROP, BATCH, cores)}

    print(json.dumps({
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': WORKLOAD,
                   'global_batch': world * BATCH, 'parallelism': 'dp%d' % world,
                   'l2': 'per-step working set (>3 GB of activations) far exceeds the 126 MB L2',
                   'launch_mode': 'cuda_graph' if use_graph else 'eager',
                   'gates': sorted(k for k, v in os.environ.items() if k.startswith('TSS_') and v not in ('', '0'))},
        'loss': float(loss),
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d_bytes,
                'd2h_bytes_per_step': 4, 'ms_per_step': float(ems) / args.steps, 'input': e2e_input,
                'loss_readback': 'loss.item() every step' if args.e2e_sync_loss else
                                 'every step reads one loss back from pinned memory: the previous step\'s (lazy_loss=True)',
                'api': 'create_segmentation_trainer(...).run(loader)'},
        'gpu_launches': launches,
        'clocks': clocks,
        'roofline': roofline,
        'kernel_families': families[:8],
        'step_roofline': {'algorithmic_bytes': STEP_ALGORITHMIC_BYTES, 'source': 'SURVEY.md section 8d: 3 x 1849 MB + 595 MB per step',
                          'achieved': STEP_ALGORITHMIC_BYTES / (ms_total / args.steps) / 1e6, 'peak': peak, 'unit': 'GB/s',
                          'frac': STEP_ALGORITHMIC_BYTES / (ms_total / args.steps) / 1e6 / peak},
        'cpu_baseline': cpu,
        **extra,
    }))
    finish()


if __name__ == '__main__':
    main()

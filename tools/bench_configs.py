#!/usr/bin/env python
"""The other BASELINE.json configurations (bench.py covers configs[1], the headline metric):

  --config 3   ContextNet-14 training, bf16, batch 8 of 1024x2048 per GPU
  --config 4   mIoU evaluation: confusion matrix over 500 synthetic 1024x2048 prediction/label maps,
               sharded over the ranks without padding, ONE int64 all-reduce, compared bit-for-bit
               with a host bincount on a sample of the maps (every map with --check-all)
  --config 5   Fast-SCNN inference batch sweep 1..64 at 1024x2048 (bf16, eval mode, CUDA graph)

One JSON line per measurement (rank 0).  Launch with torchrun for more than one GPU.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from torch_semantic_segmentation_b200 import _lib  # noqa: E402
from torch_semantic_segmentation_b200.distributed import (GradientAllReducer, broadcast_parameters,  # noqa: E402
                                                          shard_range)
from torch_semantic_segmentation_b200.engine import GraphedTrainStep  # noqa: E402
from torch_semantic_segmentation_b200.losses import CrossEntropyLoss  # noqa: E402
from torch_semantic_segmentation_b200.metrics import ConfusionMatrix, metrics_from_cm  # noqa: E402
from torch_semantic_segmentation_b200.models import contextnet14, fastscnn  # noqa: E402
from torch_semantic_segmentation_b200.optim import FlatAdamW  # noqa: E402

CLASSES = 19


def timed(fn, steps, world, device):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def config3(args, rank, world, device):
    batch, H, W = args.batch or 8, 1024, 2048
    torch.manual_seed(0)
    model = contextnet14(3, CLASSES).to(device).set_compute_dtype(torch.bfloat16)
    broadcast_parameters(model)
    opt = FlatAdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    GradientAllReducer(opt, num_buckets=4).install()
    g = torch.Generator(device=device).manual_seed(1234 + rank)
    x = torch.randn(batch, 3, H, W, generator=g, device=device)
    y = torch.randint(0, CLASSES, (batch, H, W), generator=g, device=device)
    y[torch.rand(batch, H, W, generator=g, device=device) < 0.1] = 255
    model.train()
    loss_fn = CrossEntropyLoss(ignore_index=255)
    from torch_semantic_segmentation_b200.functional import enable_deferred_logits
    enable_deferred_logits(model, loss_fn)         # gated (TSS_DEFER_LOGITS=1), like bench.py
    step = GraphedTrainStep(model, opt, loss_fn, x, y)
    for _ in range(3):
        step.graph.replay()
    ms = timed(step.graph.replay, args.steps, world, device)
    if rank == 0:
        print(json.dumps({'config': 3, 'metric': 'contextnet14_train_images_per_sec', 'unit': 'img/s',
                          'value': world * batch * args.steps / (ms / 1e3), 'ms_per_step': ms / args.steps,
                          'n_gpus': world, 'per_gpu_batch': batch, 'resolution': [H, W], 'dtype': 'bf16',
                          'loss': float(step.loss), 'kernels_per_step': step.kernels_per_step,
                          'gates': sorted(k for k, v in os.environ.items() if k.startswith('TSS_') and v not in ('', '0'))}))


def synth_map(i, device):
    g = torch.Generator(device='cpu').manual_seed(4321 + i)
    pred = torch.randint(0, CLASSES, (1024, 2048), generator=g)
    label = torch.randint(0, CLASSES, (1024, 2048), generator=g)
    label[torch.rand(1024, 2048, generator=g) < 0.1] = 255
    return pred.to(device), label.to(device)


def config4(args, rank, world, device):
    def host_cm(i):          # the definition, on the host: bincount(C*y[m] + p[m]) over 0 <= y < C
        p, l = synth_map(i, 'cpu')
        m = (l >= 0) & (l < CLASSES)
        return torch.bincount(CLASSES * l[m] + p[m], minlength=CLASSES * CLASSES).view(CLASSES, CLASSES)

    n_maps = args.maps
    lo, hi = shard_range(n_maps, world, rank)
    pairs = [synth_map(i, device) for i in range(lo, hi)]       # staged in HBM: 32 MB per pair
    # evaluation batches of up to 16 maps (one kernel launch per batch, like a val loader with batch 16)
    maps = [(torch.stack([p for p, _ in pairs[i:i + 16]]), torch.stack([l for _, l in pairs[i:i + 16]]))
            for i in range(0, len(pairs), 16)]
    del pairs
    cm = ConfusionMatrix(CLASSES, device=device)
    for p, l in maps[:1]:
        cm.update((p, l))
    cm.reset()

    def run():
        for p, l in maps:
            cm.update((p, l))
    ms = timed(run, 1, world, device)
    total = cm.compute()                       # one int64 all-reduce (SUM) across the ranks
    metrics = metrics_from_cm(total)
    if rank == 0:
        # oracle on the host: the whole set with --check-all, else the first 8 maps against a
        # single-process recount of the same maps on the GPU
        idx = range(n_maps) if args.check_all else range(min(8, n_maps))
        want = sum(host_cm(i) for i in idx)
        ref = ConfusionMatrix(CLASSES, device=device)
        for i in idx:
            ref.update(synth_map(i, device))
        got = ref.compute(sync=False)
        exact = bool((got == want).all())
        if args.check_all:
            exact = exact and bool((total == want).all()) and float(metrics['miou']) == float(metrics_from_cm(want)['miou'])
        px = n_maps * 1024 * 2048
        print(json.dumps({'config': 4, 'metric': 'miou_eval_maps_per_sec', 'unit': 'maps/s',
                          'value': n_maps / (ms / 1e3), 'ms_total': ms, 'n_gpus': world, 'maps': n_maps,
                          'gbs_per_gpu': 16.0 * px / world / (ms / 1e3) / 1e9,
                          'pixels_counted': int(total.sum()), 'miou': float(metrics['miou']),
                          'bit_exact_vs_host_bincount': exact, 'maps_checked': len(list(idx))}))


def config5(args, rank, world, device):
    if rank != 0:
        return
    torch.manual_seed(0)
    model = fastscnn(3, CLASSES).to(device).set_compute_dtype(torch.bfloat16).eval()
    for b in [int(v) for v in args.batches.split(',')]:
        x = torch.randn(b, 3, 1024, 2048, device=device)
        with torch.no_grad():
            for _ in range(3):
                out = model(x)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            with torch.cuda.graph(graph):
                out = model(x)
            kernels = _lib.launch_count() - before
        for _ in range(3):
            graph.replay()
        ms = timed(graph.replay, args.steps, 1, device)
        per = ms / args.steps
        print(json.dumps({'config': 5, 'metric': 'fastscnn_inference_fps', 'unit': 'img/s', 'batch': b,
                          'value': b / (per / 1e3), 'ms_per_batch': per, 'kernels_per_forward': kernels,
                          'hbm_ideal_fraction': (b * 549e6 / 6556.2e9) / (per / 1e3), 'dtype': 'bf16',
                          'out_shape': list(out.shape)}))
        del graph, out, x
        torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', type=int, required=True, choices=[3, 4, 5])
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--batch', type=int, default=0)
    ap.add_argument('--maps', type=int, default=500)
    ap.add_argument('--check-all', action='store_true')
    ap.add_argument('--batches', default='1,2,4,8,16,32,64')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', init_method='env://', device_id=device)
    {3: config3, 4: config4, 5: config5}[args.config](args, rank, world, device)
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)      # no NCCL teardown handshake after the result is out


if __name__ == '__main__':
    main()

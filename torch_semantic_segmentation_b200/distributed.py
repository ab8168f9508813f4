"""Data-parallel plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.

Replaces what the reference gets from apex (scripts/train_fastscnn.py:144-150;
utils/training.py:5-19): parameter broadcast at start-up and a gradient all-reduce per step.
The path shards by *batch* only (independent crops per rank); the one exchange
per iteration is the gradient sum, 4.55 MB fp32 for Fast-SCNN.

apex DDP puts this model in a single 1e7-element bucket, i.e. one all-reduce after the last
gradient.  Here the gradients already live in ONE contiguous arena (``optim.FlatAdamW``), so a
bucket is just a slice: :class:`GradientAllReducer` cuts the arena into ``num_buckets`` slices in
reverse-forward order and launches each slice's all-reduce on a side stream as soon as the
backward kernels that fill it have been enqueued (``functional.grad_ready``), overlapping the
classifier / fusion buckets with the rest of backward.  No flatten / unflatten copies; the
1/world averaging is folded into the optimizer kernel (``FlatAdamW.grad_scale``).
BatchNorm statistics stay per rank (12 crops of 768x768 per GPU).
"""
import os

import torch
import torch.distributed as dist

from . import functional as Fn

from .nn.blocks import convert_syncbn_model  # noqa: F401  (re-export: the apex name the scripts use)

__all__ = ['setup_distributed', 'broadcast_parameters', 'GradientAllReducer', 'shard_range', 'convert_syncbn_model']


def setup_distributed(enable=True, local_rank=0, backend=None):
    """reference: utils/training.py:5-19 (``env://`` rendezvous, one process per GPU)."""
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    torch.manual_seed(0)
    if enable:
        if not dist.is_initialized():
            dist.init_process_group(backend or ('nccl' if torch.cuda.is_available() else 'gloo'),
                                    init_method='env://')
        return dist.get_world_size(), dist.get_rank(), local_rank
    return 1, 0, 0


def broadcast_parameters(model, src=0):
    """Rank ``src``'s parameters and buffers everywhere (what apex DDP does at construction)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src)


def shard_range(n_items, world_size, rank):
    """Contiguous, padding-free split (the reference's DistributedSampler pads 500 -> 504 on 8
    ranks and counts four images twice)."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class GradientAllReducer:
    """Bucketed, overlapped SUM all-reduce of a ``FlatAdamW`` gradient arena."""

    def __init__(self, optimizer, num_buckets=4, process_group=None, tail_elems=None):
        self.opt = optimizer
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        optimizer.grad_scale = 1.0 / self.world
        arena = optimizer.grad_arena
        self.cuda = arena.is_cuda
        self.side = torch.cuda.Stream(arena.device) if self.cuda else None
        # buckets in reverse-forward order: the last parameters' gradients are ready first
        total = optimizer.numel
        target = (total + num_buckets - 1) // num_buckets
        self.buckets = []          # [lo, hi, n_params, n_ready, launched]
        self.bucket_of = {}
        # ... and, optionally, the parameters of the first layers (the last `tail_elems` elements to become ready: the stem and
        # the learning-to-downsample convs, ~7 k values) get a bucket of their own.  Whatever bucket holds them can only be
        # reduced after the very last weight-gradient kernel, so its all-reduce is the one that is never hidden: it should
        # be a latency-only message, not a quarter of the arena.
        if tail_elems is None:
            # off by default: built after the round's last multi-GPU visit and never measured on NVLink -- and every bucket
            # launch joins the weight-gradient lane into the main stream, so an extra bucket is not free
            tail_elems = int(os.environ.get('TSS_DDP_TAIL', '0'))           # e.g. 16384: the first layers get their own bucket
        hi, lo, count, tail_cut = total, total, 0, False
        for p, off, n in reversed(optimizer.slots):
            if not tail_cut and count > 0 and off + n <= tail_elems:
                self.buckets.append([lo, hi, count, 0, False])
                hi, count, tail_cut = lo, 0, True
            lo = off
            count += 1
            self.bucket_of[id(p)] = len(self.buckets)
            if hi - lo >= target or off == 0:
                self.buckets.append([lo, hi, count, 0, False])
                hi, count = lo, 0
        self.handles = []
        self.enabled = self.world > 1

    def install(self):
        Fn.set_grad_ready_hook(self._on_ready if self.enabled else None)
        self.opt._reducer = self     # FlatAdamW.step() waits for the buckets before updating
        return self

    def _launch(self, b):
        lo, hi = b[0], b[1]
        b[4] = True
        chunk = self.opt.grad_arena[lo:hi]
        if self.cuda:
            Fn.wgrad_lane.join(chunk.device)     # the bucket's wgrads run on the wgrad side stream
            self.side.wait_stream(torch.cuda.current_stream(chunk.device))
            with torch.cuda.stream(self.side):
                self.handles.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)

    def _on_ready(self, param):
        i = self.bucket_of.get(id(param))
        if i is None:
            return
        b = self.buckets[i]
        b[3] += 1
        if b[3] == b[2] and not b[4]:
            self._launch(b)

    def finish(self):
        """Call after ``backward()``: flush buckets that were not triggered and make the compute
        stream wait for every all-reduce."""
        if not self.enabled:
            return
        for b in self.buckets:
            if not b[4]:
                self._launch(b)
        if self.cuda:
            for h in self.handles:
                h.wait()
            torch.cuda.current_stream(self.opt.grad_arena.device).wait_stream(self.side)
        self.handles = []
        for b in self.buckets:
            b[3], b[4] = 0, False

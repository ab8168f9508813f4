"""Data-parallel plumbing with world_size 2 on CPU (gloo): the N>1 host logic of
torch_semantic_segmentation_b200/distributed.py, optim.py and metrics.py, with the kernels emulated
by tests/fake_backend.py.  Invariants:

* DP(2 ranks, per-rank batch b) gradients after the bucketed all-reduce + 1/world averaging equal
  the mean of the two ranks' single-process gradients, and every rank ends the step with identical
  parameters (parameters were broadcast from rank 0 first);
* confusion-matrix shards (padding-free contiguous split) sum to the single-process matrix, exactly.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _setup(rank, world, port):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group('gloo', init_method='env://', rank=rank, world_size=world)
    from torch_semantic_segmentation_b200 import _lib
    from tests.fake_backend import FakeBackend
    _lib.set_backend(FakeBackend())


def _batch(seed, n=2, size=64):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, size, size, generator=g)
    y = torch.randint(0, 19, (n, size, size), generator=g)
    y[torch.rand(n, size, size, generator=g) < 0.1] = 255
    return x, y


def _train_worker(rank, world, port, out):
    _setup(rank, world, port)
    from torch_semantic_segmentation_b200.distributed import GradientAllReducer, broadcast_parameters
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    from torch_semantic_segmentation_b200.optim import FlatAdamW

    def make(seed):
        torch.manual_seed(seed)
        m = fastscnn(3, 19)
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        return m.train()

    loss_fn = CrossEntropyLoss(ignore_index=255)
    # single-process gradients of BOTH ranks' batches on rank 0's weights (the reference point)
    singles = []
    for r in range(world):
        m = make(0)
        x, y = _batch(100 + r)
        loss_fn(m(x), y).backward()
        singles.append(torch.cat([p.grad.reshape(-1) for p in m.parameters()]))
    want = sum(singles) / world

    model = make(rank)                      # different init per rank: the broadcast must fix that
    broadcast_parameters(model)
    opt = FlatAdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    reducer = GradientAllReducer(opt, num_buckets=4, tail_elems=int(os.environ.get('TSS_TEST_DDP_TAIL', '0'))).install()
    assert reducer.enabled and len(reducer.buckets) >= 2
    covered = sorted((b[0], b[1]) for b in reducer.buckets)
    assert covered[0][0] == 0 and covered[-1][1] == opt.numel
    assert all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))      # a partition of the arena
    x, y = _batch(100 + rank)
    opt.zero_grad()
    loss = loss_fn(model(x), y)
    loss.backward()
    launched_in_backward = sum(1 for b in reducer.buckets if b[4])
    reducer.finish()
    got = torch.cat([p.grad.reshape(-1) for p in model.parameters()]) * opt.grad_scale
    err = float((got - want).norm() / want.norm())
    opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        torch.save({'err': err, 'same': same, 'launched_in_backward': launched_in_backward,
                    'buckets': len(reducer.buckets)}, out)
    dist.destroy_process_group()


def _eval_worker(rank, world, port, out):
    _setup(rank, world, port)
    from torch_semantic_segmentation_b200.distributed import shard_range
    from torch_semantic_segmentation_b200.metrics import ConfusionMatrix, metrics_from_cm
    n_maps = 7                                        # 7 maps on 2 ranks: 4 + 3, no padding duplicates

    def pair(i):
        g = torch.Generator().manual_seed(4321 + i)
        p = torch.randint(0, 19, (40, 56), generator=g)
        l = torch.randint(0, 19, (40, 56), generator=g)
        l[torch.rand(40, 56, generator=g) < 0.1] = 255
        return p, l

    lo, hi = shard_range(n_maps, world, rank)
    cm = ConfusionMatrix(19)
    for i in range(lo, hi):
        cm.update(pair(i))
    total = cm.compute()                              # int64 all-reduce
    if rank == 0:
        single = ConfusionMatrix(19)
        for i in range(n_maps):
            single.update(pair(i))
        ref = single.compute(sync=False)
        torch.save({'equal': bool(torch.equal(total, ref)), 'shard': (lo, hi),
                    'miou_equal': float(metrics_from_cm(total)['miou']) == float(metrics_from_cm(ref)['miou']),
                    'count': int(total.sum())}, out)
    dist.destroy_process_group()


def _syncbn_worker(rank, world, port, out):
    """DP(k ranks, per-rank batch b) with SyncBN == single process on the batch k*b
    (equal valid-pixel counts per rank, so that the mean of the local mean losses is the global mean)."""
    _setup(rank, world, port)
    from torch_semantic_segmentation_b200.distributed import convert_syncbn_model
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn

    def make():
        torch.manual_seed(0)
        m = fastscnn(3, 19)
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        return m.train()

    g = torch.Generator().manual_seed(7)
    x = torch.randn(2 * world, 3, 64, 64, generator=g)
    y = torch.randint(0, 19, (2 * world, 64, 64), generator=g)          # no ignored pixels: equal counts per rank
    loss_fn = CrossEntropyLoss(ignore_index=255)
    single = make()
    loss_fn(single(x), y).backward()
    want = torch.cat([p.grad.reshape(-1) for p in single.parameters()])
    want_rm = single.downsample[0][1].running_mean.clone()

    model = convert_syncbn_model(make())
    lo = 2 * rank
    loss = loss_fn(model(x[lo:lo + 2]), y[lo:lo + 2])
    loss.backward()
    got = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    dist.all_reduce(got)
    got /= world                                                          # what the gradient all-reduce + 1/world does
    err = float((got - want).norm() / want.norm())
    rm_err = float((model.downsample[0][1].running_mean - want_rm).abs().max())
    head, head_ref = model.classifier[3].weight.grad.clone(), single.classifier[3].weight.grad
    dist.all_reduce(head)
    head_err = float((head / world - head_ref).norm() / head_ref.norm())
    if rank == 0:
        torch.save({'err': err, 'rm_err': rm_err, 'head_err': head_err}, out)
    dist.destroy_process_group()


def _run(worker, tmp_path):
    out = str(tmp_path / 'result.pt')
    mp.spawn(worker, args=(2, _free_port(), out), nprocs=2, join=True)
    return torch.load(out)


@pytest.mark.timeout(600)
@pytest.mark.parametrize('tail', [0, 16384])
def test_bucketed_gradient_allreduce_world2(tmp_path, monkeypatch, tail):
    monkeypatch.setenv('TSS_TEST_DDP_TAIL', str(tail))       # (spawned workers inherit the environment)
    r = _run(_train_worker, tmp_path)
    assert r['err'] < 1e-5, r
    assert r['same'], 'ranks diverged after one optimizer step'
    assert r['buckets'] >= 2 and r['launched_in_backward'] >= 1      # overlap: buckets leave during backward
    if tail:
        assert r['buckets'] >= 3      # the first layers' parameters in a bucket of their own


@pytest.mark.timeout(600)
def test_confusion_matrix_shards_sum_exactly_world2(tmp_path):
    r = _run(_eval_worker, tmp_path)
    assert r['equal'] and r['miou_equal'] and r['shard'] == (0, 4) and r['count'] > 0


@pytest.mark.timeout(600)
def test_syncbn_matches_single_process_on_the_concatenated_batch_world2(tmp_path):
    r = _run(_syncbn_worker, tmp_path)
    assert r['rm_err'] < 1e-6, r           # the statistics are those of the whole batch
    assert r['head_err'] < 1e-4, r         # the well-conditioned head gradient
    assert r['err'] < 1e-2, r              # all gradients (near-cancellations in front of the BN layers: fp32 noise)


def test_shard_range_is_a_padding_free_partition():
    from torch_semantic_segmentation_b200.distributed import shard_range
    for n, w in [(500, 8), (7, 2), (3, 8), (0, 4), (64, 1)]:
        parts = [shard_range(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    assert [b - a for a, b in (shard_range(500, 8, r) for r in range(8))] == [63, 63, 63, 63, 62, 62, 62, 62]

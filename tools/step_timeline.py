"""Kernel timeline of the graph-replayed training step (BASELINE.json configs[1]) from CUPTI activity records
(torch.profiler): for every kernel of one replay its stream, start offset, duration and the gap to the previous
kernel on the same stream; then per-kernel-name totals, the busy time of each stream and the union busy time.
A number printed by this tool is NOT a bench value (the profiler is attached); it shows where the step's time is.

    python tools/step_timeline.py [--out gpurun_out/timeline.json] [--eager]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='gpurun_out/timeline.json')
    ap.add_argument('--replays', type=int, default=3)
    ap.add_argument('--batch', type=int, default=12)
    ap.add_argument('--crop', type=int, default=768)
    args = ap.parse_args()
    from bench import synthetic_batch, CLASSES
    from torch_semantic_segmentation_b200.engine import GraphedTrainStep
    from torch_semantic_segmentation_b200.functional import enable_deferred_logits
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    from torch_semantic_segmentation_b200.distributed import GradientAllReducer

    device = torch.device('cuda', 0)
    torch.manual_seed(0)
    model = fastscnn(3, CLASSES).to(device).set_compute_dtype(torch.bfloat16)
    opt = FlatAdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    GradientAllReducer(opt, num_buckets=4).install()
    loss_fn = CrossEntropyLoss(ignore_index=255)
    x, y = synthetic_batch(args.batch, args.crop, 1234, device)
    model.train()
    enable_deferred_logits(model, loss_fn)
    g = GraphedTrainStep(model, opt, loss_fn, x, y)
    for _ in range(5):
        g.graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    clean_ms = e0.elapsed_time(e1) / 20

    # launch-latency floor: a dependent chain of 300 near-empty library kernels (8 elements each) in one graph
    from torch_semantic_segmentation_b200 import ops
    tiny, one = torch.ones(8, device=device), torch.ones(1, device=device)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.scale_inplace(tiny, one)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    chain = torch.cuda.CUDAGraph()
    with torch.cuda.graph(chain):
        for _ in range(300):
            ops.scale_inplace(tiny, one)
    chain.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        chain.replay()
    e1.record()
    torch.cuda.synchronize()
    chain_us = e0.elapsed_time(e1) * 1e3 / 10 / 300
    print(json.dumps({'dependent_chain_us_per_kernel': chain_us}))

    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.replays):
            g.graph.replay()
            torch.cuda.synchronize()
    recs = []
    for e in prof.profiler.kineto_results.events():
        if e.device_type() != torch.autograd.DeviceType.CUDA:
            continue
        recs.append(dict(name=e.name(), start=e.start_ns() / 1e3, end=(e.start_ns() + e.duration_ns()) / 1e3,
                         stream=int(e.device_resource_id())))
    recs.sort(key=lambda r: r['start'])
    if not recs:
        print('no CUDA activity records (CUPTI unavailable?)')
        return
    # split into replays: a gap > 200 us separates them (synchronize + host turn-around)
    replays, cur = [], [recs[0]]
    for r in recs[1:]:
        if r['start'] - max(q['end'] for q in cur) > 200:
            replays.append(cur)
            cur = [r]
        else:
            cur.append(r)
    replays.append(cur)
    last = max(replays, key=len)
    t0 = last[0]['start']
    span = max(r['end'] for r in last) - t0
    prev_end = {}
    rows = []
    for r in last:
        gap = r['start'] - prev_end.get(r['stream'], r['start'])
        prev_end[r['stream']] = max(prev_end.get(r['stream'], 0), r['end'])
        rows.append(dict(name=r['name'][:90], stream=r['stream'], t_us=round(r['start'] - t0, 2),
                         dur_us=round(r['end'] - r['start'], 2), gap_us=round(gap, 2)))
    # union busy time
    busy, cur_s, cur_e = 0.0, None, None
    for r in sorted(last, key=lambda r: r['start']):
        if cur_s is None:
            cur_s, cur_e = r['start'], r['end']
        elif r['start'] <= cur_e:
            cur_e = max(cur_e, r['end'])
        else:
            busy += cur_e - cur_s
            cur_s, cur_e = r['start'], r['end']
    busy += cur_e - cur_s
    by_name, by_stream = {}, {}
    for r in rows:
        key = r['name'].split('<')[0].split('(')[0]
        a = by_name.setdefault(key, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += r['dur_us']
        a[2] += max(r['gap_us'], 0.0)
        s = by_stream.setdefault(str(r['stream']), [0, 0.0])
        s[0] += 1
        s[1] += r['dur_us']
    summary = dict(clean_ms_per_step=clean_ms, chain_us_per_kernel=chain_us, replays_seen=len(replays), kernels=len(rows), span_us=span, union_busy_us=busy,
                   idle_us=span - busy, by_stream={k: dict(kernels=v[0], busy_us=round(v[1], 1)) for k, v in by_stream.items()},
                   by_name=sorted(([k, v[0], round(v[1], 1), round(v[2], 1)] for k, v in by_name.items()), key=lambda t: -t[2]))
    os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
    with open(args.out, 'w') as f:
        json.dump(dict(summary=summary, kernels=rows), f)
    print(json.dumps({k: v for k, v in summary.items() if k != 'by_name'}))
    print('%-60s %6s %10s %10s' % ('kernel', 'n', 'sum_us', 'gaps_us'))
    for k, n, d, gp in summary['by_name']:
        print('%-60s %6d %10.1f %10.1f' % (k[:60], n, d, gp))


if __name__ == '__main__':
    main()

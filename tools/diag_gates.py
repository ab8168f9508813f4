"""How far do the gated training paths move the gradients, measured against the yardsticks that matter:
the fp32 run of the same step, and the run-to-run noise of the default bf16 path (fp32 atomics).  Prints one
JSON line per (shape, key).  Diagnostic, not a test.

    python tools/diag_gates.py [--full]      (--full adds the 12 x 768 x 768 benchmark shape)
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

KEYS = ('classifier.3.weight', 'features.0.0.conv1.0.weight', 'features.0.0.conv2.0.weight', 'features.0.0.conv3.0.weight',
        'downsample.1.0.weight', 'downsample.1.2.weight', 'downsample.0.0.weight', 'features.2.2.conv1.0.weight',
        'fusion.lowres.2.0.weight')


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def grads(x, y, dtype, **gates):
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    keep = {k: getattr(Fn, k) for k in gates}
    for k, v in gates.items():
        setattr(Fn, k, v)
    try:
        torch.manual_seed(0)
        model = fastscnn(3, 19).cuda().set_compute_dtype(dtype).train()
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        loss = CrossEntropyLoss(ignore_index=255)(model(x), y)
        loss.backward()
        torch.cuda.synchronize()
        return float(loss), {k: p.grad.detach().float().clone() for k, p in model.named_parameters()}
    finally:
        for k, v in keep.items():
            setattr(Fn, k, v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--full', action='store_true')
    args = ap.parse_args()
    from oracle.golden_inputs import train_batch
    shapes = [('golden 2x96x160',) + tuple(t.cuda() for t in train_batch('fastscnn'))]
    if args.full:
        from bench import synthetic_batch
        shapes.append(('bench 12x768x768',) + tuple(synthetic_batch(12, 768, 1234, torch.device('cuda'))))
    for name, x, y in shapes:
        l32, g32 = grads(x, y, torch.float32)
        runs = {'bf16 default': grads(x, y, torch.bfloat16), 'bf16 default again': grads(x, y, torch.bfloat16),
                'bf16 BNRED_EXT': grads(x, y, torch.bfloat16, FUSE_BNRED_EXT=True),
                'bf16 BNAPPLY': grads(x, y, torch.bfloat16, FUSE_BNAPPLY=True),
                'bf16 no BNRED': grads(x, y, torch.bfloat16, FUSE_BNRED=False)}
        print(json.dumps({'shape': name, 'loss_fp32': l32, **{k: v[0] for k, v in runs.items()}}))
        for key in KEYS:
            row = {'shape': name, 'key': key}
            for rn, (_, g) in runs.items():
                row[rn + ' vs fp32'] = round(rel(g[key], g32[key]), 4)
            row['BNRED_EXT vs default'] = round(rel(runs['bf16 BNRED_EXT'][1][key], runs['bf16 default'][1][key]), 4)
            row['default vs default'] = round(rel(runs['bf16 default again'][1][key], runs['bf16 default'][1][key]), 4)
            print(json.dumps(row))
        worst = {rn: max((rel(g[k], g32[k]), k) for k in g32) for rn, (_, g) in runs.items()}
        print(json.dumps({'shape': name, 'worst key vs fp32': {rn: [round(v[0], 4), v[1]] for rn, v in worst.items()}}))
        med = {rn: sorted(rel(g[k], g32[k]) for k in g32)[len(g32) // 2] for rn, (_, g) in runs.items()}
        print(json.dumps({'shape': name, 'median key vs fp32': {rn: round(v, 4) for rn, v in med.items()}}))


if __name__ == '__main__':
    main()

"""How the error of bf16 arithmetic grows with depth in TRAIN mode, measured on the UNMODIFIED reference modules
(/root/reference, stock torch ops): Fast-SCNN trained for 30 steps on synthetic scenes, then one forward of a fresh
12 x 256 x 256 batch in fp32 and under torch.autocast(bfloat16); relative L2 distance of every conv / BatchNorm output.
Runs only where /root/reference exists (this container); its output is committed as profiles/r2_bf16_depth_profile.txt
and is the evidence behind the bf16 bounds of tests/test_baseline_shapes_gpu.py.
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.golden_inputs import scene_batch                      # noqa: E402
from torch_semantic_segmentation.models import fastscnn          # noqa: E402  (the reference)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def main():
    torch.manual_seed(0)
    m = fastscnn(3, 19)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-5)
    m.train()
    for i in range(30):
        x, y = scene_batch(6, 128, 256, 100 + i % 4)
        opt.zero_grad()
        loss = F.cross_entropy(m(x), y, ignore_index=255)
        loss.backward()
        opt.step()
    print('# reference Fast-SCNN after 30 AdamW steps on scene batches, train loss %.3f' % float(loss.detach()))
    xt, _ = scene_batch(12, 256, 256, 3)
    acts = {}
    for n, mod in m.named_modules():
        if isinstance(mod, torch.nn.Conv2d):       # BatchNorm outputs are overwritten by the in-place ReLU behind them
            mod.register_forward_hook(lambda mod, inp, out, n=n: acts.setdefault(n, []).append(out.detach().float()))
    with torch.no_grad():
        m(xt)
        with torch.autocast('cpu', dtype=torch.bfloat16):
            m(xt)
    print('%-36s %10s   %s' % ('conv output (train mode)', 'bf16 rel L2', 'shape'))
    for n, (a, b) in acts.items():
        print('%-36s %10.4f   %s' % (n, rel(b, a), tuple(a.shape)))
    m.eval()
    with torch.no_grad():
        a = m(xt)
        with torch.autocast('cpu', dtype=torch.bfloat16):
            b = m(xt)
    print('eval mode (running statistics), logits: bf16 rel L2 %.4f' % rel(b.float(), a))


if __name__ == '__main__':
    main()

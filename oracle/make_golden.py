"""Generate tests/golden/* by running the UNMODIFIED reference from /root/reference.

Runs only in the build container (the GPU box has no /root/reference).  Usage:

    python oracle/make_golden.py

Writes
  tests/golden/{fastscnn,contextnet14}_spec.json   state_dict keys + shapes, in order
  tests/golden/{fastscnn,contextnet14}_eval.npz    eval-mode forward (sub-sampled + checksums)
  tests/golden/{fastscnn,contextnet14}_train.npz   train-mode fwd + CE(ignore 255) + bwd:
                                                   loss, selected grads, grad checksums,
                                                   BN running stats after the step
  tests/golden/ohem.npz                            OHEMLoss on seeded logits (both branches)
Inputs are regenerated from seeds by ``oracle.golden_inputs`` on every box, so only the
reference's *outputs* are stored.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

from oracle.golden_inputs import (eval_input, train_batch, ohem_case, SUBSAMPLE,  # noqa: E402
                                  GRAD_KEYS)
from torch_semantic_segmentation.models.fastscnn import fastscnn  # noqa: E402  (reference)
from torch_semantic_segmentation.models.contextnet import contextnet14  # noqa: E402
from torch_semantic_segmentation.losses import OHEMLoss  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
FACTORY = {'fastscnn': fastscnn, 'contextnet14': contextnet14}


def checksum(t):
    t = t.detach().double()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    for arch, factory in FACTORY.items():
        torch.manual_seed(0)
        model = factory(3, 19)
        spec = [[k, list(v.shape)] for k, v in model.state_dict().items()]
        with open(os.path.join(GOLD, '%s_spec.json' % arch), 'w') as f:
            json.dump(spec, f)

        # eval-mode forward
        model.eval()
        x = eval_input(arch)
        with torch.no_grad():
            out = model(x)
        sy, sx = SUBSAMPLE
        np.savez_compressed(
            os.path.join(GOLD, '%s_eval.npz' % arch),
            sub=out[:, :, ::sy, ::sx].numpy(), checksum=checksum(out),
            weight_checksum=np.stack([checksum(v.float()) for v in model.state_dict().values()]))

        # train-mode forward + CE + backward.  Dropout is replaced by p=0 so that no
        # RNG stream has to be matched (the mask path is tested separately).
        torch.manual_seed(0)
        model = factory(3, 19)
        model.train()
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        x, y = train_batch(arch)
        out = model(x)
        loss = F.cross_entropy(out, y, ignore_index=255)
        loss.backward()
        grads = dict((k, p.grad) for k, p in model.named_parameters())
        sd = model.state_dict()
        save = {
            'loss': np.array(loss.item()),
            'out_sub': out.detach()[:, :, ::sy, ::sx].numpy(),
            'out_checksum': checksum(out),
            'grad_checksums': np.stack([checksum(grads[k]) for k in grads]),
        }
        for k in GRAD_KEYS[arch]:
            save['grad:' + k] = grads[k].numpy()
        for k in sd:
            if k.endswith('running_mean') or k.endswith('running_var'):
                save['buf:' + k] = sd[k].numpy()
        np.savez_compressed(os.path.join(GOLD, '%s_train.npz' % arch), **save)
        print(arch, 'loss', loss.item(), 'out checksum', checksum(out))

    # OHEM (losses/ohem_loss.py) -- both branches of the `if`
    res = {}
    for name in ('many_hard', 'few_hard'):
        logits, target, kw = ohem_case(name)
        res[name] = np.array(OHEMLoss(**kw)(logits, target).item())
        print('ohem', name, res[name])
    np.savez_compressed(os.path.join(GOLD, 'ohem.npz'), **res)


if __name__ == '__main__':
    main()

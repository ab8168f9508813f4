"""OHEM loss and deep-supervision wrapper host logic on the CPU emulation of the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle.golden_inputs import ohem_case, train_batch
from oracle.init_state import GOLDEN_DIR
from oracle.losses import ohem as oracle_ohem
from torch_semantic_segmentation_b200.losses import CrossEntropyLoss, OHEMLoss, cross_entropy
from torch_semantic_segmentation_b200.models import fastscnn
from torch_semantic_segmentation_b200.models.fastscnn import Classifier
from torch_semantic_segmentation_b200.wrappers import DeepSupervisionWrapper
from torch_semantic_segmentation_b200.wrappers.deep_supervision_wrapper import AuxiliaryHead


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _untied(logits, target, kw):
    """Mask of the pixels whose loss differs from the selected order statistic: among pixels that TIE
    with it the reference's sort keeps an arbitrary subset, this package shares the slots evenly (same
    loss value, equally valid subgradient), so the comparison leaves them out."""
    pl = torch.nn.functional.cross_entropy(logits.float(), target, ignore_index=kw['ignore_index'], reduction='none')
    srt, _ = torch.sort(pl.flatten(), descending=True)
    vk = srt[int(pl.numel() * kw['numel_frac'])]
    return ((pl - vk).abs() > 1e-6 * vk.abs()).unsqueeze(1)


@pytest.mark.parametrize('name', ['many_hard', 'few_hard'])
def test_ohem_matches_golden_and_oracle(fake_backend, name):
    logits, target, kw = ohem_case(name)
    gold = float(np.load(os.path.join(GOLDEN_DIR, 'ohem.npz'))[name])
    x = logits.clone().requires_grad_(True)
    loss = OHEMLoss(**kw)(x, target)
    loss.backward()
    assert abs(float(loss) - gold) < 1e-5 * abs(gold)
    xr = logits.clone().requires_grad_(True)
    oracle_ohem(xr, target, **kw).backward()
    assert rel(x.grad * _untied(logits, target, kw), xr.grad * _untied(logits, target, kw)) < 1e-5


def test_pixel_cross_entropy_reduction_none_has_a_gradient(fake_backend):
    logits, target, _ = ohem_case('many_hard')
    x = logits.clone().requires_grad_(True)
    pix = cross_entropy(x, target, ignore_index=255, reduction='none')
    w = torch.rand_like(pix)
    (pix * w).sum().backward()
    xr = logits.clone().requires_grad_(True)
    (torch.nn.functional.cross_entropy(xr, target, ignore_index=255, reduction='none') * w).sum().backward()
    assert rel(pix, torch.nn.functional.cross_entropy(logits, target, ignore_index=255, reduction='none')) < 1e-6
    assert rel(x.grad, xr.grad) < 1e-5


def test_deep_supervision_recipe_of_the_reference_script(fake_backend):
    """scripts/train_fastscnn.py:107-137: aux heads on .downsample (x8) and .features (x32), OHEM on
    the main output + 0.4 * CE on the two auxiliary outputs; compared with stock torch ops on the same
    weights (the head convs via the oracle-style functional path)."""
    torch.manual_seed(0)
    model = fastscnn(3, 19)
    model = DeepSupervisionWrapper(model, [
        (model.downsample, AuxiliaryHead(Classifier(64, 19), 8)),
        (model.features, AuxiliaryHead(Classifier(128, 19), 32)),
    ])
    keys = list(model.state_dict().keys())
    assert keys[0] == 'module.downsample.0.0.weight' and 'auxiliary.0.0.3.bias' in keys and 'auxiliary.1.0.0.0.weight' in keys
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    x, y = train_batch('fastscnn')
    model.train()
    out, (aux1, aux2) = model(x)
    assert out.shape == aux1.shape == aux2.shape == (2, 19, 96, 160)
    ohem, ce = OHEMLoss(ignore_index=255, numel_frac=0.1), CrossEntropyLoss(ignore_index=255)
    loss = ohem(out, y) + 0.4 * ce(aux1, y) + 0.4 * ce(aux2, y)
    loss.backward()
    want = oracle_ohem(out.detach(), y, ignore_index=255, numel_frac=0.1) \
        + 0.4 * torch.nn.functional.cross_entropy(aux1.detach(), y, ignore_index=255) \
        + 0.4 * torch.nn.functional.cross_entropy(aux2.detach(), y, ignore_index=255)
    assert abs(float(loss) - float(want)) < 1e-5 * abs(float(want))
    grads = [p.grad for p in model.parameters()]
    assert all(g is not None and torch.isfinite(g).all() for g in grads)
    assert float(model.auxiliary[1][0][3].weight.grad.abs().sum()) > 0
    model.eval()
    with torch.no_grad():
        assert model(x).shape == (2, 19, 96, 160)            # eval mode: the model's output only

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


class _Recipe(torch.nn.Module):
    """scripts/train_fastscnn.py:131-137: OHEM on the main output + 0.4 * CE on each auxiliary output."""

    def __init__(self, numel_frac=0.1):
        super().__init__()
        self.ohem, self.ce = OHEMLoss(ignore_index=255, numel_frac=numel_frac), CrossEntropyLoss(ignore_index=255)

    def forward(self, y_pred, y):
        out, (aux1, aux2) = y_pred
        return self.ohem(out, y) + 0.4 * self.ce(aux1, y) + 0.4 * self.ce(aux2, y)


def _wrapped_fastscnn():
    torch.manual_seed(0)
    model = fastscnn(3, 19)
    model = DeepSupervisionWrapper(model, [
        (model.downsample, AuxiliaryHead(Classifier(64, 19), 8)),
        (model.features, AuxiliaryHead(Classifier(128, 19), 32)),
    ])
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return model


def stock_head_gradients(model, x, y, weight):
    """Gradient of ``weight * CE(upsample(conv1x1(f)))`` w.r.t. the last conv of auxiliary head 0, with stock torch ops
    on the feature ``f`` that conv saw in ``model(x)``."""
    seen = {}
    last = model.auxiliary[0][0][3]
    h = last.register_forward_hook(lambda m, args, out: seen.setdefault('f', args[0].detach().float().cpu()))
    model.train()
    model(x)
    h.remove()
    w = last.weight.detach().float().cpu().clone().requires_grad_(True)
    b = last.bias.detach().float().cpu().clone().requires_grad_(True)
    logits = torch.nn.functional.interpolate(torch.nn.functional.conv2d(seen['f'], w, b), scale_factor=8,
                                             mode='bilinear', align_corners=True)
    (weight * torch.nn.functional.cross_entropy(logits, y.cpu(), ignore_index=255)).backward()
    return w.grad, b.grad


def test_trainer_honours_the_loss_weights_of_the_recipe(fake_backend):
    """The trainer's backward must give the auxiliary terms their 0.4 (a process-wide 'the seed
    gradient is 1' promise made every loss term back-propagate with weight 1).  lr = 0 keeps the weights and
    leaves the step's gradients in ``.grad``."""
    from torch_semantic_segmentation_b200.engine import create_segmentation_trainer
    model = _wrapped_fastscnn()
    x, y = train_batch('fastscnn')
    gw, gb = stock_head_gradients(model, x, y, 0.4)
    # BatchNorm running statistics moved in that probe forward, batch statistics (training mode) did not
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    trainer = create_segmentation_trainer(model, opt, _Recipe(), 'cpu', logging=False)
    trainer.run([(x, y)], max_epochs=1)
    last = model.auxiliary[0][0][3]
    assert rel(last.weight.grad, gw) < 1e-4, rel(last.weight.grad, gw)
    assert rel(last.bias.grad, gb) < 1e-4
    # and a plain single-term loss still takes the unscaled fast path with the right value
    model2 = _wrapped_fastscnn()
    g1w, _ = stock_head_gradients(model2, x, y, 1.0)

    class OnlyAux(torch.nn.Module):
        def forward(self, y_pred, t):
            return CrossEntropyLoss(ignore_index=255)(y_pred[1][0], t)
    trainer2 = create_segmentation_trainer(model2, torch.optim.SGD(model2.parameters(), lr=0.0), OnlyAux(), 'cpu',
                                           logging=False)
    trainer2.run([(x, y)], max_epochs=1)
    assert rel(model2.auxiliary[0][0][3].weight.grad, g1w) < 1e-4

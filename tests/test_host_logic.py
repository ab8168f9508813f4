"""Host-side logic (layout bookkeeping, autograd formulas, module tree, engine) exercised on
CPU through tests/fake_backend.py.  The kernels themselves are tested on the GPU (-m gpu)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import confusion as o_cm
from oracle.golden_inputs import eval_input, train_batch
from oracle.init_state import init_state
from oracle.train_step import AdamW as OracleAdamW, loss_and_grads, model_forward, split_state, train_step
from torch_semantic_segmentation_b200 import ops
from torch_semantic_segmentation_b200.engine import (create_segmentation_evaluator,
                                                     create_segmentation_trainer)
from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
from torch_semantic_segmentation_b200.models import fastscnn


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _no_dropout(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return model


def test_geom_and_pitch():
    t = ops.empty_nhwc(2, 19, 4, 6, torch.float32, 'cpu', pitch=32)
    assert ops.geom(t) == (2, 19, 4, 6, 32)
    assert ops.geom(torch.zeros(2, 8, 4, 6)) is None                    # NCHW-contiguous is not NHWC
    assert ops.geom(ops.as_nhwc(torch.zeros(2, 8, 4, 6))) == (2, 8, 4, 6, 8)
    assert ops.geom(ops.empty_nhwc(3, 16, 1, 1, torch.float32, 'cpu')) == (3, 16, 1, 1, 16)
    cat = ops.empty_nhwc(2, 256, 3, 5, torch.float32, 'cpu')
    assert ops.geom(cat[:, 128:160]) == (2, 32, 3, 5, 256)


def test_state_dict_is_the_reference_layout(fake_backend):
    torch.manual_seed(0)
    model = fastscnn(3, 19)
    want = init_state('fastscnn', 0)
    got = model.state_dict()
    assert list(got.keys()) == list(want.keys())
    for k in want:
        assert got[k].shape == want[k].shape and torch.equal(got[k], want[k]), k
    model.load_state_dict(want, strict=True)


def test_eval_forward_matches_oracle(fake_backend):
    torch.manual_seed(0)
    model = fastscnn(3, 19).eval()
    x = eval_input('fastscnn')
    with torch.no_grad():
        out = model(x)
    assert out.shape == (1, 19, 160, 224) and out.is_contiguous()
    ref = model_forward('fastscnn', init_state('fastscnn', 0), x, False)
    assert rel(out, ref) < 1e-5


def test_forward_hooks_fire_on_downsample_and_features(fake_backend):
    torch.manual_seed(0)
    model = fastscnn(3, 19).eval()
    seen = {}
    model.downsample.register_forward_hook(lambda m, i, o: seen.__setitem__('d', o.shape))
    model.features.register_forward_hook(lambda m, i, o: seen.__setitem__('f', o.shape))
    with torch.no_grad():
        model(torch.randn(1, 3, 64, 96))
    assert seen == {'d': (1, 64, 8, 12), 'f': (1, 128, 2, 3)}


def test_train_step_matches_oracle(fake_backend):
    torch.manual_seed(0)
    model = _no_dropout(fastscnn(3, 19)).train()
    x, y = train_batch('fastscnn')
    out = model(x)
    loss = CrossEntropyLoss(ignore_index=255)(out, y)
    loss.backward()
    sd = split_state(init_state('fastscnn', 0))
    ref_loss, ref_logits, ref_grads = loss_and_grads('fastscnn', sd, x, y, dropout_mask=1.0)
    assert abs(float(loss) - float(ref_loss)) < 1e-5
    assert rel(out, ref_logits) < 1e-4
    params = dict(model.named_parameters())
    # the head is well conditioned; deep layers suffer ReLU-mask flips (the fp32 reference itself is
    # ~1e-2 away from an fp64 run there), so they get a looser bound
    for k in ('classifier.3.weight', 'classifier.3.bias', 'classifier.1.3.weight'):
        assert rel(params[k].grad, ref_grads[k]) < 1e-4, k
    for k, p in params.items():
        if k.endswith('.0.weight') or k.endswith('.2.weight'):
            assert rel(p.grad, ref_grads[k]) < 3e-2, k
    msd = model.state_dict()
    for k in sd:
        if 'running' in k:
            assert (msd[k] - sd[k]).abs().max() < 1e-5, k
        if 'num_batches' in k:
            assert int(msd[k]) == 1


def test_batch_of_one_is_rejected_in_training_like_the_reference(fake_backend):
    model = fastscnn(3, 19).train()
    with pytest.raises(RuntimeError, match='more than 1 value per channel'):
        model(torch.randn(1, 3, 32, 32))


def test_trainer_and_evaluator(fake_backend):
    torch.manual_seed(0)
    model = _no_dropout(fastscnn(3, 19))
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    trainer = create_segmentation_trainer(model, opt, CrossEntropyLoss(ignore_index=255), 'cpu', logging=False)
    x, y = train_batch('fastscnn')
    state = trainer.run([(x, y)] * 2, max_epochs=1)
    assert state.iteration == 2 and np.isfinite(state.output) and 'loss' in state.metrics and state.metrics['lr'] == 1e-3

    sd = split_state(init_state('fastscnn', 0))
    oopt = OracleAdamW(sd, lr=1e-3, weight_decay=1e-5)
    l1 = train_step('fastscnn', sd, oopt, x, y, dropout_mask=1.0)
    l2 = train_step('fastscnn', sd, oopt, x, y, dropout_mask=1.0)
    assert abs(state.output - l2) < 2e-3 * abs(l2), (state.output, l1, l2)

    evaluator = create_segmentation_evaluator(model, 'cpu', num_classes=19, loss_fn=CrossEntropyLoss(ignore_index=255))
    es = evaluator.run([(x, y)])
    with torch.no_grad():
        logits = model.eval()(x)
    want = o_cm.confusion_matrix(o_cm.argmax_classes(logits.numpy()), y.numpy(), 19)
    assert (es.metrics['confusion_matrix'].numpy() == want).all()
    met = o_cm.metrics(want)
    assert float(es.metrics['miou']) == met['miou']
    np.testing.assert_array_equal(es.metrics['iou'].numpy(), met['iou'])
    assert float(es.metrics['accuracy']) == met['accuracy']
    np.testing.assert_array_equal(es.metrics['dice'].numpy(), met['dice'])
    assert np.isfinite(es.metrics['loss'])


def test_flat_adamw_matches_oracle_adamw(fake_backend):
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    torch.manual_seed(0)
    model = _no_dropout(fastscnn(3, 19))
    opt = FlatAdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    # parameters now live in one arena and gradients accumulate in place
    p0 = next(model.parameters())
    assert p0.data_ptr() == opt.param_arena.data_ptr() and p0.grad.data_ptr() == opt.grad_arena.data_ptr()
    trainer = create_segmentation_trainer(model, opt, CrossEntropyLoss(ignore_index=255), 'cpu', logging=False)
    x, y = train_batch('fastscnn')
    state = trainer.run([(x, y)] * 3, max_epochs=1)
    sd = split_state(init_state('fastscnn', 0))
    oopt = OracleAdamW(sd, lr=1e-3, weight_decay=1e-5)
    ref = [train_step('fastscnn', sd, oopt, x, y, dropout_mask=1.0) for _ in range(3)]
    assert abs(state.output - ref[-1]) < 5e-3 * abs(ref[-1]), (state.output, ref)
    assert opt.step_count == 3
    # eval after training must see the updated weights and running statistics (cache invalidation)
    with torch.no_grad():
        out = model.eval()(x)
        want = model_forward('fastscnn', {k: v.detach() for k, v in sd.items()}, x, False)
    assert rel(out, want) < 5e-2


# ------------------------------------------------------------------ ContextNet (config 3) ----
def test_contextnet_state_dict_is_the_reference_layout(fake_backend):
    from torch_semantic_segmentation_b200.models.contextnet import contextnet14
    torch.manual_seed(0)
    model = contextnet14(3, 19)
    want = init_state('contextnet14', 0)
    got = model.state_dict()
    assert list(got.keys()) == list(want.keys())
    for k in want:
        assert got[k].shape == want[k].shape and torch.equal(got[k], want[k]), k
    model.load_state_dict(want, strict=True)


@pytest.mark.parametrize('shape', [(1, 3, 160, 224), (1, 3, 72, 104)])
def test_contextnet_eval_forward_matches_oracle(fake_backend, shape):
    from torch_semantic_segmentation_b200.models.contextnet import contextnet14
    torch.manual_seed(0)
    model = contextnet14(3, 19).eval()
    x = eval_input('contextnet14') if shape == (1, 3, 160, 224) else torch.randn(*shape, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        out = model(x)
    ref = model_forward('contextnet14', init_state('contextnet14', 0), x, False)
    assert out.shape == ref.shape and out.is_contiguous()
    assert rel(out, ref) < 1e-5


@pytest.mark.parametrize('variant', ['contextnet12', 'contextnet18'])
def test_contextnet_12_and_18_match_oracle(fake_backend, variant):
    """The other two factories of contextnet.py:13-25 (context branch on the input shrunk by 2 / 8): eval forward and one
    training step's loss against the oracle (pinned to the live reference by tests/test_oracle.py)."""
    from torch_semantic_segmentation_b200.models import contextnet as cn
    torch.manual_seed(0)
    model = _no_dropout(getattr(cn, variant)(3, 19))
    sd0 = init_state('contextnet14', 0)               # same layers, same construction order: same random init
    assert all(torch.equal(model.state_dict()[k], sd0[k]) for k in sd0)
    x = torch.randn(2, 3, 128, 192, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        out = model.eval()(x)
    ref = model_forward(variant, sd0, x, False)
    assert out.shape == ref.shape == (2, 19, 128, 192) and rel(out, ref) < 1e-5
    y = torch.randint(0, 19, (2, 128, 192), generator=torch.Generator().manual_seed(7))
    loss = CrossEntropyLoss(ignore_index=255)(model.train()(x), y)
    ref_loss, _, _ = loss_and_grads(variant, split_state(sd0), x, y, dropout_mask=1.0)
    assert abs(float(loss) - float(ref_loss)) < 1e-4 * abs(float(ref_loss))


def test_contextnet_train_step_matches_oracle(fake_backend):
    from torch_semantic_segmentation_b200.models.contextnet import contextnet14
    torch.manual_seed(0)
    model = _no_dropout(contextnet14(3, 19)).train()
    x, y = train_batch('contextnet14')
    out = model(x)
    loss = CrossEntropyLoss(ignore_index=255)(out, y)
    loss.backward()
    sd = split_state(init_state('contextnet14', 0))
    ref_loss, ref_logits, ref_grads = loss_and_grads('contextnet14', sd, x, y, dropout_mask=1.0)
    assert abs(float(loss) - float(ref_loss)) < 1e-5
    assert rel(out, ref_logits) < 1e-4
    params = dict(model.named_parameters())
    for k in ('classifier.5.weight', 'classifier.5.bias', 'classifier.3.1.weight'):
        assert rel(params[k].grad, ref_grads[k]) < 1e-4, k
    # conv weights in front of a BatchNorm (their gradient is a near-cancellation; the fp32 reference
    # itself is ~1e-2 from an fp64 run), incl. the dense 3x3 = patch GEMM + tap-major permutation
    for k in ('classifier.3.0.weight', 'context.7.0.weight', 'context.6.1.conv3.0.weight', 'context.0.0.weight', 'spatial.0.0.weight'):
        assert rel(params[k].grad, ref_grads[k]) < 3e-2, k
    msd = model.state_dict()
    for k in sd:
        if 'running' in k:
            assert (msd[k] - sd[k]).abs().max() < 1e-5, k


def test_gated_backward_fusions_are_plumbing_equivalent(fake_backend):
    """bf16 + tensor-core pointwise path on the emulated ABI: the extended fused-reduction set and the
    BatchNorm-apply-in-dgrad kernel (both off by default, functional.FUSE_BNRED_EXT / FUSE_BNAPPLY) give the very
    gradients of the default path, with fewer launches."""
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.nn.blocks import set_compute_dtype
    calls = {}
    inner = fake_backend.call

    def counting(name, kwargs):
        calls[name] = calls.get(name, 0) + 1
        return inner(name, kwargs)
    fake_backend.call = counting
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(2, 3, 64, 64, generator=g), torch.randint(0, 19, (2, 64, 64), generator=g)
    keep = Fn.FUSE_BNRED_EXT, Fn.FUSE_BNAPPLY, Fn.FUSE_BNFIN, Fn.FUSE_BNAPPLY_DW
    keep_bnin, Fn.FUSE_BNIN = Fn.FUSE_BNIN, False      # the launch counts below are those of the chain without hand-overs
    runs = {}
    try:
        for ext, fused in ((False, False), (True, False), (False, True), (True, True), ('fin', False), ('dw', False)):
            Fn.FUSE_BNFIN = ext == 'fin'
            Fn.FUSE_BNAPPLY_DW = ext == 'dw'
            Fn.FUSE_BNRED_EXT, Fn.FUSE_BNAPPLY = ext is True, fused
            calls.clear()
            torch.manual_seed(0)
            model = set_compute_dtype(_no_dropout(fastscnn(3, 19)), torch.bfloat16, pw_impl=1).train()
            CrossEntropyLoss(ignore_index=255)(model(x), y).backward()
            runs[ext, fused] = (torch.cat([p.grad.reshape(-1) for p in model.parameters()]), dict(calls))
    finally:
        Fn.FUSE_BNRED_EXT, Fn.FUSE_BNAPPLY, Fn.FUSE_BNFIN, Fn.FUSE_BNAPPLY_DW = keep
        Fn.FUSE_BNIN = keep_bnin
    base, base_calls = runs[False, False]
    # BatchNorm-backward apply folded into the stride-1 depthwise dgrads that carry a fused reduction
    assert rel(runs['dw', False][0], base) < 1e-6
    n_dw = runs['dw', False][1]['tss_dwconv3x3_bwd_fused']
    assert n_dw >= base_calls['tss_dwconv3x3_dgrad_bnred']          # also the stride-1 layers without a producer link
    assert 'tss_dwconv3x3_dgrad_bnred' not in runs['dw', False][1]
    assert runs['dw', False][1]['tss_bn_bwd_apply'] == base_calls['tss_bn_bwd_apply'] - n_dw
    # finalize folded into the apply kernel: same arithmetic, 44 launches less
    assert rel(runs['fin', False][0], base) < 1e-6
    assert runs['fin', False][1]['tss_bn_finalize_apply'] == 44 and 'tss_bn_finalize' not in runs['fin', False][1]
    assert base_calls.get('tss_pwconv_bwd_fused', 0) == 0 and base_calls.get('tss_dwconv3x3_dgrad_s2_bnred', 0) == 0
    for key, (grad, c) in runs.items():
        assert sum(c.values()) <= sum(base_calls.values())
    # the apply-in-dgrad kernel rounds exactly where the two kernels it replaces do: identical gradients
    assert rel(runs[False, True][0], base) < 1e-6 and rel(runs[True, True][0], runs[True, False][0]) < 1e-6
    # a fused reduction sums the fp32 accumulators instead of the bf16-rounded gradient: bf16 noise apart
    assert rel(runs[True, False][0], base) < 5e-2
    assert runs[False, True][1]['tss_pwconv_bwd_fused'] == 22
    assert runs[False, True][1]['tss_bn_bwd_apply'] == base_calls['tss_bn_bwd_apply'] - 22
    assert runs[True, False][1]['tss_dwconv3x3_dgrad_s2_bnred'] == 4
    assert runs[True, False][1]['tss_bn_bwd_reduce'] == base_calls['tss_bn_bwd_reduce'] - 6


def test_grouped_pyramid_pooling_is_plumbing_equivalent(fake_backend):
    """functional.FUSE_PPM (off by default): pool -> grouped branches -> concat and their backward give the
    layer-by-layer module's output, input gradient, parameter gradients and BatchNorm buffers."""
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.models.fastscnn import PyramidPoolingModule
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    keep = Fn.FUSE_PPM
    runs = {}
    try:
        for flag in (False, True):
            Fn.FUSE_PPM = flag
            torch.manual_seed(0)
            m = PyramidPoolingModule(128, 128).train()
            FlatAdamW(m.parameters(), lr=1e-3).zero_grad()
            g = torch.Generator().manual_seed(1)
            x = ops.as_nhwc(torch.randn(5, 128, 6, 9, generator=g)).requires_grad_()
            before = fake_backend.launches
            out = m(x)
            (out * torch.randn(out.shape, generator=g)).sum().backward()
            runs[flag] = (out.detach(), x.grad, {k: p.grad.clone() for k, p in m.named_parameters()},
                          {k: v.clone().float() for k, v in m.state_dict().items() if 'running' in k or 'tracked' in k},
                          fake_backend.launches - before)
    finally:
        Fn.FUSE_PPM = keep
    a, b = runs[False], runs[True]
    assert rel(b[0], a[0]) < 1e-5 and rel(b[1], a[1]) < 1e-5
    assert all(rel(b[2][k], a[2][k]) < 1e-5 for k in a[2]) and all(rel(b[3][k], a[3][k]) < 1e-5 for k in a[3])
    assert b[4] <= a[4] - 30
    # a batch of one still fails like the reference (BatchNorm over a single value in the bin-1 branch)
    Fn.FUSE_PPM = True
    try:
        with pytest.raises(RuntimeError, match='more than 1 value per channel'):
            PyramidPoolingModule(128, 128).train()(torch.randn(1, 128, 6, 9))
    finally:
        Fn.FUSE_PPM = keep


def test_tensor_core_stem_gate_is_plumbing_equivalent(fake_backend):
    """functional.STEM_TC (off by default): the training step with the stem routed to the tcgen05 entry points.
    The image is rounded to bf16 there, so the learning-to-downsample output agrees at bf16 level (the tiny
    whole network behind it -- BatchNorm over 2 values in the bin-1 pyramid branch -- amplifies that noise, so the
    final logits are not compared)."""
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.nn.blocks import set_compute_dtype
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(2, 3, 64, 64, generator=g), torch.randint(0, 19, (2, 64, 64), generator=g)
    keep = Fn.STEM_TC
    calls = {}
    inner = fake_backend.call

    def counting(name, kwargs):
        calls[name] = calls.get(name, 0) + 1
        return inner(name, kwargs)
    fake_backend.call = counting
    runs = {}
    try:
        for flag in (False, True):
            Fn.STEM_TC = flag
            calls.clear()
            torch.manual_seed(0)
            model = set_compute_dtype(_no_dropout(fastscnn(3, 19)), torch.bfloat16, pw_impl=1).train()
            seen = {}
            model.downsample.register_forward_hook(lambda m, i, o: seen.__setitem__('d', o.detach().float()))
            CrossEntropyLoss(ignore_index=255)(model(x), y).backward()
            runs[flag] = (seen['d'], model.downsample[0][0].weight.grad.clone(), dict(calls))
    finally:
        Fn.STEM_TC = keep
    assert rel(runs[True][0], runs[False][0]) < 3e-2
    assert torch.isfinite(runs[True][1]).all() and float(runs[True][1].abs().sum()) > 0
    assert runs[True][2].get('tss_stem3x3s2_fwd_tc') == 1 and (runs[True][2].get('tss_stem3x3s2_wgrad_tc') == 1 or runs[True][2].get('tss_stem3x3s2_wgrad_from_patches') == 1)
    assert 'tss_stem3x3s2_fwd' not in runs[True][2] and 'tss_stem3x3s2_fwd_tc' not in runs[False][2]


@pytest.mark.parametrize('arch', ['fastscnn', 'contextnet14'])
def test_all_gates_together_train_and_eval(fake_backend, arch):
    """Every off-by-default kernel gate switched on at once: two optimisation steps and an eval forward agree with
    the default path at bf16 level (the tensor-core stem rounds the image to bf16), with far fewer launches."""
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.models.contextnet import contextnet14
    from torch_semantic_segmentation_b200.nn.blocks import set_compute_dtype
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    flags = ['FUSE_BNRED_EXT', 'FUSE_BNAPPLY', 'FUSE_BNAPPLY_DW', 'FUSE_PPM', 'FUSE_BNFIN', 'FUSE_BNIN', 'FUSE_BNIN_PW', 'STEM_TC', 'STEM_BWD_FUSED', 'DEFER_LOGITS',
             'OWN_DROPOUT']
    keep = {f: getattr(Fn, f) for f in flags}
    factory = {'fastscnn': fastscnn, 'contextnet14': contextnet14}[arch]
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(4, 3, 64, 96, generator=g), torch.randint(0, 19, (4, 64, 96), generator=g)
    runs = {}
    try:
        for on in (False, True):
            for f in flags:
                setattr(Fn, f, on)
            torch.manual_seed(0)
            model = set_compute_dtype(_no_dropout(factory(3, 19)), torch.bfloat16, pw_impl=1).train()
            opt = FlatAdamW(model.parameters(), lr=1e-3)
            before, losses = fake_backend.launches, []
            for _ in range(2):
                opt.zero_grad()
                loss = CrossEntropyLoss(ignore_index=255)(model(x), y)
                loss.backward()
                opt.step()
                losses.append(float(loss.detach()))
            with torch.no_grad():
                out = model.eval()(x).float()
            runs[on] = (losses, out, fake_backend.launches - before)
    finally:
        for f, v in keep.items():
            setattr(Fn, f, v)
    for a, b in zip(runs[True][0], runs[False][0]):
        assert abs(a - b) < 5e-3 * abs(b), (runs[True][0], runs[False][0])
    assert rel(runs[True][1], runs[False][1]) < 6e-2
    assert runs[True][2] < runs[False][2] - 100


@pytest.mark.parametrize('arch', ['fastscnn', 'contextnet14'])
def test_one_launch_batchnorm_backward_replaces_reduce_apply_pairs(fake_backend, arch):
    """ops.BN_BWD_ONEPASS: wherever the BatchNorm backward would run its own reduction (no fused reduction in a consumer,
    no activated tensor, no residual output, no SyncBN) and the tensors are small enough, ONE tss_bn_bwd_onepass call
    replaces the tss_bn_bwd_reduce + tss_bn_bwd_apply pair: identical gradients on the emulated ABI."""
    from torch_semantic_segmentation_b200 import ops
    from torch_semantic_segmentation_b200.models.contextnet import contextnet14
    factory = {'fastscnn': fastscnn, 'contextnet14': contextnet14}[arch]
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(2, 3, 64, 96, generator=g), torch.randint(0, 19, (2, 64, 96), generator=g)
    inner, keep, runs = fake_backend.call, ops.BN_BWD_ONEPASS, {}
    try:
        for on in (False, True):
            ops.BN_BWD_ONEPASS = on
            calls = {}

            def counting(name, kwargs, calls=calls):
                calls[name] = calls.get(name, 0) + 1
                return inner(name, kwargs)
            fake_backend.call = counting
            torch.manual_seed(0)
            model = _no_dropout(factory(3, 19)).train()
            loss = CrossEntropyLoss(ignore_index=255)(model(x), y)
            loss.backward()
            runs[on] = (float(loss), [p.grad.clone() for p in model.parameters()], calls)
    finally:
        ops.BN_BWD_ONEPASS = keep
        fake_backend.call = inner
    off, on = runs[False][2], runs[True][2]
    n = on.get('tss_bn_bwd_onepass', 0)
    assert n >= 10 and 'tss_bn_bwd_onepass' not in off
    assert on.get('tss_bn_bwd_reduce', 0) == off['tss_bn_bwd_reduce'] - n
    assert on['tss_bn_bwd_apply'] == off['tss_bn_bwd_apply'] - n
    assert runs[True][0] == runs[False][0]
    for a, b in zip(runs[True][1], runs[False][1]):
        assert torch.equal(a, b)


@pytest.mark.parametrize('loss_name', ['ce', 'ohem'])
def test_deferred_logits_give_the_same_step_without_the_upsampling(fake_backend, loss_name):
    """functional.DEFER_LOGITS (off by default): in training the model hands the fused head its 1/8 scores directly;
    same loss and gradients, one full-resolution up-sampling (and its 269 MB at the benchmark shape) less."""
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.losses import OHEMLoss
    calls = {}
    inner = fake_backend.call

    def counting(name, kwargs):
        calls[name] = calls.get(name, 0) + 1
        return inner(name, kwargs)
    fake_backend.call = counting
    x, y = train_batch('fastscnn')
    loss_fn = CrossEntropyLoss(ignore_index=255) if loss_name == 'ce' else OHEMLoss(ignore_index=255)
    keep = Fn.DEFER_LOGITS
    runs = {}
    try:
        for flag in (False, True):
            Fn.DEFER_LOGITS = flag
            calls.clear()
            torch.manual_seed(0)
            model = _no_dropout(fastscnn(3, 19)).train()
            assert Fn.enable_deferred_logits(model, loss_fn) == flag
            out = model(x)
            assert isinstance(out, Fn.DeferredLogits) == flag and tuple(out.shape) == (x.shape[0], 19, x.shape[2], x.shape[3])
            loss = loss_fn(out, y)
            loss.backward()
            runs[flag] = (float(loss.detach()), torch.cat([p.grad.reshape(-1) for p in model.parameters()]), dict(calls))
            if flag:
                full = out.materialize()                    # anything else can still have the tensor
                assert tuple(full.shape) == tuple(out.shape)
                with torch.no_grad():
                    assert not isinstance(model.eval()(x), Fn.DeferredLogits)      # eval mode always materialises
    finally:
        Fn.DEFER_LOGITS = keep
    assert runs[True][0] == runs[False][0] and rel(runs[True][1], runs[False][1]) < 1e-7
    assert runs[False][2].get('tss_upsample_logits_fwd') == 1 and 'tss_upsample_logits_fwd' not in runs[True][2]
    # a loss that cannot take the handle leaves the model alone
    model = fastscnn(3, 19)
    Fn.DEFER_LOGITS = True
    try:
        assert Fn.enable_deferred_logits(model, torch.nn.CrossEntropyLoss()) is False and model.defer_logits is False
    finally:
        Fn.DEFER_LOGITS = keep


def test_own_dropout_gate(fake_backend):
    """functional.OWN_DROPOUT (off by default): the classifier's nn.Dropout(0.1) on the library kernel -- inverted
    dropout whose backward regenerates the forward's mask; a fresh mask per step; eval mode and p = 0 untouched."""
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.nn.blocks import Dropout
    keep = Fn.OWN_DROPOUT
    Fn.OWN_DROPOUT = True
    try:
        ops.rng_state('cpu', seed=1234)
        d = Dropout(0.1).train()
        x = ops.as_nhwc(torch.ones(2, 16, 5, 7)).requires_grad_()
        y1 = d(x)
        y1.sum().backward()
        kept = y1.detach() != 0
        assert 0.8 < float(kept.float().mean()) < 0.98
        assert torch.allclose(y1.detach()[kept], torch.full((), 1 / 0.9), rtol=1e-4)
        assert torch.equal(x.grad != 0, kept) and torch.allclose(x.grad[kept], torch.full((), 1 / 0.9), rtol=1e-4)
        y2 = d(x)
        assert not torch.equal(y2.detach() != 0, kept)                       # the offset advanced
        assert torch.equal(d.eval()(x), x) and torch.equal(Dropout(0.0).train()(x), x)
        ops.rng_state('cpu', seed=1234)                                      # same seed, same masks
        assert torch.equal(Dropout(0.1).train()(x).detach() != 0, kept)
        x, y = train_batch('fastscnn')
        torch.manual_seed(0)
        model = fastscnn(3, 19).train()
        before = fake_backend.launches
        CrossEntropyLoss(ignore_index=255)(model(x), y).backward()
        assert all(torch.isfinite(p.grad).all() for p in model.parameters())
        assert isinstance(model.classifier[2], torch.nn.Dropout) and list(model.classifier[2].state_dict()) == []
    finally:
        Fn.OWN_DROPOUT = keep


def test_experimental_kernel_tool_dry_run():
    """tools/experimental_kernels.py (GPU microbenchmarks of the gated kernels next to the kernels they replace):
    every call signature is exercised on the emulated ABI so that the tool cannot rot before its next GPU visit."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'tools', 'experimental_kernels.py'), '--dry'], capture_output=True,
                         text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    rows = [json.loads(line) for line in out.stdout.splitlines() if line.startswith('{')]
    assert len(rows) >= 30 and not [r for r in rows if 'error' in r], [r for r in rows if 'error' in r]


def test_forward_hooks_see_activated_outputs_with_deferred_batchnorm(fake_backend):
    """A forward hook on a block whose BatchNorm apply pass would be handed over to its consumer (functional.FUSE_BNIN)
    keeps the block on the ordinary path: the hook sees BatchNorm + ReLU output, as it does on the reference."""
    from torch_semantic_segmentation_b200 import functional as Fn
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 64, 96, generator=g)
    keep, Fn.FUSE_BNIN = Fn.FUSE_BNIN, True
    try:
        torch.manual_seed(0)
        model = _no_dropout(fastscnn(3, 19)).train()
        seen = {}
        taps = {'stem': model.downsample[0], 'conv1': model.features[0][0].conv1}
        handles = [m.register_forward_hook(lambda mod, inp, out, k=k: seen.__setitem__(k, out.detach().clone())) for k, m in taps.items()]
        calls = {}
        inner = fake_backend.call

        def counting(name, kwargs):
            calls[name] = calls.get(name, 0) + 1
            return inner(name, kwargs)
        fake_backend.call = counting
        hooked = model(x).detach()
        n_hooked = calls.get('tss_dwconv3x3_fwd_bnin', 0)
        for h in handles:
            h.remove()
        calls.clear()
        torch.manual_seed(0)
        plain = _no_dropout(fastscnn(3, 19)).train()
        out = plain(x).detach()
    finally:
        Fn.FUSE_BNIN = keep
    assert all(float(v.min()) >= 0.0 for v in seen.values())                   # ReLU outputs, not raw convolutions
    assert all(abs(float(v.float().mean())) > 1e-3 for v in seen.values())
    assert calls['tss_dwconv3x3_fwd_bnin'] == n_hooked + 2                       # exactly the two observed blocks opted out
    assert rel(hooked, out) < 1e-5


@pytest.mark.parametrize('arch', ['fastscnn', 'contextnet14'])
def test_bottleneck_without_the_expanded_activation(fake_backend, arch):
    """functional.FUSE_BNIN: conv1's BatchNorm + ReLU applied by the depthwise conv2 while it reads its
    input; same forward and gradients, one bn_apply less per bottleneck, no other change in the call list."""
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.models.contextnet import contextnet14
    factory = {'fastscnn': fastscnn, 'contextnet14': contextnet14}[arch]
    calls = {}
    inner = fake_backend.call

    def counting(name, kwargs):
        calls[name] = calls.get(name, 0) + 1
        return inner(name, kwargs)
    fake_backend.call = counting
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(3, 3, 64, 96, generator=g), torch.randint(0, 19, (3, 64, 96), generator=g)
    keep = Fn.FUSE_BNIN
    runs = {}
    try:
        for flag in (False, True):
            Fn.FUSE_BNIN = flag
            calls.clear()
            torch.manual_seed(0)
            model = _no_dropout(factory(3, 19)).train()
            out = model(x)
            CrossEntropyLoss(ignore_index=255)(out, y).backward()
            runs[flag] = (out.detach(), torch.cat([p.grad.reshape(-1) for p in model.parameters()]), dict(calls),
                          {k: v.clone().float() for k, v in model.state_dict().items() if 'running' in k})
    finally:
        Fn.FUSE_BNIN = keep
    a, b = runs[False], runs[True]
    assert rel(b[0], a[0]) < 1e-5 and rel(b[1], a[1]) < 2e-3          # fp32; gradients in front of BatchNorm layers cancel
    assert all(rel(b[3][k], a[3][k]) < 1e-5 for k in a[3])
    n = b[2]['tss_dwconv3x3_fwd_bnin']
    assert n >= 9 and b[2]['tss_dwconv3x3_wgrad_bnin'] == n
    assert b[2]['tss_bn_apply'] == a[2]['tss_bn_apply'] - n
    assert b[2]['tss_dwconv3x3_fwd'] == a[2]['tss_dwconv3x3_fwd'] - n


@pytest.mark.parametrize('arch', ['fastscnn', 'contextnet14'])
def test_bottleneck_hand_over_from_depthwise_to_pointwise(fake_backend, arch):
    """functional.FUSE_BNIN_PW (off by default; bf16 tensor-core mode): conv2's BatchNorm + ReLU applied inside conv3's
    GEMM; also together with FUSE_BNIN (conv1 -> conv2): neither activated tensor of the bottleneck is read in forward."""
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.models.contextnet import contextnet14
    from torch_semantic_segmentation_b200.nn.blocks import set_compute_dtype
    factory = {'fastscnn': fastscnn, 'contextnet14': contextnet14}[arch]
    calls = {}
    inner = fake_backend.call

    def counting(name, kwargs):
        calls[name] = calls.get(name, 0) + 1
        return inner(name, kwargs)
    fake_backend.call = counting
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(3, 3, 64, 96, generator=g), torch.randint(0, 19, (3, 64, 96), generator=g)
    keep = Fn.FUSE_BNIN, Fn.FUSE_BNIN_PW
    runs = {}
    try:
        for key in ((False, False), (False, True), (True, True)):
            Fn.FUSE_BNIN, Fn.FUSE_BNIN_PW = key
            calls.clear()
            torch.manual_seed(0)
            model = set_compute_dtype(_no_dropout(factory(3, 19)), torch.bfloat16, pw_impl=1).train()
            seen = {}
            # compared right behind the first bottleneck group: deeper in the tiny net bf16 noise is amplified (PPM bin 1)
            (model.features[0] if arch == 'fastscnn' else model.context[3]).register_forward_hook(
                lambda m, i, o: seen.__setitem__('f', o.detach().float()))
            CrossEntropyLoss(ignore_index=255)(model(x), y).backward()
            runs[key] = (seen['f'], dict(calls), [p.grad for p in model.parameters()])
    finally:
        Fn.FUSE_BNIN, Fn.FUSE_BNIN_PW = keep
    base = runs[False, False]
    for key in ((False, True), (True, True)):
        r = runs[key]
        assert rel(r[0], base[0]) < 6e-2                           # bf16: fma vs mul+add before a rounding, amplified downstream
        assert all(torch.isfinite(gr).all() for gr in r[2])
        n = r[1]['tss_pwconv_fwd_bnin']
        assert n >= 9 and r[1]['tss_pwconv_fwd'] == base[1]['tss_pwconv_fwd'] - n
    n = runs[False, True][1]['tss_pwconv_fwd_bnin']
    assert runs[False, True][1]['tss_bn_apply'] == base[1]['tss_bn_apply'] - n
    assert runs[True, True][1]['tss_bn_apply'] == base[1]['tss_bn_apply'] - n - runs[True, True][1]['tss_dwconv3x3_fwd_bnin']


def test_single_bottleneck_with_both_hand_overs_is_tight(fake_backend):
    """One BottleneckBlock (bf16, tensor-core pointwise mode) with FUSE_BNIN + FUSE_BNIN_PW against the default path.
    With the emulation rounding the on-the-fly activation like a stored tensor the two are BIT-identical (plumbing);
    without it they differ by bf16 rounding only in the forward output (the fused path is the more accurate one)."""
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.nn.blocks import BottleneckBlock, set_compute_dtype
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    keep = Fn.FUSE_BNIN, Fn.FUSE_BNIN_PW
    runs = {}
    try:
        for rounding in (True, False):
            fake_backend.round_in_act = rounding
            for key in ((False, False), (True, True)):
                Fn.FUSE_BNIN, Fn.FUSE_BNIN_PW = key
                for stride, cout in ((1, 64), (2, 96)):
                    torch.manual_seed(0)
                    blk = set_compute_dtype(BottleneckBlock(64, cout, stride=stride), torch.bfloat16, pw_impl=1).train()
                    FlatAdamW(blk.parameters(), lr=1e-3).zero_grad()
                    g = torch.Generator().manual_seed(1)
                    x = ops.as_nhwc(torch.randn(4, 64, 8, 12, generator=g).to(torch.bfloat16)).requires_grad_()
                    out = blk(x)
                    (out.float() * torch.randn(out.shape, generator=g)).sum().backward()
                    runs[rounding, key, stride] = (out.detach().float(), x.grad.float(),
                                                   torch.cat([p.grad.reshape(-1) for p in blk.parameters()]))
    finally:
        Fn.FUSE_BNIN, Fn.FUSE_BNIN_PW = keep
        fake_backend.round_in_act = False
    for stride in (1, 2):
        a, b = runs[True, (False, False), stride], runs[True, (True, True), stride]
        assert all(torch.equal(p, q) for p, q in zip(a, b))
        a, b = runs[False, (False, False), stride], runs[False, (True, True), stride]
        assert rel(b[0], a[0]) < 1e-2


def test_flat_adamw_stages_hyper_parameters_for_graph_replays(fake_backend):
    """A scheduler changes param_groups between steps; write_host_hyper() (called before every graph replay) puts the new
    values where the captured pinned->device copy reads them, and an eager step uses them too."""
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(16))
    opt = FlatAdamW([p], lr=1e-3, weight_decay=1e-5)
    opt.param_groups[0]['lr'] = 5e-4
    opt.write_host_hyper()
    assert abs(float(opt._host[0]) - 5e-4) < 1e-10 and abs(float(opt._host[4]) - 1e-5) < 1e-12
    p.grad.fill_(1.0)
    before = p.detach().clone()
    opt.step()
    assert abs(float(opt.hyper[0]) - 5e-4) < 1e-10
    assert torch.allclose(before - p.detach(), torch.full_like(before, 5e-4), rtol=1e-2)      # first Adam step = lr * sign(g)


def test_stem_backward_without_the_dy_tensor(fake_backend):
    """functional.STEM_BWD_FUSED (with STEM_TC; off by default): the stem's BatchNorm-backward apply happens inside the
    weight-gradient kernel; identical gradients (emulated ABI), one bn_bwd_apply less."""
    from torch_semantic_segmentation_b200 import functional as Fn
    from torch_semantic_segmentation_b200.nn.blocks import set_compute_dtype
    calls = {}
    inner = fake_backend.call

    def counting(name, kwargs):
        calls[name] = calls.get(name, 0) + 1
        return inner(name, kwargs)
    fake_backend.call = counting
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(2, 3, 64, 64, generator=g), torch.randint(0, 19, (2, 64, 64), generator=g)
    keep = Fn.STEM_TC, Fn.STEM_BWD_FUSED
    runs = {}
    try:
        for fused in (False, True):
            Fn.STEM_TC, Fn.STEM_BWD_FUSED = True, fused
            calls.clear()
            torch.manual_seed(0)
            model = set_compute_dtype(_no_dropout(fastscnn(3, 19)), torch.bfloat16, pw_impl=1).train()
            CrossEntropyLoss(ignore_index=255)(model(x), y).backward()
            runs[fused] = ({k: p.grad.clone() for k, p in model.named_parameters()}, dict(calls))
    finally:
        Fn.STEM_TC, Fn.STEM_BWD_FUSED = keep
    for k in runs[False][0]:
        if k == 'downsample.0.0.weight':      # the unfused default multiplies dy with a patch matrix: another summation order
            assert rel(runs[True][0][k], runs[False][0][k]) < 1e-5, k
        else:
            assert torch.equal(runs[True][0][k], runs[False][0][k]), k
    assert runs[True][1]['tss_stem3x3s2_wgrad_tc_bn'] == 1 and 'tss_stem3x3s2_wgrad_tc' not in runs[True][1]
    assert runs[True][1]['tss_bn_bwd_apply'] == runs[False][1]['tss_bn_bwd_apply'] - 1


def test_eval_operands_are_refreshed_in_place_after_training(fake_backend):
    """Folded BatchNorm scale/shift (and bf16 weight packs) are derived once and then refreshed IN PLACE: a captured eval
    graph or an address table keeps reading the same buffers and still sees the statistics of the latest training step.
    engine.GraphedTrainStep bumps ops.WEIGHTS_EPOCH after every replay for the same reason (a replay runs no Python)."""
    from torch_semantic_segmentation_b200.nn.blocks import refresh_cached_operands
    torch.manual_seed(0)
    model = _no_dropout(fastscnn(3, 19))
    x, y = train_batch('fastscnn')
    with torch.no_grad():
        out0 = model.eval()(x)
    blk = model.downsample[0]
    scale0, shift0 = blk._folded(0)
    ptr, before = scale0.data_ptr(), scale0.clone()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    model.train()
    CrossEntropyLoss(ignore_index=255)(model(x), y).backward()
    opt.step()
    refresh_cached_operands(model.eval())
    scale1, shift1 = blk._folded(0)
    assert scale1.data_ptr() == ptr and not torch.equal(scale1, before)                  # same buffer, new values
    bn = blk[1]
    want = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
    assert rel(scale1, want) < 1e-6
    with torch.no_grad():
        out1 = model(x)
    assert not torch.equal(out1, out0)
    epoch = ops.WEIGHTS_EPOCH[0]
    refresh_cached_operands(model)
    assert ops.WEIGHTS_EPOCH[0] == epoch and blk._folded(0)[0].data_ptr() == ptr          # idempotent


def test_gate_defaults_and_environment_override(monkeypatch):
    """Only what has been measured on the GPU is on by default (gates.py); TSS_<NAME> overrides for A/B runs."""
    from torch_semantic_segmentation_b200 import gates
    assert sorted(k for k, v in gates.DEFAULTS.items() if v) == ['CLASS_TC', 'DEFER_LOGITS', 'FUSE_BNIN', 'FUSE_BNRED', 'FUSE_BNRED_EXT', 'FUSE_PPM_EVAL', 'OWN_DROPOUT', 'SLOT_GRAPHS', 'STEM_TC', 'STEM_WGRAD_PATCHES']
    monkeypatch.delenv('TSS_FUSE_PPM', raising=False)
    assert gates.gate('FUSE_PPM') is False and gates.gate('FUSE_BNRED') is True
    monkeypatch.setenv('TSS_FUSE_PPM', '1')
    monkeypatch.setenv('TSS_FUSE_BNRED', '0')
    assert gates.gate('FUSE_PPM') is True and gates.gate('FUSE_BNRED') is False


def test_flat_adamw_state_dict_round_trip(fake_backend):
    """the moments and the step counter live outside Optimizer.state; a resumed run must continue with the
    same bias correction.  Two optimisers, one checkpointed and reloaded after 2 steps, end with identical parameters."""
    from torch_semantic_segmentation_b200.optim import FlatAdamW
    x, y = train_batch('fastscnn')

    def run(resume):
        torch.manual_seed(0)
        model = _no_dropout(fastscnn(3, 19)).train()
        opt = FlatAdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
        loss_fn = CrossEntropyLoss(ignore_index=255)
        for i in range(4):
            if resume and i == 2:
                ckpt = (model.state_dict(), opt.state_dict())
                torch.manual_seed(0)
                model = _no_dropout(fastscnn(3, 19)).train()
                opt = FlatAdamW(model.parameters(), lr=5e-3, weight_decay=0.0)      # other hyper-parameters: must be restored
                model.load_state_dict(ckpt[0])
                opt.load_state_dict(ckpt[1])
                assert opt.param_groups[0]['lr'] == 1e-3 and opt.step_count == 2
            opt.zero_grad()
            loss_fn(model(x), y).backward()
            opt.step()
        return opt.param_arena.clone()
    assert torch.equal(run(False), run(True))

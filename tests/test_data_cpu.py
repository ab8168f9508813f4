"""Input pipeline: the oracle against OpenCV-generated golden vectors, and the
host side of ``data.DeviceTransform`` (draws -> geometry table -> one kernel call) through the emulated ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import augment as A
from torch_semantic_segmentation_b200.data import TRAIN_MAPPING, DeviceTransform, eval_transform

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'augment.npz'))


@pytest.mark.parametrize('case', A.GOLDEN_CASES, ids=lambda c: 'seed%d' % c[0])
def test_oracle_transform_is_bit_exact_with_opencv(case):
    seed, h, w, scale, hf, wf, flip, crop = case
    img, lab = A.sample(seed, h, w)
    nh, nw = A.scaled_size(h, w, scale)
    np.testing.assert_array_equal(A.resize_linear_u8(img, nh, nw), GOLD['resized_%d' % seed])
    x, y = A.train_transform(img, lab, scale, hf, wf, flip, crop)
    np.testing.assert_array_equal(x, GOLD['image_%d' % seed])            # float32, bit for bit
    np.testing.assert_array_equal(y, GOLD['label_%d' % seed].astype(np.int64))
    assert x.dtype == np.float32 and y.dtype == np.int64 and set(np.unique(y)) <= set(range(19)) | {255}


def test_oracle_resize_checksums_at_larger_sizes():
    img, _ = A.sample(99, 256, 512)
    for nh, nw in ((384, 768), (607, 1214), (768, 1536)):
        r = A.resize_linear_u8(img, nh, nw).astype(np.int64)
        sub = r[::3, ::5]
        got = np.array([r.sum(), (r * r).sum(), (sub * np.arange(sub.size).reshape(sub.shape) % 251).sum()])
        np.testing.assert_array_equal(got, GOLD['sum_%dx%d' % (nh, nw)])


def test_oracle_properties():
    img, lab = A.sample(7, 40, 64)
    np.testing.assert_array_equal(A.resize_linear_u8(img, 40, 64), img)                  # scale 1 is the identity
    x, y = A.eval_transform(img, lab)
    np.testing.assert_array_equal(y, A.TRAIN_MAPPING[lab])
    a, la = A.train_transform(img, lab, 2.0, 0.3, 0.4, 0, (32, 48))
    b, lb = A.train_transform(img, lab, 2.0, 0.3, 0.4, 1, (32, 48))
    np.testing.assert_array_equal(a[:, :, ::-1], b)                                      # the flip mirrors the same crop
    np.testing.assert_array_equal(la[:, ::-1], lb)
    assert A.random_scale_factor(0.0) == 1.5 and abs(A.random_scale_factor(1.0) - 3.0) < 1e-12
    assert len(TRAIN_MAPPING) == 35 and np.array_equal(TRAIN_MAPPING, A.TRAIN_MAPPING[:35])
    assert sorted(set(TRAIN_MAPPING.tolist())) == list(range(19)) + [255]


def test_device_transform_host_side(fake_backend):
    cases = [c for c in A.GOLDEN_CASES if c[1:3] == (40, 64)]
    samples = [A.sample(c[0], 40, 64) for c in cases]
    images = torch.from_numpy(np.stack([s[0] for s in samples]))
    labels = torch.from_numpy(np.stack([s[1] for s in samples]))
    t = DeviceTransform(crop=(32, 48), seed=0)
    x, y = t(images, labels, draws=[(c[3], c[4], c[5], bool(c[6])) for c in cases])
    assert x.shape == (len(cases), 3, 32, 48) and x.dtype == torch.float32 and y.dtype == torch.int64
    for i, c in enumerate(cases):
        np.testing.assert_array_equal(x[i].numpy(), GOLD['image_%d' % c[0]])
        np.testing.assert_array_equal(y[i].numpy(), GOLD['label_%d' % c[0]].astype(np.int64))
    # random draws: scale factor from [1.5, 3.0] (RandomScale adds 1 to its limits), crop inside the scaled image
    draws = t.draw(64)
    assert all(1.5 <= d[0] <= 3.0 and 0 <= d[1] < 1 and 0 <= d[2] < 1 for d in draws)
    assert 10 < sum(d[3] for d in draws) < 54
    for nh, nw, cy, cx, flip in t.geometry(draws, 40, 64):
        assert 0 <= cy <= nh - 32 and 0 <= cx <= nw - 48 and flip in (0, 1)
    x2, y2 = t(images, labels)
    assert x2.shape == x.shape and torch.isfinite(x2).all() and int(y2.max()) <= 255
    with pytest.raises(ValueError, match='larger than the image'):
        DeviceTransform(crop=(512, 768), scale_limit=None)(images, labels)
    with pytest.raises(ValueError, match='1/2 shrink'):
        DeviceTransform(crop=(16, 16), scale_limit=None).geometry([(0.5, 0, 0, False)], 40, 64)


def test_eval_transform_and_trainer_on_transformed_batches(fake_backend):
    from torch_semantic_segmentation_b200.engine import create_segmentation_trainer
    from torch_semantic_segmentation_b200.losses import CrossEntropyLoss
    from torch_semantic_segmentation_b200.models import fastscnn
    samples = [A.sample(20 + i, 64, 96) for i in range(2)]
    images = torch.from_numpy(np.stack([s[0] for s in samples]))
    labels = torch.from_numpy(np.stack([s[1] for s in samples]))
    x, y = eval_transform()(images, labels)
    for i in range(2):
        ex, ey = A.eval_transform(*samples[i])
        np.testing.assert_array_equal(x[i].numpy(), ex)
        np.testing.assert_array_equal(y[i].numpy(), ey)
    torch.manual_seed(0)
    model = fastscnn(3, 19)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    trainer = create_segmentation_trainer(model, opt, CrossEntropyLoss(ignore_index=255), 'cpu', logging=False)
    t = DeviceTransform(crop=(64, 96), scale_limit=(0.5, 2.0), seed=1)
    state = trainer.run(list(t.batches([(images, labels)] * 2, 'cpu')), max_epochs=1)
    assert state.iteration == 2 and np.isfinite(state.output)
    # the same through the trainer's own hook: the loader yields the decoded uint8 batch
    trainer = create_segmentation_trainer(model, opt, CrossEntropyLoss(ignore_index=255), 'cpu', logging=False,
                                          transform=DeviceTransform(crop=(64, 96), scale_limit=(0.5, 2.0), seed=1))
    state = trainer.run([(images, labels)] * 2, max_epochs=1)
    assert state.iteration == 2 and np.isfinite(state.output)


def test_kernel_arithmetic_compiled_for_the_host_matches_opencv(tmp_path):
    """csrc/augment_math.h (the per-pixel code of the CUDA kernel) built with g++ and run over the golden cases."""
    import ctypes
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path / 'host_augment.so')
    subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-I',
                           os.path.join(root, 'torch_semantic_segmentation_b200', 'csrc'),
                           os.path.join(root, 'tests', 'host_augment.cpp'), '-o', so])
    fn = ctypes.CDLL(so).host_augment_batch
    fn.restype = None
    t = DeviceTransform(crop=(32, 48))
    norm = np.array(t.norm, dtype=np.float32)
    lut = np.ascontiguousarray(A.TRAIN_MAPPING)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    cases = list(A.GOLDEN_CASES) + [(11, 64, 96, 2.9999, 0.5, 0.5, 1, (64, 96)), (12, 33, 47, 1.7321, 0.2, 0.9, 0, (8, 12))]
    for seed, h, w, scale, hf, wf, flip, crop in cases:
        img, lab = A.sample(seed, h, w)
        t.crop = crop
        geom = np.array(t.geometry([(scale, hf, wf, bool(flip))], h, w), dtype=np.int32)
        x = np.zeros((1, 3) + crop, dtype=np.float32)
        y = np.zeros((1,) + crop, dtype=np.int64)
        fn(ptr(img), ptr(lab), ptr(geom), ptr(lut), ptr(norm), ptr(x), ptr(y), 1, h, w, crop[0], crop[1])
        ex, ey = A.train_transform(img, lab, scale, hf, wf, flip, crop)
        np.testing.assert_array_equal(x[0], ex)
        np.testing.assert_array_equal(y[0], ey)
        if 'image_%d' % seed in GOLD:
            np.testing.assert_array_equal(x[0], GOLD['image_%d' % seed])

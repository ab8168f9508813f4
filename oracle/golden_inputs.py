"""Seeded inputs shared by ``oracle/make_golden.py`` (which feeds them to the reference)
and the tests (which feed them to the oracle and to the CUDA path).  Test infrastructure only.
"""
import torch

SUBSAMPLE = (7, 5)   # stored golden outputs keep every 7th row / 5th column

# gradients stored in full in the golden files (the rest are pinned by checksums)
GRAD_KEYS = {
    'fastscnn': ['downsample.0.0.weight', 'downsample.1.0.weight', 'features.0.0.conv1.0.weight',
                 'features.2.2.conv2.0.weight', 'features.3.pyramids.2.1.0.weight',
                 'features.3.conv.1.weight', 'fusion.lowres.1.0.weight',
                 'classifier.3.weight', 'classifier.3.bias'],
    'contextnet14': ['spatial.0.0.weight', 'context.0.0.weight', 'context.2.conv2.0.weight',
                     'feature_fusion.highres.0.weight',
                     'classifier.5.weight', 'classifier.5.bias'],
}


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def eval_input(arch):
    """1x3x160x224: 1/32 map is 5x7 so that adaptive-pool windows (bins 2, 3, 6) overlap."""
    return torch.randn(1, 3, 160, 224, generator=_gen(11))


def train_batch(arch):
    """2x3x96x160 crops, labels 0..18 with ~10% ignore (255)."""
    g = _gen(12)
    x = torch.randn(2, 3, 96, 160, generator=g)
    y = torch.randint(0, 19, (2, 96, 160), generator=g)
    y[torch.rand(2, 96, 160, generator=g) < 0.1] = 255
    return x, y


def ohem_case(name):
    g = _gen(13)
    logits = torch.randn(2, 19, 24, 32, generator=g) * 3
    target = torch.randint(0, 19, (2, 24, 32), generator=g)
    target[torch.rand(2, 24, 32, generator=g) < 0.1] = 255
    if name == 'many_hard':      # loss[n] > thresh -> keep all above threshold
        return logits, target, dict(ignore_index=255, numel_frac=0.05)
    # few hard pixels: make the logits nearly perfect so that loss[n] <= thresh -> top-n
    onehot = torch.zeros_like(logits).scatter_(1, target.clamp(max=18).unsqueeze(1), 12.0)
    return logits * 0.1 + onehot, target, dict(ignore_index=255, numel_frac=0.05)


def scene_batch(n, h, w, seed, classes=19, ignore_frac=0.05):
    """Cityscapes-like synthetic scenes for whole-network parity at the benchmark shapes: piecewise-smooth class
    regions (19 smooth random fields, argmax), image = class colour + smooth shading + a little sensor noise, every
    image with its own class mix, gain and colour offset; labels carry ``ignore_frac`` of 255.  Unlike white noise
    with independent random labels (the throughput benchmark's input) the labels are a function
    of the image, so the gradients carry signal instead of being the residual of a cancellation, and per-image
    diversity keeps the batch statistics of the pooled pyramid branches away from zero variance."""
    import torch.nn.functional as F
    g = _gen(seed)
    fields = torch.randn(n, classes, max(h // 32, 2), max(w // 32, 2), generator=g)
    fields = fields + 1.5 * torch.randn(n, classes, 1, 1, generator=g)
    fields = F.interpolate(fields, size=(h, w), mode='bilinear', align_corners=False)
    y = fields.argmax(1)
    palette = torch.randn(classes, 3, generator=g)
    x = palette[y].permute(0, 3, 1, 2).contiguous()
    shade = F.interpolate(torch.randn(n, 3, max(h // 16, 2), max(w // 16, 2), generator=g), size=(h, w),
                          mode='bilinear', align_corners=False)
    x = x + 0.5 * shade + 0.1 * torch.randn(n, 3, h, w, generator=g)
    x = x * torch.exp(torch.randn(n, 1, 1, 1, generator=g) * 0.5) + torch.randn(n, 3, 1, 1, generator=g) * 0.7
    y = y.clone()
    y[torch.rand(n, h, w, generator=g) < ignore_frac] = 255
    return x.contiguous(), y

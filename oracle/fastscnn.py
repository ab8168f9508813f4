"""Oracle restatement of ``torch_semantic_segmentation/models/fastscnn.py``.

Oracle / test infrastructure only (see ``oracle/__init__.py``).
"""
import torch
import torch.nn.functional as F

from .blocks import conv_block, dw_block, ds_block, bottleneck, upsample

# (prefix, first-block stride, repeats) of the three BottleneckModules, fastscnn.py:41-44
_STAGES = (('features.0', 2, 3), ('features.1', 2, 3), ('features.2', 1, 3))
PYRAMID_BINS = (1, 2, 3, 6)  # fastscnn.py:103


def downsample(sd, x, training):
    """Learning-to-downsample, fastscnn.py:29-33."""
    x = conv_block(sd, 'downsample.0', x, training, stride=2, padding=1)
    x = ds_block(sd, 'downsample.1', x, training, stride=2, padding=1)
    x = ds_block(sd, 'downsample.2', x, training, stride=2, padding=1)
    return x


def pyramid_pooling(sd, p, x, training):
    """``PyramidPoolingModule`` fastscnn.py:101-123."""
    pools = []
    for i, b in enumerate(PYRAMID_BINS):
        y = F.adaptive_avg_pool2d(x, b)
        y = conv_block(sd, '%s.pyramids.%d.1' % (p, i), y, training)
        pools.append(upsample(y, size=x.shape[2:]))
    x = torch.cat([x, *pools], dim=1)
    return conv_block(sd, p + '.conv', x, training)


def features(sd, x, training):
    """Global feature extractor, fastscnn.py:41-46."""
    for prefix, stride, repeats in _STAGES:
        for r in range(repeats):
            x = bottleneck(sd, '%s.%d' % (prefix, r), x, training,
                           stride=stride if r == 0 else 1)
    return pyramid_pooling(sd, 'features.3', x, training)


def fusion(sd, lowres, highres, training):
    """``FeatureFusionModule`` fastscnn.py:67-89 (scale_factor 4)."""
    lowres = upsample(lowres, scale_factor=4)
    lowres = dw_block(sd, 'fusion.lowres.1', lowres, training, padding=4, dilation=4)
    lowres = conv_block(sd, 'fusion.lowres.2', lowres, training, relu=False)
    highres = conv_block(sd, 'fusion.highres.0', highres, training, relu=False)
    return F.relu(lowres + highres)


def classifier(sd, p, x, training, dropout_mask=None):
    """``Classifier`` fastscnn.py:92-98.  ``dropout_mask`` (already scaled by 1/(1-p))
    replaces ``nn.Dropout(0.1)`` so that oracle and product share one mask; with
    ``None`` in training mode stock ``F.dropout`` is used."""
    x = ds_block(sd, p + '.0', x, training, padding=1)
    x = ds_block(sd, p + '.1', x, training, padding=1)
    if training:
        x = x * dropout_mask if dropout_mask is not None else F.dropout(x, 0.1, True)
    return F.conv2d(x, sd[p + '.3.weight'], sd[p + '.3.bias'])


def forward(sd, x, training=False, dropout_mask=None, return_taps=False):
    """``FastSCNN.forward`` fastscnn.py:57-64."""
    d = downsample(sd, x, training)
    f = features(sd, d, training)
    u = fusion(sd, f, d, training)
    c = classifier(sd, 'classifier', u, training, dropout_mask)
    out = upsample(c, scale_factor=8)
    if return_taps:
        return out, {'downsample': d, 'features': f, 'fusion': u, 'classes': c}
    return out

"""Oracle restatement of ``torch_semantic_segmentation/models/contextnet.py``.

Oracle / test infrastructure only (see ``oracle/__init__.py``).
"""
import torch.nn.functional as F

from .blocks import conv_block, dw_block, bottleneck, upsample

# context branch bottleneck stacks, contextnet.py:47-56: (index, stride of first block, blocks)
_LINEAR = ((3, 2, 3), (4, 2, 3), (5, 1, 2), (6, 1, 2))


def spatial(sd, x, training):
    """Full-resolution branch, contextnet.py:37-45."""
    x = conv_block(sd, 'spatial.0', x, training, stride=2, padding=1)
    x = dw_block(sd, 'spatial.1', x, training, stride=2, padding=1)
    x = conv_block(sd, 'spatial.2', x, training)
    x = dw_block(sd, 'spatial.3', x, training, stride=2, padding=1)
    x = conv_block(sd, 'spatial.4', x, training)
    x = dw_block(sd, 'spatial.5', x, training, stride=1, padding=1)
    x = conv_block(sd, 'spatial.6', x, training)
    return x


def context(sd, x, training):
    """Context branch on the shrunk input, contextnet.py:47-56."""
    x = conv_block(sd, 'context.0', x, training, stride=2, padding=1)
    x = bottleneck(sd, 'context.1', x, training)
    x = bottleneck(sd, 'context.2', x, training)
    for idx, stride, blocks in _LINEAR:
        for b in range(blocks):
            x = bottleneck(sd, 'context.%d.%d' % (idx, b), x, training,
                           stride=stride if b == 0 else 1)
    return conv_block(sd, 'context.7', x, training, padding=1)


def feature_fusion(sd, lowres, highres, training):
    """``FeatureFusionModule`` contextnet.py:104-126."""
    lowres = upsample(lowres, size=highres.shape[2:])
    lowres = dw_block(sd, 'feature_fusion.lowres.0', lowres, training, padding=4, dilation=4)
    lowres = conv_block(sd, 'feature_fusion.lowres.1', lowres, training, relu=False)
    highres = conv_block(sd, 'feature_fusion.highres', highres, training, relu=False)
    return F.relu(lowres + highres)


def classifier(sd, p, x, training, dropout_mask=None):
    """``Classifier`` contextnet.py:79-87."""
    x = dw_block(sd, p + '.0', x, training, padding=1)
    x = conv_block(sd, p + '.1', x, training)
    x = dw_block(sd, p + '.2', x, training, padding=1)
    x = conv_block(sd, p + '.3', x, training)
    if training:
        x = x * dropout_mask if dropout_mask is not None else F.dropout(x, 0.1, True)
    return F.conv2d(x, sd[p + '.5.weight'], sd[p + '.5.bias'])


def forward(sd, x, scale_factor=4, training=False, dropout_mask=None):
    """``ContextNet.forward`` contextnet.py:62-76."""
    s = spatial(sd, x, training)
    c = upsample(x, scale_factor=1 / scale_factor)
    c = context(sd, c, training)
    u = feature_fusion(sd, c, s, training)
    k = classifier(sd, 'classifier', u, training, dropout_mask)
    return upsample(k, scale_factor=8)

from .blocks import (Conv2dBlock, DWConv2dBlock, DSConv2dBlock, ConvBNBlock, DSConvBNBlock,
                     BottleneckBlock, ClassScores, set_compute_dtype)

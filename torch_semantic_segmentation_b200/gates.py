"""Defaults of the kernel-path gates, in one place.

A gate picks between two implementations of the same step (both hand-written CUDA of this library -- never a fallback to
another backend).  ``True`` = measured on a B200 and faster; ``False`` = built and checked on the CPU (emulated ABI, SIMT /
tcgen05 emulation) but not measured yet.  ``TSS_<NAME>=1`` / ``=0`` in the environment overrides
a default for A/B runs (tools/gpu_experimental.sh).
"""
import os

DEFAULTS = {
    'FUSE_BNRED': True,         # BatchNorm-backward reduction in the consumer's dgrad epilogue (4.66 -> 4.57 ms/step)
    'FUSE_BNRED_EXT': False,    # ... also stride-2 depthwise dgrad and more single-consumer pairs
    'FUSE_BNAPPLY': False,      # BatchNorm-backward apply in the operand producer of the pointwise dgrad
    'FUSE_BNAPPLY_DW': False,   # ... and in the stride-1 depthwise dgrad
    'FUSE_BNFIN': False,        # BatchNorm finalize inside the apply kernel
    'FUSE_BNIN': False,         # a block's BatchNorm applied by the depthwise conv that reads it
    'FUSE_BNIN_PW': False,      # ... by the tensor-core pointwise conv that reads it
    'FUSE_PPM': False,          # pyramid-pooling branches as grouped launches
    'STEM_TC': False,           # stem convolution + weight gradient on tcgen05
    'STEM_BWD_FUSED': False,    # stem BatchNorm-backward apply inside its tensor-core weight gradient
    'DEFER_LOGITS': False,      # training forward without the unused full-resolution logits
    'OWN_DROPOUT': False,       # mask-free dropout kernel instead of ATen's
    'SLOT_GRAPHS': False,       # one captured training graph per staging slot of the trainer
}


def gate(name):
    env = os.environ.get('TSS_' + name)
    if env is None or env == '':
        return DEFAULTS[name]
    return env == '1'

"""Defaults of the kernel-path gates, in one place.

A gate picks between two implementations of the same step (both hand-written CUDA of this library -- never a fallback to
another backend).  ``True`` = validated on a B200 (tests/test_fused_paths_gpu.py) and faster in the step A/B; ``False`` = validated on a
B200 where noted, but slower or not faster in the A/B of round 2 (tools/gpu_experimental.sh, profiles/r2a_gate_ab.txt),
or still failing a test.  ``TSS_<NAME>=1`` / ``=0`` in the environment overrides
a default for A/B runs (tools/gpu_experimental.sh).
"""
import os

DEFAULTS = {
    'FUSE_BNRED': True,         # BatchNorm-backward reduction in the consumer's dgrad epilogue (4.66 -> 4.57 ms/step)
    'FUSE_BNRED_EXT': True,     # ... also stride-2 depthwise dgrad and more single-consumer pairs (r2: 4.58 -> 4.48 ms/step)
    'FUSE_BNAPPLY': False,      # BatchNorm-backward apply in the operand producer of the pointwise dgrad
    'FUSE_BNAPPLY_DW': False,   # ... and in the stride-1 depthwise dgrad
    'FUSE_BNFIN': False,        # BatchNorm finalize inside the apply kernel (third version, constants once per CTA through shared memory: 3.204 vs 3.185 ms/step)
    'FUSE_BNIN': True,          # a block's BatchNorm applied by the depthwise conv that reads it (r2: 3.67 -> 3.58 ms/step)
    'FUSE_BNIN_PW': False,      # ... by the tensor-core pointwise conv that reads it (r2: 3.88 vs 3.67 ms/step, not kept)
    'FUSE_PPM': False,          # pyramid-pooling branches as grouped launches (training: 4.60 vs 4.58 ms/step, not kept)
    'FUSE_PPM_EVAL': True,      # ... in eval mode (folded BatchNorm): bs1 inference 2366 -> 2425 FPS
    'STEM_TC': True,            # stem convolution + weight gradient on tcgen05 (r2: fwd 154 -> 95 us, wgrad 246 -> 202 us; 4.58 -> 4.52 ms/step)
    'STEM_BWD_FUSED': False,    # stem BatchNorm-backward apply inside its tensor-core weight gradient
    'DEFER_LOGITS': True,       # training forward without the unused full-resolution logits (r2: 4.58 -> 4.53 ms/step)
    'OWN_DROPOUT': True,        # mask-free dropout kernel instead of ATen's (r2: 28.8 -> 14.1 us per pass)
    'STEM_WGRAD_PATCHES': True,  # stem weight gradient = coalesced patch matrix + the TMA-fed pointwise wgrad GEMM (r2: 4.03 -> 3.93 ms/step)
    'CLASS_TC': True,           # class-score conv (19 classes + bias) on the tcgen05 GEMMs with zero-padded operands (r2: 4.12 -> 4.04 ms/step)
    'BN_BWD_ONEPASS': False,    # BatchNorm backward (reduce + apply) as ONE launch with a grid barrier where dz + y stay in L2
    'SLOT_GRAPHS': True,        # one captured training graph per staging slot of the trainer (no device-to-device batch copy)
}


def gate(name):
    env = os.environ.get('TSS_' + name)
    if env is None or env == '':
        return DEFAULTS[name]
    return env == '1'

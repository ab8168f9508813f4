"""torch_semantic_segmentation_b200 -- the Fast-SCNN / ContextNet training and inference hot
path of bernardomig/torch_semantic_segmentation on hand-written sm_100a (B200) CUDA kernels.

Same Python surface as the reference for this path (``models``, ``losses``, ``engine``,
``wrappers``, ``nn``); everything below the module API runs through the C-ABI library
``libtss_b200.so`` (``include/tss_b200.h``).  No CPU path, no fallback.
"""
__version__ = '0.1.0'

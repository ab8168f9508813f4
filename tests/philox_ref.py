"""Philox4x32-10 keep masks of csrc/dropout.cu restated with numpy (test infrastructure): group g of 8 elements under
(seed, offset) -> 8 keep bits; element e keeps iff its 16-bit uniform >= round(p * 65536)."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def keep_mask(seed, offset, n_elements, p):
    """-> bool array (n_elements,), scale: what tss_dropout_fwd multiplies kept elements with."""
    groups = np.arange(n_elements // 8, dtype=np.uint64)
    c = [groups & MASK, groups >> np.uint64(32), np.full_like(groups, offset & MASK), np.full_like(groups, (offset >> 32) & MASK)]
    k0, k1 = seed & MASK, (seed >> 32) & MASK
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        h0, l0, h1, l1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [h1 ^ c[1] ^ np.uint64(k0), l1, h0 ^ c[3] ^ np.uint64(k1), l0]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    thresh = int(p * 65536.0 + 0.5)
    keep = np.empty((len(groups), 8), dtype=bool)
    for e in range(8):
        keep[:, e] = ((c[e >> 1] >> np.uint64((e & 1) * 16)) & np.uint64(0xFFFF)) >= thresh
    return keep.reshape(-1), np.float32(1.0) / (np.float32(1.0) - np.float32(thresh) / np.float32(65536.0))

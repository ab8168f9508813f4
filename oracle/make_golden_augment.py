"""Generate tests/golden/augment.npz with OpenCV itself (cv2 4.13 in the build container): what the reference's
transform stack (scripts/train_fastscnn.py:62-68, through albumentations -> cv2.resize) yields for the seeded
samples of oracle.augment.GOLDEN_CASES.  albumentations is absent from the image, so its thin wrappers
(scale -> cv2.resize with the truncated size, crop, flip, normalize) are applied by hand here; the pixel
arithmetic that matters for bit-exactness -- cv2.resize INTER_LINEAR / INTER_NEAREST on uint8 -- is OpenCV's.

    python oracle/make_golden_augment.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import augment as A  # noqa: E402


def main():
    out = {}
    for seed, h, w, scale, hf, wf, flip, crop in A.GOLDEN_CASES:
        img, lab = A.sample(seed, h, w)
        nh, nw = int(h * scale), int(w * scale)
        im = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)
        # cv2 has no int64 images: resize the ids and map afterwards (a point-wise table commutes with nearest)
        lb = A.TRAIN_MAPPING[cv2.resize(lab, (nw, nh), interpolation=cv2.INTER_NEAREST)]
        y0, x0 = int((nh - crop[0]) * hf), int((nw - crop[1]) * wf)
        im, lb = im[y0:y0 + crop[0], x0:x0 + crop[1]], lb[y0:y0 + crop[0], x0:x0 + crop[1]]
        if flip:
            im, lb = im[:, ::-1], lb[:, ::-1]
        mean = np.array(A.MEAN, dtype=np.float32) * np.float32(255.0)
        den = np.reciprocal(np.array(A.STD, dtype=np.float32) * np.float32(255.0), dtype=np.float32)
        x = im.astype(np.float32)
        x -= mean
        x *= den
        out['image_%d' % seed] = np.ascontiguousarray(x.transpose(2, 0, 1))
        out['label_%d' % seed] = np.ascontiguousarray(lb).astype(np.uint8)
        out['resized_%d' % seed] = np.ascontiguousarray(cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR))
    # a full-size resize (Cityscapes frame, the reference's extreme scales) pinned by checksums
    img, _ = A.sample(99, 256, 512)
    for nh, nw in ((384, 768), (607, 1214), (768, 1536)):
        r = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR).astype(np.int64)
        out['sum_%dx%d' % (nh, nw)] = np.array([r.sum(), (r * r).sum(), (r[::3, ::5] * np.arange(r[::3, ::5].size).reshape(r[::3, ::5].shape) % 251).sum()])
    path = os.path.join(ROOT, 'tests', 'golden', 'augment.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()

"""The Cityscapes sample transform of the reference, run on the GPU for a whole batch in one kernel.

reference: ``data/cityscapes.py:17-20,81-92`` (``TRAIN_MAPPING``, ``__getitem__``) and the albumentations stacks
of ``scripts/train_fastscnn.py:62-72``::

    train_tfms = albu.Compose([albu.RandomScale([0.5, 2.0]), albu.RandomCrop(512, 768),
                               albu.HorizontalFlip(), albu.Normalize(), ToTensor()])
    val_tfms   = albu.Compose([albu.Normalize(), ToTensor()])

There the DataLoader workers do this per sample on the CPU and the trainer copies fp32 crops (4.7 MB each) plus
int64 masks (3.1 MB) to the device.  Here the loader only has to deliver the *decoded* uint8 frame and label ids
(6.3 + 2.1 MB for 1024 x 2048, one pinned copy); ``DeviceTransform`` draws the random parameters on the host and
``tss_augment_batch`` (csrc/augment.cu) produces the very tensors the reference pipeline would have produced for
those draws -- bit for bit (OpenCV's uint8 bilinear resize is reproduced exactly) -- directly in the layout the
stem kernel reads.  Decoding PNGs and the directory walk stay with the caller (out of scope).
"""
import numpy as np
import torch

from .. import ops

CLASSES = ('unlabeled', 'ego vehicle', 'rectification border', 'out of roi', 'static', 'dynamic', 'ground', 'road',
           'sidewalk', 'parking', 'rail track', 'building', 'wall', 'fence', 'guard rail', 'bridge', 'tunnel', 'pole',
           'polegroup', 'traffic light', 'traffic sign', 'vegetation', 'terrain', 'sky', 'person', 'rider', 'car',
           'truck', 'bus', 'caravan', 'trailer', 'train', 'motorcycle', 'bicycle', 'license plate')
# label id -> train id (19 classes, 255 = ignore); data/cityscapes.py:17-20
_TRAIN_IDS = {7: 0, 8: 1, 11: 2, 12: 3, 13: 4, 17: 5, 19: 6, 20: 7, 21: 8, 22: 9, 23: 10, 24: 11, 25: 12, 26: 13,
              27: 14, 28: 15, 31: 16, 32: 17, 33: 18}
TRAIN_MAPPING = np.array([_TRAIN_IDS.get(i, 255) for i in range(len(CLASSES))])


class DeviceTransform:
    """``transform(images_u8, label_ids_u8) -> (x fp32 (N,3,h,w), y int64 (N,h,w))`` on the device.

    ``crop=(height, width)`` as ``albu.RandomCrop(height, width)``; ``crop=None`` keeps the frame (evaluation).
    ``scale_limit`` as given to ``albu.RandomScale`` (which adds 1 to both ends: ``[0.5, 2.0]`` scales by a
    factor from [1.5, 3.0]); ``None`` = no scaling.  ``flip_p`` as ``albu.HorizontalFlip(p)``.
    ``mean`` / ``std`` as ``albu.Normalize`` (ImageNet defaults, max pixel value 255).
    """

    def __init__(self, crop=(512, 768), scale_limit=(0.5, 2.0), flip_p=0.5, mean=(0.485, 0.456, 0.406),
                 std=(0.229, 0.224, 0.225), mapping=TRAIN_MAPPING, seed=None):
        self.crop = tuple(crop) if crop is not None else None
        self.scale_limit = tuple(scale_limit) if scale_limit is not None else None
        self.flip_p = float(flip_p)
        m = np.array(mean, dtype=np.float32) * np.float32(255.0)
        d = np.reciprocal(np.array(std, dtype=np.float32) * np.float32(255.0), dtype=np.float32)
        self.norm = tuple(float(v) for v in m) + tuple(float(v) for v in d)
        table = np.full(256, 255, dtype=np.int64)
        if mapping is not None:
            table[:len(mapping)] = np.asarray(mapping, dtype=np.int64)
        else:
            table = np.arange(256, dtype=np.int64)
        self._table = torch.from_numpy(table)
        self._lut = {}
        self.rng = np.random.default_rng(seed)

    # ------------------------------------------------------------------ random draws (host)
    def draw(self, n):
        """-> list of (scale, h_frac, w_frac, flip): what RandomScale / RandomCrop / HorizontalFlip draw per sample."""
        out = []
        for _ in range(n):
            if self.scale_limit is not None:
                lo, hi = 1.0 + self.scale_limit[0], 1.0 + self.scale_limit[1]
                scale = lo + (hi - lo) * float(self.rng.random())
            else:
                scale = 1.0
            out.append((scale, float(self.rng.random()), float(self.rng.random()), bool(self.rng.random() < self.flip_p)))
        return out

    def geometry(self, draws, H, W):
        """Draws -> the (N,5) int table of the kernel: scaled size (truncated like albumentations' ``scale``),
        crop start ``int((size - crop) * frac)``, flip."""
        rows = []
        for scale, hf, wf, flip in draws:
            nh, nw = int(H * scale), int(W * scale)
            ch, cw = self.crop if self.crop is not None else (nh, nw)
            if ch > nh or cw > nw:
                raise ValueError('Requested crop size (%d, %d) is larger than the image size (%d, %d)' % (ch, cw, nh, nw))
            if (2 * nh, 2 * nw) == (H, W):
                raise ValueError('an exact 1/2 shrink takes another OpenCV kernel (area); not implemented')
            rows.append((nh, nw, int((nh - ch) * hf), int((nw - cw) * wf), int(bool(flip))))
        return rows

    # ------------------------------------------------------------------ the transform
    def output_size(self, H, W):
        """(height, width) of the transformed batch for frames of H x W (needs a crop or no scaling)."""
        if self.crop is not None:
            return self.crop
        if self.scale_limit is not None:
            raise ValueError('without a crop every sample of the batch needs the same scale')
        return (H, W)

    def draw_geometry(self, N, H, W, draws=None):
        """The random draws of one batch as the kernel's (N,5) int32 table (host tensor, pinned on CUDA builds)."""
        rows = self.geometry(self.draw(N) if draws is None else draws, H, W)
        if self.crop is None and len({(r[0], r[1]) for r in rows}) != 1:
            raise ValueError('without a crop every sample of the batch needs the same scale')
        geom = torch.tensor(rows, dtype=torch.int32)
        return geom.pin_memory() if torch.cuda.is_available() else geom

    def apply(self, images, label_ids, geom):
        """Run the kernel with a geometry table that already is on the device (CUDA-graph friendly: the table can be
        a static buffer refreshed before every replay)."""
        N, H, W, _ = images.shape
        crop = self.crop
        if crop is None:
            if self.scale_limit is not None:
                raise ValueError('without a crop every sample of the batch needs the same scale')
            crop = (H, W)
        dev = images.device
        lut = self._lut.get(dev)
        if lut is None:
            lut = self._lut[dev] = self._table.to(dev)
        return ops.augment_batch(images.contiguous(), label_ids.contiguous() if label_ids is not None else None,
                                 geom, lut, self.norm, crop)

    def __call__(self, images, label_ids=None, draws=None):
        if images.dim() != 4 or images.shape[3] != 3:
            raise ValueError('images: uint8 (N, H, W, 3) expected, got %s' % (tuple(images.shape),))
        N, H, W, _ = images.shape
        geom = self.draw_geometry(N, H, W, draws)
        crop = self.crop if self.crop is not None else (int(geom[0, 0]), int(geom[0, 1]))
        geom = geom.to(images.device, non_blocking=True)
        if self.crop is None and crop != (H, W):
            # a common scale for the whole batch without a crop: the output is the scaled frame
            keep, self.crop = self.crop, crop
            try:
                return self.apply(images, label_ids, geom)
            finally:
                self.crop = keep
        return self.apply(images, label_ids, geom)

    def batches(self, loader, device):
        """Wrap a loader of decoded ``(images_u8, label_ids_u8)`` batches: yields transformed device batches, the
        pinned host->device copy of the uint8 frames issued ``non_blocking``."""
        for images, label_ids in loader:
            yield self(images.to(device, non_blocking=True), label_ids.to(device, non_blocking=True))


def train_transform(seed=None):
    """scripts/train_fastscnn.py:62-68."""
    return DeviceTransform(crop=(512, 768), scale_limit=(0.5, 2.0), flip_p=0.5, seed=seed)


def eval_transform():
    """scripts/train_fastscnn.py:69-72."""
    return DeviceTransform(crop=None, scale_limit=None, flip_p=0.0)

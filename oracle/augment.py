"""Oracle for the input pipeline either side of the hot path.  Test infrastructure
only: numpy restatement of what the reference's Cityscapes training transform does to one decoded sample,

    label = TRAIN_MAPPING[label]                                         data/cityscapes.py:17-20,88
    albu.RandomScale([0.5, 2.0]) -> albu.RandomCrop(512, 768) -> albu.HorizontalFlip()
    -> albu.Normalize() -> ToTensor()                                     scripts/train_fastscnn.py:62-68

with the random draws (scale, crop start fractions, flip) given as arguments.  The arithmetic lives in two
third-party packages that are neither vendored nor pinned by the reference (setup.py:3-9): albumentations
(not installed here; its published formulas are restated below) and OpenCV (cv2 4.13 in the build container).
The uint8 bilinear resize restates OpenCV's fixed-point INTER_LINEAR path (11-bit coefficients, the
horizontal pass in int32, the vertical pass with its >>4, >>16, +2, >>2 rounding) and is PINNED bit-for-bit by
tests/golden/augment.npz, generated with cv2.resize itself by oracle/make_golden_augment.py.
"""
import numpy as np

# data/cityscapes.py:17-20 (35 label ids -> 19 train ids, 255 = ignore); ids outside the table -> ignore
TRAIN_MAPPING = np.full(256, 255, dtype=np.int64)
TRAIN_MAPPING[:35] = [255, 255, 255, 255, 255, 255, 255, 0, 1, 255, 255, 2, 3, 4, 255,
                      255, 255, 5, 255, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 255,
                      255, 16, 17, 18, 255]
MEAN = (0.485, 0.456, 0.406)      # albu.Normalize() defaults, max_pixel_value = 255
STD = (0.229, 0.224, 0.225)
COEF_BITS = 11                    # OpenCV INTER_RESIZE_COEF_BITS


def scaled_size(h, w, scale):
    """albumentations F.scale: the new size truncates."""
    return int(h * scale), int(w * scale)


def random_scale_factor(u, scale_limit=(0.5, 2.0)):
    """albu.RandomScale(scale_limit): the limits get a bias of +1 (to_tuple(scale_limit, bias=1.0)), so
    RandomScale([0.5, 2.0]) draws uniformly from [1.5, 3.0]; ``u`` in [0, 1) is the uniform draw."""
    lo, hi = 1.0 + scale_limit[0], 1.0 + scale_limit[1]
    return lo + (hi - lo) * u


def crop_start(size, crop, frac):
    """albumentations get_random_crop_coords: int((size - crop) * frac) with frac in [0, 1)."""
    return int((size - crop) * frac)


def _linear_taps(dn, sn, clamp):
    """Source index pairs and 11-bit weights of OpenCV's INTER_LINEAR for ``dn`` outputs over ``sn`` inputs.
    Horizontal taps (clamp=True) zero the fraction when the index is clamped; vertical taps keep the fraction
    and only clamp the row index."""
    scale = 1.0 / (dn / sn)                                         # double, as cv::resize computes it
    d = np.arange(dn, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp:
        lo = s < 0
        f[lo], s[lo] = 0, 0
        hi = s >= sn - 1
        f[hi], s[hi] = 0, sn - 1
    w1 = np.rint(f * np.float32(1 << COEF_BITS)).astype(np.int64)                  # saturate_cast<short>: round half even
    w0 = np.rint((np.float32(1) - f) * np.float32(1 << COEF_BITS)).astype(np.int64)
    return np.clip(s, 0, sn - 1), np.clip(s + 1, 0, sn - 1), w0, w1


def resize_linear_u8(img, dh, dw):
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR) for uint8 HxWxC, bit-exact (not the exact-1/2
    shrink, which OpenCV routes to its area kernel)."""
    sh, sw = img.shape[:2]
    if (dh, dw) == (sh, sw):
        return img.copy()
    x0, x1, ax0, ax1 = _linear_taps(dw, sw, True)
    y0, y1, by0, by1 = _linear_taps(dh, sh, False)
    src = img.astype(np.int64)
    rows = src[:, x0] * ax0[None, :, None] + src[:, x1] * ax1[None, :, None]
    r0, r1 = rows[y0], rows[y1]
    out = (((by0[:, None, None] * (r0 >> 4)) >> 16) + ((by1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def nearest_index(dn, sn):
    """OpenCV INTER_NEAREST: min(floor(d * (1 / (dn / sn))), sn - 1) in double."""
    scale = 1.0 / (dn / sn)
    return np.minimum(np.floor(np.arange(dn, dtype=np.float64) * scale).astype(np.int64), sn - 1)


def resize_nearest(img, dh, dw):
    sh, sw = img.shape[:2]
    return img[nearest_index(dh, sh)][:, nearest_index(dw, sw)]


def normalize(img_u8, mean=MEAN, std=STD):
    """albumentations F.normalize: float32, subtract mean*255, multiply by the float32 reciprocal of std*255."""
    m = np.array(mean, dtype=np.float32) * np.float32(255.0)
    d = np.reciprocal(np.array(std, dtype=np.float32) * np.float32(255.0), dtype=np.float32)
    out = img_u8.astype(np.float32)
    out -= m
    out *= d
    return out


def train_transform(image, label_ids, scale, h_frac, w_frac, flip, crop=(512, 768)):
    """One sample: uint8 RGB HxWx3, uint8 label ids HxW -> float32 (3, ch, cw), int64 (ch, cw)."""
    h, w = image.shape[:2]
    nh, nw = scaled_size(h, w, scale)
    img = resize_linear_u8(image, nh, nw)
    lab = resize_nearest(TRAIN_MAPPING[label_ids], nh, nw)
    y0, x0 = crop_start(nh, crop[0], h_frac), crop_start(nw, crop[1], w_frac)
    img, lab = img[y0:y0 + crop[0], x0:x0 + crop[1]], lab[y0:y0 + crop[0], x0:x0 + crop[1]]
    if flip:
        img, lab = img[:, ::-1], lab[:, ::-1]
    return np.ascontiguousarray(normalize(img).transpose(2, 0, 1)), np.ascontiguousarray(lab).astype(np.int64)


def eval_transform(image, label_ids):
    """albu.Normalize() + ToTensor() only (scripts/train_fastscnn.py:69-72)."""
    return (np.ascontiguousarray(normalize(image).transpose(2, 0, 1)), TRAIN_MAPPING[label_ids].astype(np.int64))


def sample(seed, h, w):
    """Seeded synthetic decoded sample: uint8 RGB image with smooth + noisy content, label ids 0..34."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = (np.sin(yy / 7.0)[..., None] * 60 + np.cos(xx / 5.0)[..., None] * 60 + 128)
    img = np.clip(base + rng.integers(-40, 41, (h, w, 3)), 0, 255).astype(np.uint8)
    lab = rng.integers(0, 35, (h // 4 + 1, w // 4 + 1), dtype=np.uint8).repeat(4, 0).repeat(4, 1)[:h, :w]
    return img, np.ascontiguousarray(lab)


# (seed, H, W, scale, h_frac, w_frac, flip, crop) of the golden cases
GOLDEN_CASES = [
    (1, 40, 64, 1.5, 0.0, 0.0, 0, (32, 48)),
    (2, 40, 64, 2.3711, 0.37, 0.81, 1, (32, 48)),
    (3, 40, 64, 3.0, 0.999, 0.999, 1, (32, 48)),
    (4, 37, 53, 1.9, 0.5, 0.25, 0, (24, 40)),
    (5, 32, 48, 1.0, 0.0, 0.0, 1, (32, 48)),
    (6, 48, 96, 0.77, 0.6, 0.1, 0, (32, 48)),
]

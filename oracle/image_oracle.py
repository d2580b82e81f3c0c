"""CPU oracle of `load_images`' per-image pixel work (TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this; the product never does).

The reference (dust3r/utils/image.py:62-71 `_resize_pil_image`, :74-128 `load_images`) calls three library routines on every
decoded image; none of their sources is part of /root/reference, so their published algorithms are restated here in scalar
Python / numpy integer arithmetic and PINNED against the libraries themselves (tests/test_image_oracle.py: Pillow 12.2's
`Image.resize` on random and natural-statistics images, enlarging and shrinking; torchvision 0.26 `ToTensor` + `Normalize`;
and the unmodified reference `load_images` end to end when /root/reference is mounted, plus a committed golden fixture):

  * Pillow `Image.resize(size, LANCZOS | BICUBIC)` on 8-bit RGB = src/libImaging/Resample.c: `precompute_coeffs` (double
    precision windowed filter, support scaled by the shrink factor, weights normalised per output pixel),
    `normalize_coeffs_8bpc` (weights rounded to 22-bit fixed point), then a HORIZONTAL pass and a VERTICAL pass, each
    accumulating int32 sums from 1 << 21, shifting right by 22 and clipping to [0, 255] (the intermediate image is uint8).
  * Pillow `Image.crop(box)` with `box` rounded by Python's round().
  * torchvision `ToTensor` (uint8 HWC -> float CHW / 255) and `Normalize(0.5, 0.5)` ((x - 0.5) / 0.5 in fp32).

Bit-exact by construction: everything after the coefficient tables is integer arithmetic, and the last step has only 256
possible inputs."""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c: coefficients are stored with 22 fractional bits
LANCZOS, BICUBIC = 'lanczos', 'bicubic'


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x):                      # Resample.c lanczos_filter, support 3
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def _bicubic(x):                      # Resample.c bicubic_filter (a = -0.5), support 2
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


_FILTERS = {LANCZOS: (_lanczos, 3.0), BICUBIC: (_bicubic, 2.0)}


def coefficients(in_size, out_size, method):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box (0, in_size).
    Returns (bounds int32 [out_size][2] = (first source index, tap count), coefs int32 [out_size][ksize])."""
    filt, fsupport = _FILTERS[method]
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coefs = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)             # C cast: truncation towards zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            coefs[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, coefs


def _pass(src, bounds, coefs, axis):
    """One separable pass along `axis` (0 = vertical, 1 = horizontal) of a uint8 (H, W, C) image."""
    src = np.moveaxis(src.astype(np.int64), axis, 0)     # resampled axis first
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.uint8)
    for o in range(bounds.shape[0]):
        lo, cnt = int(bounds[o, 0]), int(bounds[o, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for k in range(cnt):
            acc += src[lo + k] * int(coefs[o, k])
        acc = ((acc + (1 << 31)) & 0xFFFFFFFF) - (1 << 31)          # the C code accumulates in a 32-bit int
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_u8(img, out_w, out_h, method):
    """Pillow Image.resize((out_w, out_h), method) on an (H, W, 3) uint8 array: horizontal pass, then vertical pass, each
    skipped when that dimension does not change (ImagingResample: need_horizontal / need_vertical)."""
    h, w = img.shape[:2]
    if out_w != w:
        bx, kx = coefficients(w, out_w, method)
        img = _pass(img, bx, kx, axis=1)
    if out_h != h:
        by, ky = coefficients(h, out_h, method)
        img = _pass(img, by, ky, axis=0)
    return img


def resized_shape(w, h, size):
    """dust3r/utils/image.py:62-71 + :101-106: (new_w, new_h, method) for a decoded image of w x h."""
    long_edge = round(size * max(w / h, h / w)) if size == 224 else size
    s = max(w, h)
    method = LANCZOS if s > long_edge else BICUBIC
    return int(round(w * long_edge / s)), int(round(h * long_edge / s)), method


def crop_box(w, h, size, square_ok=False, patch_size=16):
    """dust3r/utils/image.py:107-117 followed by PIL.Image.crop's rounding of the box: (left, upper, right, lower) ints."""
    cx, cy = w // 2, h // 2
    if size == 224:
        half = min(cx, cy)
        box = (cx - half, cy - half, cx + half, cy + half)
    else:
        halfw = ((2 * cx) // patch_size) * patch_size / 2
        halfh = ((2 * cy) // patch_size) * patch_size / 2
        if not square_ok and w == h:
            halfh = 3 * halfw / 4
        box = (cx - halfw, cy - halfh, cx + halfw, cy + halfh)
    return tuple(int(round(v)) for v in box)


def normalise(img_u8):
    """torchvision ToTensor + Normalize((0.5,)*3, (0.5,)*3): uint8 (H, W, 3) -> float32 (3, H, W)."""
    x = img_u8.astype(np.float32) / np.float32(255)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(np.moveaxis(x, -1, 0))


def load_image_oracle(img_u8, size, square_ok=False, patch_size=16):
    """Decoded RGB uint8 (H, W, 3) -> (float32 (1, 3, H2, W2) in [-1, 1], true_shape int32 [[H2, W2]]): what one entry of
    the reference's load_images() holds for that image."""
    h, w = img_u8.shape[:2]
    nw, nh, method = resized_shape(w, h, size)
    res = resize_u8(img_u8, nw, nh, method)
    l, u, r, b = crop_box(nw, nh, size, square_ok, patch_size)
    res = res[u:b, l:r]
    return normalise(res)[None], np.int32([[res.shape[0], res.shape[1]]])

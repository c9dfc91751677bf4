"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

CPU restatement (numpy, integer arithmetic) of the image resize the reference's data loaders apply to every colour
frame: `transforms.Resize((h, w), interpolation=Image.ANTIALIAS)` on a PIL RGB image
(`DepthNetworks/monodepth2/datasets/mono_dataset.py:71, 100-104, 119-131`) -> `PIL.Image.resize(size, LANCZOS)`.

The arithmetic lives in a third-party dependency that is not under /root/reference: **Pillow** (pinned
`pillow=8.1.0` in `requirements.txt`; installed here: 12.2.0 -- the 8-bit resampling code is unchanged between the
two).  Its published algorithm (libImaging/Resample.c), restated:

  * per axis, for output index i: scale = in/out, filterscale = max(scale, 1), support = 3 * filterscale,
    center = (i + 0.5) * scale, xmin = max(int(center - support + 0.5), 0),
    xmax = min(int(center + support + 0.5), in) - xmin, weights w_j = lanczos((j + xmin - center + 0.5) / filterscale)
    normalised by their sum (double precision; lanczos(x) = sinc(x) * sinc(x / 3) on [-3, 3));
  * the weights become integers: int(+-0.5 + w * 2^22);
  * horizontal pass first, then vertical, each producing 8-bit pixels:
    out = clip8((2^21 + sum_j pixel_j * k_j) >> 22)   (arithmetic shift, then clamp to 0..255);
  * a pass whose axis keeps its size is skipped.

Pinned by `tests/test_loader_compose.py::test_oracle_resize_equals_pillow` against the installed Pillow itself
(random and structured images, the loader's four pyramid sizes and ragged ones): bit-exact.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
LANCZOS_SUPPORT = 3.0


def _sinc(x: float) -> float:
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x: float) -> float:
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def coefficients(in_size: int, out_size: int):
    """(bounds int32 [out,2] = (xmin, count), kk int32 [out, ksize]) of Resample.c precompute_coeffs +
    normalize_coeffs_8bpc for the full-image box (0, in_size)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = LANCZOS_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray) -> np.ndarray:
    """One resampling pass along the LAST axis of a uint8 array."""
    out_size, ksize = kk.shape
    src = img.astype(np.int64)
    acc = np.full(img.shape[:-1] + (out_size,), 1 << (PRECISION_BITS - 1), dtype=np.int64)
    idx = bounds[:, 0][:, None] + np.arange(ksize)[None, :]                 # [out, ksize]
    valid = np.arange(ksize)[None, :] < bounds[:, 1][:, None]
    idx = np.where(valid, idx, 0)
    k = np.where(valid, kk, 0).astype(np.int64)
    for j in range(ksize):
        acc += src[..., idx[:, j]] * k[:, j]
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_lanczos_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img: uint8 [..., H, W] (planes are independent) -> uint8 [..., out_h, out_w]; == PIL resize(LANCZOS)."""
    assert img.dtype == np.uint8
    H, W = img.shape[-2:]
    out = img
    if out_w != W:
        b, k = coefficients(W, out_w)
        out = _pass(out, b, k)
    if out_h != H:
        b, k = coefficients(H, out_h)
        out = np.swapaxes(_pass(np.swapaxes(out, -1, -2), b, k), -1, -2)
    return np.ascontiguousarray(out)


def pyramid_u8(img: np.ndarray, height: int, width: int, num_scales: int = 4):
    """MonoDataset.preprocess (mono_dataset.py:119-131): scale i is resized from scale i-1 (scale -1 = native)."""
    out = []
    cur = img
    for i in range(num_scales):
        cur = resize_lanczos_u8(cur, height // (2 ** i), width // (2 ** i))
        out.append(cur)
    return out

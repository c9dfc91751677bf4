"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.  Oracle of the colour jitter of the loader
row: `dmh_color_jitter_u8` (csrc/loader_compose.cu, csrc/jitter_math.cuh) is checked against it.

CPU restatement (numpy) of the brightness / contrast / saturation steps that `transforms.ColorJitter` applies to the
8-bit PIL frames when the reference trains with contrastive learning (`mono_dataset.py:297, 344-350`; torchvision
`functional_pil.adjust_brightness / adjust_contrast / adjust_saturation` -> `PIL.ImageEnhance` -> `Image.blend`).
The arithmetic lives in Pillow (libImaging/Blend.c, Convert.c), restated:

  grey      L = (R * 19595 + G * 38470 + B * 7471 + 0x8000) >> 16                           (Convert.c rgb2l)
  blend     out = (uint8)(a + f * (b - a)) in single precision for 0 <= f <= 1, truncating;
            outside that range the float result is clipped to [0, 255] before the truncation   (Blend.c)
  brightness  blend(black, img, f);  contrast  blend(grey level int(mean(L) + 0.5), img, f);
  saturation  blend(L replicated to RGB, img, f)

  hue       RGB -> HSV (Convert.c rgb2hsv_row, after colorsys: float quotients, double branch arithmetic, fmod,
            truncation to bytes), H shifted by uint8(f * 255) with wrap-around, HSV -> RGB (hsv2rgb: C round())

Pinned bit-exactly against the installed Pillow / torchvision by
tests/test_loader_compose.py::test_oracle_enhance_equals_pillow (random images, a dense sweep of the colour cube,
factors inside and outside [0, 1]).
"""
from __future__ import annotations

import numpy as np


def grey(img: np.ndarray) -> np.ndarray:
    """img uint8 [3,H,W] -> uint8 [H,W] (PIL convert('L'))."""
    r, g, b = (img[i].astype(np.int64) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(a: np.ndarray, b: np.ndarray, f: float) -> np.ndarray:
    """PIL.Image.blend(a, b, f) on uint8 arrays of one shape."""
    if f == 0.0:
        return a.copy()
    if f == 1.0:
        return b.copy()
    fa = np.float32(f)
    t = a.astype(np.float32) + fa * (b.astype(np.int32) - a.astype(np.int32)).astype(np.float32)
    if 0.0 <= f <= 1.0:
        return t.astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t)).astype(np.uint8)


def brightness(img: np.ndarray, f: float) -> np.ndarray:
    return blend(np.zeros_like(img), img, f)


def contrast(img: np.ndarray, f: float) -> np.ndarray:
    mean = int(grey(img).astype(np.float64).mean() + 0.5)
    return blend(np.full_like(img, mean), img, f)


def saturation(img: np.ndarray, f: float) -> np.ndarray:
    return blend(np.broadcast_to(grey(img), img.shape).copy(), img, f)


def rgb_to_hsv(img: np.ndarray) -> np.ndarray:
    """PIL convert('HSV') (Convert.c rgb2hsv_row, after colorsys): uint8 [3,H,W] -> uint8 [3,H,W]."""
    r, g, b = (img[i].astype(np.int32) for i in range(3))
    maxc = np.maximum(r, np.maximum(g, b))
    minc = np.minimum(r, np.minimum(g, b))
    flat = maxc == minc
    cr = np.where(flat, 1, maxc - minc).astype(np.float32)
    s = cr / np.where(flat, 1, maxc).astype(np.float32)                      # float / float
    rc = (maxc - r).astype(np.float32) / cr
    gc = (maxc - g).astype(np.float32) / cr
    bc = (maxc - b).astype(np.float32) / cr
    # the branch arithmetic mixes double constants with float operands: evaluated in double
    rc64, gc64, bc64 = rc.astype(np.float64), gc.astype(np.float64), bc.astype(np.float64)
    h = np.where(r == maxc, (bc - gc).astype(np.float64),                      # float - float, then widened
                 np.where(g == maxc, 2.0 + rc64 - bc64, 4.0 + gc64 - rc64))
    h = h.astype(np.float32).astype(np.float64)                                # `h` is a float variable
    h = np.fmod(h / 6.0 + 1.0, 1.0).astype(np.float32).astype(np.float64)
    uh = np.clip((h * 255.0).astype(np.int64), 0, 255)
    us = np.clip((s.astype(np.float64) * 255.0).astype(np.int64), 0, 255)
    uh = np.where(flat, 0, uh)
    us = np.where(flat, 0, us)
    return np.stack([uh, us, maxc]).astype(np.uint8)


def hsv_to_rgb(hsv: np.ndarray) -> np.ndarray:
    """PIL HSV -> RGB (Convert.c hsv2rgb): uint8 [3,H,W] -> uint8 [3,H,W]."""
    h, s, v = (hsv[i].astype(np.int32) for i in range(3))
    hf = h.astype(np.float32).astype(np.float64) * 6.0 / 255.0
    i = np.floor(hf).astype(np.int32)
    f = (hf - i.astype(np.float32).astype(np.float64)).astype(np.float32).astype(np.float64)
    fs = (s.astype(np.float32).astype(np.float64) / 255.0).astype(np.float32).astype(np.float64)
    vf = v.astype(np.float32).astype(np.float64)
    rnd = lambda x: np.where(x >= 0, np.floor(x + 0.5), np.ceil(x - 0.5)).astype(np.int64)   # C round(): half away
    p = np.clip(rnd(vf * (1.0 - fs)), 0, 255)
    q = np.clip(rnd(vf * (1.0 - fs * f)), 0, 255)
    t = np.clip(rnd(vf * (1.0 - fs * (1.0 - f))), 0, 255)
    k = i % 6
    r = np.choose(k, [v, q, p, p, t, v])
    g = np.choose(k, [t, v, v, q, p, p])
    b = np.choose(k, [p, p, t, v, v, q])
    grey_px = s == 0
    return np.stack([np.where(grey_px, v, r), np.where(grey_px, v, g), np.where(grey_px, v, b)]).astype(np.uint8)


def hue(img: np.ndarray, f: float) -> np.ndarray:
    """torchvision functional_pil.adjust_hue: H channel of the HSV image shifted by uint8(f * 255) with wrap-around."""
    hsv = rgb_to_hsv(img)
    hsv[0] = (hsv[0].astype(np.int32) + int(np.uint8(int(f * 255) & 0xff))).astype(np.uint8)   # uint8 wrap
    return hsv_to_rgb(hsv)


def jitter(img: np.ndarray, fn_idx, brightness_factor, contrast_factor, saturation_factor, hue_factor) -> np.ndarray:
    """torchvision `ColorJitter.forward` on an 8-bit RGB image [3,H,W] with the parameters `ColorJitter.get_params`
    drew (the reference draws them once per item and applies them to every pyramid level, mono_dataset.py:344-350,
    140-144): the four steps in the order `fn_idx`; a factor of None switches its step off."""
    for fn_id in [int(i) for i in fn_idx]:
        if fn_id == 0 and brightness_factor is not None:
            img = brightness(img, float(brightness_factor))
        elif fn_id == 1 and contrast_factor is not None:
            img = contrast(img, float(contrast_factor))
        elif fn_id == 2 and saturation_factor is not None:
            img = saturation(img, float(saturation_factor))
        elif fn_id == 3 and hue_factor is not None:
            img = hue(img, float(hue_factor))
    return img

"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

numpy restatement of the tube-light candidate of the reference's black-box light attack:
  * `wavelength_to_rgb`                 torchattacks/attacks/light_simulation.py:40-86
  * `tube_light`                        light_simulation.py:132-170 (tube_light_generation_by_func; the reference
                                        fills the image in a Python double loop, here the same float64 arithmetic
                                        vectorised -- operation order kept, see the comments)
  * `add_light_u8`                      phy_obj_atk_light.py:118-121 + light_simulation.py:23-28 (simple_add):
                                        light * 255.0 -> float32 -> cv2.resize to the SAME size (a copy) ->
                                        cv2.addWeighted(base, 1, light, 1, 0) (one fp32 add) -> clip -> uint8 (trunc)
  * `candidate_params`                  phy_obj_atk_light.py:75-86, 100-116: the numpy RNG walk over
                                        (wavelength, angle, intercept, beta)
  * `candidate_patch`                   ToPILImage (mul(255).byte()) -> candidate -> ToTensor (fp32 / 255)

Pinned to the reference: tests/golden/light.npz holds outputs of the reference's OWN functions
(oracle/make_golden_light.py, cv2 from this image), tests/test_oracle_golden.py compares bit for bit.
"""
from __future__ import annotations

import math

import numpy as np

# phy_obj_atk_light.py:75-86: the ten search directions over (wavelength, angle, intercept, beta)
Q = np.asarray([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [1, 1, 0, 0], [1, 0, 1, 0], [1, 0, 0, 1],
                [0, 1, 1, 0], [0, 1, 0, 1], [0, 0, 1, 1]])
Q_LO, Q_HI = [380, 0, 0, 10], [750, 180, 400, 1600]


def wavelength_to_rgb(wavelength, gamma=0.8):
    """light_simulation.py:40-86, python floats."""
    wavelength = float(wavelength)
    if 380 <= wavelength <= 440:
        att = 0.3 + 0.7 * (wavelength - 380) / (440 - 380)
        return (((-(wavelength - 440) / (440 - 380)) * att) ** gamma, 0.0, (1.0 * att) ** gamma)
    if 440 <= wavelength <= 490:
        return (0.0, ((wavelength - 440) / (490 - 440)) ** gamma, 1.0)
    if 490 <= wavelength <= 510:
        return (0.0, 1.0, (-(wavelength - 510) / (510 - 490)) ** gamma)
    if 510 <= wavelength <= 580:
        return (((wavelength - 510) / (580 - 510)) ** gamma, 1.0, 0.0)
    if 580 <= wavelength <= 645:
        return (1.0, (-(wavelength - 645) / (645 - 580)) ** gamma, 0.0)
    if 645 <= wavelength <= 750:
        att = 0.3 + 0.7 * (750 - wavelength) / (750 - 645)
        return ((1.0 * att) ** gamma, 0.0, 0.0)
    return (0.0, 0.0, 0.0)


def slope_of(angle_deg) -> float:
    """phy_obj_atk_light.py:113-114."""
    return round(math.tan(math.radians(angle_deg)), 2)


def light_ends(beta):
    """light_simulation.py:152-153: (full_light_end_y, light_end_y)."""
    return int(math.sqrt(beta) + 0.5), int(math.sqrt(beta * 20) + 0.5)


def tube_light(k, b, alpha, beta, wavelength, w=400, h=400):
    """light_simulation.py:132-170 -> (h, w, 3) float64.  Per pixel: distance = |k*x - y + b| / sqrt(1 + k*k);
    inside the core the colour * alpha, in the skirt additionally * beta / distance^2, else 0."""
    full_end, light_end = light_ends(beta)
    c = wavelength_to_rgb(wavelength)
    x = np.arange(w, dtype=np.float64)[None, :]
    y = np.arange(h, dtype=np.float64)[:, None]
    norm = math.sqrt(1 + k * k)
    dist = np.abs(np.float64(k) * x - y + np.float64(b)) / norm                 # ((k*x) - y) + b, then / sqrt
    out = np.zeros((h, w, 3))
    core = dist <= full_end
    skirt = (dist > full_end) & (dist <= light_end)
    with np.errstate(divide="ignore", invalid="ignore"):
        att = np.where(skirt, np.float64(beta) / (dist * dist), 0.0)
    for ch in range(3):
        ca = c[ch] * alpha                                                       # c * alpha, then * attenuation
        out[..., ch] = np.where(core, ca, np.where(skirt, ca * att, 0.0))
    return out


def add_light_u8(base_u8, light):
    """base_u8 (h, w, 3) uint8, light (h, w, 3) float64 in [0, ~1] -> (h, w, 3) uint8."""
    light32 = (light * 255.0).astype(np.float32)
    s = base_u8.astype(np.float32) + light32                                     # addWeighted(base, 1, light, 1, 0)
    return np.clip(s, 0.0, 255.0).astype("uint8")


def to_u8_hwc(obj):
    """ToPILImage on a float (3, h, w) tensor in [0, 1]: mul(255).byte(), HWC."""
    return np.ascontiguousarray(np.transpose((np.asarray(obj, dtype=np.float32) * np.float32(255)).astype(np.uint8), (1, 2, 0)))


def candidate_patch(base_u8, params):
    """One candidate of the search: params = (wavelength, angle, intercept, beta) after clipping ->
    (3, h, w) float32 patch (ToTensor of the lit 8-bit image)."""
    h, w = base_u8.shape[:2]
    wl, ang, icpt, beta = params
    light = tube_light(slope_of(ang), icpt, 1.0, beta, wl, w=w, h=h)
    lit = add_light_u8(base_u8, light)
    return np.transpose(lit, (2, 0, 1)).astype(np.float32) / np.float32(255)


def candidate_params(n_init=200, n_search=20):
    """The parameter walk of phy_obj_atk_light.py:96-116 on the GLOBAL numpy RNG (same draw order): yields the
    clipped int64 4-vector of every candidate."""
    inits = []
    for _ in range(n_init):
        inits.append([np.random.randint(380, 750), np.random.randint(0, 180), np.random.randint(0, 400),
                      np.random.randint(10, 1600)])
    for init_v in inits:
        for _ in range(n_search):
            q = Q[np.random.randint(len(Q))]
            q = q * np.random.randint(1, 20)
            for a in (-1, 1):
                yield np.clip(init_v + a * q, Q_LO, Q_HI)

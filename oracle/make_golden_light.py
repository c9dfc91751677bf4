"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/light.npz from the UNMODIFIED reference functions
`tube_light_generation_by_func`, `simple_add`, `wavelength_to_rgb` (/root/reference/torchattacks/attacks/
light_simulation.py:23-28, 40-86, 132-170; matplotlib stubbed by oracle/refload.py, cv2 from this image) and the
parameter walk of `Phy_obj_atk_light.forward` (phy_obj_atk_light.py:96-116, re-run here on the seeded numpy RNG with the
reference's own statements).
    python -m oracle.make_golden_light
"""
from __future__ import annotations

import importlib
import math
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import refload  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# (wavelength, angle, intercept, beta, w, h): every colour band, steep / flat / negative slopes, both clip ends
CASES = [(380, 0, 0, 10, 48, 40), (439, 35, 12, 900, 48, 40), (470, 89, 30, 1600, 40, 56), (500, 91, 399, 333, 40, 56),
         (545, 135, 20, 77, 64, 32), (600, 179, 5, 1234, 64, 32), (700, 45, 0, 10, 33, 31), (750, 160, 350, 9, 48, 40),
         (612, 72, 150, 640, 300, 260)]


def main():
    refload.load()
    old = os.getcwd()
    os.chdir(refload.M2_DIR)
    try:
        ls = importlib.import_module("torchattacks.attacks.light_simulation")
    finally:
        os.chdir(old)
    out = {"cases": np.array(CASES, dtype=np.int64)}
    rs = np.random.RandomState(3)
    for i, (wl, ang, icpt, beta, w, h) in enumerate(CASES):
        # the arguments exactly as phy_obj_atk_light.py:109-118 forms them (numpy int64 entries of the clipped vector)
        temp_q = np.clip(np.array([wl, ang, icpt, beta]), [380, 0, 0, 10], [750, 180, 400, 1600])
        k = round(math.tan(math.radians(temp_q[1])), 2)
        light = ls.tube_light_generation_by_func(k, temp_q[2], alpha=1.0, beta=temp_q[3], wavelength=temp_q[0], w=w, h=h)
        base = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        lit = np.clip(ls.simple_add(base, light * 255.0, 1.0), 0.0, 255.0).astype("uint8")
        big = w * h >= 10000        # the bases are RandomState(3) draws in case order: the test regenerates them
        out["light_%d" % i] = light[::7, ::7].copy() if big else light
        out["lit_%d" % i] = lit[::3, ::3].copy() if big else lit
        out["lit_crc_%d" % i] = np.array(zlib.crc32(np.ascontiguousarray(lit).tobytes()), dtype=np.int64)
    out["rgb"] = np.array([ls.wavelength_to_rgb(w) for w in range(370, 761)], dtype=np.float64)
    # the parameter walk: the reference's statements (phy_obj_atk_light.py:96-116) on a seeded global RNG
    Q = np.asarray([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [1, 1, 0, 0], [1, 0, 1, 0], [1, 0, 0, 1],
                    [0, 1, 1, 0], [0, 1, 0, 1], [0, 0, 1, 1]])
    np.random.seed(41)
    params_list = []
    for i in range(3):
        init_v_it = [np.random.randint(380, 750), np.random.randint(0, 180), np.random.randint(0, 400),
                     np.random.randint(10, 1600)]
        params_list.append(init_v_it)
    walk = []
    for init_v in params_list:
        for search_i in range(4):
            q_id = np.random.randint(len(Q))
            q = Q[q_id]
            step_size = np.random.randint(1, 20)
            q = q * step_size
            for a in [-1, 1]:
                temp_q = init_v + a * q
                temp_q = np.clip(temp_q, [380, 0, 0, 10], [750, 180, 400, 1600])
                walk.append(temp_q)
    out["walk"] = np.array(walk, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, "light.npz"), **out)
    print("light ok", {k: v.shape for k, v in out.items() if k.startswith(("lit", "walk"))})


if __name__ == "__main__":
    main()

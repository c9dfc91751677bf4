"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/attack_vanila.npz by running the UNMODIFIED reference class
`Phy_obj_atk_vanila` (/root/reference/torchattacks/attacks/phy_obj_atk_vanila.py:18-96) on seeded synthetic inputs
(`synth.patch_batch`), CPU, with the same import stubs as the other goldens (oracle/refload.py).
    python -m oracle.make_golden_vanila
"""
from __future__ import annotations

import importlib
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from depthmodelhardening_b200 import synth  # noqa: E402
from oracle import refload  # noqa: E402
from oracle.make_golden import TinyDepth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ref = refload.load()  # noqa: F841 (registers the torchattacks shell, my_utils, calib)
    old = os.getcwd()
    os.chdir(refload.M2_DIR)
    try:
        mod = importlib.import_module("torchattacks.attacks.phy_obj_atk_vanila")
    finally:
        os.chdir(old)
    torch.set_num_threads(8)
    pbt = synth.patch_batch(batch=3, seed=0)
    other = synth.rand((1, 3, synth.PATCH_H, synth.PATCH_W), 77)          # the object image handed to forward()
    out = {}
    for tag, ev in (("rand", False), ("eval", True)):
        random.seed(11)
        atk = mod.Phy_obj_atk_vanila(TinyDepth(), pbt.obj.clone(), pbt.mask.clone(), dist_range=list(range(5, 10, 2)))
        adv_s, ben_s, m_out, obj_adv = atk(pbt.scenes.clone(), other.clone(), 3, eval=ev)
        out[tag + "_adv_sum"] = adv_s.detach().double().sum().numpy()
        out[tag + "_ben_sum"] = ben_s.detach().double().sum().numpy()
        out[tag + "_mask_sum"] = m_out.detach().double().sum().numpy()
        out[tag + "_adv_crop"] = adv_s.detach()[:, :, 90:200:2, 380:640:2].numpy()
        out[tag + "_ben_crop"] = ben_s.detach()[:, :, 90:200:2, 380:640:2].numpy()
        out[tag + "_mask_crop"] = m_out.detach()[:, :, 90:200:2, 380:640:2].numpy()
        assert torch.equal(obj_adv, other)
    np.savez_compressed(os.path.join(GOLD, "attack_vanila.npz"), **out)
    print("vanila ok", {k: float(v) for k, v in out.items() if v.ndim == 0})


if __name__ == "__main__":
    main()

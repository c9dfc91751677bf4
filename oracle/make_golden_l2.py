"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/attack_l2.npz by running the UNMODIFIED reference class
`Phy_obj_atk_l2` (/root/reference/torchattacks/attacks/phy_obj_atk_l2.py:13-136) on seeded synthetic inputs, CPU,
batch_size 1 (the only size its `.view(batch_size, -1)` of the shared patch's gradient supports), with the import
stubs of oracle/refload.py.  Also stores one update step of the oracle restatement's inputs / output.
    python -m oracle.make_golden_l2
"""
from __future__ import annotations

import contextlib
import importlib
import io
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
    refload.load()
    old = os.getcwd()
    os.chdir(refload.M2_DIR)
    try:
        mod = importlib.import_module("torchattacks.attacks.phy_obj_atk_l2")
    finally:
        os.chdir(old)
    torch.set_num_threads(8)
    pbt = synth.patch_batch(batch=1, seed=0)
    out = {}
    for tag, ev in (("rand", False), ("eval", True)):
        random.seed(21)
        atk = mod.Phy_obj_atk_l2(TinyDepth(), pbt.obj.clone(), pbt.mask.clone(), eps=3.0, steps=3, random_start=False,
                                 dist_range=list(range(5, 10, 2)))
        with contextlib.redirect_stdout(io.StringIO()):          # the reference prints tensor sizes every iteration (:99)
            adv_s, ben_s, m_out, obj_adv = atk(pbt.scenes.clone(), 1, eval=ev)
        out[tag + "_obj_adv"] = obj_adv.detach()[:, :, ::2, ::2].numpy()
        out[tag + "_delta_norm"] = (obj_adv.detach() - pbt.obj).double().norm().numpy()
        out[tag + "_adv_sum"] = adv_s.detach().double().sum().numpy()
        out[tag + "_ben_sum"] = ben_s.detach().double().sum().numpy()
        out[tag + "_mask_sum"] = m_out.detach().double().sum().numpy()
        out[tag + "_adv_crop"] = adv_s.detach()[:, :, 90:200:2, 380:640:2].numpy()
    np.savez_compressed(os.path.join(GOLD, "attack_l2.npz"), **out)
    print("l2 ok", {k: float(v) for k, v in out.items() if v.ndim == 0})


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/md_cost_volume.npz by calling the UNMODIFIED
ManyDepth reference method `ResnetEncoderMatching.match_features`
(/root/reference/DepthNetworks/manydepth2/networks/resnet_encoder.py:157-236) unbound, with the
reference's own BackprojectDepth / Project3D, on seeded synthetic features.  Own process (module names
clash with the other trees):   python -m oracle.make_golden_md
"""
from __future__ import annotations

import importlib
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from depthmodelhardening_b200 import synth  # noqa: E402
from oracle import refload  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MD_DIR = os.path.join(refload.REF_ROOT, "DepthNetworks", "manydepth2")


cost_volume_inputs = synth.cost_volume_inputs      # seeded input factory lives with the other synthetic inputs


def load_md():
    assert "networks" not in sys.modules and "layers" not in sys.modules, "run in a fresh process"
    for p in (refload.REF_ROOT, MD_DIR):
        sys.path.insert(0, p)
    old = os.getcwd()
    os.chdir(MD_DIR)
    try:
        layers = importlib.import_module("layers")
        enc = importlib.import_module("networks.resnet_encoder")
    finally:
        os.chdir(old)
    assert enc.__file__.startswith(MD_DIR), enc.__file__
    return layers, enc


def main():
    layers, enc = load_md()
    cur, look, poses, K, inv_K, bins = cost_volume_inputs()
    B, L, C, h, w = look.shape
    D = bins.numel()
    me = SimpleNamespace(num_depth_bins=D, matching_height=h, matching_width=w, set_missing_to_max=True,
                         backprojector=layers.BackprojectDepth(batch_size=D, height=h, width=w),
                         projector=layers.Project3D(batch_size=D, height=h, width=w),
                         warp_depths=torch.stack([torch.ones((1, h, w)) * d for d in bins], 0).float())
    with torch.no_grad():
        vol, miss = enc.ResnetEncoderMatching.match_features(me, cur, look, poses, K, inv_K)
        me.set_missing_to_max = False
        vol_raw, _ = enc.ResnetEncoderMatching.match_features(me, cur, look, poses, K, inv_K)
    np.savez_compressed(os.path.join(GOLD, "md_cost_volume.npz"), cost_volume=vol.numpy(), missing=miss.numpy(),
                        cost_volume_raw=vol_raw.numpy())
    print("md cost volume", tuple(vol.shape), float(vol.mean()), "missing frac", float(miss.mean()))


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/loader_compose.npz by running the UNMODIFIED reference methods
`MonoDataset.prep_adv_data` and `MonoDataset.preprocess`
(/root/reference/DepthNetworks/monodepth2/datasets/mono_dataset.py:186-265, 119-144) -- called unbound on a
namespace that carries exactly the attributes `__init__` / `set_adv_train` would have set (:71-116, 146-175); no
KITTI files are needed -- on seeded synthetic 8-bit frames, CPU, with the import stubs of oracle/refload.py.
    python -m oracle.make_golden_loader
"""
from __future__ import annotations

import importlib
import os
import sys
import types
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from depthmodelhardening_b200 import synth  # noqa: E402
from oracle import refload  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
H, W, S = 320, 1024, 4
CASES = [("l", False, 7, -10), ("r", False, 5, 15), ("l", True, 9, 0), ("r", True, 7, 25)]


def frames_u8(seed):
    return synth.frames_u8(seed)[0].numpy()


def crc(t):
    return np.uint32(zlib.crc32(np.ascontiguousarray(t).tobytes()))


def main():
    from PIL import Image
    from torchvision import transforms
    ref = refload.load()
    old = os.getcwd()
    os.chdir(refload.M2_DIR)
    try:
        md = importlib.import_module("datasets.mono_dataset")
    finally:
        os.chdir(old)
    PT = ref.physicalTrans.PhysicalTrans
    pbt = synth.patch_batch(batch=1, seed=0)
    obj_ben, mask = pbt.obj, pbt.mask
    obj_adv = synth.rand(obj_ben.shape, 78)
    cfg = {"path": ref.calib_path}
    me = types.SimpleNamespace()
    me.to_tensor, me.to_pilimage = transforms.ToTensor(), transforms.ToPILImage()
    me.ori_H, me.ori_W = synth.ORI_H, synth.ORI_W
    me.resize_trans = transforms.Resize([me.ori_H, me.ori_W])
    me.num_scales = S
    me.resize = {i: transforms.Resize((H // 2 ** i, W // 2 ** i), interpolation=transforms.InterpolationMode.LANCZOS)
                 for i in range(S)}                        # == interpolation=Image.ANTIALIAS (:71, 100-104)
    me.half_no_synthesis = False
    me.ben_trans = PT(obj_ben, mask, cfg, (1, 3, me.ori_H, me.ori_W), dist_range=list(range(5, 10, 2)))
    me.adv_trans = PT(obj_adv, mask, cfg, (1, 3, me.ori_H, me.ori_W), dist_range=list(range(5, 10, 2)))
    me.adv_K = np.array([[0.58, 0, 0.5, 0], [0, 1.92, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
    me.adv_K[0, :] *= me.ori_W
    me.adv_K[1, :] *= me.ori_H
    me.stereo_T = np.eye(4, dtype=np.float32)
    me.stereo_T[0, 3] = -0.54
    out = {}
    for ci, (side, flip, z0, alpha) in enumerate(CASES):
        c0, cs = frames_u8(1000 + 2 * ci), frames_u8(1001 + 2 * ci)
        pil = lambda a: Image.fromarray(np.transpose(a, (1, 2, 0)))
        inputs = {("color", 0, -1): pil(c0), ("color", "s", -1): pil(cs)}
        # the placement draw of project(batch_size=1) (:207 / :216) is pinned through random.sample's two calls
        seq = iter([[z0], [alpha]])
        orig = ref.physicalTrans.sample
        ref.physicalTrans.sample = lambda rng, n: next(seq)
        try:
            md.MonoDataset.prep_adv_data(me, inputs, side, flip)
        finally:
            ref.physicalTrans.sample = orig
        md.MonoDataset.preprocess(me, inputs, (lambda x: x))
        tag = "c%d_" % ci
        for k, v in inputs.items():
            if not torch.is_tensor(v) or k[-1] == -1 or k[0] == "objdepth":
                continue
            name = tag + "%s_%s_%d" % k
            u8 = (v * 255.0).round().to(torch.uint8).numpy()
            assert torch.equal(torch.from_numpy(u8).float().div(255), v), k      # to_tensor output is k/255
            out[name + "_crc"] = crc(u8)
            out[name + "_sum"] = np.int64(u8.astype(np.int64).sum())
        out[tag + "objdepth"] = inputs[("objdepth", 0, 0)].numpy()
    np.savez_compressed(os.path.join(GOLD, "loader_compose.npz"), **out)
    print("loader golden ok:", len(out), "entries",
          os.path.getsize(os.path.join(GOLD, "loader_compose.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()

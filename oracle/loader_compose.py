"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

CPU restatement of the reference's adversarial training-item construction
(`DepthNetworks/monodepth2/datasets/mono_dataset.py`):

  prep_adv_data (:186-265)  to_tensor -> PhysicalTrans.project / project_w_trans -> flip -> composite -> to_pilimage
  preprocess    (:119-144)  Lanczos pyramid on the 8-bit images (oracle/pil_resize.py) -> to_tensor

for ONE item, with the colour jitter off.  Pinned by `tests/golden/loader_compose.npz`, produced by
`oracle/make_golden_loader.py` from the unmodified `MonoDataset.prep_adv_data` / `MonoDataset.preprocess`.
"""
from __future__ import annotations

import numpy as np
import torch

from . import patch as OQ
from . import pil_resize as R

STEREO_T = np.eye(4, dtype=np.float32)
STEREO_T[0, 3] = -0.54                                           # mono_dataset.py:111-116 (side "l")


def adv_K(ori_H=375, ori_W=1242):
    K = np.array([[0.58, 0, 0.5, 0], [0, 1.92, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
    K[0, :] *= ori_W
    K[1, :] *= ori_H
    return K                                                     # mono_dataset.py:169-175


def to_pil_u8(t: torch.Tensor) -> np.ndarray:
    """torchvision to_pil_image on a float tensor: `.mul(255).byte()` -> uint8 planes."""
    return t.mul(255).byte().numpy()


def compose_u8(scene_u8: np.ndarray, obj: torch.Tensor, mask: torch.Tensor, flip: bool) -> np.ndarray:
    """scene_u8 [3,H,W] uint8, obj [1,3,H,W], mask [1,1,H,W] fp32 -> uint8 [3,H,W] (mono_dataset.py:193, 222-236)."""
    s = torch.from_numpy(scene_u8).to(torch.float32).div(255).unsqueeze(0)     # to_tensor
    if flip:
        obj, mask = torch.flip(obj, [3]), torch.flip(mask, [3])
    return to_pil_u8((s * (1 - mask) + obj * mask).squeeze(0))


def raw_item(color_0: np.ndarray, color_s: np.ndarray, height=320, width=1024, num_scales=4):
    """An item of the `half_no_synthesis` branch that drew no synthesis (mono_dataset.py:325-328): color_aug and
    color_ben are the raw frames; then preprocess (:119-144)."""
    tt = lambda a: torch.from_numpy(a).to(torch.float32).div(255)
    out = {}
    for fid, img in ((0, color_0), ("s", color_s)):
        for i, lvl in enumerate(R.pyramid_u8(img, height, width, num_scales)):
            out[("color", fid, i)] = out[("color_aug", fid, i)] = tt(lvl)
    out[("color_ben", 0, 0)] = out[("color", 0, 0)]
    return out


def prep_item(color_0: np.ndarray, color_s: np.ndarray, side: str, do_flip: bool, z0, alpha, obj_adv, obj_ben, mask,
              P34, height=320, width=1024, num_scales=4):
    """One item: uint8 frames [3,375,1242] -> dict of fp32 tensors keyed like the reference's `inputs`."""
    K = adv_K(color_0.shape[1], color_0.shape[2])
    T0, Ts = (None, STEREO_T) if side == "l" else (STEREO_T, None)
    adv_0, mask_0, _ = OQ.project_patch(obj_adv, mask, [z0], [alpha], P34, K=K, T=T0)
    ben_0, _, _ = OQ.project_patch(obj_ben, mask, [z0], [alpha], P34, K=K, T=T0)
    ben_s, mask_s, _ = OQ.project_patch(obj_ben, mask, [z0], [alpha], P34, K=K, T=Ts)
    aug_0 = compose_u8(color_0, adv_0, mask_0, do_flip)
    aug_s = compose_u8(color_s, ben_s, mask_s, do_flip)
    ben = compose_u8(color_0, ben_0, mask_0, do_flip)
    m0 = torch.flip(mask_0, [3]) if do_flip else mask_0
    objmask = to_pil_u8(m0.expand(-1, 3, -1, -1).squeeze(0))
    tt = lambda a: torch.from_numpy(a).to(torch.float32).div(255)
    out = {}
    for name, fid, img in (("color_aug", 0, aug_0), ("color_aug", "s", aug_s), ("color", 0, ben)):
        for i, lvl in enumerate(R.pyramid_u8(img, height, width, num_scales)):
            out[(name, fid, i)] = tt(lvl)
    for i in range(num_scales):
        out[("color", "s", i)] = out[("color_aug", "s", i)]
    out[("color_ben", 0, 0)] = out[("color", 0, 0)]
    out[("color_objmask", 0, 0)] = tt(R.resize_lanczos_u8(objmask, height, width))
    out[("objdepth", 0, 0)] = torch.FloatTensor([[z0]])
    return out

"""Drop-in `PhysicalTrans` (reference: /root/reference/physicalTrans.py).

Same constructor, attributes and method signatures.  `project` /
`project_w_trans` run ONE batched homography solve on the host and ONE kernel
launch per tensor for the whole batch, instead of a Python loop of 2*Ba
`torchvision.perspective` calls (each with its own host-synchronous 8x8 solve).

CUDA tensors only (the attack-side instances, `phy_obj_atk*.py:54-55`); the
DataLoader-worker instances of the reference run on CPU tensors and stay bound to
the reference class (SURVEY.md 8(b) threading caveat) -- `install()` dispatches
on `obj_img.is_cuda`.
"""
from __future__ import annotations

from random import sample

import numpy as np
import torch

from . import patch_ops

ori_H, ori_W = patch_ops.ORI_H, patch_ops.ORI_W


def read_calib_P2(path: str) -> np.ndarray:
    """P2 (3x4) of a KITTI-object calib file (preprocessing/kitti_util.py:59-62, 80-97)."""
    with open(path, "r") as f:
        for line in f:
            line = line.rstrip()
            if not line:
                continue
            key, value = line.split(":", 1)
            if key == "P2":
                return np.array([float(x) for x in value.split()], dtype=np.float64).reshape(3, 4)
    raise KeyError("P2")


class PhysicalTrans(object):
    def __init__(self, obj_img, obj_mask, cfg, output_size, angle_range=list(range(-30, 31, 5)),
                 dist_range=list(range(5, 10, 2))) -> None:
        super().__init__()
        self.obj_img = obj_img
        self.obj_mask = obj_mask
        self.cfg = cfg
        self.P = read_calib_P2(cfg["path"])
        self.dist_range = dist_range
        self.angle_range = angle_range
        self.output_size = output_size
        assert output_size[2] == ori_H and output_size[3] == ori_W     # physicalTrans.py:32
        self.x0 = 0
        self.y0 = patch_ops.CAM_H - patch_ops.VEH_H / 2
        self.m = patch_ops.VEH_W
        self.n = patch_ops.VEH_H
        self._coeff_cache = {}
        self.padding_img()

    # -- geometry -------------------------------------------------------------
    def fromZA2Coord(self, z0, alpha):
        return patch_ops.plane_corners(z0, alpha)

    def objPosOnImage(self, z0, alpha, K=None):
        return patch_ops.project_corners(z0, alpha, self.P, K)

    def padding_img(self):
        """physicalTrans.py:107-122 -- only the start corners are needed: the
        kernels read the un-padded patch and apply the pad offsets themselves, so
        the 5.6 MB padded canvas is never materialised (nor re-padded every PGD step)."""
        _, _, H, W = self.obj_img.size()
        self.pos_obj_img_start = patch_ops.start_corners((H, W), (self.output_size[2], self.output_size[3]))

    def reset_img(self, obj_img, obj_mask):
        self.obj_img = obj_img
        self.obj_mask = obj_mask
        self.padding_img()

    # -- projection -----------------------------------------------------------
    def _coeffs(self, z0_sample, alpha_sample, K=None, T=None):
        key = (tuple(float(z) for z in z0_sample), tuple(float(a) for a in alpha_sample),
               None if K is None else np.asarray(K).tobytes(), None if T is None else np.asarray(T).tobytes(),
               tuple(map(tuple, self.pos_obj_img_start)))
        co = self._coeff_cache.get(key)
        if co is None:
            ends = np.stack([patch_ops.project_corners(z, a, self.P, K, T) for z, a in zip(z0_sample, alpha_sample)])
            _, _, h, w = self.obj_img.size()
            co = patch_ops.make_placement(self.pos_obj_img_start, ends, (h, w),
                                          (self.output_size[2], self.output_size[3])).to(self.obj_img.device)
            if len(self._coeff_cache) > 256:
                self._coeff_cache.clear()
            self._coeff_cache[key] = co
        return co

    def _warp(self, coeffs):
        hw = (self.output_size[2], self.output_size[3])
        return patch_ops.perspective_batch(self.obj_img, coeffs, hw), \
            patch_ops.perspective_batch(self.obj_mask, coeffs, hw)

    def project(self, is_all=False, batch_size=1, z0_sample=None, alpha_sample=None, K=None,
                rs: np.random.RandomState = None):
        if is_all:
            z0_sample, alpha_sample = [], []
            for z0 in self.dist_range:
                for alpha in self.angle_range:
                    z0_sample.append(z0)
                    alpha_sample.append(alpha)
        else:
            # same RNG consumption order as physicalTrans.py:146-155
            if isinstance(z0_sample, type(None)):
                z0_sample = rs.choice(self.dist_range, batch_size, replace=False) if rs else \
                    sample(self.dist_range, batch_size)
            if isinstance(alpha_sample, type(None)):
                alpha_sample = rs.choice(self.angle_range, batch_size, replace=False) if rs else \
                    sample(self.angle_range, batch_size)
        n = len(z0_sample) if is_all else batch_size       # the reference indexes range(batch_size)
        imgs, masks = self._warp(self._coeffs([z0_sample[i] for i in range(n)], [alpha_sample[i] for i in range(n)], K))
        return imgs, masks, z0_sample, alpha_sample

    def project_w_trans(self, T: np.ndarray, z0_sample, alpha_sample, K=None):
        return self._warp(self._coeffs(z0_sample, alpha_sample, K, T))

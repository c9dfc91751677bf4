"""Seeded synthetic inputs for the hot path.

Everything is drawn on the CPU from `torch.Generator().manual_seed(seed)` so
that the oracle (CPU) and the CUDA path see bit-identical inputs, then moved.
Shapes follow the reference's batch dictionary (`M2/datasets/mono_dataset.py:
333-373`, `M2/trainer.py:472-523`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

ORI_H, ORI_W = 375, 1242          # my_utils.py:12-13
PATCH_H, PATCH_W = 260, 300       # image_preprocess.py:69-83 (BMW.png resized)
SCENE_H, SCENE_W = 320, 1024      # phy_obj_atk.py:50


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed(int(seed))
    return g


def rand(shape, seed) -> torch.Tensor:
    return torch.rand(tuple(shape), generator=_gen(seed), dtype=torch.float32)


def randn(shape, seed) -> torch.Tensor:
    return torch.randn(tuple(shape), generator=_gen(seed), dtype=torch.float32)


def smooth_field(shape, seed, down=8, noise=0.05) -> torch.Tensor:
    """U[0,1) at 1/down resolution, bilinearly up-sampled, plus small noise."""
    b, c, h, w = shape
    lo = rand((b, c, max(h // down, 2), max(w // down, 2)), seed)
    up = F.interpolate(lo, size=(h, w), mode="bilinear", align_corners=False)
    if noise:
        up = up + noise * rand(shape, seed + 50000)
    return up.clamp_(0.0, 1.0).contiguous()


def intrinsics(height: int, width: int, batch: int):
    """KITTI normalised K scaled to (height,width); inv_K = pinv(K)
    (`M2/datasets/kitti_dataset.py:29-32`, `mono_dataset.py:333-342`)."""
    K = np.array([[0.58, 0, 0.5, 0], [0, 1.92, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
    K[0, :] *= width
    K[1, :] *= height
    inv_K = np.linalg.pinv(K)
    K = torch.from_numpy(K).unsqueeze(0).repeat(batch, 1, 1).contiguous()
    inv_K = torch.from_numpy(inv_K.astype(np.float32)).unsqueeze(0).repeat(batch, 1, 1).contiguous()
    return K, inv_K


def stereo_T(batch: int, sign: float = 1.0) -> torch.Tensor:
    T = torch.eye(4, dtype=torch.float32).unsqueeze(0).repeat(batch, 1, 1)
    T[:, 0, 3] = sign * 0.1           # mono_dataset.py:367-373
    return T.contiguous()


def temporal_T(batch: int, seed: int = 500) -> torch.Tensor:
    """Small random rigid motion: axis-angle ~ N(0,0.01^2), trans ~ N(0,0.05^2)
    composed as T = Trans @ Rot (`M2/layers.py:28-45, 64-103`), in fp64 then cast."""
    aa = (randn((batch, 3), seed) * 0.01).double()
    tr = (randn((batch, 3), seed + 1) * 0.05).double()
    out = torch.zeros(batch, 4, 4, dtype=torch.float64)
    for b in range(batch):
        ang = aa[b].norm().item()
        ax = aa[b] / (ang + 1e-7)
        ca, sa = math.cos(ang), math.sin(ang)
        C = 1 - ca
        x, y, z = ax.tolist()
        R = torch.tensor([[x * x * C + ca, x * y * C - z * sa, x * z * C + y * sa],
                          [x * y * C + z * sa, y * y * C + ca, y * z * C - x * sa],
                          [x * z * C - y * sa, y * z * C + x * sa, z * z * C + ca]], dtype=torch.float64)
        out[b, :3, :3] = R
        out[b, :3, 3] = tr[b]
        out[b, 3, 3] = 1.0
    return out.float().contiguous()


@dataclass
class PhotoBatch:
    """One stage-2 batch: what `Trainer.generate_images_pred/compute_losses` read."""
    batch: int
    height: int
    width: int
    frame_ids: List            # e.g. [0, "s"] or [0, -1, 1]
    scales: List[int]
    color: Dict                # (frame_id, scale) -> (B,3,h,w); all scales only for frame 0
    disp: Dict                 # scale -> (B,1,h_s,w_s), requires grad at call time
    K: torch.Tensor            # (B,4,4) scale 0
    inv_K: torch.Tensor
    T: Dict                    # frame_id -> (B,4,4)
    noise: Dict                # scale -> (B,F,H,W) tie-break noise (already * 1e-5)
    min_depth: float = 0.1
    max_depth: float = 100.0
    extras: Dict = field(default_factory=dict)

    def to(self, device, non_blocking=False):
        mv = lambda t: t.to(device, non_blocking=non_blocking)
        return PhotoBatch(self.batch, self.height, self.width, list(self.frame_ids), list(self.scales),
                          {k: mv(v) for k, v in self.color.items()}, {k: mv(v) for k, v in self.disp.items()},
                          mv(self.K), mv(self.inv_K), {k: mv(v) for k, v in self.T.items()},
                          {k: mv(v) for k, v in self.noise.items()}, self.min_depth, self.max_depth,
                          {k: (mv(v) if torch.is_tensor(v) else v) for k, v in self.extras.items()})


def photo_batch(batch=4, height=192, width=640, frame_ids: Sequence = (0, "s"), scales=(0, 1, 2, 3),
                disp_kind="realistic", image_kind="smooth", seed=0, depth_hints=False) -> PhotoBatch:
    frame_ids = list(frame_ids)
    scales = list(scales)
    color = {}
    for i, f in enumerate(frame_ids):
        shape = (batch, 3, height, width)
        if image_kind == "smooth":
            color[(f, 0)] = smooth_field(shape, seed + 200 + i)
        else:
            color[(f, 0)] = rand(shape, seed + 100 + i)
    # pyramid of the target frame (used by the smoothness term, M2/trainer.py:599,664)
    for s in scales:
        if s == 0:
            continue
        color[(0, s)] = F.interpolate(color[(0, 0)], size=(height // 2 ** s, width // 2 ** s),
                                      mode="bilinear", align_corners=False).contiguous()
    disp = {}
    for s in scales:
        shape = (batch, 1, height // 2 ** s, width // 2 ** s)
        if disp_kind == "realistic":
            # smooth field + fine texture: a bare bilinear up-sample is piecewise linear with exactly
            # constant borders, which puts |d(x)-d(x+1)| at 0 +- 1 ulp where abs() is not differentiable
            disp[s] = (0.02 + 0.3 * smooth_field(shape, seed + 300 + s, down=8, noise=0.0)
                       + 0.004 * rand(shape, seed + 350 + s)).contiguous()
        else:
            disp[s] = rand(shape, seed + 400 + s)
    K, inv_K = intrinsics(height, width, batch)
    T = {}
    for f in frame_ids[1:]:
        if f == "s":
            T[f] = stereo_T(batch)
        else:
            T[f] = temporal_T(batch, seed + 500 + 7 * int(f))
    nf = len(frame_ids) - 1
    noise = {s: (randn((batch, nf, height, width), seed + 600 + s) * 0.00001).contiguous() for s in scales}
    pb = PhotoBatch(batch, height, width, frame_ids, scales, color, disp, K, inv_K, T, noise)
    if depth_hints:
        # depth_hint = depth(disp_0) * (1 + 0.1 N(0,1)), valid mask ~ Bernoulli(0.8)
        # (DH/datasets/mono_dataset.py:367-388 loads both as (1,H,W) float maps)
        d0 = F.interpolate(disp[scales[0]], size=(height, width), mode="bilinear", align_corners=False)
        depth = 1.0 / (1.0 / pb.max_depth + (1.0 / pb.min_depth - 1.0 / pb.max_depth) * d0)
        hint = (depth * (1.0 + 0.1 * randn(depth.shape, seed + 950))).clamp_(min=1e-3)
        pb.extras["depth_hint"] = hint.contiguous()
        pb.extras["depth_hint_mask"] = (rand(depth.shape, seed + 951) < 0.8).float().contiguous()
    return pb


def ellipse_mask(h=PATCH_H, w=PATCH_W) -> torch.Tensor:
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    cy, cx = (h - 1) / 2.0, (w - 1) / 2.0
    inside = ((yy - cy) / (0.46 * h)) ** 2 + ((xx - cx) / (0.48 * w)) ** 2 <= 1.0
    return inside.float().view(1, 1, h, w).contiguous()


@dataclass
class PatchBatch:
    """One stage-1 batch: what `Phy_obj_atk*.forward` reads each PGD iteration."""
    batch: int
    obj: torch.Tensor          # (1,3,260,300)
    mask: torch.Tensor         # (1,1,260,300)
    scenes: torch.Tensor       # (Ba,3,375,1242)
    z0: List[float]
    alpha: List[float]
    upstream: torch.Tensor     # (Ba,3,320,1024) d(cost)/d(adv_scene)
    pattern_pos: torch.Tensor  # (1,3,260,300) L0 init
    pattern_neg: torch.Tensor

    def to(self, device, non_blocking=False):
        mv = lambda t: t.to(device, non_blocking=non_blocking)
        return PatchBatch(self.batch, mv(self.obj), mv(self.mask), mv(self.scenes), list(self.z0), list(self.alpha),
                          mv(self.upstream), mv(self.pattern_pos), mv(self.pattern_neg))


def patch_batch(batch=4, seed=0, scene_kind="smooth") -> PatchBatch:
    obj = rand((1, 3, PATCH_H, PATCH_W), seed + 700)
    mask = ellipse_mask()
    if scene_kind == "smooth":
        scenes = smooth_field((batch, 3, ORI_H, ORI_W), seed + 800)
    else:
        scenes = rand((batch, 3, ORI_H, ORI_W), seed + 800)
    z0 = [5.0 + 0.2 * (i % 25) for i in range(batch)]
    alpha = [-30.0 + 5.0 * (i % 13) for i in range(batch)]
    upstream = randn((batch, 3, SCENE_H, SCENE_W), seed + 900) * 1e-3
    ppos = rand((1, 3, PATCH_H, PATCH_W), seed + 901)
    pneg = rand((1, 3, PATCH_H, PATCH_W), seed + 902)
    return PatchBatch(batch, obj, mask, scenes, z0, alpha, upstream, ppos, pneg)


def frames_u8(seed: int, batch: int = 1) -> torch.Tensor:
    """Seeded native-resolution 8-bit frames (batch,3,375,1242): smooth field + noise, quantised -- the raw images
    a KITTI loader decodes (inputs of the training-batch compositing, loader.py)."""
    f = smooth_field((batch, 3, ORI_H, ORI_W), seed)
    return (f * 255.0).round().clamp(0, 255).to(torch.uint8)


def cost_volume_inputs(B=2, L=2, C=16, h=24, w=40, D=12, seed=80):
    """Seeded matching-resolution inputs: smooth non-negative features (post-ReLU), small temporal poses,
    item 1's second lookup frame missing (all-zero pose), KITTI intrinsics at (h, w), linear depth bins."""
    cur = smooth_field((B, C, h, w), seed, down=4, noise=0.02)
    look = torch.stack([smooth_field((B, C, h, w), seed + 1 + l, down=4, noise=0.02) for l in range(L)], 1)
    poses = torch.stack([temporal_T(B, seed + 10 + 7 * l) for l in range(L)], 1).contiguous()
    poses[:, :, 0, 3] += 0.3                      # a visible baseline so that the bins sweep along the epipolar line
    if B > 1 and L > 1:
        poses[1, 1] = 0.0
    K, inv_K = intrinsics(h, w, B)
    bins = torch.linspace(0.5, 20.0, D)
    return cur.contiguous(), look.contiguous(), poses, K, inv_K, bins

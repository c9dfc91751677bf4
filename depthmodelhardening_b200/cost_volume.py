"""ManyDepth cost volume: host-side mirror of
`ResnetEncoderMatching.match_features`
(`DepthNetworks/manydepth2/networks/resnet_encoder.py:157-236`).

`match_features(self, current_feats, lookup_feats, relative_poses, K, invK)` keeps the
reference's signature and return values (batch_cost_volume (B,D,h,w), cost_volume_masks
(B,D,h,w)) so that `install()` can bind it onto the reference class; the D x L x ~25
ATen launches per batch item become one kernel (`dmh_cost_volume`).  CUDA tensors only.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, f32c, ptr, stream


def cost_volume(current_feats, lookup_feats, relative_poses, K, invK, depth_bins, set_missing_to_max=True):
    """current_feats (B,C,h,w), lookup_feats (B,L,C,h,w), relative_poses (B,L,4,4), K / invK (B,4,4) at the
    matching resolution, depth_bins (D,) -> (cost_volume (B,D,h,w), missing_mask (B,D,h,w)).  No gradient."""
    with torch.no_grad():
        cur, look, pose = f32c(current_feats), f32c(lookup_feats), f32c(relative_poses)
        B, C, h, w = cur.shape
        from .ops import _mat_batch
        k, ik = f32c(_mat_batch(K, B, "K")), f32c(_mat_batch(invK, B, "invK"))   # the kernel indexes K + b*16
        bins = f32c(depth_bins.to(cur.device)).reshape(-1)
        if pose.dim() != 4 or pose.shape[0] != B or tuple(pose.shape[2:]) != (4, 4):
            raise RuntimeError("relative_poses must be (B, L, 4, 4), got %s" % (tuple(pose.shape),))
        if look.dim() != 5 or look.shape[0] != B or look.shape[2:] != (C, h, w):
            raise RuntimeError("lookup_feats must be (B, L, C, h, w) matching current_feats %s, got %s" %
                               (tuple(cur.shape), tuple(look.shape)))
        L, D = look.shape[1], bins.numel()
        lib = _lib.load()
        ws = torch.empty(lib.dmh_cost_volume_workspace_floats(B, L, C, h, w), device=cur.device, dtype=torch.float32)
        cost = torch.empty(B, D, h, w, device=cur.device, dtype=torch.float32)
        miss = torch.empty(B, D, h, w, device=cur.device, dtype=torch.float32)
        check(lib.dmh_cost_volume(ptr(cur), ptr(look), ptr(pose), ptr(k), ptr(ik), ptr(bins), B, L, C, D, h, w,
                                  int(bool(set_missing_to_max)), ptr(ws), ptr(cost), ptr(miss), stream()),
              "cost_volume")
    return cost, miss


def match_features(self, current_feats, lookup_feats, relative_poses, K, invK):
    """Drop-in for `ResnetEncoderMatching.match_features`: reads `self.depth_bins` (kept current by
    `compute_depth_bins`, also with adaptive bins), `self.set_missing_to_max`."""
    return cost_volume(current_feats, lookup_feats, relative_poses, K, invK, self.depth_bins,
                       getattr(self, "set_missing_to_max", True))

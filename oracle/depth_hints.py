"""ORACLE (test infrastructure only): CPU restatement of the depth-hints objective
 in plain torch ops.  Paths below are relative to
/root/reference/DepthNetworks/depth-hints.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline legs may
import this module.  Pinned by `oracle/make_golden_dh.py`, which imports the
UNMODIFIED depth-hints `trainer.py` and runs its `generate_images_pred` +
`compute_losses` on seeded synthetic inputs (`tests/golden/dh_*.npz`).  The reference
ships no tests / golden vectors for this path.  dtype-generic (fp64 arbiter).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import photometric as P


def hint_warp(src, depth_hint, K, inv_K, T):
    """trainer.py:510-525 -- warp of the stereo source with the depth hint; NOTE the
    reference calls F.grid_sample(..., padding_mode="border") with the DEFAULT
    align_corners (False) here, unlike the main warp (:500-504, align_corners=True)."""
    b, _, h, w = depth_hint.shape
    pts = P.backproject(depth_hint, inv_K)
    grid = P.project3d(pts, K, T, h, w)
    return F.grid_sample(src, grid, padding_mode="border", align_corners=False)


def proxy_supervised_loss(pred, target, valid_pixels, loss_mask):
    """trainer.py:525-539 -- log(|target - pred| + 1) on valid hint pixels where the hint wins."""
    return torch.log(torch.abs(target - pred) + 1) * valid_pixels * loss_mask


def loss_masks(reprojection_loss, identity_reprojection_loss, depth_hint_reprojection_loss):
    """trainer.py:541-590 -- argmin over [reprojection, identity, depth-hint reprojection]."""
    if identity_reprojection_loss is None:
        reprojection_loss_mask = torch.ones_like(reprojection_loss)
        if depth_hint_reprojection_loss:          # (sic) raises for a multi-element tensor, as the reference does
            all_losses = torch.cat([reprojection_loss, depth_hint_reprojection_loss], dim=1)
            idxs = torch.argmin(all_losses, dim=1, keepdim=True)
            depth_hint_loss_mask = (idxs == 1).to(reprojection_loss.dtype)
        idxs = torch.zeros_like(reprojection_loss, dtype=torch.long)
    else:
        if depth_hint_reprojection_loss is not None:
            all_losses = torch.cat([reprojection_loss, identity_reprojection_loss, depth_hint_reprojection_loss], dim=1)
        else:
            all_losses = torch.cat([reprojection_loss, identity_reprojection_loss], dim=1)
        idxs = torch.argmin(all_losses, dim=1, keepdim=True)
        reprojection_loss_mask = (idxs != 1).to(reprojection_loss.dtype)
        depth_hint_loss_mask = (idxs == 2).to(reprojection_loss.dtype)
    if depth_hint_reprojection_loss is None:
        depth_hint_loss_mask = None
    return reprojection_loss_mask, depth_hint_loss_mask, idxs


def depth_hints_objective(colors: Dict, disps: Dict, K, inv_K, Ts: Dict, frame_ids: List, noise: Optional[Dict],
                          depth_hint, depth_hint_mask, use_depth_hints=True, opts=None, return_aux=False):
    """generate_images_pred (trainer.py:476-525) + the per-scale loop of compute_losses
    (:629-727) without the adv / supervised / contrastive / predictive-mask terms.
    noise: {scale: (B,1,H,W)} tie-break noise already * 1e-5 (:687-690, injected)."""
    opts = opts or P.default_opts()
    srcs = frame_ids[1:]
    target = colors[(0, 0)]
    H, W = target.shape[2], target.shape[3]
    aux, losses = {}, {}
    if use_depth_hints:
        pred_h = hint_warp(colors[("s", 0)], depth_hint, K, inv_K, Ts["s"])
        aux[("color_depth_hint", "s", 0)] = pred_h
        hint_rl = P.reprojection_loss(pred_h, target, opts.no_ssim) + 1000 * (1 - depth_hint_mask)   # :629-634
    else:
        hint_rl = None
    total = 0
    for scale in opts.scales:
        disp = disps[scale]
        disp_full = F.interpolate(disp, [H, W], mode="bilinear", align_corners=False)
        reproj = []
        for f in srcs:
            pred, _, depth = P.warp_from_disp(disp_full, colors[(f, 0)], K, inv_K, Ts[f], opts.min_depth,
                                              opts.max_depth)
            reproj.append(P.reprojection_loss(pred, target, opts.no_ssim))
        reproj = torch.cat(reproj, 1)
        if not opts.disable_automasking:
            ident = torch.cat([P.reprojection_loss(colors[(f, 0)], target, opts.no_ssim) for f in srcs], 1)
            if opts.avg_reprojection:
                ident = ident.mean(1, keepdim=True)
            else:
                ident, _ = torch.min(ident, dim=1, keepdim=True)              # :670-672
        else:
            ident = None
        if opts.avg_reprojection:
            reproj = reproj.mean(1, keepdim=True)
        else:
            reproj, _ = torch.min(reproj, dim=1, keepdim=True)                # :683-685
        if ident is not None:
            ident = ident + noise[scale].to(ident.dtype)                      # :687-690
        m_r, m_h, idxs = loss_masks(reproj, ident, hint_rl)
        loss_r = (reproj * m_r).sum() / (m_r.sum() + 1e-7)                    # :699-700
        losses["reproj_loss/{}".format(scale)] = loss_r
        loss = loss_r
        aux[("argmin", scale)] = idxs
        if use_depth_hints:
            hl = proxy_supervised_loss(depth, depth_hint, depth_hint_mask, m_h)
            loss_h = hl.sum() / (m_h.sum() + 1e-7)                            # :712-713
            losses["depth_hint_loss/{}".format(scale)] = loss_h
            loss = loss + loss_h
        sm = P.normalised_smooth_loss(disp, colors[(0, scale)])
        loss = loss + opts.disparity_smoothness * sm / (2 ** scale)
        losses["loss/{}".format(scale)] = loss
        total = total + loss
    total = total / len(opts.scales)
    losses["loss"] = total
    if return_aux:
        return total, losses, aux
    return total, losses


def objective_from_batch(pb, use_depth_hints=True, opts=None, dtype=torch.float32, return_aux=False):
    """Run on a `synth.PhotoBatch` built with depth_hints=True; returns (loss, losses, grads[, aux])."""
    cast = lambda t: t.detach().to(dtype)
    colors = {k: cast(v) for k, v in pb.color.items()}
    disps = {s: cast(v).requires_grad_(True) for s, v in pb.disp.items()}
    Ts = {k: cast(v) for k, v in pb.T.items()}
    noise = {k: cast(v[:, :1]) for k, v in pb.noise.items()}
    opts = opts or P.default_opts(scales=list(pb.scales), min_depth=pb.min_depth, max_depth=pb.max_depth)
    out = depth_hints_objective(colors, disps, cast(pb.K), cast(pb.inv_K), Ts, pb.frame_ids, noise,
                                cast(pb.extras["depth_hint"]), cast(pb.extras["depth_hint_mask"]), use_depth_hints,
                                opts, return_aux=return_aux)
    out[0].backward()
    grads = {s: d.grad for s, d in disps.items()}
    return (out[0].detach(), out[1], grads) + ((out[2],) if return_aux else ())

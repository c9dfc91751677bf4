"""ORACLE (test infrastructure only): CPU restatement of the reference's stage-1
hot path -- physical patch placement / compositing and the PGD / L0 updates.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.  The product never does.

Third-party arithmetic on this path lives in torchvision (absent from
/root/reference; pinned there as torchvision==0.8.2, requirements.txt:93; the
oracle of record is the INSTALLED torchvision 0.26.0):
  * transforms.functional.perspective  -> `perspective_coeffs`, `perspective_grid`
    restate functional.py:674-704 and _functional_tensor.py:672-698, 545-561
  * transforms.Resize (bilinear, antialias=True) -> F.interpolate(antialias=True)
  * transforms.Pad -> F.pad
Pinned by goldens generated from the reference itself (`oracle/make_golden.py`).
Paths below are relative to /root/reference.
"""
from __future__ import annotations

from math import cos, radians, sin
from typing import List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

ORI_H, ORI_W = 375, 1242      # my_utils.py:12-13
VEH_H, VEH_W, CAM_H = 1.6, 1.82, 1.65   # physicalTrans.py:40-42


# ----------------------------------------------------------------------------- A1
def plane_corners(z0: float, alpha: float) -> np.ndarray:
    """physicalTrans.py:83-105 -- 4 world corners (tl,tr,br,bl) of the vertical
    plane at distance z0 and yaw alpha (degrees); camera frame x right, y down."""
    half_w = VEH_W / 2
    dx = cos(radians(alpha)) * half_w
    dz = sin(radians(alpha)) * half_w
    y_c = CAM_H - VEH_H / 2
    top, bot = y_c - VEH_H / 2, y_c + VEH_H / 2
    return np.array([[-dx, top, z0 - dz], [dx, top, z0 + dz], [dx, bot, z0 + dz], [-dx, bot, z0 - dz]])


def corners_on_image(z0, alpha, P34: np.ndarray, K=None, T=None) -> np.ndarray:
    """physicalTrans.py:62-81 (`objPosOnImage`) and :175-189 (`project_w_trans`):
    project the 4 corners and TRUNCATE to int32.  With K given the monodepth2
    intrinsics path (+1e-7 in the divide) is used, otherwise the KITTI P2 path
    (preprocessing/kitti_util.py:144-152)."""
    world = plane_corners(z0, alpha)
    hom = np.concatenate((world.T, np.ones((1, 4))), axis=0)
    if K is not None:
        P = K[:3, :] if T is None else np.matmul(K, T)[:3, :]
        cam = np.matmul(P, hom)
        return (cam[:2, :] / (cam[[2], :] + 1e-7)).T.astype(np.int32)
    if T is not None:
        world = np.matmul(T, hom).T[:, :3]
        hom = np.concatenate((world.T, np.ones((1, 4))), axis=0)
    pts = np.dot(hom.T, P34.T)
    pts[:, 0] /= pts[:, 2]
    pts[:, 1] /= pts[:, 2]
    return pts[:, :2].astype(np.int32)


# ----------------------------------------------------------------------------- A2
def pad_to_canvas(img):
    """physicalTrans.py:107-122 -- centre zero-pad to 375x1242; returns padded
    tensor and the start corners [tl,tr,br,bl] as (u,v)."""
    _, _, h, w = img.shape
    l = (ORI_W - w) // 2
    r = ORI_W - w - l
    t = (ORI_H - h) // 2
    b = ORI_H - h - t
    corners = [[l, t], [l + w, t], [l + w, t + h], [l, t + h]]
    return F.pad(img, (l, r, t, b)), corners


# ----------------------------------------------------------------------------- A3
def perspective_coeffs(start: Sequence, end: Sequence) -> List[float]:
    """torchvision functional.py:674-704 -- 8 homography coefficients mapping
    OUTPUT pixel -> INPUT pixel; fp64 least squares, result cast to fp32."""
    A = torch.zeros(8, 8, dtype=torch.float64)
    for i, (pe, ps) in enumerate(zip(end, start)):
        pe = [float(pe[0]), float(pe[1])]
        ps = [float(ps[0]), float(ps[1])]
        A[2 * i, :] = torch.tensor([pe[0], pe[1], 1, 0, 0, 0, -ps[0] * pe[0], -ps[0] * pe[1]], dtype=torch.float64)
        A[2 * i + 1, :] = torch.tensor([0, 0, 0, pe[0], pe[1], 1, -ps[1] * pe[0], -ps[1] * pe[1]], dtype=torch.float64)
    rhs = torch.tensor([[float(p[0]), float(p[1])] for p in start], dtype=torch.float64).view(8)
    sol = torch.linalg.lstsq(A, rhs, driver="gels").solution.to(torch.float32)
    return sol.tolist()


def perspective_grid(coeffs, oh, ow, dtype=torch.float32, device=None):
    """torchvision _functional_tensor.py:672-698 -- sampling grid in [-1,1]
    built from pixel centres (x+0.5, y+0.5)."""
    c = coeffs
    th1 = torch.tensor([[[c[0], c[1], c[2]], [c[3], c[4], c[5]]]], dtype=dtype, device=device)
    th2 = torch.tensor([[[c[6], c[7], 1.0], [c[6], c[7], 1.0]]], dtype=dtype, device=device)
    base = torch.empty(1, oh, ow, 3, dtype=dtype, device=device)
    base[..., 0].copy_(torch.linspace(0.5, ow + 0.5 - 1.0, steps=ow, device=device))
    base[..., 1].copy_(torch.linspace(0.5, oh + 0.5 - 1.0, steps=oh, device=device).unsqueeze(-1))
    base[..., 2].fill_(1)
    r1 = th1.transpose(1, 2) / torch.tensor([0.5 * ow, 0.5 * oh], dtype=dtype, device=device)
    g1 = base.view(1, oh * ow, 3).bmm(r1)
    g2 = base.view(1, oh * ow, 3).bmm(th2.transpose(1, 2))
    return (g1 / g2 - 1.0).view(1, oh, ow, 2)


def perspective_warp(img, coeffs):
    """torchvision _functional_tensor.py:545-561 -- bilinear, zeros padding,
    align_corners=False."""
    grid = perspective_grid(coeffs, img.shape[-2], img.shape[-1], img.dtype, img.device)
    return F.grid_sample(img, grid.expand(img.shape[0], -1, -1, -1), mode="bilinear", padding_mode="zeros",
                         align_corners=False)


def project_patch(obj, mask, z0s, alphas, P34, K=None, T=None):
    """physicalTrans.py:130-166 / :168-196 -- per item: corners -> homography ->
    warp padded patch and padded mask; concatenated over the batch."""
    obj_pad, start = pad_to_canvas(obj)
    mask_pad, _ = pad_to_canvas(mask)
    imgs, masks, coeffs_all = [], [], []
    for z0, a in zip(z0s, alphas):
        end = corners_on_image(z0, a, P34, K, T).tolist()
        co = perspective_coeffs(start, end)
        coeffs_all.append(co)
        imgs.append(perspective_warp(obj_pad, co))
        masks.append(perspective_warp(mask_pad, co))
    return torch.cat(imgs, 0), torch.cat(masks, 0), coeffs_all


# ----------------------------------------------------------------------------- A4 / A5
def resize_aa(x, size=(320, 1024)):
    """torchvision Resize on a tensor (phy_obj_atk.py:51,89-90): bilinear,
    antialias=True (default in the installed torchvision), align_corners=False."""
    return F.interpolate(x, size=list(size), mode="bilinear", align_corners=False, antialias=True)


def apply_patch(obj, mask, scenes, z0s, alphas, P34, K=None, T=None, size=(320, 1024)):
    """phy_obj_atk.py:86-90 -- project, composite scene*(1-m)+obj*m, resize both."""
    o, m, _ = project_patch(obj, mask, z0s, alphas, P34, K, T)
    adv = scenes * (1 - m) + o * m
    return resize_aa(adv, size), resize_aa(m, size)


# ----------------------------------------------------------------------------- A6
def pgd_linf_step(adv, grad, clean, alpha, eps):
    """phy_obj_atk.py:98-100 -- ascent on sign(grad), project to the eps-ball
    around the clean patch, clamp to [0,1]."""
    adv = adv + alpha * grad.sign()
    delta = torch.clamp(adv - clean, min=-eps, max=eps)
    return torch.clamp(clean + delta, min=0, max=1)


def pgd_l2_step(adv, grad, clean, alpha, eps, eps_div=1e-10):
    """torchattacks/attacks/phy_obj_atk_l2.py:108-120 for the one shared patch (the reference's batch_size is 1
    there: it views the (1,3,h,w) gradient as (batch_size, -1))."""
    n = grad.shape[0]
    grad_norms = torch.norm(grad.view(n, -1), p=2, dim=1) + eps_div
    grad = grad / grad_norms.view(n, 1, 1, 1)
    adv = adv + alpha * grad
    delta = adv - clean
    delta_norms = torch.norm(delta.view(n, -1), p=2, dim=1)
    factor = eps / delta_norms
    factor = torch.min(factor, torch.ones_like(delta_norms))
    delta = delta * factor.view(-1, 1, 1, 1)
    return torch.clamp(clean + delta, min=0, max=1)


# ----------------------------------------------------------------------------- A7
def l0_compose(obj, p_pos, p_neg, clip_max=1.0):
    """phy_obj_atk_l0.py:94-99 -- adv patch from the positive/negative patterns."""
    pos = torch.clamp(p_pos * clip_max, min=0.0, max=clip_max)
    neg = -torch.clamp(p_neg * clip_max, min=0.0, max=clip_max)
    return torch.clamp(obj + (pos + neg), min=0.0, max=clip_max), pos, neg


def l0_count(pos, neg, thr=1.0 / 255.0):
    """phy_obj_atk_l0.py:43-52 -- pixels whose thresholded pattern is non-zero in
    any channel (channel-summed |.| != 0).  Returns (count int64, survivor mask)."""
    p = pos.detach().clone()
    n = neg.detach().clone()
    p[p < thr] = 0
    n[n > -thr] = 0
    s = torch.sum(torch.abs(p + n), dim=1)
    return torch.count_nonzero(s), (s != 0)


# ----------------------------------------------------------------------------- A8
def l0_mask_cost(p_pos, p_neg):
    """phy_obj_atk_l0.py:130-132 -- mean over pixels of max_c(tanh(p/10)/(2-1e-7)+0.5)."""
    mp = torch.max(torch.tanh(p_pos / 10) / (2 - 1e-7) + 0.5, dim=1)[0]
    mn = torch.max(torch.tanh(p_neg / 10) / (2 - 1e-7) + 0.5, dim=1)[0]
    return torch.mean(mp) + torch.mean(mn)


def adam_step(p, g, m, v, step, lr=0.5, b1=0.5, b2=0.9, eps=1e-8):
    """torch.optim.Adam single-tensor update as used at phy_obj_atk_l0.py:86,138
    (betas=(0.5,0.9), no weight decay, no amsgrad).  `step` is 1-based.
    Returns (p, m, v) new tensors."""
    m = torch.lerp(m, g, 1 - b1)
    v = v * b2 + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / (bc2 ** 0.5) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def l0_finalize(obj, p_pos, p_neg, clip_max=1.0, thr=1.0 / 255.0):
    """phy_obj_atk_l0.py:143-150 -- hard threshold then compose."""
    pos = torch.clamp(p_pos * clip_max, min=0.0, max=clip_max)
    neg = -torch.clamp(p_neg * clip_max, min=0.0, max=clip_max)
    pos = torch.where(pos < thr, torch.zeros_like(pos), pos)
    neg = torch.where(neg > -thr, torch.zeros_like(neg), neg)
    return torch.clamp(obj + (pos + neg), min=0.0, max=clip_max), pos, neg


def topk_l0_project(p_pos, p_neg, k):
    """EXTENSION (not reference behaviour; SURVEY.md headline fact 3): keep the k
    pixels with the largest channel-max pattern magnitude, zero the rest.
    Ties are broken towards the lower flat pixel index.  Oracle: sort."""
    mag = torch.maximum(p_pos.clamp(0, 1), p_neg.clamp(0, 1)).amax(dim=1).reshape(-1)
    n = mag.numel()
    k = max(0, min(int(k), n))
    # stable descending order by (value, -index): sort index ascending among equals
    order = torch.argsort(-mag, stable=True)
    keep = torch.zeros(n, dtype=torch.bool)
    keep[order[:k]] = True
    keep = keep.view(1, 1, *p_pos.shape[2:])
    return p_pos * keep, p_neg * keep, keep


# ----------------------------------------------------------------------------- whole attack loops
def linf_attack(model, obj, mask, scenes, placements, final_placement, P34, eps, alpha, step=None):
    """phy_obj_atk.py:73-123 with the random placements injected:
    placements = [(z0s, alphas)] per step.  Device agnostic (runs where the tensors live)."""
    import torch.nn as nn
    loss = nn.MSELoss()
    adv = obj.clone().detach()
    for z0s, als in placements:
        adv.requires_grad_()
        scene, m = apply_patch(adv, mask, scenes, z0s, als, P34)
        depth = model(scene)
        cost = -loss(depth * m, torch.zeros_like(depth))
        grad = torch.autograd.grad(cost, adv)[0]
        with torch.no_grad():
            adv = (step or pgd_linf_step)(adv, grad, obj, alpha, eps)
    z0s, als = final_placement
    with torch.no_grad():
        adv_s, m_out = apply_patch(adv, mask, scenes, z0s, als, P34)
        ben_s, _ = apply_patch(obj, mask, scenes, z0s, als, P34)
    return adv_s, ben_s, m_out, adv


def l2_attack(model, obj, mask, scenes, placements, final_placement, P34, eps, steps):
    """phy_obj_atk_l2.py:73-136 (random_start off) with the random placements injected: the L-inf loop with the L2
    update and the step 2.5 * eps / steps (:44)."""
    return linf_attack(model, obj, mask, scenes, placements, final_placement, P34, eps, 2.5 * eps / steps,
                       step=pgd_l2_step)


def l0_attack(model, obj, mask, scenes, init_pos, init_neg, placements, final_placement, P34, steps, lr, mask_wt,
              l0_thresh):
    """phy_obj_atk_l0.py:73-174 with the random inits / placements injected."""
    import torch.nn as nn
    loss = nn.MSELoss()
    pp = init_pos.clone().requires_grad_(True)
    pn = init_neg.clone().requires_grad_(True)
    opt = torch.optim.Adam([pp, pn], lr=lr, betas=(0.5, 0.9))
    l0_init = None
    it = iter(placements)
    for stp in range(steps * 2):
        adv, pos, neg = l0_compose(obj, pp, pn)
        l0, _ = l0_count(pos, neg)
        if stp == 0:
            l0_init = l0
        if (l0 / l0_init) <= l0_thresh:
            w = 0
            if stp >= steps:
                break
        else:
            w = mask_wt
        z0s, als = next(it)
        scene, m = apply_patch(adv, mask, scenes, z0s, als, P34)
        depth = model(scene)
        cost = loss(depth * m, torch.zeros_like(depth)) + w * l0_mask_cost(pp, pn)
        opt.zero_grad()
        cost.backward()
        opt.step()
    adv, pos, neg = l0_finalize(obj, pp.detach(), pn.detach())
    z0s, als = final_placement
    with torch.no_grad():
        adv_s, m_out = apply_patch(adv, mask, scenes, z0s, als, P34)
    return adv_s, m_out, adv, pp.detach(), pn.detach(), pos + neg

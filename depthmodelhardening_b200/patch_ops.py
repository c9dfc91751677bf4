"""Stage-1 host side: patch placement geometry (host, fp64), homography solve
(one batched fp64 least-squares instead of 2*Ba sequential ones), and the
autograd wrappers of the patch kernels.

Reference lines mirrored (paths under /root/reference):
  physicalTrans.py:62-105   fromZA2Coord / objPosOnImage (corner projection, int32 truncation)
  physicalTrans.py:107-122  padding_img (centre pad -> start corners)
  torchvision functional.py:674-704  _get_perspective_coeffs (fp64 gels, cast to fp32)
"""
from __future__ import annotations

from math import cos, radians, sin
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, f32c, ptr, stream

ORI_H, ORI_W = 375, 1242            # my_utils.py:12-13
VEH_H, VEH_W, CAM_H = 1.6, 1.82, 1.65   # physicalTrans.py:40-42 (BMW)
# KITTI object calib 003086 P2 as printed in physicalTrans.py:208-213 (used for synthetic runs)
KITTI_P2_003086 = (7.215377e+02, 0.0, 6.095593e+02, 4.485728e+01,
                   0.0, 7.215377e+02, 1.728540e+02, 2.163791e-01,
                   0.0, 0.0, 1.0, 2.745884e-03)


def _lib_():
    return _lib.load()


# ----------------------------------------------------------------------------- geometry (host)
def plane_corners(z0: float, alpha: float) -> np.ndarray:
    """World corners tl,tr,br,bl of the vehicle plane (physicalTrans.py:83-105)."""
    off_x = cos(radians(alpha)) * VEH_W / 2
    off_z = sin(radians(alpha)) * VEH_W / 2
    yc = CAM_H - VEH_H / 2
    y1, y2 = yc - VEH_H / 2, yc + VEH_H / 2
    return np.array([[-off_x, y1, z0 - off_z], [off_x, y1, z0 + off_z], [off_x, y2, z0 + off_z],
                     [-off_x, y2, z0 - off_z]], dtype=np.float64)


def project_corners(z0: float, alpha: float, P34: np.ndarray, K: Optional[np.ndarray] = None,
                    T: Optional[np.ndarray] = None) -> np.ndarray:
    """(4,2) int32 image corners: objPosOnImage (physicalTrans.py:62-81) and the
    project_w_trans variant (:175-189)."""
    pts = plane_corners(z0, alpha)
    hom = np.concatenate((pts.T, np.ones((1, 4))), axis=0)
    if K is not None:
        P = K[:3, :] if T is None else np.matmul(K, T)[:3, :]
        cam = np.matmul(P, hom)
        return (cam[:2, :] / (cam[[2], :] + 1e-7)).T.astype(np.int32)
    if T is not None:
        pts = np.matmul(T, hom).T[:, :3]
        hom = np.concatenate((pts.T, np.ones((1, 4))), axis=0)
    uvw = np.dot(hom.T, np.transpose(P34))
    uvw[:, 0] /= uvw[:, 2]
    uvw[:, 1] /= uvw[:, 2]
    return uvw[:, 0:2].astype(np.int32)


def start_corners(obj_hw, canvas_hw=(ORI_H, ORI_W)):
    """Corners of the centred, zero-padded patch on the canvas (physicalTrans.py:107-122)."""
    h, w = obj_hw
    H, W = canvas_hw
    l, t = (W - w) // 2, (H - h) // 2
    return [[l, t], [l + w, t], [l + w, t + h], [l, t + h]]


def solve_homographies(start: Sequence, ends: np.ndarray) -> torch.Tensor:
    """(Ba,8) fp32 coefficients mapping output pixel -> input pixel; one batched
    fp64 `gels` solve (torchvision solves them one by one, on the host, per call)."""
    ends = np.asarray(ends, dtype=np.float64).reshape(-1, 4, 2)
    n = ends.shape[0]
    A = torch.zeros(n, 8, 8, dtype=torch.float64)
    st = torch.tensor(start, dtype=torch.float64)          # (4,2)
    en = torch.from_numpy(ends)                             # (n,4,2)
    for i in range(4):
        ex, ey = en[:, i, 0], en[:, i, 1]
        sx, sy = st[i, 0], st[i, 1]
        A[:, 2 * i, 0], A[:, 2 * i, 1], A[:, 2 * i, 2] = ex, ey, 1.0
        A[:, 2 * i, 6], A[:, 2 * i, 7] = -sx * ex, -sx * ey
        A[:, 2 * i + 1, 3], A[:, 2 * i + 1, 4], A[:, 2 * i + 1, 5] = ex, ey, 1.0
        A[:, 2 * i + 1, 6], A[:, 2 * i + 1, 7] = -sy * ex, -sy * ey
    rhs = st.reshape(8).unsqueeze(0).expand(n, 8).unsqueeze(-1)
    sol = torch.linalg.lstsq(A, rhs, driver="gels").solution.squeeze(-1)
    return sol.to(torch.float32).contiguous()


class Placement:
    """Per-item placement of the patch on the canvas: the (Ba,8) perspective
    coefficients plus a conservative (Ba,4) int32 bounding box {x0,y0,x1,y1}
    (inclusive canvas pixels) outside of which item b cannot sample the patch.
    The box is only an optimisation hint for the fused apply kernels (they skip
    the perspective maths / launch smaller grids); results do not depend on it."""

    def __init__(self, coeffs: torch.Tensor, bbox: Optional[torch.Tensor] = None, bbox_wh=(0, 0)):
        self.coeffs, self.bbox, self.bbox_wh = coeffs, bbox, (int(bbox_wh[0]), int(bbox_wh[1]))

    @property
    def shape(self):
        return self.coeffs.shape

    @property
    def device(self):
        return self.coeffs.device

    def to(self, device):
        return Placement(self.coeffs.to(device), None if self.bbox is None else self.bbox.to(device), self.bbox_wh)


def placement_bbox(coeffs: torch.Tensor, ends: np.ndarray, obj_hw, canvas_hw=(ORI_H, ORI_W)):
    """Bounding box of the canvas pixels that can sample the patch.  The patch
    rectangle is convex and the homography's denominator g*x+h*y+1 is positive at
    the four projected corners, hence on the whole quadrilateral: the pre-image
    of the rectangle is exactly the convex hull of those corners.  A margin covers
    the half-pixel reach of the bilinear taps (magnified when the patch is
    enlarged) and the fp32 rounding of the coefficients.  Items whose denominator
    changes sign get the full canvas."""
    H, W = canvas_hw
    ends = np.asarray(ends, dtype=np.float64).reshape(-1, 4, 2)
    co = coeffs.detach().cpu().double().numpy()
    den = co[:, [6]] * ends[:, :, 0] + co[:, [7]] * ends[:, :, 1] + 1.0        # (n,4)
    ok = (den > 1e-3).all(axis=1)
    ext_w = ends[:, :, 0].max(1) - ends[:, :, 0].min(1)
    ext_h = ends[:, :, 1].max(1) - ends[:, :, 1].min(1)
    margin = 3.0 + np.ceil(np.maximum(ext_w / obj_hw[1], ext_h / obj_hw[0]))
    x0 = np.clip(np.floor(ends[:, :, 0].min(1) - margin), 0, W - 1)
    x1 = np.clip(np.ceil(ends[:, :, 0].max(1) + margin), 0, W - 1)
    y0 = np.clip(np.floor(ends[:, :, 1].min(1) - margin), 0, H - 1)
    y1 = np.clip(np.ceil(ends[:, :, 1].max(1) + margin), 0, H - 1)
    box = np.stack([x0, y0, x1, y1], axis=1)
    box[~ok] = (0, 0, W - 1, H - 1)
    box = box.astype(np.int32)
    wh = (int((box[:, 2] - box[:, 0] + 1).max()), int((box[:, 3] - box[:, 1] + 1).max()))
    return torch.from_numpy(box).contiguous(), wh


def make_placement(start, ends, obj_hw, canvas_hw=(ORI_H, ORI_W)) -> Placement:
    co = solve_homographies(start, ends)
    box, wh = placement_bbox(co, ends, obj_hw, canvas_hw)
    return Placement(co, box, wh)


def homographies(z0s, alphas, P34, K=None, T=None, obj_hw=(260, 300), canvas_hw=(ORI_H, ORI_W)) -> Placement:
    ends = np.stack([project_corners(z, a, P34, K, T) for z, a in zip(z0s, alphas)])
    return make_placement(start_corners(obj_hw, canvas_hw), ends, obj_hw, canvas_hw)


def _unpack(place):
    """(coeffs, bbox or None, max bbox width, max bbox height) of a Placement or a bare (Ba,8) tensor."""
    if isinstance(place, Placement):
        bb = place.bbox
        if bb is not None:
            bb = bb.to(device=place.coeffs.device, dtype=torch.int32).contiguous()
        return f32c(place.coeffs), bb, place.bbox_wh[0], place.bbox_wh[1]
    return f32c(place), None, 0, 0


# ----------------------------------------------------------------------------- perspective (canvas resolution)
class _Perspective(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, coeffs, oh, ow):
        im, co = f32c(img), f32c(coeffs)
        _, Cc, ph, pw = im.shape
        B = co.shape[0]
        out = torch.empty(B, Cc, oh, ow, device=im.device, dtype=torch.float32)
        check(_lib_().dmh_perspective_fwd(ptr(im), ptr(co), B, Cc, ph, pw, oh, ow, ptr(out), stream()),
              "perspective_fwd")
        ctx.save_for_backward(co)
        ctx.dims = (B, Cc, ph, pw, oh, ow)
        return out

    @staticmethod
    def backward(ctx, g_out):
        (co,) = ctx.saved_tensors
        B, Cc, ph, pw, oh, ow = ctx.dims
        g = f32c(g_out)
        gi = torch.zeros(1, Cc, ph, pw, device=g.device, dtype=torch.float32)
        check(_lib_().dmh_perspective_bwd(ptr(g), ptr(co), B, Cc, ph, pw, oh, ow, ptr(gi), stream()),
              "perspective_bwd")
        return gi, None, None, None


def perspective_batch(img, coeffs, canvas_hw=(ORI_H, ORI_W)):
    """img (1,C,h,w) -> (Ba,C,H,W): Pad + torchvision.perspective for every item."""
    if img.shape[0] != 1:
        raise RuntimeError("perspective_batch expects a single (1,C,h,w) image shared by the batch")
    if isinstance(coeffs, Placement):
        coeffs = coeffs.coeffs
    return _Perspective.apply(img, coeffs, int(canvas_hw[0]), int(canvas_hw[1]))


# ----------------------------------------------------------------------------- fused apply
class _PatchApply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, obj, mask, scenes, place, oh, ow):
        o, m, s = f32c(obj), f32c(mask), f32c(scenes)
        co, bb, bw, bh = _unpack(place)
        _, _, ph, pw = o.shape
        B, _, ih, iw = s.shape
        adv = torch.empty(B, 3, oh, ow, device=o.device, dtype=torch.float32)
        mout = torch.empty(B, 1, oh, ow, device=o.device, dtype=torch.float32)
        check(_lib_().dmh_patch_apply_fwd(ptr(o), ptr(m), ptr(s), ptr(co), ptr(bb), B, ph, pw, ih, iw, oh, ow,
                                          ptr(adv), ptr(mout), stream()), "patch_apply_fwd")
        ctx.save_for_backward(m, co, *([bb] if bb is not None else []))
        ctx.dims = (B, ph, pw, ih, iw, oh, ow, bw, bh)
        ctx.mark_non_differentiable(mout)
        return adv, mout

    @staticmethod
    def backward(ctx, g_adv, _g_mask):
        m, co, *rest = ctx.saved_tensors
        bb = rest[0] if rest else None
        B, ph, pw, ih, iw, oh, ow, bw, bh = ctx.dims
        g = f32c(g_adv)
        gp = torch.zeros(1, 3, ph, pw, device=g.device, dtype=torch.float32)
        check(_lib_().dmh_patch_apply_bwd(ptr(g), ptr(m), ptr(co), ptr(bb), bw, bh, B, ph, pw, ih, iw, oh, ow,
                                          ptr(gp), stream()), "patch_apply_bwd")
        return gp, None, None, None, None, None


def apply_patch(obj, mask, scenes, coeffs, size=(320, 1024)):
    """obj (1,3,h,w) [grad], mask (1,1,h,w), scenes (Ba,3,375,1242), coeffs (Ba,8) tensor
    or Placement -> adv scenes (Ba,3,320,1024), resized masks (Ba,1,320,1024)."""
    if scenes.shape[0] != coeffs.shape[0]:
        raise RuntimeError("Batch size doesn't match!")
    return _PatchApply.apply(obj, mask, scenes, coeffs, int(size[0]), int(size[1]))


def apply_patch_fwd_bwd(obj, mask, scenes, coeffs, upstream, size=(320, 1024), grad_out=None):
    """No-autograd fast path for a PGD iteration whose upstream gradient
    d(cost)/d(adv_scene) is already known: returns (adv, mask_out, grad_patch).
    grad_out: optional flat fp32 buffer of >= obj.numel() elements that receives the patch gradient in its head
    (zeroed here).  Multi-GPU callers hand in the all-reduce buffer itself, with room for the scalar attack loss
    in the tail: the one collective of the step then runs in place, without staging launches."""
    o, m, s, up = f32c(obj), f32c(mask), f32c(scenes), f32c(upstream)
    co, bb, bw, bh = _unpack(coeffs)
    _, _, ph, pw = o.shape
    B, _, ih, iw = s.shape
    oh, ow = int(size[0]), int(size[1])
    adv = torch.empty(B, 3, oh, ow, device=o.device, dtype=torch.float32)
    mout = torch.empty(B, 1, oh, ow, device=o.device, dtype=torch.float32)
    if grad_out is None:
        gp = torch.zeros(1, 3, ph, pw, device=o.device, dtype=torch.float32)
    else:
        if (grad_out.dtype != torch.float32 or not grad_out.is_contiguous() or grad_out.dim() != 1
                or grad_out.numel() < o.numel() or grad_out.device != o.device):
            raise RuntimeError("grad_out must be a contiguous 1-D fp32 buffer of >= %d elements on %s" % (o.numel(), o.device))
        grad_out.zero_()
        gp = grad_out[:o.numel()].view(1, 3, ph, pw)
    lib = _lib_()
    check(lib.dmh_patch_apply_fwd(ptr(o), ptr(m), ptr(s), ptr(co), ptr(bb), B, ph, pw, ih, iw, oh, ow, ptr(adv),
                                  ptr(mout), stream()), "patch_apply_fwd")
    check(lib.dmh_patch_apply_bwd(ptr(up), ptr(m), ptr(co), ptr(bb), bw, bh, B, ph, pw, ih, iw, oh, ow, ptr(gp),
                                  stream()), "patch_apply_bwd")
    return adv, mout, gp


# ----------------------------------------------------------------------------- update rules
def pgd_linf_step(adv, grad, clean, alpha, eps):
    """phy_obj_atk.py:98-100 in one launch."""
    a, g, c = f32c(adv), f32c(grad), f32c(clean)
    out = torch.empty_like(a)
    check(_lib_().dmh_pgd_linf_step(ptr(a), ptr(g), ptr(c), a.numel(), float(alpha), float(eps), ptr(out), stream()),
          "pgd_linf_step")
    return out


def apgd_linf_step(x_adv, x_adv_old, grad, x0, step, a, eps):
    """phy_obj_atk_apgd.py:214-222 (L-inf branch) in one launch: momentum step with the double eps-ball / [0,1]
    projection; returns the new iterate."""
    xa, xo, g, c = f32c(x_adv), f32c(x_adv_old), f32c(grad), f32c(x0)
    out = torch.empty_like(xa)
    check(_lib_().dmh_apgd_linf_step(ptr(xa), ptr(xo), ptr(g), ptr(c), xa.numel(), float(step), float(a), float(eps),
                                     ptr(out), stream()), "apgd_linf_step")
    return out


def pgd_l2_step(adv, grad, clean, alpha, eps, eps_div=1e-10):
    """phy_obj_atk_l2.py:108-120 in one launch: normalised-gradient ascent, projection onto the L2 ball of radius
    eps around the clean patch, clamp to [0,1]."""
    a, g, c = f32c(adv), f32c(grad), f32c(clean)
    out = torch.empty_like(a)
    check(_lib_().dmh_pgd_l2_step(ptr(a), ptr(g), ptr(c), a.numel(), float(alpha), float(eps), float(eps_div), ptr(out),
                                  stream()), "pgd_l2_step")
    return out


def tube_light_patch(base_u8, k, b, beta, rgb, alpha=1.0, out=None):
    """One candidate of the tube-light search (light_simulation.py:132-170 + phy_obj_atk_light.py:118-122) in one
    launch: base_u8 = planar (3,h,w) uint8 object image on the device, (k, b) the beam axis y = k*x + b, `beta` the
    attenuation, `rgb` = wavelength_to_rgb(wavelength) (python floats) -> (1,3,h,w) fp32 candidate patch.  The beam
    scalars are formed here exactly as the reference forms them (python floats)."""
    import math
    if base_u8.dtype != torch.uint8 or not base_u8.is_cuda or base_u8.dim() != 3 or base_u8.shape[0] != 3:
        raise RuntimeError("base_u8 must be a (3,h,w) uint8 CUDA tensor")
    bu = base_u8.contiguous()
    _, h, w = bu.shape
    if out is None:
        out = torch.empty(1, 3, h, w, device=bu.device, dtype=torch.float32)
    k, b, beta = float(k), float(b), float(beta)
    check(_lib_().dmh_tube_light_patch(ptr(bu), h, w, k, b, math.sqrt(1 + k * k), beta, int(math.sqrt(beta) + 0.5),
                                       int(math.sqrt(beta * 20) + 0.5), float(rgb[0] * alpha), float(rgb[1] * alpha),
                                       float(rgb[2] * alpha), ptr(out), None, stream()), "tube_light_patch")
    return out


def square_linf_candidate(x_best, x, vh, vw, s, delta3, eps, out=None):
    """phy_obj_atk_square.py:263-274 in one launch: x_best / x (1,3,H,W), window [vh,vh+s) x [vw,vw+s), delta3 = the
    three per-channel moves (fp32 values of 2 * eps * sign)."""
    xb, xc = f32c(x_best), f32c(x)
    _, _, H, W = xb.shape
    if out is None:
        out = torch.empty_like(xb)
    check(_lib_().dmh_square_linf_candidate(ptr(xb), ptr(xc), H, W, int(vh), int(vw), int(s), float(delta3[0]),
                                            float(delta3[1]), float(delta3[2]), float(eps), ptr(out), stream()),
          "square_linf_candidate")
    return out


class BestKeeper:
    """`if cost < best_cost: best_cost, best = cost, candidate` on the device (dmh_keep_best): the search loops of the
    black-box attacks never wait for the host.  `best` starts as `init` (or zeros), `best_cost` as `init_cost`."""

    def __init__(self, like, init_cost=1e10, init=None):
        self.best = f32c(init).clone() if init is not None else torch.zeros_like(f32c(like))
        self._cost = [torch.full((1,), float(init_cost), device=self.best.device, dtype=torch.float32),
                      torch.empty(1, device=self.best.device, dtype=torch.float32)]
        self._cur = 0

    @property
    def best_cost(self):
        return self._cost[self._cur]

    def set_cost(self, cost):
        self._cost[self._cur].copy_(cost.detach().reshape(1))

    def offer(self, cost, cand):
        c, x = f32c(cost.detach()).reshape(1), f32c(cand)
        if x.numel() != self.best.numel():
            raise RuntimeError("candidate / best size mismatch")
        check(_lib_().dmh_keep_best(ptr(c), ptr(self._cost[self._cur]), ptr(self._cost[1 - self._cur]), ptr(x),
                                    ptr(self.best), x.numel(), stream()), "keep_best")
        self._cur = 1 - self._cur


class L0State:
    """Device-resident state of the L0 attack (patterns, Adam moments, counts)."""

    BIAS_TABLE_LEN = 4096

    def __init__(self, obj, pattern_pos, pattern_neg, lr=0.5, betas=(0.5, 0.9), eps=1e-8, clip_max=1.0,
                 device_step=True):
        """device_step (default): Adam's step index lives on the device (dmh_l0_adam_step_dev) -- an attack iteration
        can be captured into a CUDA graph and replayed with the correct bias correction at every replay; False: the
        host-side index of dmh_l0_adam_step (frozen by a capture).  Both give the same bits."""
        self.obj = f32c(obj)
        self.ppos = f32c(pattern_pos).clone()
        self.pneg = f32c(pattern_neg).clone()
        self.m_pos, self.v_pos = torch.zeros_like(self.ppos), torch.zeros_like(self.ppos)
        self.m_neg, self.v_neg = torch.zeros_like(self.ppos), torch.zeros_like(self.ppos)
        self.counts = torch.zeros(2, device=self.obj.device, dtype=torch.int64)   # [now, init]
        self.lr, self.betas, self.eps, self.clip_max = float(lr), betas, float(eps), float(clip_max)
        self.step = 0
        self.thr = self.clip_max / 255.0
        _, self.C, self.H, self.W = self.obj.shape
        self.step_state = self.bias_table = None
        if device_step:
            import ctypes as C
            n = self.BIAS_TABLE_LEN
            host = (C.c_float * (2 * n))()
            check(_lib_().dmh_l0_adam_bias_table(self.lr, float(betas[0]), float(betas[1]), n, host), "l0_adam_bias_table")
            self.bias_table = torch.tensor(list(host), dtype=torch.float32).to(self.obj.device)
            self.step_state = torch.zeros(2, device=self.obj.device, dtype=torch.int32)   # [steps done, ticket]

    def compose_count(self, first=False):
        """phy_obj_atk_l0.py:94-111: adv patch + l0 count (stays on the device)."""
        adv = torch.empty_like(self.obj)
        check(_lib_().dmh_l0_compose_count(ptr(self.obj), ptr(self.ppos), ptr(self.pneg), self.C, self.H, self.W,
                                           self.clip_max, self.thr, ptr(adv), ptr(self.counts), stream()),
              "l0_compose_count")
        if first:
            self.counts[1:2].copy_(self.counts[0:1])
        return adv

    def adam_step(self, grad_adv, mask_weight, l0_thresh):
        self.step += 1
        g = f32c(grad_adv) if grad_adv is not None else None
        if self.step_state is not None:
            check(_lib_().dmh_l0_adam_step_dev(ptr(self.obj), ptr(g), ptr(self.ppos), ptr(self.pneg), ptr(self.m_pos),
                                               ptr(self.v_pos), ptr(self.m_neg), ptr(self.v_neg), self.C, self.H, self.W,
                                               self.clip_max, ptr(self.counts), float(l0_thresh), float(mask_weight),
                                               self.betas[0], self.betas[1], self.eps, ptr(self.bias_table),
                                               self.BIAS_TABLE_LEN, ptr(self.step_state), stream()), "l0_adam_step_dev")
            return
        check(_lib_().dmh_l0_adam_step(ptr(self.obj), ptr(g), ptr(self.ppos), ptr(self.pneg), ptr(self.m_pos),
                                       ptr(self.v_pos), ptr(self.m_neg), ptr(self.v_neg), self.C, self.H, self.W,
                                       self.clip_max, ptr(self.counts), float(l0_thresh), float(mask_weight), self.lr,
                                       self.betas[0], self.betas[1], self.eps, self.step, stream()), "l0_adam_step")

    def finalize(self):
        adv = torch.empty_like(self.obj)
        pattern = torch.empty_like(self.obj)
        check(_lib_().dmh_l0_finalize(ptr(self.obj), ptr(self.ppos), ptr(self.pneg), self.obj.numel(), self.clip_max,
                                      self.thr, ptr(adv), ptr(pattern), stream()), "l0_finalize")
        return adv, pattern

    def l0(self) -> torch.Tensor:
        return self.counts[0]


def l0_count(obj, pattern_pos, pattern_neg, clip_max=1.0):
    """cal_l0 (phy_obj_atk_l0.py:43-52) as a device int64 scalar."""
    o, pp, pn = f32c(obj), f32c(pattern_pos), f32c(pattern_neg)
    cnt = torch.zeros(1, device=o.device, dtype=torch.int64)
    _, Cc, H, W = o.shape
    check(_lib_().dmh_l0_compose_count(ptr(o), ptr(pp), ptr(pn), Cc, H, W, float(clip_max), float(clip_max) / 255.0,
                                       None, ptr(cnt), stream()), "l0_compose_count")
    return cnt[0]


def topk_l0_project(pattern_pos, pattern_neg, k):
    """EXTENSION: keep the k pixels with the largest channel-max magnitude (radix
    select on the device); returns (P+, P-, keep mask) -- inputs are not modified."""
    pp, pn = f32c(pattern_pos).clone(), f32c(pattern_neg).clone()
    _, Cc, H, W = pp.shape
    keep = torch.empty(H * W, device=pp.device, dtype=torch.uint8)
    check(_lib_().dmh_topk_select(ptr(pp), ptr(pn), Cc, H, W, int(k), ptr(keep), None, stream()), "topk_select")
    return pp, pn, keep.view(1, 1, H, W).bool()

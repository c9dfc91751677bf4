"""Training-batch compositing on the GPU.

The reference builds every adversarial training item inside its DataLoader workers, on the CPU
(`DepthNetworks/monodepth2/datasets/mono_dataset.py`):

  prep_adv_data (:186-265)  to_tensor(frame) -> PhysicalTrans.project / project_w_trans of the adversarial and the
                            benign patch -> optional flip -> scene*(1-m) + obj*m -> to_pilimage (8-bit again)
  preprocess    (:119-144)  4-level pyramid, level i resized from level i-1 with PIL ANTIALIAS (Lanczos), to_tensor

`AdvBatchComposer` does the same for a whole collated batch of raw 8-bit frames that is already on the device:
one batched perspective launch per warped tensor (`physical.PhysicalTrans`), `dmh_compose_u8`, `dmh_lanczos_u8`
(Pillow's fixed-point resampling, bit-exact; `to_tensor` fused into its last pass).  It is a *host-code change* for a
maintainer (a post-collate hook instead of `prep_adv_data`), which is why it sits outside
`install()`.  There is no CPU path: tensors must be CUDA tensors.

Not mirrored: the colour jitter (`color_aug`; off in the reference's adversarial configuration,
`mono_dataset.py:297`, `adv_args['color_aug']`).
"""
from __future__ import annotations

import math
from random import random, sample
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib, patch_ops
from .physical import PhysicalTrans
from .staging import unpack_u8

_PRECISION_BITS = 32 - 8 - 2          # Resample.c: PRECISION_BITS
_coeff_cache: Dict = {}


def _lanczos(x: float) -> float:
    def sinc(v):
        if v == 0.0:
            return 1.0
        v *= math.pi
        return math.sin(v) / v
    return sinc(x) * sinc(x / 3) if -3.0 <= x < 3.0 else 0.0


def lanczos_coefficients(in_size: int, out_size: int):
    """Pillow's `precompute_coeffs` + `normalize_coeffs_8bpc` for the Lanczos filter (support 3) over the whole axis:
    (bounds int32 [out, 2] = (first input index, tap count), weights int32 [out, ksize]).  Double precision with the
    C library's sin (math.sin) and Pillow's summation order, so the integers are Pillow's."""
    scale = in_size / out_size
    fscale = scale if scale > 1.0 else 1.0
    support = 3.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    inv = 1.0 / fscale
    one = float(1 << _PRECISION_BITS)
    for o in range(out_size):
        center = (o + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        n = min(int(center + support + 0.5), in_size) - lo
        ws = [_lanczos((j + lo - center + 0.5) * inv) for j in range(n)]
        total = 0.0
        for w in ws:
            total += w
        for j, w in enumerate(ws):
            if total != 0.0:
                w = w / total
            kk[o, j] = int(w * one - 0.5) if w < 0 else int(w * one + 0.5)
        bounds[o, 0], bounds[o, 1] = lo, n
    return bounds, kk


def _device_coefficients(in_size: int, out_size: int, device):
    key = (in_size, out_size, str(device))
    hit = _coeff_cache.get(key)
    if hit is None:
        b, k = lanczos_coefficients(in_size, out_size)
        # the kernels want the weights tap-major, (ksize, out): neighbouring outputs read neighbouring integers
        hit = (torch.from_numpy(b).to(device), torch.from_numpy(np.ascontiguousarray(k.T)).to(device), k.shape[1])
        _coeff_cache[key] = hit
    return hit


def _need_u8_cuda(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("dmh_b200: %s is on %s; the compositing path is CUDA-only (no CPU fallback)" % (what, t.device))
    if t.dtype != torch.uint8:
        raise RuntimeError("%s: expected uint8, got %s" % (what, t.dtype))
    return t.contiguous()


def resize_lanczos_u8(img: torch.Tensor, out_h: int, out_w: int, want_f32: bool = False):
    """uint8 CUDA tensor [..., H, W] -> [..., out_h, out_w]; == `PIL.Image.resize((out_w, out_h), LANCZOS)` plane by
    plane, bit for bit (`transforms.Resize(..., interpolation=Image.ANTIALIAS)`, mono_dataset.py:100-104).
    want_f32: also return `to_tensor` of the result (fp32 k/255, written by the last resize pass): (u8, f32)."""
    img = _need_u8_cuda(img, "resize_lanczos_u8 input")
    H, W = img.shape[-2:]
    planes = img.numel() // max(H * W, 1)
    if planes == 0 or H == 0 or W == 0 or out_h <= 0 or out_w <= 0:
        raise RuntimeError("resize_lanczos_u8: empty image (%s -> %dx%d)" % (tuple(img.shape), out_h, out_w))
    lib = _lib.load()
    out = torch.empty(img.shape[:-2] + (out_h, out_w), dtype=torch.uint8, device=img.device)
    bx = kx = by = ky = None
    nx = ny = 0
    if out_w != W:
        bx, kx, nx = _device_coefficients(W, out_w, img.device)
    if out_h != H:
        by, ky, ny = _device_coefficients(H, out_h, img.device)
    tmp = torch.empty((planes, H, out_w), dtype=torch.uint8, device=img.device) if (nx and ny) else None
    f32 = torch.empty(out.shape, dtype=torch.float32, device=img.device) if want_f32 else None
    _lib.check(lib.dmh_lanczos_u8(_lib.ptr(img), planes, H, W, out_h, out_w, _lib.ptr(bx), _lib.ptr(kx), nx,
                                  _lib.ptr(by), _lib.ptr(ky), ny, _lib.ptr(tmp), _lib.ptr(out), _lib.ptr(f32),
                                  _lib.stream()), "lanczos_u8")
    return (out, f32) if want_f32 else out


def pyramid_u8(img: torch.Tensor, height: int, width: int, num_scales: int = 4, want_f32: bool = False):
    """MonoDataset.preprocess (mono_dataset.py:126-131): level i = Resize(height // 2^i, width // 2^i) of level i-1
    (level -1 = the native-resolution image); 8-bit levels -- with want_f32 the list of their `to_tensor` images
    (:137-144) instead, each written by the last pass of its resize."""
    out, cur = [], img
    for i in range(num_scales):
        if want_f32:
            cur, f = resize_lanczos_u8(cur, height // (2 ** i), width // (2 ** i), want_f32=True)
            out.append(f)
        else:
            cur = resize_lanczos_u8(cur, height // (2 ** i), width // (2 ** i))
            out.append(cur)
    return out


def compose_u8(scene: Optional[torch.Tensor], obj: torch.Tensor, mask: Optional[torch.Tensor],
               flip: Optional[torch.Tensor] = None) -> torch.Tensor:
    """to_pilimage(to_tensor(scene) * (1 - mask) + obj * mask) as 8-bit planes (mono_dataset.py:227-236); `flip`
    (B,) int32: items whose warped patch and mask are mirrored first (:222-225).  scene None: to_pilimage(obj)."""
    lib = _lib.load()
    obj = _lib.f32c(obj)
    B, C, H, W = obj.shape
    if scene is not None:
        scene = _need_u8_cuda(scene, "compose_u8 scene")
        if tuple(scene.shape) != (B, C, H, W):
            raise RuntimeError("compose_u8: scene %s does not match the warped patch %s" % (tuple(scene.shape), (B, C, H, W)))
        if mask is None:
            raise RuntimeError("compose_u8: a scene needs a mask")
    if mask is not None:
        mask = _lib.f32c(mask)
        if tuple(mask.shape) != (B, 1, H, W):
            raise RuntimeError("compose_u8: mask must be (B,1,H,W)")
    if flip is not None:
        flip = flip.to(device=obj.device, dtype=torch.int32).contiguous()
        if flip.numel() != B:
            raise RuntimeError("compose_u8: one flip flag per item")
    out = torch.empty((B, C, H, W), dtype=torch.uint8, device=obj.device)
    _lib.check(lib.dmh_compose_u8(_lib.ptr(scene), _lib.ptr(obj), _lib.ptr(mask), _lib.ptr(flip), B, C, H, W,
                                  _lib.ptr(out), _lib.stream()), "compose_u8")
    return out


def compose_patch_u8(scene: torch.Tensor, patch_a: torch.Tensor, patch_b: Optional[torch.Tensor], patch_mask: torch.Tensor,
                     place, flip: Optional[torch.Tensor] = None, want_mask: bool = False, out_a=None, out_b=None,
                     active: Optional[torch.Tensor] = None):
    """`compose_u8` with the perspective warp inside (`dmh_compose_patch_u8`): scene (B,3,H,W) uint8, patches
    (1,3,h,w) and mask (1,1,h,w) fp32, `place` a `patch_ops.Placement` (or bare (B,8) coefficients) for the canvas
    (H,W).  Returns (out_a, out_b or None, warped mask as 8-bit (B,1,H,W) or None).  Bit-identical to
    `compose_u8(scene, perspective_batch(patch, place), perspective_batch(mask, place), flip)` without the fp32
    canvases.  `active` (B,) int32: items with 0 get no patch (their frame passes through byte for byte)."""
    lib = _lib.load()
    scene = _need_u8_cuda(scene, "compose_patch_u8 scene")
    B, C, H, W = scene.shape
    patch_a, patch_mask = _lib.f32c(patch_a), _lib.f32c(patch_mask)
    if patch_b is not None:
        patch_b = _lib.f32c(patch_b)
        if patch_b.shape != patch_a.shape:
            raise RuntimeError("compose_patch_u8: the two patches must have one shape")
    _, pc, ph, pw = patch_a.shape
    if C != 3 or pc != 3 or tuple(patch_mask.shape) != (1, 1, ph, pw):
        raise RuntimeError("compose_patch_u8: expected (B,3,H,W) scenes, (1,3,h,w) patches and a (1,1,h,w) mask")
    coeffs, bbox, _, _ = patch_ops._unpack(place)
    if coeffs.shape[0] != B:
        raise RuntimeError("Batch size doesn't match!")
    if flip is not None:
        flip = flip.to(device=scene.device, dtype=torch.int32).contiguous()
        if flip.numel() != B:
            raise RuntimeError("compose_patch_u8: one flip flag per item")
    if active is not None:
        active = active.to(device=scene.device, dtype=torch.int32).contiguous()
        if active.numel() != B:
            raise RuntimeError("compose_patch_u8: one synthesis flag per item")
    for o in (out_a, out_b):
        if o is not None and (o.shape != scene.shape or o.dtype != torch.uint8 or not o.is_contiguous()
                              or o.device != scene.device):
            raise RuntimeError("compose_patch_u8: preallocated outputs must be contiguous uint8 of the scene's shape")
    if out_a is None:
        out_a = torch.empty_like(scene)
    if patch_b is None:
        out_b = None
    elif out_b is None:
        out_b = torch.empty_like(scene)
    m_out = torch.empty((B, 1, H, W), dtype=torch.uint8, device=scene.device) if want_mask else None
    _lib.check(lib.dmh_compose_patch_u8(_lib.ptr(scene), _lib.ptr(patch_a), _lib.ptr(patch_b), _lib.ptr(patch_mask),
                                        _lib.ptr(coeffs), _lib.ptr(bbox), _lib.ptr(flip), _lib.ptr(active), B, ph, pw, H, W,
                                        _lib.ptr(out_a), _lib.ptr(out_b), _lib.ptr(m_out), _lib.stream()),
               "compose_patch_u8")
    return out_a, out_b, m_out


def color_jitter_u8(img: torch.Tensor, params: Sequence, want_u8: bool = True, want_f32: bool = False):
    """`transforms.ColorJitter` with drawn parameters on a batch of 8-bit frames (`dmh_color_jitter_u8`).

    img (B,3,H,W) uint8 CUDA; params: one entry per item -- the tuple `ColorJitter.get_params(...)` returns,
    (fn_idx, brightness_factor, contrast_factor, saturation_factor, hue_factor) with factors possibly None, or None
    for an item without augmentation (`color_aug = lambda x: x`, mono_dataset.py:347-348).  The same parameters
    apply to every pyramid level of an item (:120-125): call once per level.  Returns (uint8 or None, fp32 or None),
    fp32 = byte / 255 (`to_tensor`).  Bit-exact against Pillow / torchvision (oracle/pil_enhance.py)."""
    lib = _lib.load()
    img = _need_u8_cuda(img, "color_jitter_u8")
    B, C, H, W = img.shape
    if C != 3 or len(params) != B:
        raise RuntimeError("color_jitter_u8: expected (B,3,H,W) frames and one parameter tuple per item")
    order = np.full((B, 4), -1, dtype=np.int32)
    fac = np.ones((B, 3), dtype=np.float32)
    shift = np.zeros((B,), dtype=np.int32)
    for i, p in enumerate(params):
        if p is None:
            continue
        fn_idx, bf, cf, sf, hf = p
        for k, fn in enumerate([int(v) for v in fn_idx]):
            val = (bf, cf, sf, hf)[fn]
            order[i, k] = fn if val is not None else -1
        for k, val in enumerate((bf, cf, sf)):
            if val is not None:
                fac[i, k] = np.float32(float(val))
        if hf is not None:
            shift[i] = int(float(hf) * 255) & 0xff           # np.uint8(hue_factor * 255), functional_pil.adjust_hue
    dev = img.device
    order_d = torch.from_numpy(order).to(dev)
    fac_d = torch.from_numpy(fac).to(dev)
    shift_d = torch.from_numpy(shift).to(dev)
    sums = torch.empty(B, dtype=torch.int64, device=dev)
    out_u8 = torch.empty_like(img) if want_u8 else None
    out_f32 = torch.empty(img.shape, dtype=torch.float32, device=dev) if want_f32 else None
    _lib.check(lib.dmh_color_jitter_u8(_lib.ptr(img), B, H, W, _lib.ptr(order_d), _lib.ptr(fac_d), _lib.ptr(shift_d),
                                       _lib.ptr(sums), _lib.ptr(out_u8), _lib.ptr(out_f32), _lib.stream()),
               "color_jitter_u8")
    return out_u8, out_f32


class AdvBatchComposer:
    """`MonoDataset.set_adv_train` / `update_adv_obj` / `prep_adv_data` / `preprocess` for a collated batch on the
    device (mono_dataset.py:146-265, 119-144).

    obj_tensor / mask_tensor: the benign patch (1,3,h,w) and its mask (1,1,h,w), CUDA; cfg: {'path': calib file};
    height, width: network input size (scale 0); dist_range as `train_dist_range`."""

    def __init__(self, obj_tensor, mask_tensor, cfg, height, width, num_scales=4, dist_range=list(range(5, 10, 2)),
                 ori_H=patch_ops.ORI_H, ori_W=patch_ops.ORI_W, half_no_synthesis=False):
        self.height, self.width, self.num_scales = height, width, num_scales
        self.half_no_synthesis = half_no_synthesis            # args['half_no_synthesis'], mono_dataset.py:160
        self.ori_H, self.ori_W = ori_H, ori_W
        self.obj_mask = mask_tensor
        self.obj_img_ben = obj_tensor
        self.obj_img_adv = obj_tensor.clone()
        size = (1, 3, ori_H, ori_W)
        self.ben_trans = PhysicalTrans(self.obj_img_ben, self.obj_mask, cfg, size, dist_range=dist_range)
        self.adv_trans = PhysicalTrans(self.obj_img_adv, self.obj_mask, cfg, size, dist_range=dist_range)
        # mono_dataset.py:169-175, 111-116
        self.adv_K = np.array([[0.58, 0, 0.5, 0], [0, 1.92, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
        self.adv_K[0, :] *= ori_W
        self.adv_K[1, :] *= ori_H
        self.stereo_T = np.eye(4, dtype=np.float32)
        self.stereo_T[0, 3] = -0.54
        self._place_cache: Dict = {}

    def update_adv_obj(self, obj_img_adv: torch.Tensor) -> None:
        """mono_dataset.py:178-184 with the attack's result handed in (the attack itself is `attacks.Phy_obj_atk*`)."""
        self.obj_img_adv = obj_img_adv
        self.adv_trans.reset_img(self.obj_img_adv, self.obj_mask)

    def _placement(self, trans: PhysicalTrans, z0, alpha, with_T):
        """Per-item placements (mono_dataset.py:207-223): homography + conservative box, cached per
        (distance, angle, camera) -- the reference draws them from 3 x 13 values -- and sent as one small copy."""
        _, _, h, w = trans.obj_img.size()
        keys = [(float(z), float(a), bool(t), h, w) for z, a, t in zip(z0, alpha, with_T)]
        miss = [k for k in dict.fromkeys(keys) if k not in self._place_cache]
        if miss:
            ends = np.stack([patch_ops.project_corners(z, a, trans.P, self.adv_K, self.stereo_T if t else None)
                             for (z, a, t, _, _) in miss])
            pl = patch_ops.make_placement(trans.pos_obj_img_start, ends, (h, w), (self.ori_H, self.ori_W))
            if len(self._place_cache) > 4096:
                self._place_cache.clear()
            for i, k in enumerate(miss):
                self._place_cache[k] = (pl.coeffs[i].clone(), pl.bbox[i].clone())
        co = torch.stack([self._place_cache[k][0] for k in keys])
        bb = torch.stack([self._place_cache[k][1] for k in keys])
        wh = (int((bb[:, 2] - bb[:, 0] + 1).max()), int((bb[:, 3] - bb[:, 1] + 1).max()))
        dev = trans.obj_img.device
        return patch_ops.Placement(co.to(dev, non_blocking=True), bb.to(dev, non_blocking=True), wh)

    def __call__(self, color_0: torch.Tensor, color_s: torch.Tensor, sides: Sequence[str], do_flip: Sequence[bool],
                 z0_sample: Optional[Sequence[float]] = None, alpha_sample: Optional[Sequence[float]] = None,
                 synthesize: Optional[Sequence[bool]] = None, color_aug: Optional[Sequence] = None):
        """color_0 / color_s: (B,3,ori_H,ori_W) uint8 CUDA -- frame 0 and its stereo partner at native resolution, as
        `get_color` returns them (already mirrored for the items with do_flip, mono_dataset.py:325-329); sides[i]:
        'l' / 'r', the side frame 0 of item i was taken from; do_flip[i]: mirror the warped patch too (:222-225);
        z0_sample / alpha_sample: one placement per item (drawn like `PhysicalTrans.project` if None);
        synthesize[i] (only with half_no_synthesis, :321-328; drawn with `random.random() > 0.5` if None): items
        with False keep their raw frames -- ("color_objmask", 0, 0) / ("objdepth", 0, 0) are then not produced at all,
        as in the reference (:253-255).
        color_aug: per item, the tuple `transforms.ColorJitter.get_params(...)` drew for it (`do_color_aug`, :344-348)
        or None (identity); None for the whole batch = no colour augmentation (the adversarial configuration).  With
        it the "color_aug" entries and ("color_ben", 0, 0) carry the jittered levels (:132-133, 143-144).

        Returns the dictionary entries `prep_adv_data` + `preprocess` produce, as fp32 CUDA tensors:
        ("color_aug", 0 | "s", 0..S-1), ("color", 0 | "s", 0..S-1), ("color_ben", 0, 0), ("color_objmask", 0, 0),
        ("objdepth", 0, 0)."""
        color_0 = _need_u8_cuda(color_0, "color_0")
        color_s = _need_u8_cuda(color_s, "color_s")
        B = color_0.shape[0]
        if tuple(color_0.shape) != (B, 3, self.ori_H, self.ori_W) or color_s.shape != color_0.shape:
            raise RuntimeError("AdvBatchComposer: frames must be (B,3,%d,%d) uint8" % (self.ori_H, self.ori_W))
        if len(sides) != B or len(do_flip) != B:
            raise RuntimeError("Batch size doesn't match!")
        if z0_sample is None:
            z0_sample = [sample(self.ben_trans.dist_range, 1)[0] for _ in range(B)]
        if alpha_sample is None:
            alpha_sample = [sample(self.ben_trans.angle_range, 1)[0] for _ in range(B)]
        active = None
        if self.half_no_synthesis:
            if synthesize is None:
                synthesize = [random() > 0.5 for _ in range(B)]
            if len(synthesize) != B:
                raise RuntimeError("Batch size doesn't match!")
            active = torch.tensor([1 if a else 0 for a in synthesize], dtype=torch.int32).to(color_0.device)
        right = [s != "l" for s in sides]
        # frame 0 sees the placement of its own camera: left camera (project) for side 'l', right camera
        # (project_w_trans with stereo_T) for side 'r'; the stereo partner sees the other one (:207-223)
        place_0 = self._placement(self.ben_trans, z0_sample, alpha_sample, right)
        place_s = self._placement(self.ben_trans, z0_sample, alpha_sample, [not r for r in right])
        flip = torch.tensor([1 if f else 0 for f in do_flip], dtype=torch.int32).to(color_0.device, non_blocking=True)
        with torch.no_grad():
            # frame 0: the adversarial frame, the benign frame (:239-251) and the warped mask image (:254) share the
            # scene, the placement and the mask -> one pass; the stereo partner carries the benign patch
            comp = torch.empty((3,) + tuple(color_0.shape), dtype=torch.uint8, device=color_0.device)
            _, _, objmask = compose_patch_u8(color_0, self.obj_img_adv, self.obj_img_ben, self.obj_mask, place_0, flip,
                                             want_mask=not self.half_no_synthesis, out_a=comp[0], out_b=comp[2],
                                             active=active)
            compose_patch_u8(color_s, self.obj_img_ben, None, self.obj_mask, place_s, flip, out_a=comp[1],
                             active=active)
            out = {}
            S = self.num_scales
            # the three composites go through the pyramid as one stack: 2 resize passes + 1 unpack per level
            names = (("color_aug", 0), ("color_aug", "s"), ("color", 0))
            stack = comp.view((3 * B,) + tuple(color_0.shape[1:]))
            if color_aug is None:
                for i, f in enumerate(pyramid_u8(stack, self.height, self.width, S, want_f32=True)):
                    for j, (name, fid) in enumerate(names):
                        out[(name, fid, i)] = f[j * B:(j + 1) * B]
                for i in range(S):                                    # :258: color['s'] is color_aug['s']
                    out[("color", "s", i)] = out[("color_aug", "s", i)]
                out[("color_ben", 0, 0)] = out[("color", 0, 0)]       # :132-133 with the identity colour jitter
            else:
                if len(color_aug) != B:
                    raise RuntimeError("Batch size doesn't match!")
                pa = list(color_aug)
                for i, u in enumerate(pyramid_u8(stack, self.height, self.width, S, want_f32=False)):
                    # "color" entries: to_tensor of the level; "color_aug": to_tensor(color_aug(level)) (:140-144)
                    plain = unpack_u8(u[B:])                          # [stereo partner, benign frame 0]
                    out[("color", "s", i)], out[("color", 0, i)] = plain[:B], plain[B:]
                    _, jit = color_jitter_u8(u[:2 * B], pa + pa, want_u8=False, want_f32=True)
                    out[("color_aug", 0, i)], out[("color_aug", "s", i)] = jit[:B], jit[B:]
                    if i == 0:                                        # :132-133
                        _, out[("color_ben", 0, 0)] = color_jitter_u8(u[2 * B:], pa, want_u8=False, want_f32=True)
            if not self.half_no_synthesis:
                # mask.expand(-1, 3, -1, -1): three identical planes -> resize one, replicate
                _, om = resize_lanczos_u8(objmask, self.height, self.width, want_f32=True)
                # (a stride-0 view: the trainer reads channel 0 only, M2/trainer.py:552; .contiguous() if needed)
                out[("color_objmask", 0, 0)] = om.expand(-1, 3, -1, -1)
                out[("objdepth", 0, 0)] = torch.tensor([[[float(z)]] for z in z0_sample], dtype=torch.float32,
                                                       device=color_0.device)
        return out

"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the
UNMODIFIED reference (/root/reference, imported through oracle/refload.py) on
the seeded synthetic inputs of depthmodelhardening_b200/synth.py.

Run in the build container (the reference tree does not exist on the GPU box):
    python -m oracle.make_golden
The fixtures are committed; tests compare the oracle restatement AND the CUDA
path against them.
"""
from __future__ import annotations

import os
import random
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

PHOTO_CASES = {
    # name: (synth kwargs, option overrides)
    "stereo_small": (dict(batch=2, height=64, width=96, frame_ids=(0, "s"), seed=11), {}),
    "stereo_iid": (dict(batch=2, height=64, width=96, frame_ids=(0, "s"), seed=12, disp_kind="iid",
                        image_kind="iid"), {}),
    "mono_small": (dict(batch=2, height=64, width=96, frame_ids=(0, -1, 1), seed=13), {}),
    "mono_stereo": (dict(batch=1, height=32, width=64, frame_ids=(0, -1, 1, "s"), seed=14), {}),
    "no_ssim": (dict(batch=2, height=32, width=64, frame_ids=(0, "s"), seed=15), dict(no_ssim=True)),
    "avg_reproj": (dict(batch=2, height=32, width=64, frame_ids=(0, -1, 1), seed=16), dict(avg_reprojection=True)),
    "no_automask": (dict(batch=2, height=32, width=64, frame_ids=(0, -1, 1), seed=17),
                    dict(disable_automasking=True)),
    "no_automask_f1": (dict(batch=2, height=32, width=64, frame_ids=(0, "s"), seed=18),
                       dict(disable_automasking=True)),
}


def ref_opts(pb, **over):
    o = dict(scales=list(pb.scales), v1_multiscale=False, height=pb.height, width=pb.width,
             min_depth=pb.min_depth, max_depth=pb.max_depth, frame_ids=list(pb.frame_ids),
             pose_model_type="separate_resnet", disable_automasking=False, no_ssim=False, adv_train=False,
             supervised_adv=False, contrastive_learning=False, no_original_train=False, avg_reprojection=False,
             predictive_mask=False, disparity_smoothness=1e-3, batch_size=pb.batch)
    o.update(over)
    return SimpleNamespace(**o)


class _InjectedRandn:
    """Stands in for torch.randn inside Trainer.compute_losses (trainer.py:644)
    so the tie-break noise is the injected tensor (noise/1e-5 per scale)."""

    def __init__(self, pb, n_ident):
        self.queue = [pb.noise[s][:, :n_ident] / 0.00001 for s in pb.scales]
        self.real = torch.randn

    def __call__(self, *a, **k):
        t = self.queue.pop(0)
        shape = tuple(a[0]) if len(a) == 1 and not isinstance(a[0], int) else tuple(a)
        assert tuple(t.shape) == shape, (t.shape, shape)
        return t.clone()


def run_reference_objective(pb, **over):
    """Call the reference's own generate_images_pred + compute_losses unbound."""
    ref = refload.load()
    Trainer, L = ref.trainer.Trainer, ref.layers
    opt = ref_opts(pb, **over)
    me = SimpleNamespace(opt=opt, ssim=L.SSIM(), num_scales=len(opt.scales),
                         backproject_depth={0: L.BackprojectDepth(pb.batch, pb.height, pb.width)},
                         project_3d={0: L.Project3D(pb.batch, pb.height, pb.width)})
    me.compute_reprojection_loss = lambda pred, target: Trainer.compute_reprojection_loss(me, pred, target)
    inputs = {("K", 0): pb.K, ("inv_K", 0): pb.inv_K}
    for (f, s), v in pb.color.items():
        inputs[("color", f, s)] = v
    if "s" in pb.T:
        inputs["stereo_T"] = pb.T["s"]
    disps = {s: pb.disp[s].clone().requires_grad_(True) for s in pb.scales}
    outputs = {("disp", s): disps[s] for s in pb.scales}
    for f in pb.frame_ids[1:]:
        if f != "s":
            outputs[("cam_T_cam", 0, f)] = pb.T[f]
    n_src = len(pb.frame_ids) - 1
    n_ident = 1 if opt.avg_reprojection else n_src
    inj = _InjectedRandn(pb, n_ident)
    cuda_attr = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self     # trainer.py:634,645 hard-code .cuda()
    torch.randn = inj
    try:
        Trainer.generate_images_pred(me, inputs, outputs)
        losses = Trainer.compute_losses(me, inputs, outputs)
    finally:
        torch.randn = inj.real
        torch.Tensor.cuda = cuda_attr
    losses["loss"].backward()
    return losses, outputs, disps


def golden_photo(name, skw, over):
    pb = synth.photo_batch(**skw)
    losses, outputs, disps = run_reference_objective(pb, **over)
    out = {"loss": losses["loss"].detach().numpy()}
    for s in pb.scales:
        out["loss_%d" % s] = losses["loss/%d" % s].detach().numpy()
        out["grad_disp_%d" % s] = disps[s].grad.numpy()
        key = "identity_selection/%d" % s
        if key in outputs:
            out["ident_sel_%d" % s] = outputs[key].numpy().astype(np.uint8)
    for f in pb.frame_ids[1:]:
        tag = str(f)
        out["warped_%s_0" % tag] = outputs[("color", f, 0)].detach().numpy()
        out["grid_%s_0" % tag] = outputs[("sample", f, 0)].detach().numpy()
        out["warped_%s_3" % tag] = outputs[("color", f, 3)].detach().numpy()
    out["depth_0"] = outputs[("depth", 0, 0)].detach().numpy()
    np.savez_compressed(os.path.join(GOLD, "photo_%s.npz" % name), **out)
    print("photo", name, float(out["loss"]))


def golden_layers():
    """Op-level goldens for the drop-in classes (layers.py)."""
    ref = refload.load()
    L = ref.layers
    pb = synth.photo_batch(batch=2, height=48, width=80, frame_ids=(0, -1), seed=21)
    B, H, W = pb.batch, pb.height, pb.width
    depth = (1.0 / (0.01 + 9.99 * pb.disp[0])).clone().requires_grad_(True)
    bp, pj, ss = L.BackprojectDepth(B, H, W), L.Project3D(B, H, W), L.SSIM()
    pts = bp(depth, pb.inv_K)
    T = pb.T[-1].clone().requires_grad_(True)
    grid = pj(pts, pb.K, T)
    g_up = synth.randn(grid.shape, 22)
    (grid * g_up).sum().backward()
    x = pb.color[(0, 0)].clone().requires_grad_(True)
    y = pb.color[(-1, 0)].clone().requires_grad_(True)
    s = ss(x, y)
    s_up = synth.randn(s.shape, 23)
    (s * s_up).sum().backward()
    d = pb.disp[0].clone().requires_grad_(True)
    img = pb.color[(0, 0)].clone().requires_grad_(True)
    sm = L.get_smooth_loss(d, img)
    sm.backward()
    sd, dp = L.disp_to_depth(pb.disp[0], 0.1, 100.0)
    np.savez_compressed(os.path.join(GOLD, "layers.npz"), points=pts.detach().numpy(), grid=grid.detach().numpy(),
                        grad_depth=depth.grad.numpy(), grad_T=T.grad.numpy(), ssim=s.detach().numpy(),
                        grad_x=x.grad.numpy(), grad_y=y.grad.numpy(), smooth=sm.detach().numpy(),
                        grad_disp=d.grad.numpy(), grad_img=img.grad.numpy(), scaled_disp=sd.numpy(),
                        depth_from_disp=dp.numpy())
    print("layers ok")


def _sample_idx(n, count, seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g)[:count].numpy()


def _crop(t):
    # dense window around the projected patch (canvas centre region)
    return t[..., 100:228, 440:760]


class TinyDepth(torch.nn.Module):
    """Deterministic stand-in for the depth network (outside the graft)."""

    def __init__(self):
        super().__init__()
        self.c1 = torch.nn.Conv2d(3, 4, 3, padding=1)
        self.c2 = torch.nn.Conv2d(4, 1, 3, padding=1)
        g = torch.Generator()
        g.manual_seed(31)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)

    def forward(self, x):
        return torch.sigmoid(self.c2(torch.tanh(self.c1(x))))


def golden_patch():
    ref = refload.load()
    PT = ref.physicalTrans.PhysicalTrans
    pbt = synth.patch_batch(batch=3, seed=0)
    conf = {"path": ref.calib_path}
    out = {}
    # --- A1-A3: PhysicalTrans.project (KITTI P2 path) and project_w_trans (K path)
    pt = PT(pbt.obj.clone().requires_grad_(True), pbt.mask, conf, (1, 3, synth.ORI_H, synth.ORI_W))
    z0 = [5, 7, 9]
    al = [-30, 0, 25]
    imgs, masks, _, _ = pt.project(batch_size=3, z0_sample=z0, alpha_sample=al)
    out["z0"], out["alpha"] = np.array(z0, dtype=np.float64), np.array(al, dtype=np.float64)
    out["proj_img_crop"] = _crop(imgs).detach().numpy()
    out["proj_mask_crop"] = _crop(masks).detach().numpy()
    out["proj_img_sum"] = imgs.detach().double().sum().numpy()
    out["proj_mask_sum"] = masks.detach().double().sum().numpy()
    out["corners"] = np.stack([pt.objPosOnImage(z, a) for z, a in zip(z0, al)])
    K = np.array([[0.58 * 1242, 0, 0.5 * 1242, 0], [0, 1.92 * 375, 0.5 * 375, 0], [0, 0, 1, 0], [0, 0, 0, 1]],
                 dtype=np.float32)
    T = np.eye(4, dtype=np.float32)
    T[0, 3] = -0.1
    imgs_k, masks_k = pt.project_w_trans(T, z0, al, K=K)
    out["projk_img_crop"] = _crop(imgs_k).detach().numpy()
    out["projk_mask_sum"] = masks_k.detach().double().sum().numpy()
    out["corners_k"] = np.stack([pt.objPosOnImage(z, a, K) for z, a in zip(z0, al)])
    # --- A4-A5 composite + Resize, and gradient to the patch
    from torchvision.transforms import Resize
    rs = Resize([320, 1024])
    adv = rs(pbt.scenes * (1 - masks) + imgs * masks)
    m_rs = rs(masks)
    up = pbt.upstream
    (adv * up).sum().backward()
    idx = _sample_idx(adv.numel(), 4096, 41)
    out["adv_idx"] = idx
    out["adv_samples"] = adv.detach().reshape(-1)[idx].numpy()
    out["adv_sum"] = adv.detach().double().sum().numpy()
    out["adv_crop"] = adv.detach()[:, :, 90:200, 380:640].numpy()
    out["mask_rs_crop"] = m_rs.detach()[:, :, 90:200, 380:640].numpy()
    out["mask_rs_sum"] = m_rs.detach().double().sum().numpy()
    out["grad_patch"] = pt.obj_img.grad.numpy()[:, :, ::3, ::3].copy()
    out["grad_patch_sum"] = pt.obj_img.grad.double().sum().numpy()
    out["grad_patch_abs_sum"] = pt.obj_img.grad.double().abs().sum().numpy()
    np.savez_compressed(os.path.join(GOLD, "patch.npz"), **out)
    print("patch ok")

    # --- A6: full L-inf attack, 2 steps, no random start, tiny model
    model = TinyDepth()
    random.seed(5)
    atk = ref.atk_linf.Phy_obj_atk(model, pbt.obj.clone(), pbt.mask.clone(), eps=0.1, alpha=0.02, steps=2,
                                   random_start=False, dist_range=list(range(5, 10, 2)))
    adv_s, ben_s, m_out, obj_adv = atk(pbt.scenes.clone(), 3)
    o2 = {"obj_adv": obj_adv.detach().numpy()[:, :, ::2, ::2].copy(),
          "obj_adv_sum": obj_adv.detach().double().sum().numpy(),
          "adv_scene_sum": adv_s.detach().double().sum().numpy(),
          "ben_scene_sum": ben_s.detach().double().sum().numpy(),
          "mask_out_sum": m_out.detach().double().sum().numpy(),
          "adv_scene_crop": adv_s.detach()[:, :, 90:200, 380:640].numpy()}
    np.savez_compressed(os.path.join(GOLD, "attack_linf.npz"), **o2)
    print("linf ok")

    # --- A7-A8: full L0 attack, steps=2 (4 Adam iterations), tiny model
    random.seed(6)
    np.random.seed(7)
    atk0 = ref.atk_l0.Phy_obj_atk_l0(model, pbt.obj.clone(), pbt.mask.clone(), adam_lr=0.5, steps=2, mask_wt=0.06,
                                     l0_thresh=0.1, dist_range=list(range(5, 10, 2)))
    adv_s, ben_s, m_out, obj_adv = atk0(pbt.scenes.clone(), 3)
    thr = 1.0 / 255.0
    surv = (torch.sum(torch.abs(atk0.pattern), dim=1) != 0)
    o3 = {"obj_adv": obj_adv.detach().numpy()[:, :, ::2, ::2].copy(),
          "obj_adv_sum": obj_adv.detach().double().sum().numpy(),
          "l0_count": np.array(int(atk0.cal_l0()), dtype=np.int64),
          "survivors": np.packbits(surv.numpy().astype(np.uint8)),
          "pattern_pos_tensor": atk0.pattern_pos_tensor.detach().numpy()[:, :, ::2, ::2].copy(),
          "adv_scene_sum": adv_s.detach().double().sum().numpy(), "thr": np.array(thr)}
    np.savez_compressed(os.path.join(GOLD, "attack_l0.npz"), **o3)
    print("l0 ok", int(o3["l0_count"]))


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    for name, (skw, over) in PHOTO_CASES.items():
        golden_photo(name, skw, over)
    golden_layers()
    golden_patch()


if __name__ == "__main__":
    main()

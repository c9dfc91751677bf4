"""TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product package.

One step of the hot path executed by the UNMODIFIED reference (loaded by oracle/refload.py from /root/reference or
from the copy oracle/build_ref.py ships to the GPU box), on any device:

  stage 2  `Trainer.generate_images_pred` + `Trainer.compute_losses` (DepthNetworks/monodepth2/trainer.py:472-523,
           539-674) called unbound on a namespace that carries exactly the attributes they read, + `.backward()` to
           the four disparity maps -- the reference's methods, its `layers.BackprojectDepth / Project3D / SSIM`.
  stage 1  one iteration of `Phy_obj_atk_l0.forward`'s loop (torchattacks/attacks/phy_obj_atk_l0.py:94-138) with the
           depth network's gradient supplied: the reference's own `PhysicalTrans.reset_img / project` and
           torchvision `Resize`, the loop body's statements in the reference's order, `torch.optim.Adam`.  The class
           itself cannot be called at batch 32: it draws `random.sample(dist_range, batch_size)` from 13 values
           (physicalTrans.py:150,155), so the placements are passed explicitly as `project()` allows.

Used by bench.py (`--impl reference` on the host cores; `gpu_eager_baseline` on cuda) and by the GPU integration
tests.  The tie-break noise (`torch.randn(...).cuda()`, trainer.py:644-645) is drawn by the reference itself unless
`inject_noise` replaces it with the batch's seeded noise (parity runs).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from oracle import refload


def ref_opts(pb, **over):
    o = dict(scales=list(pb.scales), v1_multiscale=False, height=pb.height, width=pb.width,
             min_depth=pb.min_depth, max_depth=pb.max_depth, frame_ids=list(pb.frame_ids),
             pose_model_type="separate_resnet", disable_automasking=False, no_ssim=False, adv_train=False,
             supervised_adv=False, contrastive_learning=False, no_original_train=False, avg_reprojection=False,
             predictive_mask=False, disparity_smoothness=1e-3, batch_size=pb.batch)
    o.update(over)
    return SimpleNamespace(**o)


class _InjectedRandn:
    def __init__(self, pb, n_ident):
        self.queue = [pb.noise[s][:, :n_ident] / 0.00001 for s in pb.scales]

    def __call__(self, *a, **k):
        t = self.queue.pop(0)
        return t.clone()


def reference_trainer(pb, device, **over):
    """The namespace `Trainer.generate_images_pred / compute_losses` run on (what `Trainer.__init__` sets up,
    trainer.py:157-170, for the attributes the two methods read)."""
    ref = refload.load()
    Trainer, L = ref.trainer.Trainer, ref.layers
    opt = ref_opts(pb, **over)
    me = SimpleNamespace(opt=opt, ssim=L.SSIM().to(device), num_scales=len(opt.scales), device=device,
                         backproject_depth={0: L.BackprojectDepth(pb.batch, pb.height, pb.width).to(device)},
                         project_3d={0: L.Project3D(pb.batch, pb.height, pb.width).to(device)})
    me.compute_reprojection_loss = lambda pred, target: Trainer.compute_reprojection_loss(me, pred, target)
    return ref, me


def batch_inputs(pb):
    inputs = {("K", 0): pb.K, ("inv_K", 0): pb.inv_K}
    for (f, s), v in pb.color.items():
        inputs[("color", f, s)] = v
    if "s" in pb.T:
        inputs["stereo_T"] = pb.T["s"]
    return inputs


class Stage2Reference:
    """generate_images_pred + compute_losses + backward of the reference on `device` (pb already there)."""

    def __init__(self, pb, device, inject_noise=False, **over):
        self.pb = pb
        self.device = torch.device(device)
        self.ref, self.me = reference_trainer(pb, self.device, **over)
        self.inputs = batch_inputs(pb)
        self.disps = {s: pb.disp[s].clone().requires_grad_(True) for s in pb.scales}
        self.inject = inject_noise

    def step(self):
        Trainer = self.ref.trainer.Trainer
        pb = self.pb
        for d in self.disps.values():
            d.grad = None
        outputs = {("disp", s): self.disps[s] for s in pb.scales}
        for f in pb.frame_ids[1:]:
            if f != "s":
                outputs[("cam_T_cam", 0, f)] = pb.T[f]
        patched = []
        if self.device.type != "cuda":                 # trainer.py:634,645 hard-code .cuda()
            patched.append((torch.Tensor, "cuda", torch.Tensor.cuda))
            torch.Tensor.cuda = lambda t, *a, **k: t
        if self.inject:
            n_src = len(pb.frame_ids) - 1
            patched.append((torch, "randn", torch.randn))
            torch.randn = _InjectedRandn(pb, 1 if self.me.opt.avg_reprojection else n_src)
        try:
            Trainer.generate_images_pred(self.me, self.inputs, outputs)
            losses = Trainer.compute_losses(self.me, self.inputs, outputs)
        finally:
            for obj, name, val in patched:
                setattr(obj, name, val)
        losses["loss"].backward()
        self.outputs = outputs
        return losses


class Stage1Reference:
    """One L0 PGD iteration (phy_obj_atk_l0.py:94-138) on the reference's PhysicalTrans, network gradient supplied."""

    def __init__(self, pt, device, mask_weight=0.06, lr=0.5):
        from torchvision.transforms import Resize
        self.ref = refload.load()
        self.pt = pt
        self.device = torch.device(device)
        conf = {"path": self.ref.calib_path}
        size = (1, 3, int(self.ref.my_utils.ori_H), int(self.ref.my_utils.ori_W))
        self.phy = self.ref.physicalTrans.PhysicalTrans(pt.obj.clone(), pt.mask, conf, size)
        self.pp = pt.pattern_pos.clone().requires_grad_(True)
        self.pn = pt.pattern_neg.clone().requires_grad_(True)
        self.opt = torch.optim.Adam([self.pp, self.pn], lr=lr, betas=(0.5, 0.9))
        self.resize = Resize([320, 1024])
        self.mask_weight = mask_weight
        self.z0 = [float(v) for v in pt.z0]
        self.alpha = [float(v) for v in pt.alpha]
        self.l0_clip = 1.0 / 255.0

    def cal_l0(self, pos, neg):                            # phy_obj_atk_l0.py:43-52
        p, n = pos.detach().clone(), neg.detach().clone()
        p[p < self.l0_clip] = 0
        n[n > -self.l0_clip] = 0
        return torch.count_nonzero(torch.sum(torch.abs(p + n), dim=1))

    def step(self):
        pt = self.pt
        B = pt.scenes.shape[0]
        pos = torch.clamp(self.pp * 1.0, min=0.0, max=1.0)
        neg = -torch.clamp(self.pn * 1.0, min=0.0, max=1.0)
        adv = torch.clamp(pt.obj + (pos + neg), min=0.0, max=1.0)
        self.cal_l0(pos, neg)
        self.phy.reset_img(adv, pt.mask)
        imgs, masks, _, _ = self.phy.project(batch_size=B, z0_sample=list(self.z0[:B]), alpha_sample=list(self.alpha[:B]))
        scenes = pt.scenes * (1 - masks) + imgs * masks
        scenes = self.resize(scenes)
        self.resize(masks)
        adv_cost = (scenes * pt.upstream).sum()            # the depth network's gradient, supplied
        mask_pos = torch.max(torch.tanh(self.pp / 10) / (2 - 1e-7) + 0.5, axis=1)[0]
        mask_neg = torch.max(torch.tanh(self.pn / 10) / (2 - 1e-7) + 0.5, axis=1)[0]
        total = adv_cost + self.mask_weight * (torch.mean(mask_pos) + torch.mean(mask_neg))
        self.opt.zero_grad()
        total.backward()
        self.opt.step()
        return scenes

"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/dh_*.npz by running the
UNMODIFIED depth-hints reference (/root/reference/DepthNetworks/depth-hints/trainer.py:
Trainer.generate_images_pred + Trainer.compute_losses, called unbound) on the seeded
synthetic inputs of depthmodelhardening_b200/synth.py.

Run in the build container, in its OWN process (the depth-hints tree and the
monodepth2 tree both define top-level modules `trainer`, `layers`, `networks`, ...):
    python -m oracle.make_golden_dh
Import-level stubs only (same list as oracle/refload.py plus numpy.lib.function_base,
gone in numpy 2 -- DH/datasets/mono_dataset.py:15); none touches arithmetic.
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from depthmodelhardening_b200 import synth  # noqa: E402
from oracle import refload  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DH_DIR = os.path.join(refload.REF_ROOT, "DepthNetworks", "depth-hints")

CASES = {
    # name: (synth kwargs, use_depth_hints, option overrides)
    "stereo_hints": (dict(batch=2, height=64, width=96, frame_ids=(0, "s"), seed=31, depth_hints=True), True, {}),
    "mono_stereo_hints": (dict(batch=1, height=32, width=64, frame_ids=(0, -1, 1, "s"), seed=32, depth_hints=True),
                          True, {}),
    "stereo_nohints": (dict(batch=2, height=32, width=64, frame_ids=(0, "s"), seed=33, depth_hints=True), False, {}),
    "avg_hints": (dict(batch=2, height=32, width=64, frame_ids=(0, -1, "s"), seed=34, depth_hints=True), True,
                  dict(avg_reprojection=True)),
    "no_ssim_hints": (dict(batch=2, height=32, width=64, frame_ids=(0, "s"), seed=35, depth_hints=True), True,
                      dict(no_ssim=True)),
}


def load_dh_trainer():
    assert "trainer" not in sys.modules, "run in a fresh process (module names clash with the monodepth2 tree)"
    fake = refload._fake
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = fake("matplotlib")
        mpl.pyplot = fake("matplotlib.pyplot", axes=None, axis=None, get=None)
        mpl.figure = fake("matplotlib.figure", Figure=object)
    for name, attrs in (("numpy.lib.utils", {}), ("numpy.lib.function_base", {"flip": np.flip})):
        try:
            importlib.import_module(name)
        except Exception:
            fake(name, **attrs)
    try:
        import tensorboardX  # noqa: F401
    except Exception:
        fake("tensorboardX", SummaryWriter=object)
    try:
        import skimage.transform  # noqa: F401
    except Exception:
        sk = fake("skimage")
        sk.transform = fake("skimage.transform", resize=None)
    import PIL.Image
    if not hasattr(PIL.Image, "ANTIALIAS"):
        PIL.Image.ANTIALIAS = PIL.Image.LANCZOS
    for p in (refload.REF_ROOT, DH_DIR):
        sys.path.insert(0, p)
    old = os.getcwd()
    os.chdir(DH_DIR)
    try:
        tmp = tempfile.mkdtemp(prefix="dmh_ref_calib_")
        refload.write_calib(tmp)
        importlib.import_module("my_utils").object_dataset_root = tmp
        ta_dir = os.path.join(refload.REF_ROOT, "torchattacks")
        ta = types.ModuleType("torchattacks")
        ta.__path__ = [ta_dir]
        sys.modules["torchattacks"] = ta
        att = types.ModuleType("torchattacks.attacks")
        att.__path__ = [os.path.join(ta_dir, "attacks")]
        sys.modules["torchattacks.attacks"] = att
        trainer = importlib.import_module("trainer")
        layers = importlib.import_module("layers")
    finally:
        os.chdir(old)
    assert trainer.__file__.startswith(DH_DIR), trainer.__file__
    return trainer, layers


def run_reference(trainer, L, pb, use_hints, **over):
    Trainer = trainer.Trainer
    o = dict(scales=list(pb.scales), v1_multiscale=False, height=pb.height, width=pb.width, min_depth=pb.min_depth,
             max_depth=pb.max_depth, frame_ids=list(pb.frame_ids), pose_model_type="separate_resnet",
             disable_automasking=False, no_ssim=False, adv_train=False, supervised_adv=False,
             contrastive_learning=False, no_original_train=False, avg_reprojection=False, predictive_mask=False,
             disparity_smoothness=1e-3, batch_size=pb.batch, use_depth_hints=use_hints)
    o.update(over)
    opt = SimpleNamespace(**o)
    me = SimpleNamespace(opt=opt, ssim=L.SSIM(), num_scales=len(opt.scales),
                         backproject_depth={0: L.BackprojectDepth(pb.batch, pb.height, pb.width)},
                         project_3d={0: L.Project3D(pb.batch, pb.height, pb.width)},
                         compute_proxy_supervised_loss=Trainer.compute_proxy_supervised_loss,
                         compute_loss_masks=Trainer.compute_loss_masks)
    me.compute_reprojection_loss = lambda pred, target: Trainer.compute_reprojection_loss(me, pred, target)
    inputs = {("K", 0): pb.K, ("inv_K", 0): pb.inv_K, "depth_hint": pb.extras["depth_hint"],
              "depth_hint_mask": pb.extras["depth_hint_mask"]}
    for (f, s), v in pb.color.items():
        inputs[("color", f, s)] = v
    if "s" in pb.T:
        inputs["stereo_T"] = pb.T["s"]
    disps = {s: pb.disp[s].clone().requires_grad_(True) for s in pb.scales}
    outputs = {("disp", s): disps[s] for s in pb.scales}
    for f in pb.frame_ids[1:]:
        if f != "s":
            outputs[("cam_T_cam", 0, f)] = pb.T[f]
    queue = [pb.noise[s][:, :1] / 0.00001 for s in pb.scales]      # trainer.py:688-690: ONE plane per scale
    real_randn, cuda_attr = torch.randn, torch.Tensor.cuda

    def injected(*a, **k):
        t = queue.pop(0)
        shape = tuple(a[0]) if len(a) == 1 and not isinstance(a[0], int) else tuple(a)
        assert tuple(t.shape) == shape, (t.shape, shape)
        return t.clone()

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.randn = injected
    try:
        Trainer.generate_images_pred(me, inputs, outputs)
        losses = Trainer.compute_losses(me, inputs, outputs)
    finally:
        torch.randn = real_randn
        torch.Tensor.cuda = cuda_attr
    losses["loss"].backward()
    return losses, outputs, disps


def main():
    trainer, L = load_dh_trainer()
    for name, (skw, use_hints, over) in CASES.items():
        pb = synth.photo_batch(**skw)
        losses, outputs, disps = run_reference(trainer, L, pb, use_hints, **over)
        out = {"loss": losses["loss"].detach().numpy()}
        for s in pb.scales:
            out["loss_%d" % s] = losses["loss/%d" % s].detach().numpy()
            out["reproj_loss_%d" % s] = losses["reproj_loss/%d" % s].detach().numpy()
            out["grad_disp_%d" % s] = disps[s].grad.numpy()
            out["ident_sel_%d" % s] = outputs["identity_selection/%d" % s].numpy().astype(np.uint8)
            if use_hints:
                out["depth_hint_loss_%d" % s] = losses["depth_hint_loss/%d" % s].detach().numpy()
                out["hint_pixels_%d" % s] = outputs["depth_hint_pixels/%d" % s].numpy().astype(np.uint8)
        if use_hints:
            out["color_depth_hint"] = outputs[("color_depth_hint", "s", 0)].detach().numpy()
        np.savez_compressed(os.path.join(GOLD, "dh_%s.npz" % name), **out)
        print("dh", name, float(out["loss"]), {k: float(v) for k, v in out.items() if k.startswith("depth_hint_loss")})


if __name__ == "__main__":
    main()

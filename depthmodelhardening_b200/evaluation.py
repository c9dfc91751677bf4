"""Attack evaluation harness on the grafted path (next-4): the device side of the reference's
`evaluate_attacks` / `compute_errors` (DepthNetworks/monodepth2/evaluate_depth.py:57-99, 113-214).

What the reference does per scene batch: run the chosen patch attack (`depth_atk(scene_img, batch_size, eval=True)`),
push the adversarial and the benign scenes through the depth network, convert both disparity maps to metric depth
(`clamp(disp_to_depth(|disp|, 0.1, 100)[1] * 5.4, 1e-3, 80)`), copy depths and masks to the host and reduce eight
error statistics over the pixels of the pasted object in numpy.  Here the attack classes are the CUDA drop-ins
(`attacks.py`), and the depth conversion + the eight masked reductions are ONE launch on the device
(`dmh_depth_errors`): nothing but nine doubles per batch crosses PCIe.  Data loading (the KITTI object loader, the car
image) stays with the caller: `evaluate_attacks` takes any iterable of scene batches.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional, Sequence

import numpy as np
import torch

from . import _lib, attacks
from ._lib import check, f32c, ptr, stream

STEREO_SCALE_FACTOR = 5.4          # evaluate_depth.py:44-46
MIN_DEPTH = 1e-3
MAX_DEPTH = 80
ERROR_NAMES = ("abs_err", "abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3")


def depth_error_sums(disp_gt: torch.Tensor, disp_atk: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The nine device sums of `dmh_depth_errors` for one batch (float64 tensor on the device, mask total first)."""
    g, a = f32c(disp_gt), f32c(disp_atk)
    if g.shape != a.shape:
        raise AssertionError("disparity maps differ in shape: %s vs %s" % (tuple(g.shape), tuple(a.shape)))
    m = None
    if mask is not None:
        m = f32c(mask)
        assert m.shape == g.shape and m.shape == a.shape          # compute_errors :78
    out = torch.empty(9, dtype=torch.float64, device=g.device)
    check(_lib.load().dmh_depth_errors(ptr(g), ptr(a), ptr(m), g.numel(), 0.1, 100.0, STEREO_SCALE_FACTOR, MIN_DEPTH,
                                       float(MAX_DEPTH), ptr(out), stream()), "depth_errors")
    return out


def errors_from_sums(sums) -> tuple:
    """compute_errors' return tuple (abs_err, abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3) from the nine sums."""
    s = [float(v) for v in (sums.tolist() if torch.is_tensor(sums) else sums)]
    n = s[0]
    return (s[1] / n, s[2] / n, s[3] / n, math.sqrt(s[4] / n), math.sqrt(s[5] / n), s[6] / n, s[7] / n, s[8] / n)


def compute_errors(disp_gt: torch.Tensor, disp_atk: torch.Tensor, mask: Optional[torch.Tensor] = None) -> tuple:
    """Drop-in for the pair evaluate_depth.py:193-196 + `compute_errors(gt_depth, atk_depth, mask)`: takes the two
    DISPARITY maps the network returned (CUDA tensors) and the resized object mask (or None)."""
    return errors_from_sums(depth_error_sums(disp_gt, disp_atk, mask))


def build_attack(model2atk, args: Dict, obj_tensor: torch.Tensor, mask_tensor: torch.Tensor):
    """The `norm_type` switch of evaluate_attacks (:119-155) over the drop-in classes."""
    nt = args["norm_type"]
    if nt == "l_inf":
        return attacks.Phy_obj_atk(model2atk, obj_tensor, mask_tensor, eps=args["epsilon"], alpha=args["alpha"],
                                   steps=args["step"])
    if nt == "l_0":
        return attacks.Phy_obj_atk_l0(model2atk, obj_tensor, mask_tensor, adam_lr=args["adam_lr"], steps=args["step"],
                                      mask_wt=args["mask_wt"], l0_thresh=args["l0_thresh"])
    if nt == "l_2":
        return attacks.Phy_obj_atk_l2(model2atk, obj_tensor, mask_tensor, eps=args["epsilon"], alpha=args["alpha"],
                                      steps=args["step"])
    if nt == "APGD":
        return attacks.Phy_obj_atk_APGD(model2atk, obj_tensor, mask_tensor, eps=args["epsilon"], steps=args["step"])
    if nt == "arbi":
        return attacks.Phy_obj_atk_arbi(model2atk, obj_tensor, mask_tensor)
    if nt == "guassian":
        return attacks.Phy_obj_atk_guassian(model2atk, obj_tensor, mask_tensor, steps=args["step"])
    if nt == "vanila":
        return attacks.Phy_obj_atk_vanila(model2atk, obj_tensor, mask_tensor)
    if nt == "Square":
        return attacks.Phy_obj_atk_Square(model2atk, obj_tensor, mask_tensor, eps=args["epsilon"],
                                          n_queries=args["n_queries"])
    if nt == "light":
        return attacks.Phy_obj_atk_light(model2atk, obj_tensor, mask_tensor, **args.get("light_kwargs", {}))
    raise NotImplementedError("norm_type %r: the whole-image PGD (`PGD_depth`, no patch, no PhysicalTrans) is not on "
                              "the grafted path" % (nt,))


def evaluate_attacks(model2atk, args: Dict, scenes: Iterable[torch.Tensor], obj_tensor: torch.Tensor,
                     mask_tensor: torch.Tensor, eval_count: int = 25, start_idx: int = 0, verbose: bool = True):
    """evaluate_attacks (:113-214) over an iterable of (B,3,375,1242) scene batches; returns (mean_errors,
    max_errors) as numpy arrays of the eight statistics.  The iterable is cycled like the reference's loader."""
    atk = build_attack(model2atk, args, obj_tensor, mask_tensor)
    vanila = attacks.Phy_obj_atk_vanila(model2atk, obj_tensor, mask_tensor)
    dev = obj_tensor.device
    per_batch = []
    it = iter(scenes)
    i = -1
    while True:
        try:
            scene_img = next(it)
        except StopIteration:
            it = iter(scenes)
            scene_img = next(it)
        i += 1
        if i < start_idx:
            continue
        if i - start_idx >= eval_count:
            break
        scene_img = scene_img.to(dev)
        if args["norm_type"] == "vanila":
            adv_images, ben_images, obj_masks_out, _ = vanila(scene_img, obj_tensor, args["batch_size"], eval=True)
        elif args["norm_type"] == "light" and i != start_idx:
            # evaluate_depth.py:178-182: the light search runs on the first batch only, its patch is then placed
            adv_images, ben_images, obj_masks_out, obj_img_adv = vanila(scene_img, obj_img_adv, args["batch_size"],
                                                                        eval=True)
        else:
            adv_images, ben_images, obj_masks_out, obj_img_adv = atk(scene_img, args["batch_size"], eval=True)
        with torch.no_grad():
            disp_gt = model2atk(ben_images)
            disp_atk = model2atk(adv_images)
        per_batch.append(depth_error_sums(disp_gt, disp_atk, obj_masks_out))     # stays on the device
    errors = np.array([errors_from_sums(s) for s in torch.stack(per_batch).cpu()])
    mean_errors, max_errors = errors.mean(0), errors.max(0)
    if verbose:
        for title, e in (("Mean Error:", mean_errors), ("Max Error:", max_errors)):
            print(title)
            print("\n  " + ("{:>8} | " * 8).format(*ERROR_NAMES))
            print(("&{: 8.3f}  " * 8).format(*e.tolist()) + "\\\\")
    return mean_errors, max_errors

"""`install()` rebinds the reference's hot-path symbols to the CUDA drop-ins so
that `train.py --adv_train`, `simple_adv_training.py` and
`physical_adv_training.py` run unchanged (the reference's
boundary for this path is module attribute substitution -- there is no FFI).

Call it BEFORE the reference's `trainer` module is imported (it star-imports
`layers` at import time, `M2/trainer.py:24`); modules that are already imported
are patched in place as well.

    import depthmodelhardening_b200.install as dmh
    dmh.install(mode="fused")      # or mode="ops": op-level drop-ins only
"""
from __future__ import annotations

import sys

from . import attacks as _attacks
from . import cost_volume as _cost_volume
from . import depth_hints as _depth_hints
from . import layers as _layers
from . import objective as _objective
from . import physical as _physical

_LAYER_SYMBOLS = ("BackprojectDepth", "Project3D", "SSIM", "get_smooth_loss", "disp_to_depth")


def _dispatching_physical_trans(ref_cls):
    """CUDA tensors -> kernel drop-in; CPU tensors (DataLoader workers,
    mono_dataset.py:163-168) -> the reference class, untouched."""

    def factory(obj_img, obj_mask, *a, **k):
        if getattr(obj_img, "is_cuda", False):
            return _physical.PhysicalTrans(obj_img, obj_mask, *a, **k)
        return ref_cls(obj_img, obj_mask, *a, **k)

    factory.__name__ = "PhysicalTrans"
    factory._dmh_reference = ref_cls
    return factory


def install(mode: str = "fused", dataset_root: str = None) -> dict:
    """Returns {symbol: patched?}.  `mode`: 'ops' (layers/PhysicalTrans/attacks)
    or 'fused' (additionally Trainer.generate_images_pred/compute_losses/compute_reprojection_loss)."""
    if mode not in ("ops", "fused"):
        raise ValueError(mode)
    done = {}
    # --- layers.*  (M2/DH/MD trainers, MD resnet_encoder, DH precompute star-import it)
    layers_mod = sys.modules.get("layers")
    if layers_mod is None:
        import importlib
        try:
            layers_mod = importlib.import_module("layers")
        except ImportError:
            layers_mod = None
    if layers_mod is not None:
        for name in _LAYER_SYMBOLS:
            if not hasattr(layers_mod, "_dmh_ref_" + name):
                setattr(layers_mod, "_dmh_ref_" + name, getattr(layers_mod, name))
            setattr(layers_mod, name, getattr(_layers, name))
            done["layers." + name] = True
    # modules that already star-imported layers
    for mod_name in ("trainer", "trainer_contras"):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            for name in _LAYER_SYMBOLS:
                if hasattr(mod, name):
                    setattr(mod, name, getattr(_layers, name))
                    done["%s.%s" % (mod_name, name)] = True
    # --- physicalTrans.PhysicalTrans
    pt_mod = sys.modules.get("physicalTrans")
    if pt_mod is not None:
        ref_cls = getattr(pt_mod.PhysicalTrans, "_dmh_reference", pt_mod.PhysicalTrans)
        pt_mod.PhysicalTrans = _dispatching_physical_trans(ref_cls)
        done["physicalTrans.PhysicalTrans"] = True
    # --- torchattacks.Phy_obj_atk / Phy_obj_atk_l0
    if dataset_root is None:
        mu = sys.modules.get("my_utils")
        dataset_root = getattr(mu, "object_dataset_root", None)
    if dataset_root is not None:
        _attacks.object_dataset_root = dataset_root
    for mod_name, cls in (("torchattacks.attacks.phy_obj_atk", "Phy_obj_atk"),
                          ("torchattacks.attacks.phy_obj_atk_l0", "Phy_obj_atk_l0"),
                          ("torchattacks.attacks.phy_obj_atk_vanila", "Phy_obj_atk_vanila"),
                          ("torchattacks.attacks.phy_obj_atk_l2", "Phy_obj_atk_l2"), ("torchattacks", "Phy_obj_atk"),
                          ("torchattacks", "Phy_obj_atk_l0"), ("torchattacks", "Phy_obj_atk_vanila"),
                          ("torchattacks", "Phy_obj_atk_l2")):
        mod = sys.modules.get(mod_name)
        if mod is not None and hasattr(mod, cls):
            setattr(mod, cls, getattr(_attacks, cls))
            done["%s.%s" % (mod_name, cls)] = True
    # --- Trainer methods (fused fast path)
    for mod_name in ("trainer", "trainer_contras"):
        mod = sys.modules.get(mod_name)
        if mod is None or not hasattr(mod, "Trainer"):
            continue
        T = mod.Trainer
        if not hasattr(T, "_dmh_ref_compute_losses"):
            T._dmh_ref_compute_losses = T.compute_losses
            T._dmh_ref_generate_images_pred = T.generate_images_pred
            T._dmh_ref_compute_reprojection_loss = T.compute_reprojection_loss
        T.compute_reprojection_loss = _layers.compute_reprojection_loss
        done[mod_name + ".Trainer.compute_reprojection_loss"] = True
        if mode == "fused":
            if hasattr(T, "compute_loss_masks"):
                # the depth-hints trainer (DH/trainer.py:541): masked-mean objective + depth-hint terms
                T.generate_images_pred = _depth_hints.dh_generate_images_pred
                T.compute_losses = _depth_hints.dh_compute_losses
                done[mod_name + ".Trainer.compute_losses(depth-hints)"] = True
            else:
                T.generate_images_pred = _objective.fused_generate_images_pred
                T.compute_losses = _objective.fused_compute_losses
            done[mod_name + ".Trainer.compute_losses"] = True
    # --- ManyDepth cost volume (MD/networks/resnet_encoder.py:157): ResnetEncoderMatching.match_features
    for mod_name, mod in list(sys.modules.items()):
        if mod is not None and mod_name.endswith("resnet_encoder") and hasattr(mod, "ResnetEncoderMatching"):
            cls = mod.ResnetEncoderMatching
            if not hasattr(cls, "_dmh_ref_match_features"):
                cls._dmh_ref_match_features = cls.match_features
            cls.match_features = _cost_volume.match_features
            done[mod_name + ".ResnetEncoderMatching.match_features"] = True
    return done


def uninstall() -> None:
    for mod_name, mod in list(sys.modules.items()):
        if mod is not None and mod_name.endswith("resnet_encoder") and hasattr(mod, "ResnetEncoderMatching"):
            cls = mod.ResnetEncoderMatching
            if hasattr(cls, "_dmh_ref_match_features"):
                cls.match_features = cls._dmh_ref_match_features
    layers_mod = sys.modules.get("layers")
    if layers_mod is not None:
        for name in _LAYER_SYMBOLS:
            if hasattr(layers_mod, "_dmh_ref_" + name):
                setattr(layers_mod, name, getattr(layers_mod, "_dmh_ref_" + name))
    pt_mod = sys.modules.get("physicalTrans")
    if pt_mod is not None and hasattr(pt_mod.PhysicalTrans, "_dmh_reference"):
        pt_mod.PhysicalTrans = pt_mod.PhysicalTrans._dmh_reference
    for mod_name in ("trainer", "trainer_contras"):
        mod = sys.modules.get(mod_name)
        if mod is not None and hasattr(mod, "Trainer") and hasattr(mod.Trainer, "_dmh_ref_compute_losses"):
            T = mod.Trainer
            T.compute_losses = T._dmh_ref_compute_losses
            T.generate_images_pred = T._dmh_ref_generate_images_pred
            T.compute_reprojection_loss = T._dmh_ref_compute_reprojection_loss

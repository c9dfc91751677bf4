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


# every (object, attribute, original value) install() has replaced, in patch order: uninstall() is the exact inverse
_SAVED = []
_MISSING = object()


def _patch(obj, name, new):
    """setattr that remembers the original the first time (obj, name) is patched."""
    if not any(o is obj and n == name for o, n, _ in _SAVED):
        _SAVED.append((obj, name, obj.__dict__.get(name, _MISSING) if isinstance(obj, type) else
                       getattr(obj, name, _MISSING)))
    setattr(obj, name, new)


def install(mode: str = "fused", dataset_root: str = None) -> dict:
    """Returns {symbol: patched?}.  `mode`: 'ops' (layers/PhysicalTrans/attacks)
    or 'fused' (additionally Trainer.generate_images_pred/compute_losses/compute_reprojection_loss).
    `uninstall()` restores every symbol replaced here (A/B runs against the reference)."""
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
                _patch(layers_mod, "_dmh_ref_" + name, getattr(layers_mod, name))
            _patch(layers_mod, name, getattr(_layers, name))
            done["layers." + name] = True
    # modules that already star-imported layers (trainers, the ManyDepth encoder, ...)
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or mod is layers_mod or not (mod_name in ("trainer", "trainer_contras") or
                                                    mod_name.endswith("resnet_encoder")):
            continue
        for name in _LAYER_SYMBOLS:
            if hasattr(mod, name):
                _patch(mod, name, getattr(_layers, name))
                done["%s.%s" % (mod_name, name)] = True
    # --- physicalTrans.PhysicalTrans
    pt_mod = sys.modules.get("physicalTrans")
    if pt_mod is not None:
        ref_cls = getattr(pt_mod.PhysicalTrans, "_dmh_reference", pt_mod.PhysicalTrans)
        _patch(pt_mod, "PhysicalTrans", _dispatching_physical_trans(ref_cls))
        done["physicalTrans.PhysicalTrans"] = True
    # --- torchattacks.Phy_obj_atk / Phy_obj_atk_l0
    if dataset_root is None:
        mu = sys.modules.get("my_utils")
        dataset_root = getattr(mu, "object_dataset_root", None)
    if dataset_root is not None:
        _patch(_attacks, "object_dataset_root", dataset_root)
    for mod_name, cls in (("torchattacks.attacks.phy_obj_atk", "Phy_obj_atk"),
                          ("torchattacks.attacks.phy_obj_atk_l0", "Phy_obj_atk_l0"),
                          ("torchattacks.attacks.phy_obj_atk_vanila", "Phy_obj_atk_vanila"),
                          ("torchattacks.attacks.phy_obj_atk_l2", "Phy_obj_atk_l2"),
                          ("torchattacks.attacks.phy_obj_atk_apgd", "Phy_obj_atk_APGD"),
                          ("torchattacks.attacks.phy_obj_atk_guassian", "Phy_obj_atk_guassian"),
                          ("torchattacks.attacks.phy_obj_atk_arbi", "Phy_obj_atk_arbi"),
                          ("torchattacks.attacks.phy_obj_atk_square", "Phy_obj_atk_Square"),
                          ("torchattacks.attacks.phy_obj_atk_light", "Phy_obj_atk_light"),
                          ("torchattacks", "Phy_obj_atk_Square"), ("torchattacks", "Phy_obj_atk_light"),
                          ("torchattacks", "Phy_obj_atk_arbi"),
                          ("torchattacks", "Phy_obj_atk_guassian"),
                          ("torchattacks", "Phy_obj_atk_APGD"), ("torchattacks", "Phy_obj_atk"),
                          ("torchattacks", "Phy_obj_atk_l0"), ("torchattacks", "Phy_obj_atk_vanila"),
                          ("torchattacks", "Phy_obj_atk_l2")):
        mod = sys.modules.get(mod_name)
        if mod is not None and hasattr(mod, cls):
            _patch(mod, cls, getattr(_attacks, cls))
            done["%s.%s" % (mod_name, cls)] = True
    # --- Trainer methods (fused fast path)
    for mod_name in ("trainer", "trainer_contras"):
        mod = sys.modules.get(mod_name)
        if mod is None or not hasattr(mod, "Trainer"):
            continue
        T = mod.Trainer
        if not hasattr(T, "_dmh_ref_compute_losses"):
            _patch(T, "_dmh_ref_compute_losses", T.compute_losses)
            _patch(T, "_dmh_ref_generate_images_pred", T.generate_images_pred)
            _patch(T, "_dmh_ref_compute_reprojection_loss", T.compute_reprojection_loss)
        _patch(T, "compute_reprojection_loss", _layers.compute_reprojection_loss)
        done[mod_name + ".Trainer.compute_reprojection_loss"] = True
        if mode == "fused":
            if hasattr(T, "compute_loss_masks"):
                # the depth-hints trainer (DH/trainer.py:541): masked-mean objective + depth-hint terms
                _patch(T, "generate_images_pred", _depth_hints.dh_generate_images_pred)
                _patch(T, "compute_losses", _depth_hints.dh_compute_losses)
                done[mod_name + ".Trainer.compute_losses(depth-hints)"] = True
            else:
                _patch(T, "generate_images_pred", _objective.fused_generate_images_pred)
                _patch(T, "compute_losses", _objective.fused_compute_losses)
            done[mod_name + ".Trainer.compute_losses"] = True
    # --- ManyDepth cost volume (MD/networks/resnet_encoder.py:157): ResnetEncoderMatching.match_features
    _patch_matching_encoders(done)
    _install_import_hook()
    return done


def _patch_matching_encoders(done=None):
    for mod_name, mod in list(sys.modules.items()):
        if mod is not None and mod_name.endswith("resnet_encoder") and hasattr(mod, "ResnetEncoderMatching"):
            cls = mod.ResnetEncoderMatching
            if cls.__dict__.get("match_features") is _cost_volume.match_features:
                continue
            if not hasattr(cls, "_dmh_ref_match_features"):
                _patch(cls, "_dmh_ref_match_features", cls.match_features)
            _patch(cls, "match_features", _cost_volume.match_features)
            if done is not None:
                done[mod_name + ".ResnetEncoderMatching.match_features"] = True


class _EncoderImportHook:
    """install() is documented to run BEFORE the trainer is imported; `resnet_encoder` (ManyDepth) is then not in
    sys.modules yet, and its `from layers import BackprojectDepth, Project3D` would bind the drop-ins while
    `match_features` -- which calls them with batch_size = 96 depth bins and (1,4,4) matrices -- stayed unpatched.
    This meta-path finder patches the encoder class right after such a module has been executed."""

    def find_spec(self, fullname, path=None, target=None):
        if not fullname.endswith("resnet_encoder"):
            return None
        import importlib.machinery
        import importlib.util
        for finder in sys.meta_path:
            if finder is self or not hasattr(finder, "find_spec"):
                continue
            spec = finder.find_spec(fullname, path, target)
            if spec is not None and spec.loader is not None and hasattr(spec.loader, "exec_module"):
                loader = spec.loader
                orig_exec = loader.exec_module

                def exec_module(module, _orig=orig_exec):
                    _orig(module)
                    if _HOOK[0] is not None:
                        _patch_matching_encoders()
                try:
                    loader.exec_module = exec_module
                except AttributeError:
                    return None
                return spec
        return None


_HOOK = [None]


def _install_import_hook():
    if _HOOK[0] is None:
        _HOOK[0] = _EncoderImportHook()
        sys.meta_path.insert(0, _HOOK[0])


def uninstall() -> None:
    """Exact inverse of install(): every replaced attribute gets its original value back (attributes install()
    created are deleted), in reverse order; the import hook is removed."""
    if _HOOK[0] is not None:
        try:
            sys.meta_path.remove(_HOOK[0])
        except ValueError:
            pass
        _HOOK[0] = None
    while _SAVED:
        obj, name, orig = _SAVED.pop()
        if orig is _MISSING:
            try:
                delattr(obj, name)
            except AttributeError:
                pass
        else:
            setattr(obj, name, orig)

"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Imports the *unmodified* reference (`/root/reference`, present only in the
build container, never on the GPU box) so that `oracle/make_golden.py` can
generate the golden vectors under `tests/golden/` (`tests/test_oracle_golden.py`
pins the oracle restatement to them).  On the GPU box the copy made by
`oracle/build_ref.py` (`oracle/_ref/reference`, git-ignored) is loaded instead:
`oracle/ref_step.py` runs the reference's own code there for the bench's
reference arm, its eager-PyTorch-on-B200 baseline and the integration tests.

The reference needs six import-level stubs; none of
them touches arithmetic:
  1. fake matplotlib / matplotlib.pyplot / matplotlib.figure   (my_utils.py:3-5)
  2. fake numpy.lib.utils                                      (phy_obj_atk_l0.py:3)
  3. empty `torchattacks` package shell so TA/__init__.py is not executed
  4. fake tensorboardX.SummaryWriter                           (M2/trainer.py:16)
  5. fake skimage.transform + PIL.Image.ANTIALIAS              (kitti_dataset.py:10)
  6. my_utils.object_dataset_root -> temp dir with a synthetic calib file
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# the reference where it lies (build container), else the copy `oracle/build_ref.py` shipped to the GPU box
_SHIPPED = os.path.join(_HERE, "_ref", "reference")
REF_ROOT = os.environ.get("DMH_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference") else _SHIPPED)
M2_DIR = os.path.join(REF_ROOT, "DepthNetworks", "monodepth2")

# KITTI object calib 003086 values as printed in physicalTrans.py:208-213.
CALIB_P2 = (7.215377e+02, 0.0, 6.095593e+02, 4.485728e+01,
            0.0, 7.215377e+02, 1.728540e+02, 2.163791e-01,
            0.0, 0.0, 1.0, 2.745884e-03)


def available() -> bool:
    return os.path.isdir(M2_DIR)


def write_calib(root: str) -> str:
    """Write a synthetic KITTI-object calib file; returns its path."""
    d = os.path.join(root, "training", "calib")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "003086.txt")
    ident34 = "1 0 0 0 0 1 0 0 0 0 1 0"
    with open(path, "w") as f:
        f.write("P0: " + " ".join(repr(v) for v in CALIB_P2) + "\n")
        f.write("P1: " + " ".join(repr(v) for v in CALIB_P2) + "\n")
        f.write("P2: " + " ".join(repr(v) for v in CALIB_P2) + "\n")
        f.write("P3: " + " ".join(repr(v) for v in CALIB_P2) + "\n")
        f.write("R0_rect: 1 0 0 0 1 0 0 0 1\n")
        f.write("Tr_velo_to_cam: " + ident34 + "\n")
        f.write("Tr_imu_to_velo: " + ident34 + "\n")
    return path


def _fake(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_LOADED: dict = {}


def load():
    """Return a namespace with the reference modules (imported once)."""
    if _LOADED:
        return types.SimpleNamespace(**_LOADED)
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)

    # (1) matplotlib
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = _fake("matplotlib")
            plt = _fake("matplotlib.pyplot", axes=None, axis=None, get=None)
            fig = _fake("matplotlib.figure", Figure=object)
            mpl.pyplot, mpl.figure = plt, fig
    # (2) numpy.lib.utils (removed in numpy 2) and numpy.core.numeric.zeros_like
    import numpy as np
    if "numpy.lib.utils" not in sys.modules:
        try:
            importlib.import_module("numpy.lib.utils")
        except Exception:
            np.lib.utils = _fake("numpy.lib.utils")
    # (4) tensorboardX
    if "tensorboardX" not in sys.modules:
        try:
            import tensorboardX  # noqa: F401
        except Exception:
            _fake("tensorboardX", SummaryWriter=object)
    # (5) skimage + PIL.Image.ANTIALIAS
    if "skimage" not in sys.modules:
        try:
            import skimage.transform  # noqa: F401
        except Exception:
            sk = _fake("skimage")
            sk.transform = _fake("skimage.transform", resize=None)
    import PIL.Image
    if not hasattr(PIL.Image, "ANTIALIAS"):
        PIL.Image.ANTIALIAS = PIL.Image.LANCZOS

    for p in (REF_ROOT, M2_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the reference uses relative sys.path.append("../.."): cwd must be M2/
    old_cwd = os.getcwd()
    os.chdir(M2_DIR)
    try:
        # (6) dataset root with a synthetic calib
        tmp = tempfile.mkdtemp(prefix="dmh_ref_calib_")
        write_calib(tmp)
        my_utils = importlib.import_module("my_utils")
        my_utils.object_dataset_root = tmp
        # (3) torchattacks shell (do not execute TA/__init__.py)
        ta_dir = os.path.join(REF_ROOT, "torchattacks")
        ta = types.ModuleType("torchattacks")
        ta.__path__ = [ta_dir]
        sys.modules["torchattacks"] = ta
        ta_att = types.ModuleType("torchattacks.attacks")
        ta_att.__path__ = [os.path.join(ta_dir, "attacks")]
        sys.modules["torchattacks.attacks"] = ta_att

        layers = importlib.import_module("layers")
        physicalTrans = importlib.import_module("physicalTrans")
        atk_linf = importlib.import_module("torchattacks.attacks.phy_obj_atk")
        atk_l0 = importlib.import_module("torchattacks.attacks.phy_obj_atk_l0")
        trainer = importlib.import_module("trainer")
        depth_model = importlib.import_module("depth_model")
        networks = importlib.import_module("networks")
    finally:
        os.chdir(old_cwd)

    _LOADED.update(dict(layers=layers, physicalTrans=physicalTrans, atk_linf=atk_linf,
                        atk_l0=atk_l0, trainer=trainer, depth_model=depth_model,
                        networks=networks, my_utils=my_utils, calib_root=tmp,
                        calib_path=os.path.join(tmp, "training", "calib", "003086.txt")))
    return types.SimpleNamespace(**_LOADED)

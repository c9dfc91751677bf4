"""ManyDepth cost volume.
CPU: oracle restatement vs the golden generated from the UNMODIFIED reference method.
GPU: dmh_cost_volume (through the C ABI) vs oracle and golden; 1e-5 relative, missing masks exact."""
import numpy as np
import pytest
import torch

from oracle import cost_volume as OC
from depthmodelhardening_b200.synth import cost_volume_inputs
from tests.util import assert_close, load_golden

TOL = 1e-5


def test_oracle_matches_reference_golden():
    g = load_golden("md_cost_volume")
    cur, look, poses, K, inv_K, bins = cost_volume_inputs()
    vol, miss = OC.match_features(cur, look, poses, K, inv_K, bins, True)
    assert_close(vol, g["cost_volume"], TOL, "cost volume")
    assert np.array_equal(miss.numpy(), g["missing"])
    vol_raw, _ = OC.match_features(cur, look, poses, K, inv_K, bins, False)
    assert_close(vol_raw, g["cost_volume_raw"], TOL, "cost volume (no fill)")
    assert 0.05 < float(miss.mean()) < 0.95          # the case exercises both branches


@pytest.fixture(scope="module")
def dev():
    from depthmodelhardening_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


@pytest.mark.gpu
def test_cuda_matches_oracle_and_golden(dev):
    from depthmodelhardening_b200 import cost_volume as CV
    g = load_golden("md_cost_volume")
    cur, look, poses, K, inv_K, bins = cost_volume_inputs()
    mv = lambda t: t.to(dev)
    for fill, key in ((True, "cost_volume"), (False, "cost_volume_raw")):
        vol, miss = CV.cost_volume(mv(cur), mv(look), mv(poses), mv(K), mv(inv_K), bins, fill)
        ref_vol, ref_miss = OC.match_features(cur, look, poses, K, inv_K, bins, fill)
        assert np.array_equal(miss.cpu().numpy(), ref_miss.numpy())
        assert np.array_equal(miss.cpu().numpy(), g["missing"])
        assert_close(vol, ref_vol, TOL, "cost volume vs oracle")
        assert_close(vol, g[key], TOL, "cost volume vs golden")


@pytest.mark.gpu
def test_cuda_dropin_method_and_edge_cases(dev):
    from types import SimpleNamespace
    from depthmodelhardening_b200 import cost_volume as CV
    cur, look, poses, K, inv_K, bins = cost_volume_inputs(B=1, L=1, h=16, w=20, D=5, seed=81)
    me = SimpleNamespace(depth_bins=bins, set_missing_to_max=True)
    vol, miss = CV.match_features(me, cur.to(dev), look.to(dev), poses.to(dev), K.to(dev), inv_K.to(dev))
    ref_vol, ref_miss = OC.match_features(cur, look, poses, K, inv_K, bins, True)
    assert_close(vol, ref_vol, TOL, "drop-in")
    # every lookup frame missing (start of a sequence): all-zero volume, everything flagged
    zero = torch.zeros_like(poses).to(dev)
    vol0, miss0 = CV.match_features(me, cur.to(dev), look.to(dev), zero, K.to(dev), inv_K.to(dev))
    assert float(vol0.abs().max()) == 0.0 and float(miss0.min()) == 1.0
    with pytest.raises(RuntimeError):                     # no CPU fallback
        CV.match_features(me, cur, look, poses, K, inv_K)
    with pytest.raises(RuntimeError):                     # shape contract
        CV.cost_volume(cur.to(dev), look[:, :, :8].to(dev), poses.to(dev), K.to(dev), inv_K.to(dev), bins)


@pytest.mark.gpu
def test_cuda_full_size(dev):
    """BASELINE config 5 matching resolution (1024x320 -> 80x256, 96 bins, 16 channels, B=4, 2 lookups):
    deterministic, identical lookup == current under the identity pose gives an all-zero (all-missing) volume,
    finite everywhere."""
    from depthmodelhardening_b200 import cost_volume as CV
    cur, look, poses, K, inv_K, bins = cost_volume_inputs(B=4, L=2, h=80, w=256, D=96, seed=82)
    mv = lambda t: t.to(dev)
    a, ma = CV.cost_volume(mv(cur), mv(look), mv(poses), mv(K), mv(inv_K), bins)
    b, mb = CV.cost_volume(mv(cur), mv(look), mv(poses), mv(K), mv(inv_K), bins)
    assert torch.equal(a, b) and torch.equal(ma, mb) and torch.isfinite(a).all()
    same = cur.unsqueeze(1).repeat(1, 2, 1, 1, 1).contiguous()
    eye = torch.eye(4).view(1, 1, 4, 4).repeat(4, 2, 1, 1).contiguous()
    v, m = CV.cost_volume(mv(cur), mv(same), mv(eye), mv(K), mv(inv_K), bins, set_missing_to_max=False)
    assert float(v.abs().max()) < 1e-5


@pytest.mark.reference
def test_install_patches_manydepth_encoder():
    """Own process: install() binds the drop-in onto the reference's ResnetEncoderMatching."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from oracle.make_golden_md import load_md\n"
        "layers, enc = load_md()\n"
        "import depthmodelhardening_b200.install as dmh\n"
        "done = dmh.install(mode='ops')\n"
        "assert enc.ResnetEncoderMatching.match_features.__module__ == 'depthmodelhardening_b200.cost_volume'\n"
        "assert any(k.endswith('ResnetEncoderMatching.match_features') for k in done)\n"
        "dmh.uninstall()\n"
        "assert enc.ResnetEncoderMatching.match_features.__module__.endswith('resnet_encoder')\n"
        "print('ok')\n") % root
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.reference
def test_install_before_import_patches_manydepth_encoder():
    """The documented order -- install() BEFORE the trainer / encoder modules are imported: `resnet_encoder` is not
    in sys.modules yet, its `from layers import BackprojectDepth, Project3D` then binds the drop-ins, and
    `match_features` (which hands them (1,4,4) matrices with batch_size = 96 depth bins,
    MD/networks/resnet_encoder.py:182,194) must be patched by the import hook; uninstall() restores everything."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, os, importlib; sys.path.insert(0, %r)\n"
        "from oracle import make_golden_md as M, refload\n"
        "for p in (refload.REF_ROOT, M.MD_DIR): sys.path.insert(0, p)\n"
        "os.chdir(M.MD_DIR)\n"
        "layers = importlib.import_module('layers')\n"
        "ref_bp = layers.BackprojectDepth\n"
        "import depthmodelhardening_b200.install as dmh\n"
        "done = dmh.install(mode='ops')\n"
        "assert 'networks.resnet_encoder' not in sys.modules\n"
        "enc = importlib.import_module('networks.resnet_encoder')\n"
        "assert enc.BackprojectDepth.__module__ == 'depthmodelhardening_b200.layers'\n"
        "assert enc.ResnetEncoderMatching.match_features.__module__ == 'depthmodelhardening_b200.cost_volume', enc.ResnetEncoderMatching.match_features.__module__\n"
        "dmh.uninstall()\n"
        "assert enc.ResnetEncoderMatching.match_features.__module__.endswith('resnet_encoder')\n"
        "assert layers.BackprojectDepth is ref_bp and not hasattr(layers, '_dmh_ref_SSIM')\n"
        "print('ok')\n") % root
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_camera_matrix_batch_is_validated():
    """ops._mat_batch: a batch-1 matrix broadcasts (ManyDepth passes (1,4,4) with batch_size = num_depth_bins), any
    other mismatch raises instead of letting the kernel index past the buffer (`K + b*16`)."""
    import torch
    from depthmodelhardening_b200 import ops
    K = torch.eye(4).unsqueeze(0)
    assert tuple(ops._mat_batch(K, 96, "K").shape) == (96, 4, 4)
    assert tuple(ops._mat_batch(torch.eye(4), 3, "K").shape) == (3, 4, 4)
    assert ops._mat_batch(K.repeat(5, 1, 1), 5, "K").shape[0] == 5
    with pytest.raises(RuntimeError):
        ops._mat_batch(K.repeat(2, 1, 1), 96, "K")
    with pytest.raises(RuntimeError):
        ops._mat_batch(torch.zeros(1, 3, 4), 1, "K")

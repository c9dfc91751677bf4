"""CPU: the reference-step harness (oracle/ref_step.py, bench.py's reference arm / eager baseline and the GPU
integration tests run on it) is pinned to the committed goldens, and the evaluation-metric post-processing."""
import math

import numpy as np
import pytest
import torch

from depthmodelhardening_b200 import synth
from oracle.make_golden import PHOTO_CASES
from tests.util import load_golden


@pytest.mark.reference
@pytest.mark.parametrize("name", ["stereo_small", "mono_small"])
def test_ref_step_reproduces_the_reference_golden(name):
    """Stage2Reference = Trainer.generate_images_pred + compute_losses + backward of the UNMODIFIED reference, unbound
    on the attribute namespace: must reproduce the golden oracle/make_golden.py wrote from the same code, exactly."""
    from oracle import ref_step
    skw, over = PHOTO_CASES[name]
    pb = synth.photo_batch(**skw)
    g = load_golden("photo_" + name)
    s2 = ref_step.Stage2Reference(pb, "cpu", inject_noise=True, **over)
    losses = s2.step()
    assert float(losses["loss"]) == float(g["loss"])
    for s in pb.scales:
        assert np.array_equal(s2.disps[s].grad.numpy(), g["grad_disp_%d" % s])


@pytest.mark.reference
def test_ref_stage1_step_matches_the_oracle_restatement():
    """Stage1Reference (the reference's PhysicalTrans + the loop body of phy_obj_atk_l0.py:94-138) against the oracle
    restatement of the same iteration: composited scenes to 1e-6, updated patterns equal."""
    from oracle import patch as OQ
    from oracle import ref_step
    from oracle.refload import CALIB_P2
    pt = synth.patch_batch(batch=2, seed=5)
    s1 = ref_step.Stage1Reference(pt, "cpu")
    scenes = s1.step()
    P34 = np.array(CALIB_P2, dtype=np.float64).reshape(3, 4)
    pp = pt.pattern_pos.clone().requires_grad_(True)
    pn = pt.pattern_neg.clone().requires_grad_(True)
    opt = torch.optim.Adam([pp, pn], lr=0.5, betas=(0.5, 0.9))
    adv, pos, neg = OQ.l0_compose(pt.obj, pp, pn)
    sc, _ = OQ.apply_patch(adv, pt.mask, pt.scenes, pt.z0, pt.alpha, P34)
    cost = (sc * pt.upstream).sum() + 0.06 * OQ.l0_mask_cost(pp, pn)
    opt.zero_grad()
    cost.backward()
    opt.step()
    assert float((scenes - sc).abs().max()) < 1e-6
    assert float((s1.pp - pp).abs().max()) < 1e-6 and float((s1.pn - pn).abs().max()) < 1e-6


def test_errors_from_sums():
    from depthmodelhardening_b200 import evaluation
    sums = [4.0, 2.0, 1.0, 0.5, 16.0, 0.36, 3.0, 4.0, 4.0]
    e = evaluation.errors_from_sums(sums)
    assert e == (0.5, 0.25, 0.125, 2.0, math.sqrt(0.09), 0.75, 1.0, 1.0)
    assert evaluation.ERROR_NAMES[3] == "rmse" and len(evaluation.ERROR_NAMES) == 8

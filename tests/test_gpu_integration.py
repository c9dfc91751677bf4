"""GPU integration of the Trainer-method drop-ins: the symbols `train.py --adv_train` actually hits after
`install(mode="fused")` are EXECUTED here (not merely checked for their names):

  * `objective.fused_generate_images_pred` + `fused_compute_losses` called the way `Trainer.process_batch` calls the
    methods they replace (DepthNetworks/monodepth2/trainer.py:335-375 -> :472-523, :539-674), on a namespace that
    carries the attributes `Trainer.__init__` sets, against the goldens generated from the unmodified reference --
    plain, with `_dmh_materialise`, and with the `supervised_adv` (:545-563), `contrastive_learning` (:568-575) and
    `no_original_train` (:577-579) branches on;
  * the same through `install()` on the REFERENCE'S OWN `Trainer` class (the copy `oracle/build_ref.py` ships to the
    GPU box): reference methods on cuda (stock eager PyTorch) vs the patched methods, same inputs, and `uninstall()`
    restoring every symbol (round-trip).
"""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from depthmodelhardening_b200 import synth
from oracle.make_golden import PHOTO_CASES
from tests.util import assert_close, assert_close_arb, assert_grad_close, load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    from depthmodelhardening_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


def _opts(pb, **over):
    o = dict(scales=list(pb.scales), v1_multiscale=False, height=pb.height, width=pb.width,
             min_depth=pb.min_depth, max_depth=pb.max_depth, frame_ids=list(pb.frame_ids),
             pose_model_type="separate_resnet", disable_automasking=False, no_ssim=False, adv_train=False,
             supervised_adv=False, contrastive_learning=False, no_original_train=False, avg_reprojection=False,
             predictive_mask=False, disparity_smoothness=1e-3, batch_size=pb.batch, gt_depth=False)
    o.update(over)
    return SimpleNamespace(**o)


def _inputs_outputs(pb):
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
    return inputs, outputs, disps


def _noise(pb, over):
    n_src = len(pb.frame_ids) - 1
    n_ident = 1 if over.get("avg_reprojection") else n_src
    return {s: pb.noise[s][:, :n_ident].contiguous() for s in pb.scales}


@pytest.mark.parametrize("name", sorted(PHOTO_CASES))
@pytest.mark.parametrize("materialise", [False, True])
def test_trainer_dropins_executed_vs_reference_golden(dev, name, materialise):
    """fused_generate_images_pred + fused_compute_losses as `Trainer.process_batch` calls them, vs the golden the
    unmodified `Trainer.generate_images_pred / compute_losses` produced (oracle/make_golden.py)."""
    from depthmodelhardening_b200 import objective
    from oracle import photometric as OP
    skw, over = PHOTO_CASES[name]
    g = load_golden("photo_" + name)
    pb_cpu = synth.photo_batch(**skw)
    _, _, g64 = OP.objective_from_batch(pb_cpu, OP.default_opts(scales=list(pb_cpu.scales), **over), dtype=torch.float64)
    pb = pb_cpu.to(dev)
    me = SimpleNamespace(opt=_opts(pb, **over), num_scales=len(pb.scales))
    me._dmh_noise = pb.noise
    me._dmh_materialise = materialise
    inputs, outputs, disps = _inputs_outputs(pb)
    objective.fused_generate_images_pred(me, inputs, outputs)
    losses = objective.fused_compute_losses(me, inputs, outputs)
    losses["loss"].backward()
    assert_close(losses["loss"], g["loss"], TOL, "loss")
    for s in pb.scales:
        assert_close(losses["loss/%d" % s], g["loss_%d" % s], TOL, "loss/%d" % s)
        assert_grad_close(disps[s].grad, g["grad_disp_%d" % s], g64[s], TOL, "grad_disp_%d" % s,
                          outlier_frac=5e-3 if name == "stereo_iid" else (4e-3 if name == "mono_small" else 2e-3))
    assert_close(outputs[("depth", 0, 0)], g["depth_0"], TOL, "depth_0")
    if materialise:
        for f in pb.frame_ids[1:]:
            assert_close(outputs[("color", f, 0)], g["warped_%s_0" % f], TOL, "warped %s" % f, max_outlier_frac=1e-3,
                         outlier_rtol=1.0)
            assert_close(outputs[("sample", f, 0)], g["grid_%s_0" % f], TOL, "grid %s" % f)
        if not over.get("disable_automasking") and "ident_sel_0" in g and "identity_selection/0" in outputs:
            sel = outputs["identity_selection/0"].cpu().numpy().astype(np.uint8)
            assert float(np.mean(sel != g["ident_sel_0"])) < 1e-3


class _TinyNet(torch.nn.Module):
    def __init__(self, seed):
        super().__init__()
        self.c = torch.nn.Conv2d(3, 1, 3, padding=1)
        gen = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=gen) * 0.2)

    def forward(self, x):
        return torch.sigmoid(self.c(x))


class _Contrast(torch.nn.Module):
    def forward(self, a, b):
        return ((a - b) ** 2).mean()


def _adv_branches_reference(me, inputs, outputs, base_losses):
    """trainer.py:545-579 restated in three lines for the expected value (the branch bodies are network-side torch
    ops, outside the graft; what is under test is that the drop-in routes through them and adds them up)."""
    tot = 0
    exp = {}
    if me.opt.supervised_adv:
        with torch.no_grad():
            dgt = me.gt_model(inputs[("color_ben", 0, 0)])
        exp["sup_loss"] = me.sup_loss_creteria(dgt, outputs[("disp", 0)])
        tot = tot + exp["sup_loss"]
    if me.opt.contrastive_learning:
        exp["contras_loss"] = me.models["contrastive_learning"](outputs["middle_features_aug"], outputs["middle_features_ben"])
        tot = tot + exp["contras_loss"]
    exp["loss"] = tot if me.opt.no_original_train else tot + base_losses["loss"]
    return exp


@pytest.mark.parametrize("sup,con,no_orig", [(True, False, False), (False, True, False), (True, True, False),
                                             (True, True, True)])
def test_trainer_dropin_adversarial_branches(dev, sup, con, no_orig):
    """`--adv_train` with `--supervised_adv` / `--contrastive_learning` / `--no_original_train`: the branches of
    compute_losses that ride on top of the photometric objective (trainer.py:545-579) are executed by the drop-in."""
    from depthmodelhardening_b200 import objective
    skw, over = PHOTO_CASES["stereo_small"]
    pb = synth.photo_batch(**skw).to(dev)
    g = load_golden("photo_stereo_small")
    me = SimpleNamespace(opt=_opts(pb, adv_train=True, supervised_adv=sup, contrastive_learning=con,
                                   no_original_train=no_orig), num_scales=len(pb.scales))
    me._dmh_noise = pb.noise
    me.gt_model = _TinyNet(3).to(dev)
    me.sup_loss_creteria = torch.nn.MSELoss()
    me.models = {"contrastive_learning": _Contrast()}
    inputs, outputs, disps = _inputs_outputs(pb)
    inputs[("color_ben", 0, 0)] = pb.color[("s", 0)]
    gen = torch.Generator().manual_seed(4)
    outputs["middle_features_aug"] = torch.randn(2, 8, generator=gen).to(dev)
    outputs["middle_features_ben"] = torch.randn(2, 8, generator=gen).to(dev)
    objective.fused_generate_images_pred(me, inputs, outputs)
    losses = objective.fused_compute_losses(me, inputs, outputs)
    exp = _adv_branches_reference(me, inputs, outputs, {"loss": torch.as_tensor(g["loss"]).to(dev)})
    for k, v in exp.items():
        assert k in losses, k
        assert_close(losses[k], v, TOL, k)
    assert ("loss/0" in losses) == (not no_orig)
    if sup or not no_orig:
        losses["loss"].backward()
        assert disps[0].grad is not None and float(disps[0].grad.abs().max()) > 0


def _reference_or_skip():
    from oracle import refload
    if not refload.available():
        pytest.skip("reference copy not present (run __graft_entry__.build() where /root/reference exists)")
    return refload.load()


@pytest.mark.parametrize("shape", [(2, 64, 96), (2, 192, 640)])
def test_install_on_the_reference_trainer_class(dev, shape):
    """The unmodified reference `Trainer` (oracle/_ref copy) on cuda: its own methods (stock eager PyTorch) first,
    then `install(mode="fused")` and the SAME unbound calls -- now the drop-ins --, then `uninstall()`.  Losses to
    1e-5, disparity gradients arbitrated by the fp64 oracle."""
    from depthmodelhardening_b200 import install as dmh_install
    from oracle import photometric as OP
    from oracle import ref_step
    ref = _reference_or_skip()
    B, H, W = shape
    pb_cpu = synth.photo_batch(batch=B, height=H, width=W, frame_ids=(0, "s"), seed=61)
    pb = pb_cpu.to(dev)
    # --- the reference's own methods on the GPU
    s2 = ref_step.Stage2Reference(pb, dev, inject_noise=True)
    ref_losses = s2.step()
    ref_grads = {s: s2.disps[s].grad.clone() for s in pb.scales}
    Trainer = ref.trainer.Trainer
    orig = (Trainer.generate_images_pred, Trainer.compute_losses, Trainer.compute_reprojection_loss,
            ref.layers.BackprojectDepth, ref.layers.SSIM)
    # --- patched
    done = dmh_install.install(mode="fused")
    try:
        assert done.get("trainer.Trainer.compute_losses") and done.get("layers.SSIM")
        assert Trainer.compute_losses is not orig[1] and Trainer.generate_images_pred is not orig[0]
        me = s2.me
        me._dmh_noise = {s: pb.noise[s][:, :1].contiguous() for s in pb.scales}
        inputs, outputs, disps = _inputs_outputs(pb)
        Trainer.generate_images_pred(me, inputs, outputs)
        losses = Trainer.compute_losses(me, inputs, outputs)
        losses["loss"].backward()
    finally:
        dmh_install.uninstall()
    assert Trainer.generate_images_pred is orig[0] and Trainer.compute_losses is orig[1]
    assert Trainer.compute_reprojection_loss is orig[2]
    assert ref.layers.BackprojectDepth is orig[3] and ref.layers.SSIM is orig[4]
    # fp64 arbiter: the oracle restatement in double on the CPU
    _, _, g64 = OP.objective_from_batch(pb_cpu, dtype=torch.float64)
    assert_close(losses["loss"], ref_losses["loss"], TOL, "loss vs reference-on-cuda")
    for s in pb.scales:
        assert_close(losses["loss/%d" % s], ref_losses["loss/%d" % s], TOL, "loss/%d vs reference-on-cuda" % s)
        assert_grad_close(disps[s].grad, ref_grads[s], g64[s], TOL, "grad_disp_%d vs reference-on-cuda" % s)
    assert_close(outputs[("depth", 0, 0)], s2.outputs[("depth", 0, 0)], TOL, "depth")


def test_reference_attack_iteration_vs_fused_apply(dev):
    """Stage 1 against the reference's own `PhysicalTrans` on cuda (oracle/ref_step.Stage1Reference: one L0 iteration
    of phy_obj_atk_l0.py:94-138, network gradient supplied) at Ba = 32 -- the benchmarked attack batch: composited +
    resized scenes and the patterns after the Adam step."""
    import numpy as np
    from depthmodelhardening_b200 import patch_ops
    from oracle import ref_step
    from oracle.refload import CALIB_P2
    _reference_or_skip()
    Ba = 32
    pt_cpu = synth.patch_batch(batch=Ba, seed=5)
    pt = pt_cpu.to(dev)
    s1 = ref_step.Stage1Reference(pt, dev)
    scenes_ref = s1.step()
    P34 = np.array(CALIB_P2, dtype=np.float64).reshape(3, 4)
    coeffs = patch_ops.homographies(pt_cpu.z0, pt_cpu.alpha, P34, obj_hw=(synth.PATCH_H, synth.PATCH_W)).to(dev)
    l0 = patch_ops.L0State(pt.obj, pt.pattern_pos, pt.pattern_neg, lr=0.5, betas=(0.5, 0.9))
    adv = l0.compose_count(first=True)
    scenes, mask_out, grad_patch = patch_ops.apply_patch_fwd_bwd(adv, pt.mask, pt.scenes, coeffs, pt.upstream)
    l0.adam_step(grad_patch, 0.06, 0.1)
    torch.cuda.synchronize()
    assert_close(scenes, scenes_ref, TOL, "adv scenes (Ba=32) vs reference PhysicalTrans on cuda", max_outlier_frac=1e-4,
                 outlier_rtol=1.0)
    # Adam's first step moves every element by lr * sign(g) wherever |g| >> eps: compare the updated patterns
    assert_close(l0.ppos, s1.pp.detach(), 1e-4, "pattern_pos after the Adam step", max_outlier_frac=2e-3, outlier_rtol=2.0)
    assert_close(l0.pneg, s1.pn.detach(), 1e-4, "pattern_neg after the Adam step", max_outlier_frac=2e-3, outlier_rtol=2.0)


def test_apgd_dropin_vs_reference_class_same_device(dev):
    """next-4: `Phy_obj_atk_apgd` (Auto-PGD with momentum and the step-size schedule) -- the drop-in against the
    reference's OWN class (oracle/_ref copy) on the same GPU, same stand-in network, same torch / numpy / random seeds.
    Both consume the RNGs identically, so placements and the random start agree; what differs is kernels vs torch ops
    (sign flips of the step where |grad| ~ 0 are bounded as in the other attack-loop tests)."""
    import importlib
    import os
    import random
    from depthmodelhardening_b200 import attacks
    from tests.test_gpu_patch import _tiny
    ref = _reference_or_skip()
    ref_apgd = importlib.import_module("torchattacks.attacks.phy_obj_atk_apgd")
    attacks.object_dataset_root = ref.calib_root
    pbt = synth.patch_batch(batch=3, seed=4).to(dev)
    model = _tiny(dev).eval()
    dist = list(range(5, 10, 2))
    outs = []
    for cls in (ref_apgd.Phy_obj_atk_APGD, attacks.Phy_obj_atk_APGD):
        torch.manual_seed(123)
        torch.cuda.manual_seed_all(123)
        random.seed(11)
        np.random.seed(12)
        atk = cls(model, pbt.obj.clone(), pbt.mask.clone(), norm="Linf", eps=0.1, steps=8, seed=17, dist_range=dist)
        outs.append(atk(pbt.scenes.clone(), 3))
    (adv_r, ben_r, m_r, x_r), (adv_o, ben_o, m_o, x_o) = outs
    assert x_o.shape == x_r.shape == pbt.obj.shape
    assert_close(ben_o, ben_r, TOL, "benign scenes", max_outlier_frac=1e-4, outlier_rtol=1.0)
    assert_close(m_o, m_r, TOL, "masks", max_outlier_frac=1e-4, outlier_rtol=1.0)
    # the adversarial patch: identical except where a ~0 gradient flipped the sign of a step
    frac = float(((x_o - x_r).abs() > 1e-6).float().mean())
    assert frac < 2e-2, frac
    assert float((x_o - pbt.obj).abs().max()) <= 0.1 + 1e-6 and float(x_o.min()) >= 0.0 and float(x_o.max()) <= 1.0
    assert_close(adv_o.double().sum(), adv_r.double().sum(), 1e-3, "adv scenes (sum)")


def test_apgd_step_kernel_is_the_torch_expression_bit_for_bit(dev):
    """dmh_apgd_linf_step against phy_obj_atk_apgd.py:214-222 evaluated by torch, element for element."""
    from depthmodelhardening_b200 import patch_ops
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(1, 3, 260, 300, generator=gen).to(dev)
    x_adv = (x + 0.1 * (2 * torch.rand(x.shape, generator=gen).to(dev) - 1)).clamp(0, 1)
    x_old = (x + 0.1 * (2 * torch.rand(x.shape, generator=gen).to(dev) - 1)).clamp(0, 1)
    grad = torch.randn(x.shape, generator=gen).to(dev)
    grad.view(-1)[::7] = 0.0                                   # sign(0) = 0
    eps = 0.1
    for a, step in ((1.0, 0.2), (0.75, 0.05), (0.75, 0.0125)):
        step_size = step * torch.ones([1, 1, 1, 1], device=dev)
        grad2 = x_adv - x_old
        z = x_adv + step_size * torch.sign(grad)
        z = torch.clamp(torch.min(torch.max(z, x - eps), x + eps), 0.0, 1.0)
        want = torch.clamp(torch.min(torch.max(x_adv + (z - x_adv) * a + grad2 * (1 - a), x - eps), x + eps), 0.0, 1.0)
        got = patch_ops.apgd_linf_step(x_adv, x_old, grad, x, step, a, eps)
        assert torch.equal(got, want), int((got != want).sum())


def _numpy_compute_errors(gt, pred, mask):
    """evaluate_depth.py:57-99 verbatim semantics in numpy (float32 in, as the harness feeds it)."""
    if mask is None:
        mask = np.ones_like(gt)
    total = mask.sum()
    thresh = np.maximum(gt / pred, pred / gt)
    a = [((thresh < 1.25 ** k) * mask).sum() / total for k in (1, 2, 3)]
    abs_err = (np.abs(gt - pred) * mask).sum() / total
    rmse = np.sqrt((((gt - pred) ** 2) * mask).sum() / total)
    rmse_log = np.sqrt((((np.log(gt) - np.log(pred)) ** 2) * mask).sum() / total)
    abs_rel = np.sum(np.abs(gt - pred) / gt * mask) / total
    sq_rel = np.sum(((gt - pred) ** 2) / gt * mask) / total
    return abs_err, abs_rel, sq_rel, rmse, rmse_log, a[0], a[1], a[2]


@pytest.mark.parametrize("masked", [True, False])
def test_depth_error_metrics_vs_numpy(dev, masked):
    """evaluation.compute_errors (one launch: disparity -> clamped metric depth -> eight masked reductions) against the
    harness' own sequence -- torch disp_to_depth / clamp on the device, numpy reductions on the host
    (evaluate_depth.py:193-196, 57-99)."""
    from depthmodelhardening_b200 import evaluation, layers
    gen = torch.Generator().manual_seed(8)
    B, H, W = 3, 320, 1024
    disp_gt = (torch.rand(B, 1, H, W, generator=gen) * 0.6 + 0.01).to(dev)
    disp_atk = (disp_gt.cpu() * (0.5 + torch.rand(B, 1, H, W, generator=gen))).to(dev)
    disp_atk.view(-1)[::1001] *= -1.0                      # the harness takes |disp|
    mask = None
    if masked:
        mask = torch.zeros(B, 1, H, W)
        mask[:, :, 100:220, 300:700] = torch.rand(B, 1, 120, 400, generator=gen)   # resized masks are fractional
        mask = mask.to(dev)
    got = evaluation.compute_errors(disp_gt, disp_atk, mask)
    gt_depth = torch.clamp(layers.disp_to_depth(torch.abs(disp_gt), 0.1, 100)[1] * 5.4, max=80, min=1e-3)
    atk_depth = torch.clamp(layers.disp_to_depth(torch.abs(disp_atk), 0.1, 100)[1] * 5.4, max=80, min=1e-3)
    want = _numpy_compute_errors(gt_depth.cpu().numpy().astype(np.float64), atk_depth.cpu().numpy().astype(np.float64),
                                 None if mask is None else mask.cpu().numpy().astype(np.float64))
    for name, g, w in zip(evaluation.ERROR_NAMES, got, want):
        assert abs(g - w) <= 1e-5 * max(abs(w), 1e-12), (name, g, w)


def test_evaluate_attacks_harness_runs_on_the_dropins(dev):
    """evaluation.evaluate_attacks (evaluate_depth.py:113-214 on the drop-in attack classes): two scene batches, L-inf
    and vanila, finite statistics; the attacked depth error on the object is larger than the vanila one."""
    import os
    from depthmodelhardening_b200 import attacks, evaluation
    from oracle import refload
    from tests.test_gpu_patch import _tiny
    import tempfile
    root = tempfile.mkdtemp(prefix="dmh_eval_")
    refload.write_calib(root)
    attacks.object_dataset_root = root
    pbt = synth.patch_batch(batch=2, seed=6).to(dev)
    model = _tiny(dev).eval()
    scenes = [pbt.scenes, pbt.scenes.flip(0)]
    base = dict(batch_size=2, epsilon=0.1, alpha=0.02, step=3)
    torch.manual_seed(0)
    mean_v, max_v = evaluation.evaluate_attacks(model, dict(base, norm_type="vanila"), scenes, pbt.obj, pbt.mask,
                                                eval_count=2, verbose=False)
    mean_a, max_a = evaluation.evaluate_attacks(model, dict(base, norm_type="l_inf"), scenes, pbt.obj, pbt.mask,
                                                eval_count=2, verbose=False)
    assert mean_v.shape == (8,) and np.isfinite(mean_v).all() and np.isfinite(max_a).all()
    assert mean_v[0] < 1e-6                                  # vanila: adversarial == benign object
    assert mean_a[0] > mean_v[0]
    with pytest.raises(NotImplementedError):
        evaluation.build_attack(model, dict(base, norm_type="image"), pbt.obj, pbt.mask)
    # the black-box searches through the same harness: Square (reference constructor arguments) and light (search on
    # the first batch, its patch placed on the following ones, evaluate_depth.py:178-182)
    mean_s, _ = evaluation.evaluate_attacks(model, dict(base, norm_type="Square", n_queries=3), scenes, pbt.obj,
                                            pbt.mask, eval_count=2, verbose=False)
    np.random.seed(3)
    mean_l, _ = evaluation.evaluate_attacks(model, dict(base, norm_type="light", light_kwargs=dict(n_init=2, n_search=2)),
                                            scenes, pbt.obj, pbt.mask, eval_count=2, verbose=False)
    assert np.isfinite(mean_s).all() and np.isfinite(mean_l).all()
    assert mean_s[0] > mean_v[0] and mean_l[0] >= mean_v[0]



def test_gaussian_blur_attack_dropin_vs_reference_class_same_device(dev):
    """next-4: `Phy_obj_atk_guassian` -- the drop-in against the reference's own class on the same GPU, same network
    and seeds: same candidates (scipy blur), same placements, so the chosen patch is the same candidate and the
    returned scenes agree to the patch-apply tolerance."""
    import importlib
    import random
    from depthmodelhardening_b200 import attacks
    from tests.test_gpu_patch import _tiny
    ref = _reference_or_skip()
    ref_g = importlib.import_module("torchattacks.attacks.phy_obj_atk_guassian")
    attacks.object_dataset_root = ref.calib_root
    pbt = synth.patch_batch(batch=3, seed=4).to(dev)
    model = _tiny(dev).eval()
    outs = []
    for cls in (ref_g.Phy_obj_atk_guassian, attacks.Phy_obj_atk_guassian):
        random.seed(21)
        atk = cls(model, pbt.obj.clone(), pbt.mask.clone(), steps=5, dist_range=list(range(5, 10, 2)))
        outs.append(atk(pbt.scenes.clone(), 3))
    (adv_r, ben_r, m_r, x_r), (adv_o, ben_o, m_o, x_o) = outs
    assert torch.equal(x_o, x_r)                           # the same blurred candidate won
    assert_close(adv_o, adv_r, TOL, "adv scenes", max_outlier_frac=1e-4, outlier_rtol=1.0)
    assert_close(ben_o, ben_r, TOL, "benign scenes", max_outlier_frac=1e-4, outlier_rtol=1.0)
    assert_close(m_o, m_r, TOL, "masks", max_outlier_frac=1e-4, outlier_rtol=1.0)



def test_arbitrary_pattern_attack_dropin_vs_reference_class_same_device(dev):
    """next-4: `Phy_obj_atk_arbi` against the reference's own class: same RandomState(17) stream (two calls: the
    stream carries over), same placements; patches equal bit for bit, scenes to the patch-apply tolerance."""
    import importlib
    from depthmodelhardening_b200 import attacks
    from tests.test_gpu_patch import _tiny
    ref = _reference_or_skip()
    ref_a = importlib.import_module("torchattacks.attacks.phy_obj_atk_arbi")
    attacks.object_dataset_root = ref.calib_root
    pbt = synth.patch_batch(batch=3, seed=4).to(dev)
    model = _tiny(dev).eval()
    a_ref = ref_a.Phy_obj_atk_arbi(model, pbt.obj.clone(), pbt.mask.clone())
    a_our = attacks.Phy_obj_atk_arbi(model, pbt.obj.clone(), pbt.mask.clone())
    for call in range(2):
        adv_r, ben_r, m_r, x_r = a_ref(pbt.scenes.clone(), 3, eval=bool(call))
        adv_o, ben_o, m_o, x_o = a_our(pbt.scenes.clone(), 3, eval=bool(call))
        assert torch.equal(x_o, x_r)
        assert_close(adv_o, adv_r, TOL, "adv scenes", max_outlier_frac=1e-4, outlier_rtol=1.0)
        assert_close(ben_o, ben_r, TOL, "benign scenes", max_outlier_frac=1e-4, outlier_rtol=1.0)
        assert_close(m_o, m_r, TOL, "masks", max_outlier_frac=1e-4, outlier_rtol=1.0)



def test_tube_light_kernel_vs_oracle(dev):
    """`dmh_tube_light_patch` against oracle/light.py (pinned bit for bit to the reference's
    tube_light_generation_by_func + simple_add, tests/golden/light.npz): candidate patches of a seeded beam walk at
    the attack's patch size, compared bit for bit."""
    from depthmodelhardening_b200 import attacks, patch_ops
    from oracle import light as OL
    obj = synth.patch_batch(batch=1, seed=2).obj
    base_hwc = OL.to_u8_hwc(obj[0].numpy())
    base_dev = obj[0].to(dev).mul(255).byte().contiguous()
    assert np.array_equal(base_dev.cpu().numpy(), np.transpose(base_hwc, (2, 0, 1)))
    np.random.seed(7)
    cases = [tuple(int(v) for v in q) for q in OL.candidate_params(n_init=3, n_search=2)]
    cases += [(380, 0, 0, 10), (470, 90, 30, 1600), (500, 91, 399, 333), (750, 180, 400, 9)]
    n_lit = 0
    for wl, ang, icpt, beta in cases:
        want = OL.candidate_patch(base_hwc, (wl, ang, icpt, beta))
        got = patch_ops.tube_light_patch(base_dev, OL.slope_of(ang), icpt, beta, attacks.wavelength_to_rgb(wl))
        assert got.shape == (1, 3, synth.PATCH_H, synth.PATCH_W)
        assert np.array_equal(got[0].cpu().numpy(), want), (wl, ang, icpt, beta)
        n_lit += int((want != np.transpose(base_hwc, (2, 0, 1)).astype(np.float32) / np.float32(255)).any())
    assert n_lit >= 6


def test_keep_best_and_square_candidate_kernels_bit_exact(dev):
    """`dmh_square_linf_candidate` against the torch expression of phy_obj_atk_square.py:263-274 and `dmh_keep_best`
    against the host-side `if cost < best_cost` (strict, NaN never accepted), bit for bit."""
    from depthmodelhardening_b200 import patch_ops
    g = torch.Generator().manual_seed(3)
    H, W, eps = synth.PATCH_H, synth.PATCH_W, 0.1
    x = torch.rand(1, 3, H, W, generator=g).to(dev)
    x_best = torch.clamp(x + eps * torch.sign(2 * torch.rand(1, 3, 1, W, generator=g) - 1).to(dev), 0., 1.)
    for vh, vw, s in ((0, 0, 250), (5, 7, 99), (259, 299, 1), (3, 4, 0)):
        sign = torch.sign(2 * torch.rand(3, 1, 1, generator=g) - 1)
        new_deltas = torch.zeros(3, H, W, device=dev)
        new_deltas[:, vh:vh + s, vw:vw + s] = (2. * eps * sign).to(dev)
        want = torch.clamp(torch.min(torch.max(x_best + new_deltas, x - eps), x + eps), 0., 1.)
        got = patch_ops.square_linf_candidate(x_best, x, vh, vw, s, (2. * eps * sign).reshape(3).tolist(), eps)
        assert torch.equal(got, want), (vh, vw, s)
    keeper = patch_ops.BestKeeper(x, init_cost=1e10)
    best_cost, best = 1e10, None
    cands = [torch.rand(1, 3, H, W, generator=g).to(dev) for _ in range(6)]
    for cost, cand in zip((0.5, 0.7, 0.5, float("nan"), 0.25, 0.25), cands):
        keeper.offer(torch.tensor(cost, device=dev), cand)
        if cost < best_cost:
            best_cost, best = cost, cand
        assert torch.equal(keeper.best, best) and float(keeper.best_cost) == np.float32(best_cost)


def test_square_attack_dropin_vs_reference_class_same_device(dev):
    """next-4: `Phy_obj_atk_Square` -- the drop-in against the reference's own class on the same GPU, same network
    and seeds.  Both draw the stripes, windows and signs from the torch CPU generator and the final placements from
    `random` in the same order; the returned patch must be the SAME tensor (the reference's search evaluates the
    current best at every query and therefore keeps its striped start, see the class docstring)."""
    import importlib
    import random
    from depthmodelhardening_b200 import attacks
    from tests.test_gpu_patch import _tiny
    ref = _reference_or_skip()
    ref_s = importlib.import_module("torchattacks.attacks.phy_obj_atk_square")
    attacks.object_dataset_root = ref.calib_root
    pbt = synth.patch_batch(batch=3, seed=4).to(dev)
    model = _tiny(dev).eval()
    outs, rng_after = [], []
    for cls in (ref_s.Phy_obj_atk_Square, attacks.Phy_obj_atk_Square):
        torch.manual_seed(31)
        random.seed(32)
        atk = cls(model, pbt.obj.clone(), pbt.mask.clone(), eps=0.1, n_queries=6, seed=5, dist_range=list(range(5, 10, 2)))
        outs.append(atk(pbt.scenes.clone(), 3, eval=True))
        rng_after.append((float(torch.rand(1)), random.random()))
    (adv_r, ben_r, m_r, x_r), (adv_o, ben_o, m_o, x_o) = outs
    assert rng_after[0] == rng_after[1]                    # the same number of draws from both generators
    assert torch.equal(x_o, x_r)
    assert float((x_o - pbt.obj).abs().max()) > 0.05       # (the stripes)
    assert_close(adv_o, adv_r, TOL, "adv scenes", max_outlier_frac=1e-4, outlier_rtol=1.0)
    assert_close(ben_o, ben_r, TOL, "benign scenes", max_outlier_frac=1e-4, outlier_rtol=1.0)
    assert_close(m_o, m_r, TOL, "masks", max_outlier_frac=1e-4, outlier_rtol=1.0)
    with pytest.raises(NameError):
        attacks.Phy_obj_atk_Square(model, pbt.obj.clone(), pbt.mask.clone(), norm="L2", n_queries=2)(pbt.scenes.clone(), 3)


def test_light_attack_dropin_vs_reference_class_same_device(dev, monkeypatch):
    """next-4: `Phy_obj_atk_light` -- the drop-in against the reference's own class on the same GPU, same network and
    seeds, on a shortened search (the reference hard-codes 200 x 20 x 2 candidates at ~0.1 s of Python loops each:
    its module-level `range` is shadowed so that both loops run twice -- 8 candidates -- and the drop-in gets
    n_init = n_search = 2).  Same beam walk (numpy RNG) and placements (`random`): the same candidate must win, and
    its patch -- built by `dmh_tube_light_patch` -- must equal the reference's PIL / OpenCV result bit for bit."""
    import builtins
    import importlib
    import random
    from depthmodelhardening_b200 import attacks
    from tests.test_gpu_patch import _tiny
    ref = _reference_or_skip()
    try:
        ref_l = importlib.import_module("torchattacks.attacks.phy_obj_atk_light")
    except ImportError as e:                               # cv2 / scipy.ndimage.filters of the reference's imports
        pytest.skip("reference light attack not importable here: %s" % e)
    monkeypatch.setattr(ref_l, "range", lambda n: builtins.range(min(n, 2)), raising=False)
    attacks.object_dataset_root = ref.calib_root
    pbt = synth.patch_batch(batch=3, seed=4).to(dev)
    model = _tiny(dev).eval()
    outs, rng_after = [], []
    for cls, kw in ((ref_l.Phy_obj_atk_light, {}), (attacks.Phy_obj_atk_light, dict(n_init=2, n_search=2))):
        np.random.seed(41)
        random.seed(42)
        atk = cls(model, pbt.obj.clone(), pbt.mask.clone(), dist_range=list(range(5, 10, 2)), **kw)
        outs.append(atk(pbt.scenes.clone(), 3, eval=True))
        rng_after.append((np.random.randint(1 << 30), random.random()))
    (adv_r, ben_r, m_r, x_r), (adv_o, ben_o, m_o, x_o) = outs
    assert rng_after[0] == rng_after[1]
    assert torch.equal(x_o, x_r)                           # the same candidate won, the same bytes / 255
    assert float((x_o - pbt.obj).abs().max()) > 0.0
    assert_close(adv_o, adv_r, TOL, "adv scenes", max_outlier_frac=1e-4, outlier_rtol=1.0)
    assert_close(ben_o, ben_r, TOL, "benign scenes", max_outlier_frac=1e-4, outlier_rtol=1.0)
    assert_close(m_o, m_r, TOL, "masks", max_outlier_frac=1e-4, outlier_rtol=1.0)

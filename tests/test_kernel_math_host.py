"""CPU: the per-pixel formulas of the CUDA kernels (csrc/dmh_math.cuh), compiled
for the host by tests/host_emul.cpp, against the oracle and the reference
goldens.  This validates the MATH of the kernels without a GPU (coordinate
chain, bilinear gradient, coefficient form of the SSIM backward, reflection
multiplicities, automask argmin); the real kernels' tiling is checked by the
`-m gpu` tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from depthmodelhardening_b200 import synth
from oracle import photometric as OP
from oracle.make_golden import PHOTO_CASES
from tests.util import assert_close, assert_grad_close, load_golden

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
SO = os.path.join(BUILD, "libdmh_hostemu.so")
SRC = os.path.join(HERE, "host_emul.cpp")
HDR = os.path.join(os.path.dirname(HERE), "depthmodelhardening_b200", "csrc", "dmh_math.cuh")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(BUILD, exist_ok=True)
    if (not os.path.exists(SO)) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", "-o", SO, SRC])
    return C.CDLL(SO)


def fp(t):
    return t.contiguous().data_ptr()


def vp(t):
    return C.c_void_p(t.contiguous().data_ptr())


def test_warp_forward_backward_math(emu):
    pb = synth.photo_batch(batch=2, height=48, width=80, frame_ids=(0, -1), seed=3)
    B, H, W = pb.batch, pb.height, pb.width
    disp = pb.disp[0].clone()
    src = pb.color[(-1, 0)].contiguous()
    T = pb.T[-1].contiguous()
    warped = torch.empty(B, 3, H, W)
    emu.emu_warp_fwd(vp(disp), vp(src), vp(pb.K), vp(pb.inv_K), vp(T), B, 3, H, W, C.c_float(0.1), C.c_float(100.0),
                     vp(warped))
    d = disp.clone().requires_grad_(True)
    Tr = T.clone().requires_grad_(True)
    ref, _, _ = OP.warp_from_disp(d, src, pb.K, pb.inv_K, Tr, 0.1, 100.0)
    assert_close(warped, ref, 1e-5, "warped")
    up = synth.randn(ref.shape, 4)
    (ref * up).sum().backward()
    gd = torch.empty(B, 1, H, W)
    gP = torch.empty(B, 12)
    emu.emu_warp_bwd(vp(up), vp(disp), vp(src), vp(pb.K), vp(pb.inv_K), vp(T), B, 3, H, W, C.c_float(0.1),
                     C.c_float(100.0), vp(gd), vp(gP))
    assert_close(gd, d.grad, 1e-5, "grad_disp", max_outlier_frac=2e-3)
    gT = torch.matmul(pb.K[:, :3, :].transpose(1, 2), gP.view(B, 3, 4))
    assert_close(gT, Tr.grad, 1e-4, "grad_T")


def test_ssim_coefficient_backward_math(emu):
    g = load_golden("layers")
    pb = synth.photo_batch(batch=2, height=48, width=80, frame_ids=(0, -1), seed=21)
    x, y = pb.color[(0, 0)].contiguous(), pb.color[(-1, 0)].contiguous()
    up = synth.randn(x.shape, 23)
    out, gx, gy = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    emu.emu_ssim(vp(x), vp(y), vp(up), x.shape[0] * x.shape[1], 48, 80, vp(out), vp(gx), vp(gy))
    assert_close(out, g["ssim"], 1e-5, "ssim")
    assert_close(gx, g["grad_x"], 1e-5, "grad_x")
    assert_close(gy, g["grad_y"], 1e-5, "grad_y")


@pytest.mark.parametrize("name", sorted(PHOTO_CASES))
def test_fused_objective_math_vs_reference_golden(emu, name):
    skw, over = PHOTO_CASES[name]
    pb = synth.photo_batch(**skw)
    g = load_golden("photo_" + name)
    B, H, W = pb.batch, pb.height, pb.width
    srcs_ids = pb.frame_ids[1:]
    F_ = len(srcs_ids)
    flags = (1 if over.get("no_ssim") else 0) | (2 if over.get("avg_reprojection") else 0)
    automask = not over.get("disable_automasking", False)
    target = pb.color[(0, 0)].contiguous()
    srcs = [pb.color[(f, 0)].contiguous() for f in srcs_ids]
    Ts = [pb.T[f].contiguous() for f in srcs_ids]
    ident = None
    if automask:
        ident = torch.cat([OP.reprojection_loss(s, target, bool(over.get("no_ssim"))) for s in srcs], 1).contiguous()
    src_arr = (C.c_void_p * F_)(*[fp(s) for s in srcs])
    T_arr = (C.c_void_p * F_)(*[fp(t) for t in Ts])
    n_ident = 0 if not automask else (1 if over.get("avg_reprojection") else F_)
    total = 0.0
    for s in pb.scales:
        dlow = pb.disp[s].clone().requires_grad_(True)
        dfull = torch.nn.functional.interpolate(dlow, [H, W], mode="bilinear", align_corners=False)
        nz = pb.noise[s][:, :n_ident].contiguous() if automask else None
        loss_sum = C.c_double(0.0)
        gd = torch.empty(B, 1, H, W)
        sel = torch.empty(B, H, W, dtype=torch.uint8)
        emu.emu_photo_scale(vp(target), src_arr, T_arr, F_, vp(dfull.detach()), vp(pb.K), vp(pb.inv_K),
                            vp(ident) if automask else None, vp(nz) if automask else None, B, H, W, C.c_float(0.1),
                            C.c_float(100.0), flags, C.byref(loss_sum), vp(gd), vp(sel), None)
        # compose with the parts outside the fused kernel (interpolate, smoothness) via the oracle
        sm = OP.normalised_smooth_loss(dlow, pb.color[(0, s)])
        smw = 1e-3 / (2 ** s)
        loss_s = loss_sum.value / (B * H * W) + smw * float(sm)
        assert_close(loss_s, g["loss_%d" % s], 1e-5, "loss/%d" % s)
        total += loss_s
        (dfull * gd / (B * H * W) / len(pb.scales)).sum().backward(retain_graph=True)
        (sm * smw / len(pb.scales)).backward()
        assert_close(dlow.grad, g["grad_disp_%d" % s], 1e-5, "grad_disp_%d" % s, max_outlier_frac=2e-3,
                     outlier_rtol=0.5)
        if automask:
            assert np.array_equal((sel.numpy() > n_ident - 1).astype(np.uint8), g["ident_sel_%d" % s])
    assert_close(total / len(pb.scales), g["loss"], 1e-5, "loss")


@pytest.mark.parametrize("name", ["stereo_small", "stereo_iid", "no_ssim", "no_automask_f1"])
def test_fast_path_arithmetic_vs_reference_golden(emu, name):
    """photo_fast.cu's arithmetic (separable sums, sum*(1/9), collapsed chain) on the host."""
    skw, over = PHOTO_CASES[name]
    pb = synth.photo_batch(**skw)
    g = load_golden("photo_" + name)
    B, H, W = pb.batch, pb.height, pb.width
    flags = 1 if over.get("no_ssim") else 0
    automask = not over.get("disable_automasking", False)
    target, src, T = pb.color[(0, 0)].contiguous(), pb.color[("s", 0)].contiguous(), pb.T["s"].contiguous()
    ident = OP.reprojection_loss(src, target, bool(over.get("no_ssim"))).contiguous() if automask else None
    _, _, g64 = OP.objective_from_batch(pb, OP.default_opts(scales=list(pb.scales), **over), dtype=torch.float64)
    for s in pb.scales:
        dlow = pb.disp[s].clone().requires_grad_(True)
        dfull = torch.nn.functional.interpolate(dlow, [H, W], mode="bilinear", align_corners=False)
        nz = pb.noise[s][:, :1].contiguous() if automask else None
        loss_sum = C.c_double(0.0)
        gd = torch.empty(B, 1, H, W)
        sel = torch.empty(B, H, W, dtype=torch.uint8)
        emu.emu_photo_scale_fast(vp(target), vp(src), vp(T), vp(dfull.detach()), vp(pb.K), vp(pb.inv_K),
                                 vp(ident) if automask else None, vp(nz) if automask else None, B, H, W,
                                 C.c_float(0.1), C.c_float(100.0), flags, C.byref(loss_sum), vp(gd), vp(sel))
        sm = OP.normalised_smooth_loss(dlow, pb.color[(0, s)])
        smw = 1e-3 / (2 ** s)
        assert_close(loss_sum.value / (B * H * W) + smw * float(sm.detach()), g["loss_%d" % s], 1e-5, "loss/%d" % s)
        (dfull * gd / (B * H * W) / len(pb.scales)).sum().backward(retain_graph=True)
        (sm * smw / len(pb.scales)).backward()
        # iid inputs ("adversarial gather") put many pixels on floor()/argmin knife edges: allow 0.5 % there
        assert_grad_close(dlow.grad, g["grad_disp_%d" % s], g64[s], 1e-5, "grad_disp_%d" % s,
                          outlier_frac=5e-3 if name == "stereo_iid" else 2e-3)
        if automask:
            assert np.mean((sel.numpy() > 0).astype(np.uint8) != g["ident_sel_%d" % s]) < 1e-3


# ----------------------------------------------------------------------------- colour jitter (groundwork)
@pytest.fixture(scope="module")
def emu_jitter():
    src = os.path.join(HERE, "host_emul_jitter.cpp")
    hdr = os.path.join(os.path.dirname(HERE), "depthmodelhardening_b200", "csrc", "jitter_math.cuh")
    so = os.path.join(BUILD, "libdmh_hostemu_jitter.so")
    os.makedirs(BUILD, exist_ok=True)
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", "-o", so, src])
    lib = C.CDLL(so)
    lib.emu_grey_sum.restype = C.c_longlong
    return lib


def test_colour_jitter_math_vs_oracle(emu_jitter):
    """csrc/jitter_math.cuh (the per-pixel formulas a jitter kernel will use) compiled with g++ equals
    oracle/pil_enhance.py -- which equals Pillow / torchvision bit for bit -- over a sweep of the colour cube, for
    factors inside and outside [0, 1] and every hue shift class."""
    import numpy as np
    from oracle import pil_enhance as E
    r, g, b = np.meshgrid(np.arange(256), np.arange(256), np.arange(0, 256, 15), indexing="ij")
    img = np.ascontiguousarray(np.stack([r.reshape(1024, -1), g.reshape(1024, -1), b.reshape(1024, -1)]).astype(np.uint8))
    n = img.shape[1] * img.shape[2]

    def run(op, f=0.0, aux=0):
        out = np.empty_like(img)
        emu_jitter.emu_jitter(C.c_void_p(img.ctypes.data), C.c_longlong(n), C.c_int(op), C.c_float(f), C.c_int(aux),
                              C.c_void_p(out.ctypes.data))
        return out

    assert emu_jitter.emu_grey_sum(C.c_void_p(img.ctypes.data), C.c_longlong(n)) == int(E.grey(img).astype(np.int64).sum())
    mean = int(E.grey(img).astype(np.float64).mean() + 0.5)
    for f in (0.8, 0.8731, 1.0, 1.1999, 1.2, 0.0, 0.5, 1.7):
        assert np.array_equal(run(0, f), E.brightness(img, f)), ("brightness", f)
        assert np.array_equal(run(1, f, mean), E.contrast(img, f)), ("contrast", f)
        assert np.array_equal(run(2, f), E.saturation(img, f)), ("saturation", f)
    hsv = E.rgb_to_hsv(img)
    assert np.array_equal(run(4), hsv)
    img_keep, img = img, hsv
    assert np.array_equal(run(5), E.hsv_to_rgb(hsv))
    img = img_keep
    for f in (-0.1, -0.0371, 0.0, 0.05, 0.1, 0.5, -0.5):
        assert np.array_equal(run(3, aux=int(f * 255) & 0xff), E.hue(img, f)), ("hue", f)


def test_pose_sum_reduce_scatter_network_host_emulation():
    """csrc/photo_mf.cu pose_epilogue: 12 per-lane sums are reduced over a warp by a reduce-scatter (12 -> 6 -> 3 -> 2 -> 1
    values per lane, 13 shuffles) instead of 12 butterflies (60).  The same network emulated on 32 numpy lanes: every
    one of the 12 totals ends in exactly the lane / index the kernel writes from, and equals the plain sum."""
    rng = np.random.RandomState(5)
    acc = rng.randn(32, 12)                     # acc[lane][i]
    lanes = np.arange(32)
    xor = lambda v, m: v[lanes ^ m]             # __shfl_xor_sync
    b4, b3, b2, b1 = [(lanes & m) != 0 for m in (16, 8, 4, 2)]
    v6 = np.zeros((32, 6))
    for i in range(6):
        keep = np.where(b4, acc[:, 6 + i], acc[:, i]); send = np.where(b4, acc[:, i], acc[:, 6 + i])
        v6[:, i] = keep + xor(send, 16)
    v3 = np.zeros((32, 3))
    for i in range(3):
        keep = np.where(b3, v6[:, 3 + i], v6[:, i]); send = np.where(b3, v6[:, i], v6[:, 3 + i])
        v3[:, i] = keep + xor(send, 8)
    k0 = np.where(b2, v3[:, 2], v3[:, 0]); s0 = np.where(b2, v3[:, 0], v3[:, 2])
    k1 = np.where(b2, 0.0, v3[:, 1]); s1 = np.where(b2, v3[:, 1], 0.0)
    v2 = np.stack([k0 + xor(s0, 4), k1 + xor(s1, 4)], 1)
    k = np.where(b1, v2[:, 1], v2[:, 0]); sd = np.where(b1, v2[:, 0], v2[:, 1])
    v1 = k + xor(sd, 2)
    v1 = v1 + xor(v1, 1)
    vi = np.where(b4, 6, 0) + np.where(b3, 3, 0) + np.where(b2, 2, np.where(b1, 1, 0))
    writers = ((lanes & 1) == 0) & ~(b2 & b1)
    assert sorted(vi[writers].tolist()) == list(range(12))
    want = acc.sum(0)
    for lane in lanes[writers]:
        assert abs(v1[lane] - want[vi[lane]]) < 1e-12


# ----------------------------------------------------------------------------- black-box candidates (next-4)
@pytest.fixture(scope="module")
def emu_light():
    src = os.path.join(HERE, "host_emul_light.cpp")
    hdr = os.path.join(os.path.dirname(HERE), "depthmodelhardening_b200", "csrc", "light_math.cuh")
    so = os.path.join(BUILD, "libdmh_hostemu_light.so")
    os.makedirs(BUILD, exist_ok=True)
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", "-o", so, src])
    return C.CDLL(so)


def test_tube_light_math_vs_oracle(emu_light):
    """csrc/light_math.cuh (the per-pixel body of dmh_tube_light_patch) compiled with g++ equals oracle/light.py --
    which equals the reference's tube_light_generation_by_func + simple_add bit for bit (tests/golden/light.npz) --
    over the golden cases and a seeded walk of the search, lit bytes and fp32 patch."""
    import math
    from oracle import light as OL
    rs = np.random.RandomState(9)
    np.random.seed(5)
    cases = [(380, 0, 0, 10), (470, 89, 30, 1600), (500, 91, 399, 333), (600, 179, 5, 1234), (750, 160, 350, 9)]
    cases += [tuple(int(v) for v in q) for q in OL.candidate_params(n_init=3, n_search=3)]
    h, w = 52, 60
    for wl, ang, icpt, beta in cases:
        base = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        want = OL.candidate_patch(base, (wl, ang, icpt, beta))
        k = OL.slope_of(ang)
        full_end, light_end = OL.light_ends(beta)
        ca = [c * 1.0 for c in OL.wavelength_to_rgb(wl)]
        planar = np.ascontiguousarray(np.transpose(base, (2, 0, 1)))
        patch = np.empty((3, h, w), np.float32)
        lit = np.empty((3, h, w), np.uint8)
        emu_light.emu_tube_light(C.c_void_p(planar.ctypes.data), C.c_int(h), C.c_int(w), C.c_double(k),
                                 C.c_double(float(icpt)), C.c_double(math.sqrt(1 + k * k)), C.c_double(float(beta)),
                                 C.c_int(full_end), C.c_int(light_end), C.c_double(ca[0]), C.c_double(ca[1]),
                                 C.c_double(ca[2]), C.c_void_p(patch.ctypes.data), C.c_void_p(lit.ctypes.data))
        assert np.array_equal(patch, want), (wl, ang, icpt, beta)
        assert np.array_equal(lit, (want * np.float32(255) + np.float32(0.5)).astype(np.uint8))


def test_square_candidate_math_vs_torch(emu_light):
    """The body of dmh_square_linf_candidate against the torch expression of phy_obj_atk_square.py:263-274."""
    import torch
    g = torch.Generator().manual_seed(3)
    H, W, eps = 26, 30, 0.1
    x = torch.rand(1, 3, H, W, generator=g)
    x_best = torch.clamp(x + eps * torch.sign(2 * torch.rand(1, 3, 1, W, generator=g) - 1), 0., 1.)
    for vh, vw, s in ((0, 0, 23), (5, 7, 9), (25, 29, 1), (3, 4, 0)):
        sign = torch.sign(2 * torch.rand(3, 1, 1, generator=g) - 1)
        new_deltas = torch.zeros(3, H, W)
        new_deltas[:, vh:vh + s, vw:vw + s] = 2. * eps * sign
        want = torch.clamp(torch.min(torch.max(x_best + new_deltas, x - eps), x + eps), 0., 1.)
        d = (2. * eps * sign).reshape(3).tolist()
        xb, xc = x_best.numpy().copy(), x.numpy().copy()
        out = np.empty((1, 3, H, W), np.float32)
        emu_light.emu_square_candidate(C.c_void_p(xb.ctypes.data), C.c_void_p(xc.ctypes.data), C.c_int(H), C.c_int(W),
                                       C.c_int(vh), C.c_int(vw), C.c_int(s), C.c_float(d[0]), C.c_float(d[1]),
                                       C.c_float(d[2]), C.c_float(eps), C.c_void_p(out.ctypes.data))
        assert np.array_equal(out, want.numpy()), (vh, vw, s)

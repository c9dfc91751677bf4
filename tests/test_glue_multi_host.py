"""CPU: host-side restatement of the launch tables of the all-scales glue kernels (csrc/objective_fused.cu:
dmh_smooth_fused_multi / dmh_disp_grad_multi) and of the kernels' block decoding: every (scale, image, tile) is visited
exactly once, in the scale order the launcher lays out, and the scale classification (same size / integer factor 2, 4,
8 / refused) follows the rule of dmh_disp_grad.  The GPU tests compare the kernels' OUTPUTS bit for bit with the
per-scale launches (tests/test_gpu_photometric.py::test_glue_multi_launches_are_bit_identical_to_per_scale)."""
import itertools

import pytest

SF_TW, SF_TH = 128, 32


def ceil_div(a, b):
    return (a + b - 1) // b


def smooth_table(sizes, B):
    """dmh_smooth_fused_multi: (gx, gy, blk0) per scale, total blocks."""
    tab, blocks = [], 0
    for (h, w) in sizes:
        gx, gy = ceil_div(w, SF_TW), ceil_div(h, SF_TH)
        tab.append((gx, gy, blocks))
        blocks += gx * gy * B
    return tab, blocks


def decode(tab, blk):
    """sf_main_multi_kernel / disp_grad_multi_kernel: linear block -> (scale, image, by, bx)."""
    s = 0
    while s + 1 < len(tab) and blk >= tab[s + 1][2]:
        s += 1
    gx, gy, blk0 = tab[s]
    r = blk - blk0
    per = gx * gy
    b, t = divmod(r, per)
    by, bx = divmod(t, gx)
    return s, b, by, bx


def dgrad_table(sizes, H, W, B, aligned=True):
    """dmh_disp_grad_multi: None when a scale needs the generic kernel, else ((R, gx, gy, blk0) per scale, blocks)."""
    tab, blocks = [], 0
    for (h, w) in sizes:
        R = H // h if (H % h == 0 and W % w == 0 and H // h == W // w) else 0
        same_ok = R == 1 and (h * w) % 4 == 0 and aligned
        up_ok = R in (2, 4, 8) and aligned and W % 4 == 0
        if not (same_ok or up_ok):
            return None
        gx, gy = (ceil_div((h * w) // 4, 1024), 1) if R == 1 else (ceil_div(w, 32), ceil_div(h, 64 // R))
        tab.append((R, gx, gy, blocks))
        blocks += gx * gy * B
    return tab, blocks


@pytest.mark.parametrize("HW,B", [((320, 1024), 3), ((96, 160), 2), ((64, 96), 1), ((50, 70), 2), ((640, 2048), 1)])
def test_smooth_multi_blocks_cover_every_tile_once(HW, B):
    H, W = HW
    sizes = [(H >> s, W >> s) for s in range(4)]
    tab, blocks = smooth_table(sizes, B)
    seen = [decode(tab, blk) for blk in range(blocks)]
    want = [(s, b, by, bx) for s, (gx, gy, _) in enumerate(tab) for b in range(B) for by, bx in
            itertools.product(range(gy), range(gx))]
    assert seen == want                       # exactly once, scale 0 (the largest) first
    for s, (h, w) in enumerate(sizes):        # the tiles of a scale cover its image
        gx, gy, _ = tab[s]
        assert gx * SF_TW >= w > (gx - 1) * SF_TW and gy * SF_TH >= h > (gy - 1) * SF_TH


@pytest.mark.parametrize("HW,B", [((320, 1024), 4), ((96, 160), 2), ((192, 640), 1)])
def test_disp_grad_multi_blocks_cover_every_output_once(HW, B):
    H, W = HW
    sizes = [(H >> s, W >> s) for s in range(4)]
    tab, blocks = dgrad_table(sizes, H, W, B)
    assert [t[0] for t in tab] == [1, 2, 4, 8]
    dec_tab = [(gx, gy, blk0) for (_, gx, gy, blk0) in tab]
    covered = [set() for _ in sizes]
    for blk in range(blocks):
        s, b, by, bx = decode(dec_tab, blk)
        R, gx, gy, _ = tab[s]
        h, w = sizes[s]
        if R == 1:                            # 1024 float4 chunks per block, four per thread
            n4 = (h * w) // 4
            items = {(b, i) for i in range(bx * 1024, min(bx * 1024 + 1024, n4))}
        else:                                 # tall tile: 32 x (64 / R) low-resolution pixels
            items = {(b, y, x) for y in range(by * (64 // R), min((by + 1) * (64 // R), h))
                     for x in range(bx * 32, min(bx * 32 + 32, w))}
        assert not (covered[s] & items)
        covered[s] |= items
    for s, (h, w) in enumerate(sizes):
        assert len(covered[s]) == (B * (h * w) // 4 if tab[s][0] == 1 else B * h * w)


def test_disp_grad_multi_refuses_what_needs_the_generic_kernel():
    H, W = 96, 160
    assert dgrad_table([(H, W), (H // 2 + 3, W // 2 + 5)], H, W, 1) is None      # non-integer factor
    assert dgrad_table([(H // 16, W // 16)], H, W, 1) is None                  # factor 16
    assert dgrad_table([(H // 2, W // 4)], H, W, 1) is None                    # anisotropic
    assert dgrad_table([(25, 35)], 50, 70, 1) is None                          # factor 2 but rows not 16-byte aligned
    assert dgrad_table([(H, W)], H, W, 1, aligned=False) is None
    assert dgrad_table([(H, W), (H // 2, W // 2)], H, W, 1) is not None


# ---- stage 1: the span table of the patch kernels (csrc/patch.cu aa_table_kernel / aa_span3) against the per-pixel
# weight function the backward used before (aa_weight_of), restated op by op in IEEE float32
import numpy as np

f32 = np.float32


def _aa_filter(x):
    x = abs(x)
    return f32(1.0) - x if x < f32(1.0) else f32(0.0)


def _span3(i, in_size, scale):
    """aa_span3: (lo, w0, w1, w2) of output index i."""
    support = scale if scale >= f32(1.0) else f32(1.0)
    invscale = f32(1.0) / scale if scale >= f32(1.0) else f32(1.0)
    center = f32(scale * f32(f32(i) + f32(0.5)))
    lo = max(int(f32(f32(center - support) + f32(0.5))), 0)
    n = min(min(int(f32(f32(center + support) + f32(0.5))), in_size) - lo, 3)
    w, total = [], f32(0.0)
    for j in range(3):
        wj = _aa_filter(f32(f32(f32(f32(j) + f32(f32(lo) - center)) + f32(0.5)) * invscale)) if j < n else f32(0.0)
        w.append(wj)
        total = f32(total + wj)
    if total != 0:
        w = [f32(x / total) for x in w]
    return lo, w


def _weight_of(o, in_size, scale, ci, maxt=8):
    """aa_weight_of: weight with which output index o reads input index ci."""
    support = scale if scale >= f32(1.0) else f32(1.0)
    invscale = f32(1.0) / scale if scale >= f32(1.0) else f32(1.0)
    center = f32(scale * f32(f32(o) + f32(0.5)))
    lo = max(int(f32(f32(center - support) + f32(0.5))), 0)
    n = min(min(int(f32(f32(center + support) + f32(0.5))), in_size) - lo, maxt)
    if ci < lo or ci >= lo + n:
        return f32(0.0)
    total, mine = f32(0.0), f32(0.0)
    for j in range(n):
        w = _aa_filter(f32(f32(f32(f32(j) + f32(f32(lo) - center)) + f32(0.5)) * invscale))
        total = f32(total + w)
        if lo + j == ci:
            mine = w
    return f32(mine / total) if total != 0 else mine


@pytest.mark.parametrize("in_size,out_size", [(1242, 1024), (375, 320), (100, 81), (64, 64), (97, 70)])
def test_span_table_equals_the_per_pixel_weight_function(in_size, out_size):
    """Every (output index, input index) pair: the table entry the patch kernels read == aa_weight_of, bit for bit, for
    scale factors below 1.5 (3-tap spans); the weights of a span sum to 1 within rounding."""
    scale = f32(f32(in_size) / f32(out_size))
    assert scale < 1.5
    for o in range(out_size):
        lo, w = _span3(o, in_size, scale)
        assert abs(float(sum(w, f32(0.0))) - 1.0) < 1e-6
        for ci in range(max(lo - 2, 0), min(lo + 5, in_size)):
            d = ci - lo
            tab = w[d] if 0 <= d < 3 else f32(0.0)
            ref = _weight_of(o, in_size, scale, ci)
            assert tab.tobytes() == ref.tobytes(), (o, ci, float(tab), float(ref))

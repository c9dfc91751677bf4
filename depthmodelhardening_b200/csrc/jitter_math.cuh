// Per-pixel arithmetic of the colour jitter on 8-bit frames (dmh_color_jitter_u8, loader_compose.cu).
// `transforms.ColorJitter` on PIL images
// (DepthNetworks/monodepth2/datasets/mono_dataset.py:297, 344-350 -> torchvision functional_pil ->
// PIL.ImageEnhance / convert('HSV')) restated from Pillow's libImaging (Blend.c, Convert.c) as DMH_HD functions so
// that tests/host_emul_jitter.cpp can run the SAME code with g++ against the oracle (oracle/pil_enhance.py, itself
// bit-exact against Pillow) on a machine without a GPU.
#pragma once

#include <stdint.h>

#include "dmh_math.cuh"

namespace dmh {

struct Rgb8 { uint8_t r, g, b; };

// Convert.c rgb2l: ITU-R 601-2 luma in 16.16 fixed point
DMH_HD uint8_t jit_grey(Rgb8 p) { return (uint8_t)(((int)p.r * 19595 + (int)p.g * 38470 + (int)p.b * 7471 + 0x8000) >> 16); }

// Blend.c: (UINT8)(a + f * (b - a)) in single precision, each operation rounded; outside 0 <= f <= 1 the float is
// clipped to [0, 255] first.  f == 0 / f == 1 need no special case: the formula returns a / b exactly.
DMH_HD uint8_t jit_blend(uint8_t a, uint8_t b, float f) {
    const float t = add_rn((float)a, mul_rn(f, (float)((int)b - (int)a)));
    if (f >= 0.0f && f <= 1.0f) return (uint8_t)(int)t;
    return t <= 0.0f ? (uint8_t)0 : (t >= 255.0f ? (uint8_t)255 : (uint8_t)(int)t);
}

// ImageEnhance.Brightness / Contrast / Color: blend against black, the rounded mean grey level, the pixel's grey
DMH_HD Rgb8 jit_blend3(Rgb8 a, Rgb8 p, float f) {
    Rgb8 o;
    o.r = jit_blend(a.r, p.r, f); o.g = jit_blend(a.g, p.g, f); o.b = jit_blend(a.b, p.b, f);
    return o;
}
DMH_HD Rgb8 jit_brightness(Rgb8 p, float f) { const Rgb8 z = {0, 0, 0}; return jit_blend3(z, p, f); }
DMH_HD Rgb8 jit_contrast(Rgb8 p, float f, uint8_t mean_grey) { const Rgb8 m = {mean_grey, mean_grey, mean_grey}; return jit_blend3(m, p, f); }
DMH_HD Rgb8 jit_saturation(Rgb8 p, float f) { const uint8_t l = jit_grey(p); const Rgb8 m = {l, l, l}; return jit_blend3(m, p, f); }

DMH_HD int jit_clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// Convert.c rgb2hsv_row (after colorsys.rgb_to_hsv): float quotients, the branch arithmetic in double (the C source
// mixes double constants with float operands), result stored as float, fmod in double, truncation to bytes
DMH_HD Rgb8 jit_rgb2hsv(Rgb8 p) {
    const int r = p.r, g = p.g, b = p.b;
    const int maxc = r > g ? (r > b ? r : b) : (g > b ? g : b);
    const int minc = r < g ? (r < b ? r : b) : (g < b ? g : b);
    Rgb8 o;
    o.b = (uint8_t)maxc;                                   // V
    if (minc == maxc) { o.r = 0; o.g = 0; return o; }
    const float cr = (float)(maxc - minc);
    const float s = div_rn(cr, (float)maxc);
    const float rc = div_rn((float)(maxc - r), cr), gc = div_rn((float)(maxc - g), cr), bc = div_rn((float)(maxc - b), cr);
    float h;
    if (r == maxc) h = sub_rn(bc, gc);
    else if (g == maxc) h = (float)(2.0 + (double)rc - (double)bc);
    else h = (float)(4.0 + (double)gc - (double)rc);
    h = (float)fmod((double)h / 6.0 + 1.0, 1.0);
    o.r = (uint8_t)jit_clip8((int)((double)h * 255.0));    // H
    o.g = (uint8_t)jit_clip8((int)((double)s * 255.0));    // S
    return o;
}

// Convert.c hsv2rgb (after colorsys.hsv_to_rgb): sector / remainder in double, C round() (half away from zero)
DMH_HD Rgb8 jit_hsv2rgb(Rgb8 q) {
    const int hh = q.r, ss = q.g, v = q.b;
    Rgb8 o;
    if (ss == 0) { o.r = o.g = o.b = (uint8_t)v; return o; }
    const double hf = (double)(float)hh * 6.0 / 255.0;
    const int i = (int)floor(hf);
    const float f = (float)(hf - (double)(float)i);
    const float fs = (float)((double)(float)ss / 255.0);
    const double vf = (double)(float)v;
    const uint8_t up = (uint8_t)jit_clip8((int)round(vf * (1.0 - (double)fs)));
    const uint8_t uq = (uint8_t)jit_clip8((int)round(vf * (1.0 - (double)fs * (double)f)));
    const uint8_t ut = (uint8_t)jit_clip8((int)round(vf * (1.0 - (double)fs * (1.0 - (double)f))));
    const uint8_t uv = (uint8_t)v;
    switch (i % 6) {
        case 0: o.r = uv; o.g = ut; o.b = up; break;
        case 1: o.r = uq; o.g = uv; o.b = up; break;
        case 2: o.r = up; o.g = uv; o.b = ut; break;
        case 3: o.r = up; o.g = uq; o.b = uv; break;
        case 4: o.r = ut; o.g = up; o.b = uv; break;
        default: o.r = uv; o.g = up; o.b = uq; break;
    }
    return o;
}

// torchvision functional_pil.adjust_hue: H += uint8(f * 255) with 8-bit wrap-around (`shift` computed by the host)
DMH_HD Rgb8 jit_hue(Rgb8 p, uint8_t shift) {
    Rgb8 hsv = jit_rgb2hsv(p);
    hsv.r = (uint8_t)(hsv.r + shift);
    return jit_hsv2rgb(hsv);
}

}  // namespace dmh

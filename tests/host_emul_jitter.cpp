// TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product.
// Runs the per-pixel colour-jitter formulas of depthmodelhardening_b200/csrc/jitter_math.cuh with g++ over planar
// 8-bit images so that they can be checked against oracle/pil_enhance.py (bit-exact against Pillow) without a GPU.
// Build: g++ -O1 -ffp-contract=off -shared -fPIC -o tests/_build/libdmh_hostemu_jitter.so tests/host_emul_jitter.cpp
#include <cmath>
#include <cstdint>

#include "../depthmodelhardening_b200/csrc/jitter_math.cuh"

using namespace dmh;

extern "C" {

// op: 0 brightness, 1 contrast (aux = rounded mean grey), 2 saturation, 3 hue (aux = uint8 shift), 4 rgb->hsv,
// 5 hsv->rgb.  img / out: planar (3, n) bytes.
void emu_jitter(const uint8_t* img, long long n, int op, float f, int aux, uint8_t* out) {
    for (long long i = 0; i < n; ++i) {
        const Rgb8 p = {img[i], img[n + i], img[2 * n + i]};
        Rgb8 o;
        switch (op) {
            case 0: o = jit_brightness(p, f); break;
            case 1: o = jit_contrast(p, f, (uint8_t)aux); break;
            case 2: o = jit_saturation(p, f); break;
            case 3: o = jit_hue(p, (uint8_t)aux); break;
            case 4: o = jit_rgb2hsv(p); break;
            default: o = jit_hsv2rgb(p); break;
        }
        out[i] = o.r; out[n + i] = o.g; out[2 * n + i] = o.b;
    }
}

// sum of the grey levels (ImageStat.Stat(L).mean = sum / n in double; the contrast level is int(mean + 0.5))
long long emu_grey_sum(const uint8_t* img, long long n) {
    long long s = 0;
    for (long long i = 0; i < n; ++i) {
        const Rgb8 p = {img[i], img[n + i], img[2 * n + i]};
        s += jit_grey(p);
    }
    return s;
}

}  // extern "C"

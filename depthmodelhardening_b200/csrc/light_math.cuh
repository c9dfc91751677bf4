// Per-pixel arithmetic of the black-box attacks' candidate patches (next-4):
//   * tube light: torchattacks/attacks/light_simulation.py:132-170 (tube_light_generation_by_func) followed by
//     phy_obj_atk_light.py:118-121 / light_simulation.py:23-28 (light * 255 -> float32, one fp32 add onto the 8-bit
//     object image, clip, truncate to 8 bits);
//   * Square attack, L-inf candidate: phy_obj_atk_square.py:268-274.
// DMH_HD so that tests/host_emul_light.cpp runs the SAME code with g++ against oracle/light.py (pinned to the
// reference's own functions) on a machine without a GPU.  The reference computes the light field in Python floats:
// every operation below is a separately rounded float64 operation in the reference's order.
#pragma once

#include <stdint.h>

#include "dmh_math.cuh"

namespace dmh {

#if defined(__CUDA_ARCH__)
DMH_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
DMH_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
DMH_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
DMH_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
#else
DMH_HD double dmul(double a, double b) { volatile double r = a * b; return r; }
DMH_HD double dadd(double a, double b) { volatile double r = a + b; return r; }
DMH_HD double dsub(double a, double b) { volatile double r = a - b; return r; }
DMH_HD double ddiv(double a, double b) { volatile double r = a / b; return r; }
#endif

// The scalars of one candidate, formed on the host exactly as the reference forms them (python floats / ints).
struct TubeLight {
    double k, b;          // the beam's axis y = k*x + b
    double norm;          // math.sqrt(1 + k*k)
    double beta;          // attenuation parameter
    double full_end;      // int(math.sqrt(beta) + 0.5): half-width of the saturated core
    double light_end;     // int(math.sqrt(beta * 20) + 0.5): where the skirt ends
    double ca[3];         // wavelength_to_rgb(wavelength)[c] * alpha
};

// light_simulation.py:157-168 for one pixel: which zone it lies in and, in the skirt, the attenuation beta / d^2
enum { TUBE_DARK = 0, TUBE_CORE = 1, TUBE_SKIRT = 2 };
DMH_HD int tube_light_zone(const TubeLight& t, int x, int y, double* att) {
    const double lin = dadd(dsub(dmul(t.k, (double)x), (double)y), t.b);      // k*x - y + b
    const double dist = ddiv(lin < 0.0 ? -lin : lin, t.norm);                 // abs(.) / math.sqrt(1 + k*k)
    if (dist <= t.full_end) return TUBE_CORE;
    if (dist <= t.light_end) { *att = ddiv(t.beta, dmul(dist, dist)); return TUBE_SKIRT; }
    return TUBE_DARK;
}

// One channel of the lit 8-bit image (ca = colour * alpha of the channel).
DMH_HD uint8_t lit_u8(uint8_t base, double ca, int zone, double att) {
    const double light = zone == TUBE_CORE ? ca : (zone == TUBE_SKIRT ? dmul(ca, att) : 0.0);
    const float l32 = (float)dmul(light, 255.0);                  // (tube_light * 255.0).astype(float32)
    const float s = add_rn((float)base, l32);                     // cv2.addWeighted(base, 1, light, 1, 0)
    const float c = s < 0.0f ? 0.0f : (s > 255.0f ? 255.0f : s);  // np.clip(., 0, 255)
    return (uint8_t)(int)c;                                       // .astype('uint8')
}

// Square attack (L-inf) candidate for one element: phy_obj_atk_square.py:271-274
//   x_new = clamp(min(max(x_best + delta, x - eps), x + eps), 0, 1)
DMH_HD float square_linf_candidate(float x_best, float x, float delta, float eps) {
    float v = add_rn(x_best, delta);
    const float lo = sub_rn(x, eps), hi = add_rn(x, eps);
    v = v > lo ? v : lo;                                          // torch.max
    v = v < hi ? v : hi;                                          // torch.min
    return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
}

}  // namespace dmh

// TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product.
// Runs the per-pixel formulas of depthmodelhardening_b200/csrc/light_math.cuh (tube-light candidate, Square-attack
// candidate) with g++ so that they can be checked against oracle/light.py (pinned to the reference's own functions)
// and the torch expression without a GPU.
// Build: g++ -O1 -ffp-contract=off -shared -fPIC -o tests/_build/libdmh_hostemu_light.so tests/host_emul_light.cpp
#include <cstdint>

#include "../depthmodelhardening_b200/csrc/light_math.cuh"

using namespace dmh;

extern "C" {

// base / lit: planar (3, h, w) bytes; patch: (3, h, w) floats -- the loop body of tube_light_kernel
void emu_tube_light(const uint8_t* base, int h, int w, double k, double b, double norm, double beta, int full_end,
                    int light_end, double ca0, double ca1, double ca2, float* patch, uint8_t* lit) {
    TubeLight t;
    t.k = k; t.b = b; t.norm = norm; t.beta = beta; t.full_end = (double)full_end; t.light_end = (double)light_end;
    t.ca[0] = ca0; t.ca[1] = ca1; t.ca[2] = ca2;
    const int n = h * w;
    for (int i = 0; i < n; ++i) {
        const int y = i / w, x = i - y * w;
        double att = 0.0;
        const int zone = tube_light_zone(t, x, y, &att);
        for (int c = 0; c < 3; ++c) {
            const uint8_t v = lit_u8(base[c * n + i], t.ca[c], zone, att);
            patch[c * n + i] = div_rn((float)v, 255.0f);
            lit[c * n + i] = v;
        }
    }
}

// the loop body of square_candidate_kernel
void emu_square_candidate(const float* x_best, const float* x, int H, int W, int vh, int vw, int s, float d0, float d1,
                          float d2, float eps, float* x_new) {
    const int n = H * W;
    const float d[3] = {d0, d1, d2};
    for (int i = 0; i < n; ++i) {
        const int py = i / W, px = i - py * W;
        const bool in = py >= vh && py < vh + s && px >= vw && px < vw + s;
        for (int c = 0; c < 3; ++c)
            x_new[c * n + i] = square_linf_candidate(x_best[c * n + i], x[c * n + i], in ? d[c] : 0.0f, eps);
    }
}

}  // extern "C"

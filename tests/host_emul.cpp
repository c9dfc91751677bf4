// TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product.
//
// Compiles the SAME per-pixel formulas the CUDA kernels use
// (depthmodelhardening_b200/csrc/dmh_math.cuh) with g++ and runs them with plain
// loops over whole images, so that the maths (coordinate chain, bilinear
// gradient, SSIM coefficient form of the backward, reflection multiplicities,
// argmin/automask, depth chain rule) can be checked against the oracle on a
// machine without a GPU.  The tiling / shared-memory indexing of the real
// kernels is NOT exercised here; the `-m gpu` parity tests do that.
//
// Build: g++ -O1 -ffp-contract=off -shared -fPIC -o tests/_build/libdmh_hostemu.so tests/host_emul.cpp
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../depthmodelhardening_b200/csrc/dmh_math.cuh"

using namespace dmh;

static inline float tap(const float* s, long long o, bool in) { return in ? s[o] : 0.0f; }

struct TapSet { long long o00; bool nw, ne, sw, se; };
static TapSet taps(const Bilinear& bl, int H, int W) {
    TapSet t;
    bool x0 = bl.x0 >= 0 && bl.x0 < W, x1 = bl.x0 + 1 >= 0 && bl.x0 + 1 < W;
    bool y0 = bl.y0 >= 0 && bl.y0 < H, y1 = bl.y0 + 1 >= 0 && bl.y0 + 1 < H;
    t.o00 = (long long)bl.y0 * W + bl.x0;
    t.nw = y0 && x0; t.ne = y0 && x1; t.sw = y1 && x0; t.se = y1 && x1;
    return t;
}

static float mult(int p, int q, int n) {
    float m = 1.0f;
    if (p == 1 && q == 0) m += 1.0f;
    if (p == n - 2 && q == n - 1) m += 1.0f;
    return m;
}

extern "C" {

// A9-A12 forward: warped (B,C,H,W)
void emu_warp_fwd(const float* disp, const float* src, const float* K, const float* inv_K, const float* T, int B, int C,
                  int H, int W, float min_depth, float max_depth, float* warped) {
    DepthScale ds{(float)(1.0 / (double)max_depth), (float)(1.0 / (double)min_depth - 1.0 / (double)max_depth)};
    const long long N = (long long)H * W;
    for (int b = 0; b < B; ++b) {
        Camera cam;
        compose_camera(K + b * 16, T + b * 16, inv_K + b * 16, cam);
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const long long n = (long long)y * W + x;
                const float depth = disp_to_depth(disp[b * N + n], ds);
                WarpCoord wc = warp_coord(cam, (float)x, (float)y, depth, W, H, 1e-7f);
                Bilinear bl = bilinear_setup(wc.ix, wc.iy);
                TapSet t = taps(bl, H, W);
                for (int c = 0; c < C; ++c) {
                    const float* s = src + ((long long)b * C + c) * N;
                    float acc = 0.f;
                    if (t.nw) acc = fmaf(s[t.o00], bl.wnw, acc);
                    if (t.ne) acc = fmaf(s[t.o00 + 1], bl.wne, acc);
                    if (t.sw) acc = fmaf(s[t.o00 + W], bl.wsw, acc);
                    if (t.se) acc = fmaf(s[t.o00 + W + 1], bl.wse, acc);
                    warped[((long long)b * C + c) * N + n] = acc;
                }
            }
    }
}

// A9-A12 backward: grad_disp (B,1,H,W), grad_P (B,12)
void emu_warp_bwd(const float* gw, const float* disp, const float* src, const float* K, const float* inv_K,
                  const float* T, int B, int C, int H, int W, float min_depth, float max_depth, float* gdisp,
                  float* gP) {
    DepthScale ds{(float)(1.0 / (double)max_depth), (float)(1.0 / (double)min_depth - 1.0 / (double)max_depth)};
    const long long N = (long long)H * W;
    for (int b = 0; b < B; ++b) {
        Camera cam;
        compose_camera(K + b * 16, T + b * 16, inv_K + b * 16, cam);
        double accP[12] = {0};
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const long long n = (long long)y * W + x;
                const float depth = disp_to_depth(disp[b * N + n], ds);
                WarpCoord wc = warp_coord(cam, (float)x, (float)y, depth, W, H, 1e-7f);
                Bilinear bl = bilinear_setup(wc.ix, wc.iy);
                TapSet t = taps(bl, H, W);
                float gix = 0.f, giy = 0.f;
                for (int c = 0; c < C; ++c) {
                    const float* s = src + ((long long)b * C + c) * N;
                    const float go = gw[((long long)b * C + c) * N + n];
                    if (t.nw) { float v = s[t.o00];         gix -= v * bl.ty1 * go; giy -= v * bl.tx1 * go; }
                    if (t.ne) { float v = s[t.o00 + 1];     gix += v * bl.ty1 * go; giy -= v * bl.tx0 * go; }
                    if (t.sw) { float v = s[t.o00 + W];     gix -= v * bl.ty0 * go; giy += v * bl.tx1 * go; }
                    if (t.se) { float v = s[t.o00 + W + 1]; gix += v * bl.ty0 * go; giy += v * bl.tx0 * go; }
                }
                float dp[3];
                const float gd = warp_coord_bwd(cam, wc, gix, giy, W, H, dp);
                gdisp[b * N + n] = gd * ddepth_ddisp(depth, ds);
                const float pt[4] = {depth * wc.ray[0], depth * wc.ray[1], depth * wc.ray[2], 1.0f};
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 4; ++j) accP[i * 4 + j] += (double)dp[i] * pt[j];
            }
        if (gP)
            for (int k = 0; k < 12; ++k) gP[b * 12 + k] = (float)accP[k];
    }
}

static SsimStats stats_at(const float* xp, const float* yp, int px, int py, int H, int W) {
    float s1 = 0, s2 = 0, s11 = 0, s22 = 0, s12 = 0;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const int ry = reflect1(py + dy, H), rx = reflect1(px + dx, W);
            const float a = xp[(long long)ry * W + rx], d = yp[(long long)ry * W + rx];
            s1 = add_rn(s1, a); s2 = add_rn(s2, d);
            s11 = add_rn(s11, mul_rn(a, a)); s22 = add_rn(s22, mul_rn(d, d)); s12 = add_rn(s12, mul_rn(a, d));
        }
    return ssim_stats(s1, s2, s11, s22, s12);
}

// A13 forward + backward with a dense upstream gradient: out, gx, gy (B*C planes)
void emu_ssim(const float* x, const float* y, const float* gout, int planes, int H, int W, float* out, float* gx,
              float* gy) {
    const long long N = (long long)H * W;
    std::vector<float> kax(N), kay(N), kb(N), kc(N);
    for (int pl = 0; pl < planes; ++pl) {
        const float* xp = x + pl * N;
        const float* yp = y + pl * N;
        for (int py = 0; py < H; ++py)
            for (int px = 0; px < W; ++px) {
                const long long n = (long long)py * W + px;
                SsimStats st = stats_at(xp, yp, px, py, H, W);
                float pass;
                out[pl * N + n] = ssim_value(st, pass);
                SsimCoef k = ssim_coef(st);
                const float g = gout[pl * N + n] * pass;
                kax[n] = g * k.ax; kay[n] = g * k.ay; kb[n] = g * k.b; kc[n] = g * k.c;
            }
        for (int py = 0; py < H; ++py)
            for (int px = 0; px < W; ++px) {
                float ax = 0, ay = 0, sb = 0, sc = 0;
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int qy = py + dy, qx = px + dx;
                        if (qy < 0 || qy >= H || qx < 0 || qx >= W) continue;
                        const float w = mult(py, qy, H) * mult(px, qx, W);
                        const long long q = (long long)qy * W + qx;
                        ax = fmaf(w, kax[q], ax); ay = fmaf(w, kay[q], ay);
                        sb = fmaf(w, kb[q], sb); sc = fmaf(w, kc[q], sc);
                    }
                const long long n = (long long)py * W + px;
                gx[pl * N + n] = ax + sb * xp[n] + sc * yp[n];
                gy[pl * N + n] = ay + sb * yp[n] + sc * xp[n];
            }
    }
}

// A9-A15 for one scale: sum of to_optimise, grad wrt full-res disp, argmin.
// src: F pointers, T: F pointers; ident (B,F,H,W) or null; noise (B,Fi,H,W) or null.
void emu_photo_scale(const float* target, const float* const* src, const float* const* T, int F, const float* disp,
                     const float* K, const float* inv_K, const float* ident, const float* noise, int B, int H, int W,
                     float min_depth, float max_depth, int flags, double* loss_sum, float* gdisp, uint8_t* sel,
                     float* gP /* (F,B,12) or null */) {
    const bool no_ssim = flags & 1, avg = flags & 2;
    const float w_ssim = no_ssim ? 0.f : 0.85f / 3.f, w_l1 = no_ssim ? 1.f / 3.f : 0.15f / 3.f;
    DepthScale ds{(float)(1.0 / (double)max_depth), (float)(1.0 / (double)min_depth - 1.0 / (double)max_depth)};
    const long long N = (long long)H * W;
    const int Fi = ident ? (avg ? 1 : F) : 0;
    std::vector<float> pred((size_t)F * 3 * N), win(N), ka(3 * N), kb(3 * N), kc(3 * N);
    double total = 0.0;
    for (int b = 0; b < B; ++b) {
        std::vector<Camera> cams(F);
        for (int f = 0; f < F; ++f) {
            compose_camera(K + b * 16, T[f] + b * 16, inv_K + b * 16, cams[f]);
            for (long long n = 0; n < N; ++n) {
                const int x = (int)(n % W), y = (int)(n / W);
                const float depth = disp_to_depth(disp[b * N + n], ds);
                WarpCoord wc = warp_coord(cams[f], (float)x, (float)y, depth, W, H, 1e-7f);
                Bilinear bl = bilinear_setup(wc.ix, wc.iy);
                TapSet t = taps(bl, H, W);
                for (int c = 0; c < 3; ++c) {
                    const float* s = src[f] + ((long long)b * 3 + c) * N;
                    float acc = 0.f;
                    if (t.nw) acc = fmaf(s[t.o00], bl.wnw, acc);
                    if (t.ne) acc = fmaf(s[t.o00 + 1], bl.wne, acc);
                    if (t.sw) acc = fmaf(s[t.o00 + W], bl.wsw, acc);
                    if (t.se) acc = fmaf(s[t.o00 + W + 1], bl.wse, acc);
                    pred[((size_t)f * 3 + c) * N + n] = acc;
                }
            }
        }
        for (int qy = 0; qy < H; ++qy)
            for (int qx = 0; qx < W; ++qx) {
                const long long q = (long long)qy * W + qx;
                float rp[8], rp_avg = 0.f;
                for (int f = 0; f < F; ++f) {
                    float l1 = 0.f, ss = 0.f;
                    for (int c = 0; c < 3; ++c) {
                        const float* xs = &pred[((size_t)f * 3 + c) * N];
                        const float* ys = target + ((long long)b * 3 + c) * N;
                        l1 = add_rn(l1, fabsf(sub_rn(ys[q], xs[q])));
                        if (!no_ssim) { float pass; ss = add_rn(ss, ssim_value(stats_at(xs, ys, qx, qy, H, W), pass)); }
                    }
                    l1 = div_rn(l1, 3.f);
                    rp[f] = no_ssim ? l1 : add_rn(mul_rn(0.85f, div_rn(ss, 3.f)), mul_rn(0.15f, l1));
                    rp_avg = add_rn(rp_avg, rp[f]);
                }
                rp_avg = div_rn(rp_avg, (float)F);
                float best = 3.4e38f, w = -1.f;
                int best_idx = 0, idx = 0;
                if (Fi > 0) {
                    if (avg) {
                        float s = 0.f;
                        for (int f = 0; f < F; ++f) s = add_rn(s, ident[((long long)b * F + f) * N + q]);
                        float v = div_rn(s, (float)F);
                        if (noise) v = add_rn(v, noise[(long long)b * N + q]);
                        best = v; idx = 1;
                    } else {
                        for (int f = 0; f < F; ++f) {
                            float v = ident[((long long)b * F + f) * N + q];
                            if (noise) v = add_rn(v, noise[((long long)b * F + f) * N + q]);
                            if (idx == 0 || v < best) { best = v; best_idx = idx; }
                            ++idx;
                        }
                    }
                }
                if (avg) {
                    if (idx == 0 || rp_avg < best) { best = rp_avg; best_idx = idx; w = 0.f; }
                } else {
                    for (int f = 0; f < F; ++f) {
                        if (idx == 0 || rp[f] < best) { best = rp[f]; best_idx = idx; w = (float)f; }
                        ++idx;
                    }
                }
                total += best;
                win[q] = w;
                if (sel) sel[b * N + q] = (uint8_t)best_idx;
            }
        std::vector<float> gacc(N, 0.f);
        for (int f = 0; f < F; ++f) {
            for (int qy = 0; qy < H; ++qy)
                for (int qx = 0; qx < W; ++qx) {
                    const long long q = (long long)qy * W + qx;
                    const float gate = avg ? (win[q] >= 0.f ? 1.f / (float)F : 0.f) : (win[q] == (float)f ? 1.f : 0.f);
                    for (int c = 0; c < 3; ++c) {
                        float a = 0, bq = 0, cq = 0;
                        if (gate != 0.f && !no_ssim) {
                            SsimStats st = stats_at(&pred[((size_t)f * 3 + c) * N], target + ((long long)b * 3 + c) * N,
                                                    qx, qy, H, W);
                            float pass;
                            ssim_value(st, pass);
                            SsimCoef k = ssim_coef(st);
                            const float g = gate * w_ssim * pass;
                            a = g * k.ax; bq = g * k.b; cq = g * k.c;
                        }
                        ka[c * N + q] = a; kb[c * N + q] = bq; kc[c * N + q] = cq;
                    }
                }
            double accP[12] = {0};
            for (int py = 0; py < H; ++py)
                for (int px = 0; px < W; ++px) {
                    const long long n = (long long)py * W + px;
                    const float gate = avg ? (win[n] >= 0.f ? 1.f / (float)F : 0.f) : (win[n] == (float)f ? 1.f : 0.f);
                    float g_pred[3];
                    for (int c = 0; c < 3; ++c) {
                        float sa = 0, sb = 0, sc = 0;
                        for (int dy = -1; dy <= 1; ++dy)
                            for (int dx = -1; dx <= 1; ++dx) {
                                const int qy = py + dy, qx = px + dx;
                                if (qy < 0 || qy >= H || qx < 0 || qx >= W) continue;
                                const float wt = mult(py, qy, H) * mult(px, qx, W);
                                const long long q = (long long)qy * W + qx;
                                sa = fmaf(wt, ka[c * N + q], sa); sb = fmaf(wt, kb[c * N + q], sb);
                                sc = fmaf(wt, kc[c * N + q], sc);
                            }
                        const float xv = pred[((size_t)f * 3 + c) * N + n], yv = target[((long long)b * 3 + c) * N + n];
                        const float d = xv - yv;
                        const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
                        g_pred[c] = sa + sb * xv + sc * yv + gate * w_l1 * sg;
                    }
                    const float depth = disp_to_depth(disp[b * N + n], ds);
                    WarpCoord wc = warp_coord(cams[f], (float)px, (float)py, depth, W, H, 1e-7f);
                    Bilinear bl = bilinear_setup(wc.ix, wc.iy);
                    TapSet t = taps(bl, H, W);
                    float gix = 0.f, giy = 0.f;
                    for (int c = 0; c < 3; ++c) {
                        const float* s = src[f] + ((long long)b * 3 + c) * N;
                        const float go = g_pred[c];
                        if (t.nw) { float v = s[t.o00];         gix -= v * bl.ty1 * go; giy -= v * bl.tx1 * go; }
                        if (t.ne) { float v = s[t.o00 + 1];     gix += v * bl.ty1 * go; giy -= v * bl.tx0 * go; }
                        if (t.sw) { float v = s[t.o00 + W];     gix -= v * bl.ty0 * go; giy += v * bl.tx1 * go; }
                        if (t.se) { float v = s[t.o00 + W + 1]; gix += v * bl.ty0 * go; giy += v * bl.tx0 * go; }
                    }
                    float dp[3];
                    gacc[n] += warp_coord_bwd(cams[f], wc, gix, giy, W, H, dp);
                    const float pt[4] = {depth * wc.ray[0], depth * wc.ray[1], depth * wc.ray[2], 1.0f};
                    for (int i = 0; i < 3; ++i)
                        for (int j = 0; j < 4; ++j) accP[i * 4 + j] += (double)dp[i] * pt[j];
                }
            if (gP)
                for (int k = 0; k < 12; ++k) gP[((size_t)f * B + b) * 12 + k] = (float)accP[k];
        }
        for (long long n = 0; n < N; ++n) {
            const float depth = disp_to_depth(disp[b * N + n], ds);
            gdisp[b * N + n] = gacc[n] * ddepth_ddisp(depth, ds);
        }
    }
    *loss_sum = total;
}


// Single-source fast path arithmetic (photo_fast.cu): separable sliding sums,
// sum*(1/9) means, shared reciprocal, chain collapsed into 3 scalars per pixel.
void emu_photo_scale_fast(const float* target, const float* src, const float* T, const float* disp, const float* K,
                          const float* inv_K, const float* ident, const float* noise, int B, int H, int W,
                          float min_depth, float max_depth, int flags, double* loss_sum, float* gdisp, uint8_t* sel) {
    const bool no_ssim = flags & 1;
    const float w_ssim = no_ssim ? 0.f : 0.85f / 3.f, w_l1 = no_ssim ? 1.f / 3.f : 0.15f / 3.f;
    DepthScale ds{(float)(1.0 / (double)max_depth), (float)(1.0 / (double)min_depth - 1.0 / (double)max_depth)};
    const long long N = (long long)H * W;
    std::vector<float> pred(3 * N), D(3 * N), coef(9 * N), gate(N);
    double total = 0.0;
    auto PX = [&](const std::vector<float>& a, int ch, int y, int x) { return a[ch * N + (long long)reflect1(y, H) * W + reflect1(x, W)]; };
    for (int b = 0; b < B; ++b) {
        Camera cam;
        compose_camera(K + b * 16, T + b * 16, inv_K + b * 16, cam);
        const float* tg = target + (long long)b * 3 * N;
        for (long long n = 0; n < N; ++n) {
            const int x = (int)(n % W), y = (int)(n / W);
            const float depth = disp_to_depth(disp[b * N + n], ds);
            WarpCoord wc = warp_coord(cam, (float)x, (float)y, depth, W, H, 1e-7f);
            Bilinear bl = bilinear_setup(wc.ix, wc.iy);
            TapSet t = taps(bl, H, W);
            float ax, ay;
            warp_chain_factors(cam, wc, W, H, ax, ay);
            const float dd = ddepth_ddisp(depth, ds);
            for (int c = 0; c < 3; ++c) {
                const float* s = src + ((long long)b * 3 + c) * N;
                const float nw = tap(s, t.o00, t.nw), ne = tap(s, t.o00 + 1, t.ne), sw = tap(s, t.o00 + W, t.sw),
                            se = tap(s, t.o00 + W + 1, t.se);
                float acc = nw * bl.wnw;
                acc = fmaf(ne, bl.wne, acc); acc = fmaf(sw, bl.wsw, acc); acc = fmaf(se, bl.wse, acc);
                pred[c * N + n] = acc;
                const float dix = (ne - nw) * bl.ty1 + (se - sw) * bl.ty0, diy = (sw - nw) * bl.tx1 + (se - ne) * bl.tx0;
                D[c * N + n] = (dix * ax + diy * ay) * dd;
            }
        }
        for (int qy = 0; qy < H; ++qy)
            for (int qx = 0; qx < W; ++qx) {
                const long long q = (long long)qy * W + qx;
                float l1 = 0.f, ss = 0.f, ka[3] = {0, 0, 0}, kb[3] = {0, 0, 0}, kc[3] = {0, 0, 0};
                for (int c = 0; c < 3; ++c) {
                    l1 += fabsf(tg[c * N + q] - pred[c * N + q]);
                    if (!no_ssim) {
                        Row5 rows[3];
                        for (int dy = -1; dy <= 1; ++dy) {
                            auto TG = [&](int yy, int xx) { return tg[c * N + (long long)reflect1(yy, H) * W + reflect1(xx, W)]; };
                            rows[dy + 1] = row5(PX(pred, c, qy + dy, qx - 1), PX(pred, c, qy + dy, qx), PX(pred, c, qy + dy, qx + 1),
                                                TG(qy + dy, qx - 1), TG(qy + dy, qx), TG(qy + dy, qx + 1));
                        }
                        SsimStatsRows st = ssim_stats_rows(rows[0], rows[1], rows[2]);
                        float pass; SsimCoef k;
                        ss += ssim_value_coef(st, pass, k);
                        const float g = w_ssim * pass;
                        ka[c] = g * k.ax; kb[c] = g * k.b; kc[c] = g * k.c;
                    }
                }
                l1 *= (1.0f / 3.0f);
                const float rp = no_ssim ? l1 : fmaf(0.85f, ss * (1.0f / 3.0f), 0.15f * l1);
                float best = rp; int best_idx = 0; bool win = true;
                if (ident) {
                    float idv = ident[b * N + q];
                    if (noise) idv = add_rn(idv, noise[b * N + q]);
                    win = rp < idv; best = win ? rp : idv; best_idx = win ? 1 : 0;
                }
                gate[q] = win ? 1.f : 0.f;
                for (int c = 0; c < 3; ++c) {
                    coef[(c * 3 + 0) * N + q] = win ? ka[c] : 0.f;
                    coef[(c * 3 + 1) * N + q] = win ? kb[c] : 0.f;
                    coef[(c * 3 + 2) * N + q] = win ? kc[c] : 0.f;
                }
                total += best;
                if (sel) sel[b * N + q] = (uint8_t)best_idx;
            }
        auto CF = [&](int pl, int y, int x) -> float { return (y < 0 || y >= H || x < 0 || x >= W) ? 0.f : coef[pl * N + (long long)y * W + x]; };
        for (int py = 0; py < H; ++py)
            for (int px = 0; px < W; ++px) {
                const long long n = (long long)py * W + px;
                const float wl = px == 1 ? 2.f : 1.f, wr = px == W - 2 ? 2.f : 1.f, wu = py == 1 ? 2.f : 1.f, wd = py == H - 2 ? 2.f : 1.f;
                float g = 0.f;
                for (int c = 0; c < 3; ++c) {
                    float s3[3];
                    for (int j = 0; j < 3; ++j) {
                        const int pl = c * 3 + j;
                        float h[3];
                        for (int dy = -1; dy <= 1; ++dy)
                            h[dy + 1] = fmaf(wl, CF(pl, py + dy, px - 1), fmaf(wr, CF(pl, py + dy, px + 1), CF(pl, py + dy, px)));
                        s3[j] = fmaf(wu, h[0], fmaf(wd, h[2], h[1]));
                    }
                    const float xv = pred[c * N + n], yv = tg[c * N + n];
                    const float d = xv - yv;
                    const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
                    const float g_pred = fmaf(s3[1], xv, fmaf(s3[2], yv, s3[0])) + gate[n] * w_l1 * sg;
                    g = fmaf(g_pred, D[c * N + n], g);
                }
                gdisp[b * N + n] = g;
            }
    }
    *loss_sum = total;
}

}  // extern "C"

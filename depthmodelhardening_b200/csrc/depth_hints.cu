// A18 -- depth-hints objective (DepthNetworks/depth-hints/trainer.py), per scale:
//
//   dmh_hint_select   trainer.py:666-712 in ONE pass over the pixels:
//       reprojection_loss = min_f (or mean_f) of the per-frame reprojection losses   (:681-685)
//       identity          = min_f (or mean_f) of the identity losses + tie-break noise (:666-672, 687-690)
//       idx               = argmin [reprojection, identity, depth-hint reprojection]   (:541-590 compute_loss_masks)
//       reprojection mask = idx != 1, depth-hint mask = idx == 2
//       numerators / denominators of the two masked means                              (:699-700, 712-713)
//       proxy loss        = log(|hint - depth| + 1) * valid * hint mask                (:525-539)
//   and their derivatives w.r.t. every per-frame reprojection loss and w.r.t. the
//   predicted depth (the masks are piecewise constant).  The masked-mean
//   denominators are only known after the whole pass, so the derivatives are those of
//   the NUMERATORS; the autograd wrapper divides by (sum(mask) + 1e-7) on the device.
//
// Roofline: HBM.  Algorithmic bytes per pixel: 4F (reproj) + 4F (ident) + 4 (noise)
// + 4 (hint loss) + 12 (depth, hint, valid) read; 4F + 4 (+1 sel) written.
#include "../../include/dmh_b200.h"
#include "dmh_common.cuh"

using namespace dmh;

namespace {

#define HS_THREADS 256
#define HS_PER_THREAD 4
#define HS_MAXF DMH_PHOTO_MAX_FRAMES

struct HintParams {
    const float* reproj[HS_MAXF];
    float* g_reproj[HS_MAXF];
    const float* ident;
    const float* noise;
    const float* hint_reproj;
    const float* depth;
    const float* hint_depth;
    const float* hint_valid;
    float* g_depth;
    float* part;              // [4][nblk]
    uint8_t* sel;
    long long n, HW;
    int F, avg;
};

__global__ void __launch_bounds__(HS_THREADS)
hint_select_kernel(const HintParams p) {
    __shared__ float red[32];
    float s_r = 0.f, s_rm = 0.f, s_h = 0.f, s_hm = 0.f;
    const long long base = ((long long)blockIdx.x * HS_THREADS) * HS_PER_THREAD + threadIdx.x;
#pragma unroll
    for (int k = 0; k < HS_PER_THREAD; ++k) {
        const long long i = base + (long long)k * HS_THREADS;
        if (i >= p.n) continue;
        const long long b = i / p.HW, px = i - b * p.HW;
        // reprojection candidate: first minimum over the frames (torch.min) or their mean
        float rp = __ldg(p.reproj[0] + i);
        int fbest = 0;
        if (p.avg) {
            for (int f = 1; f < p.F; ++f) rp = add_rn(rp, __ldg(p.reproj[f] + i));
            rp = div_rn(rp, (float)p.F);
        } else {
            for (int f = 1; f < p.F; ++f) {
                const float v = __ldg(p.reproj[f] + i);
                if (v < rp) { rp = v; fbest = f; }
            }
        }
        int idx = 0;
        float best = rp;
        if (p.ident) {
            const float* ip = p.ident + (b * p.F) * p.HW + px;
            float idv = __ldg(ip);
            if (p.avg) {
                for (int f = 1; f < p.F; ++f) idv = add_rn(idv, __ldg(ip + (long long)f * p.HW));
                idv = div_rn(idv, (float)p.F);
            } else {
                for (int f = 1; f < p.F; ++f) idv = fminf(idv, __ldg(ip + (long long)f * p.HW));
            }
            if (p.noise) idv = add_rn(idv, __ldg(p.noise + i));
            if (idv < best) { best = idv; idx = 1; }
        }
        if (p.hint_reproj) {
            const float hv = __ldg(p.hint_reproj + i);
            if (hv < best) { best = hv; idx = 2; }
        }
        const float m_r = (idx != 1) ? 1.f : 0.f;
        s_r += rp * m_r;
        s_rm += m_r;
        for (int f = 0; f < p.F; ++f)
            p.g_reproj[f][i] = p.avg ? m_r / (float)p.F : ((f == fbest) ? m_r : 0.f);
        if (p.hint_reproj) {
            const float m_h = (idx == 2) ? 1.f : 0.f;
            const float d = __ldg(p.depth + i), hd = __ldg(p.hint_depth + i), va = __ldg(p.hint_valid + i);
            const float diff = sub_rn(hd, d);
            const float a1 = add_rn(fabsf(diff), 1.0f);
            s_h += mul_rn(mul_rn(logf(a1), va), m_h);
            s_hm += m_h;
            // d/d(depth) of log(|hint - depth| + 1): -sign(hint - depth) / (|hint - depth| + 1)
            const float sg = diff > 0.f ? -1.f : (diff < 0.f ? 1.f : 0.f);
            p.g_depth[i] = m_h * va * sg / a1;
        }
        if (p.sel) p.sel[i] = (uint8_t)idx;
    }
    const int nblk = gridDim.x;
    float t = block_sum(s_r, red);
    if (threadIdx.x == 0) p.part[blockIdx.x] = t;
    t = block_sum(s_rm, red);
    if (threadIdx.x == 0) p.part[nblk + blockIdx.x] = t;
    t = block_sum(s_h, red);
    if (threadIdx.x == 0) p.part[2 * nblk + blockIdx.x] = t;
    t = block_sum(s_hm, red);
    if (threadIdx.x == 0) p.part[3 * nblk + blockIdx.x] = t;
}

__global__ void axpby_dev_kernel(const float* __restrict__ a, const float* __restrict__ x, const float* __restrict__ b,
                                 const float* __restrict__ y, long long n, float* __restrict__ out) {
    const float av = a[0], bv = b[0];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = fmaf(av, x[i], bv * y[i]);
}

}  // namespace

extern "C" {

int dmh_axpby_dev(const float* a, const float* x, const float* b, const float* y, long long n, float* out,
                  dmh_stream_t stream) {
    DMH_REQUIRE(a && x && b && y && out && n > 0, "dmh_axpby_dev: null pointer or empty");
    const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    DMH_LAUNCH(axpby_dev_kernel, blocks, 256, 0, (cudaStream_t)stream)(a, x, b, y, n, out);
    DMH_CHECK_LAUNCH("dmh_axpby_dev");
    return DMH_OK;
}

int dmh_hint_select_blocks(int B, int H, int W) {
    return ceil_div((long long)B * H * W, (long long)HS_THREADS * HS_PER_THREAD);
}

int dmh_hint_select(const float* const* reproj_host, int F, const float* ident, const float* noise,
                    const float* hint_reproj, const float* depth, const float* hint_depth, const float* hint_valid,
                    int avg_reprojection, int B, int H, int W, float* part, float* const* g_reproj_host,
                    float* g_depth, unsigned char* sel, dmh_stream_t stream) {
    DMH_REQUIRE(reproj_host && g_reproj_host && part, "dmh_hint_select: null pointer");
    DMH_REQUIRE(F >= 1 && F <= HS_MAXF, "dmh_hint_select: F=%d outside [1,%d]", F, HS_MAXF);
    DMH_REQUIRE(B > 0 && H > 0 && W > 0, "dmh_hint_select: bad shape");
    DMH_REQUIRE(!hint_reproj || (depth && hint_depth && hint_valid && g_depth),
                "dmh_hint_select: depth hints need depth, hint_depth, hint_valid and g_depth");
    DMH_REQUIRE(!noise || ident, "dmh_hint_select: noise without identity losses");
    HintParams p;
    for (int f = 0; f < HS_MAXF; ++f) {
        p.reproj[f] = f < F ? reproj_host[f] : nullptr;
        p.g_reproj[f] = f < F ? g_reproj_host[f] : nullptr;
        DMH_REQUIRE(f >= F || (p.reproj[f] && p.g_reproj[f]), "dmh_hint_select: null buffer for frame %d", f);
    }
    p.ident = ident; p.noise = noise; p.hint_reproj = hint_reproj; p.depth = depth; p.hint_depth = hint_depth;
    p.hint_valid = hint_valid; p.g_depth = g_depth; p.part = part; p.sel = sel;
    p.n = (long long)B * H * W; p.HW = (long long)H * W; p.F = F; p.avg = avg_reprojection ? 1 : 0;
    DMH_LAUNCH(hint_select_kernel, dmh_hint_select_blocks(B, H, W), HS_THREADS, 0, (cudaStream_t)stream)(p);
    DMH_CHECK_LAUNCH("dmh_hint_select");
    return DMH_OK;
}

}  // extern "C"

// next-3 -- ManyDepth cost volume (DepthNetworks/manydepth2/networks/resnet_encoder.py:157-236,
// ResnetEncoderMatching.match_features): for every depth bin d and every lookup frame l,
//   back-project the 1/4-resolution pixel grid with the constant depth plane d, project it into the lookup
//   view (layers.py BackprojectDepth / Project3D, same op-by-op rounding as the photometric warp),
//   bilinearly sample the C-channel lookup features (grid_sample zeros, align_corners=True), take the
//   channel mean of |warped - current|, mask samples landing within 2 px of either border, average over the
//   lookup frames that contributed, flag empty cells and (set_missing_to_max) fill them with the per-pixel
//   maximum over the bins.
// The reference does this with B x L x ~25 ATen launches over (D, C, h, w) temporaries (96 x 16 x 80 x 256
// floats = 126 MB per lookup frame at 1024x320); here one thread owns one pixel, keeps its C current
// features in registers and walks the D bins, so nothing but the (B, D, h, w) volume is written.
// Lookup features are first transposed to channel-last so that one tap is 4 x 128-bit loads.
//
// Roofline: L1/L2 gather bandwidth (each pixel issues D * L * 4 taps * C loads against ~1.3 MB of lookup
// features that stay cache resident); algorithmic HBM bytes: read (1 + L) * C * 4 per pixel, write 2 * D * 4.
#include "../../include/dmh_b200.h"
#include "dmh_common.cuh"

using namespace dmh;

namespace {

#define CV_C 16
#define CV_MAXL 4

// (B*L, C, h*w) -> (B*L, h*w, C)
__global__ void cv_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int hw) {
    __shared__ float tile[CV_C][33];
    const int n = blockIdx.y;
    const int p0 = blockIdx.x * 32;
    const float* src = in + (size_t)n * CV_C * hw;
    for (int i = threadIdx.x; i < CV_C * 32; i += blockDim.x) {
        const int c = i >> 5, p = i & 31;
        tile[c][p] = (p0 + p < hw) ? __ldg(src + (size_t)c * hw + p0 + p) : 0.f;
    }
    __syncthreads();
    float* dst = out + ((size_t)n * hw + p0) * CV_C;
    for (int i = threadIdx.x; i < CV_C * 32; i += blockDim.x) {
        const int p = i / CV_C, c = i % CV_C;
        if (p0 + p < hw) dst[i] = tile[c][p];
    }
}

struct CvParams {
    const float* cur;        // (B, C, h, w)
    const float* lookT;      // (B, L, h, w, C) channel-last
    const float* poses;      // (B, L, 4, 4)
    const float* K;          // (B, 4, 4)
    const float* inv_K;
    const float* bins;       // (D)
    float* cost;             // (B, D, h, w)
    float* missing;          // (B, D, h, w)
    int B, L, D, h, w, fill_max;
};

__global__ void __launch_bounds__(128)
cost_volume_kernel(const CvParams p) {
    __shared__ float cams[CV_MAXL][24];
    __shared__ int live[CV_MAXL];
    const int b = blockIdx.z;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid < p.L * 21) {
        const int l = tid / 21, t = tid % 21;
        const float* T = p.poses + ((size_t)b * p.L + l) * 16;
        if (t < 12) {
            const int i = t / 4, j = t % 4;
            const float* k = p.K + b * 16 + i * 4;
            float acc = __ldg(k) * __ldg(T + j);
            acc = fmaf(__ldg(k + 1), __ldg(T + 4 + j), acc);
            acc = fmaf(__ldg(k + 2), __ldg(T + 8 + j), acc);
            acc = fmaf(__ldg(k + 3), __ldg(T + 12 + j), acc);
            cams[l][t] = acc;
        } else {
            const int i = (t - 12) / 3, j = (t - 12) % 3;
            cams[l][t] = __ldg(p.inv_K + b * 16 + i * 4 + j);
        }
    }
    if (tid < p.L) {
        // "ignore missing images": lookup_pose.sum() == 0 (resnet_encoder.py:190-192)
        const float* T = p.poses + ((size_t)b * p.L + tid) * 16;
        float s = 0.f;
        for (int i = 0; i < 16; ++i) s += __ldg(T + i);
        live[tid] = (s != 0.f) ? 1 : 0;
    }
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int h = p.h, w = p.w;
    if (x >= w || y >= h) return;
    const size_t hw = (size_t)h * w;
    const size_t pix = (size_t)y * w + x;
    float cur[CV_C];
#pragma unroll
    for (int c = 0; c < CV_C; ++c) cur[c] = __ldg(p.cur + ((size_t)b * CV_C + c) * hw + pix);
    // masking of the current image border (resnet_encoder.py:210-212)
    const bool cur_ok = x >= 2 && x < w - 2 && y >= 2 && y < h - 2;
    float* cost = p.cost + (size_t)b * p.D * hw + pix;
    float* miss = p.missing + (size_t)b * p.D * hw + pix;
    float vmax = -INFINITY;
    bool any_missing = false;
    for (int d = 0; d < p.D; ++d) {
        const float depth = __ldg(p.bins + d);
        float total = 0.f, count = 0.f;
        for (int l = 0; l < p.L; ++l) {
            if (!live[l]) continue;
            Camera cam;
#pragma unroll
            for (int i = 0; i < 12; ++i) cam.P[i] = cams[l][i];
#pragma unroll
            for (int i = 0; i < 9; ++i) cam.iK[i] = cams[l][12 + i];
            // BackprojectDepth + Project3D, op by op (layers.py:163-168, 182-198)
            float ray[3], pr[3];
            pixel_ray(cam, (float)x, (float)y, ray);
            const float pt[3] = {mul_rn(depth, ray[0]), mul_rn(depth, ray[1]), mul_rn(depth, ray[2])};
            project_point(cam, pt, pr);
            const float z = add_rn(pr[2], 1e-7f);
            const float gx = normalise_coord(pr[0], z, w), gy = normalise_coord(pr[1], z, h);
            // edge mask on the lookup side (resnet_encoder.py:203-208): its own un-normalisation formula
            const float xv = mul_rn(add_rn(mul_rn(gx, 0.5f), 0.5f), (float)(w - 1));
            const float yv = mul_rn(add_rn(mul_rn(gy, 0.5f), 0.5f), (float)(h - 1));
            const bool ok = cur_ok && xv >= 2.0f && xv <= (float)(w - 2) && yv >= 2.0f && yv <= (float)(h - 2);
            float diff = 0.f;
            if (ok) {   // (masked samples contribute exactly 0 whatever the features are)
                const float ix = safe_coord(unnormalise_coord(gx, w, true));
                const float iy = safe_coord(unnormalise_coord(gy, h, true));
                const Bilinear bl = bilinear_setup(ix, iy);
                // inside the mask 2 <= ix <= w-2: all four taps are in bounds
                const float4* t00 = reinterpret_cast<const float4*>(
                    p.lookT + ((((size_t)b * p.L + l) * h + bl.y0) * w + bl.x0) * CV_C);
                const bool x1in = bl.x0 + 1 < w, y1in = bl.y0 + 1 < h;
                const float4* t01 = t00 + (x1in ? CV_C / 4 : 0);
                const float4* t10 = t00 + (y1in ? (size_t)w * (CV_C / 4) : 0);
                const float4* t11 = t10 + (x1in ? CV_C / 4 : 0);
                const float wne = x1in ? bl.wne : 0.f, wsw = y1in ? bl.wsw : 0.f, wse = (x1in && y1in) ? bl.wse : 0.f;
                float acc = 0.f;
#pragma unroll
                for (int q = 0; q < CV_C / 4; ++q) {
                    const float4 a = __ldg(t00 + q), bq = __ldg(t01 + q), c = __ldg(t10 + q), e = __ldg(t11 + q);
                    float v;
                    v = a.x * bl.wnw; v = fmaf(bq.x, wne, v); v = fmaf(c.x, wsw, v); v = fmaf(e.x, wse, v);
                    acc += fabsf(v - cur[4 * q + 0]);
                    v = a.y * bl.wnw; v = fmaf(bq.y, wne, v); v = fmaf(c.y, wsw, v); v = fmaf(e.y, wse, v);
                    acc += fabsf(v - cur[4 * q + 1]);
                    v = a.z * bl.wnw; v = fmaf(bq.z, wne, v); v = fmaf(c.z, wsw, v); v = fmaf(e.z, wse, v);
                    acc += fabsf(v - cur[4 * q + 2]);
                    v = a.w * bl.wnw; v = fmaf(bq.w, wne, v); v = fmaf(c.w, wsw, v); v = fmaf(e.w, wse, v);
                    acc += fabsf(v - cur[4 * q + 3]);
                }
                diff = acc * (1.0f / CV_C);
            }
            total = add_rn(total, diff);
            count += (diff > 0.f) ? 1.f : 0.f;
        }
        const float cv = div_rn(total, add_rn(count, 1e-7f));
        cost[(size_t)d * hw] = cv;
        vmax = fmaxf(vmax, cv);
        any_missing |= (cv == 0.f);
    }
    // missing cells (resnet_encoder.py:224-231): flag, and fill with the per-pixel maximum over the bins
    for (int d = 0; d < p.D; ++d) {
        float cv = 0.f;
        bool m = false;
        if (any_missing) {
            cv = cost[(size_t)d * hw];
            m = (cv == 0.f);
            if (m && p.fill_max) cost[(size_t)d * hw] = vmax;
        }
        miss[(size_t)d * hw] = m ? 1.f : 0.f;
    }
}

}  // namespace

extern "C" {

long long dmh_cost_volume_workspace_floats(int B, int L, int C, int h, int w) {
    return (long long)B * L * C * h * w;
}

int dmh_cost_volume(const float* current_feats, const float* lookup_feats, const float* poses, const float* K,
                    const float* inv_K, const float* depth_bins, int B, int L, int C, int D, int h, int w,
                    int set_missing_to_max, float* workspace, float* cost_volume, float* missing_mask,
                    dmh_stream_t stream) {
    DMH_REQUIRE(current_feats && lookup_feats && poses && K && inv_K && depth_bins && workspace && cost_volume &&
                    missing_mask, "dmh_cost_volume: null pointer");
    DMH_REQUIRE(C == CV_C, "dmh_cost_volume: C=%d unsupported (the matching features have %d channels)", C, CV_C);
    DMH_REQUIRE(L >= 1 && L <= CV_MAXL, "dmh_cost_volume: L=%d outside [1,%d]", L, CV_MAXL);
    DMH_REQUIRE(B > 0 && B <= 65535 && D > 0 && h >= 5 && w >= 5, "dmh_cost_volume: bad shape");
    DMH_REQUIRE(((uintptr_t)workspace & 15) == 0, "dmh_cost_volume: workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = h * w;
    DMH_LAUNCH(cv_transpose_kernel, dim3(ceil_div(hw, 32), B * L), 256, 0, st)(lookup_feats, workspace, hw);
    CvParams p;
    p.cur = current_feats; p.lookT = workspace; p.poses = poses; p.K = K; p.inv_K = inv_K; p.bins = depth_bins;
    p.cost = cost_volume; p.missing = missing_mask; p.B = B; p.L = L; p.D = D; p.h = h; p.w = w;
    p.fill_max = set_missing_to_max ? 1 : 0;
    dim3 block(32, 4), grid(ceil_div(w, 32), ceil_div(h, 4), B);
    DMH_LAUNCH(cost_volume_kernel, grid, block, 0, st)(p);
    DMH_CHECK_LAUNCH("dmh_cost_volume");
    return DMH_OK;
}

}  // extern "C"

// Glue kernels of the fused multi-scale objective (M2/trainer.py:589-674):
//
//   dmh_smooth_fused      A16 forward AND the gradient w.r.t. the normalised
//                         disparity in one pass per scale (per-block partials)
//   dmh_objective_finish  one launch for ALL scales: per-image mean / correction
//                         scalars of the normalisation backward, the per-scale
//                         losses and the total (deterministic, fp64 accumulate)
//   dmh_disp_grad         backward: one launch per scale turns the speculative
//                         gradients into d(total)/d(disp_s) =
//                           u_s * [ F.interpolate^T(grad wrt up-sampled disp)
//                                   + w_s * (gN / (mean+eps) - corr) ]
//                         where u_s is the incoming autograd scalar(s), read on
//                         the device (no host sync, no extra elementwise pass).
//
// With these the whole objective is 4 x (photo + 2 smooth launches) + ident +
// finish forward and 4 launches backward, instead of ~250 ATen kernels per
// (scale, frame) round trip in the reference.
#include "../../include/dmh_b200.h"
#include "dmh_common.cuh"

using namespace dmh;

namespace {

#define SF_NB1 64
#define SF_THREADS 256

__global__ void sf_mean_kernel(const float* __restrict__ disp, int hw, float* __restrict__ part) {
    __shared__ float red[32];
    const int b = blockIdx.y;
    const float* d = disp + (size_t)b * hw;
    float s = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) s += d[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) part[b * SF_NB1 + blockIdx.x] = s;
}

// mean(disp_b) + 1e-7 from the 64 partials; called by the 32 lanes of one warp, same
// summation order wherever it is used (forward normalisation and its backward)
__device__ __forceinline__ float mean_eps_warp(const float* __restrict__ mean_part, int b, int hw) {
    const int l = threadIdx.x & 31;
    float s = mean_part[b * SF_NB1 + l] + mean_part[b * SF_NB1 + 32 + l];
    s = warp_sum(s);
    return s / (float)hw + 1e-7f;
}

__device__ __forceinline__ float edge_w(const float* __restrict__ im, int C, size_t plane, size_t a, size_t b2) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += fabsf(__ldg(im + c * plane + a) - __ldg(im + c * plane + b2));
    return expf(-(s / (float)C));
}
__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

__global__ void __launch_bounds__(SF_THREADS)
sf_main_kernel(const float* __restrict__ disp, const float* __restrict__ img, int C, int h, int w,
               const float* __restrict__ mean_part, float inv_nx, float inv_ny, float* __restrict__ gN,
               float* __restrict__ part) {
    __shared__ float red[32];
    __shared__ float s_m;
    const int b = blockIdx.y;
    const int hw = h * w;
    if (threadIdx.x < 32) {
        const float m0 = mean_eps_warp(mean_part, b, hw);
        if (threadIdx.x == 0) s_m = m0;
    }
    __syncthreads();
    const float m = s_m;
    const float* d = disp + (size_t)b * hw;
    const float* im = img + (size_t)b * C * hw;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    float tx = 0.f, ty = 0.f, g = 0.f, dv = 0.f;
    if (n < hw) {
        const int x = n % w, y = n / w;
        dv = d[n];
        const float dn = div_rn(dv, m);
        if (x < w - 1) {
            const float diff = dn - div_rn(d[n + 1], m);
            const float e = edge_w(im, C, hw, n, n + 1);
            tx = fabsf(diff) * e;
            g += sgn(diff) * e * inv_nx;
        }
        if (x > 0) {
            const float diff = div_rn(d[n - 1], m) - dn;
            g -= sgn(diff) * edge_w(im, C, hw, n - 1, n) * inv_nx;
        }
        if (y < h - 1) {
            const float diff = dn - div_rn(d[n + w], m);
            const float e = edge_w(im, C, hw, n, n + w);
            ty = fabsf(diff) * e;
            g += sgn(diff) * e * inv_ny;
        }
        if (y > 0) {
            const float diff = div_rn(d[n - w], m) - dn;
            g -= sgn(diff) * edge_w(im, C, hw, n - w, n) * inv_ny;
        }
        gN[(size_t)b * hw + n] = g;
    }
    const float sx = block_sum(tx, red);
    const float sy = block_sum(ty, red);
    const float sg = block_sum(g * dv, red);
    if (threadIdx.x == 0) {
        float* o = part + ((size_t)b * gridDim.x + blockIdx.x) * 3;
        o[0] = sx; o[1] = sy; o[2] = sg;
    }
}

#define FIN_MAXS 8
struct FinishParams {
    const float* smooth_ws[FIN_MAXS];     // [B*64 mean partials][B*nb*3 partials]
    const float* photo_part[FIN_MAXS];
    int photo_n[FIN_MAXS];
    int h[FIN_MAXS], w[FIN_MAXS];
    float smooth_weight[FIN_MAXS];
    int S, B;
    double inv_photo_den;
    float* img_scalars;                   // (S,B,2): 1/(mean+eps), corr
    float* losses;                        // (S+1): per-scale, total
};

__device__ __forceinline__ double block_sum_d(double v, double* red) {
    v = warp_sum_d(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    return t;
}

__global__ void __launch_bounds__(1024)
finish_kernel(const FinishParams p) {
    __shared__ double red[32];
    const int blk = blockIdx.x;
    if (blk < p.S * p.B) {
        const int s = blk / p.B, b = blk % p.B;
        const int hw = p.h[s] * p.w[s];
        const int nb = (hw + SF_THREADS - 1) / SF_THREADS;
        const float* mean_part = p.smooth_ws[s];
        const float* part = p.smooth_ws[s] + (size_t)p.B * SF_NB1;
        double sg = 0.0;
        for (int i = threadIdx.x; i < nb; i += blockDim.x) sg += (double)part[((size_t)b * nb + i) * 3 + 2];
        sg = block_sum_d(sg, red);
        float m = 0.f;
        if (threadIdx.x < 32) m = mean_eps_warp(mean_part, b, hw);
        if (threadIdx.x == 0) {
            const float inv_m = 1.0f / m;
            p.img_scalars[((size_t)s * p.B + b) * 2 + 0] = inv_m;
            p.img_scalars[((size_t)s * p.B + b) * 2 + 1] = (float)(sg / (double)hw) * inv_m * inv_m;
        }
        return;
    }
    // last block: the scalar losses
    double total = 0.0;
    for (int s = 0; s < p.S; ++s) {
        const int hw = p.h[s] * p.w[s];
        const int nb = (hw + SF_THREADS - 1) / SF_THREADS;
        const float* part = p.smooth_ws[s] + (size_t)p.B * SF_NB1;
        double ph = 0.0, sx = 0.0, sy = 0.0;
        for (int i = threadIdx.x; i < p.photo_n[s]; i += blockDim.x) ph += (double)p.photo_part[s][i];
        for (int i = threadIdx.x; i < p.B * nb; i += blockDim.x) {
            sx += (double)part[(size_t)i * 3 + 0];
            sy += (double)part[(size_t)i * 3 + 1];
        }
        ph = block_sum_d(ph, red);
        sx = block_sum_d(sx, red);
        sy = block_sum_d(sy, red);
        if (threadIdx.x == 0) {
            const double inv_nx = 1.0 / ((double)p.B * p.h[s] * (p.w[s] - 1));
            const double inv_ny = 1.0 / ((double)p.B * (p.h[s] - 1) * p.w[s]);
            const double loss = ph * p.inv_photo_den + (double)p.smooth_weight[s] * (sx * inv_nx + sy * inv_ny);
            p.losses[s] = (float)loss;
            total += loss;
        }
    }
    if (threadIdx.x == 0) p.losses[p.S] = (float)(total / (double)p.S);
}

// one thread per low-res pixel
__global__ void disp_grad_kernel(const float* __restrict__ G_full, const float* __restrict__ gN,
                                 const float* __restrict__ img_scalars, float smooth_weight,
                                 const float* __restrict__ g_total, const float* __restrict__ g_scale, float inv_S,
                                 int h, int w, int H, int W, float sh, float sw, float* __restrict__ grad) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int b = blockIdx.z;
    const float up = (g_total ? g_total[0] * inv_S : 0.0f) + (g_scale ? g_scale[0] : 0.0f);
    const float* g = G_full + (size_t)b * H * W;
    float acc;
    if (h == H && w == W) {
        acc = __ldg(g + (size_t)y * W + x);
    } else {
        const float rh = 1.0f / sh, rw = 1.0f / sw;
        const int Y0 = max(0, (int)floorf(((float)y - 0.5f) * rh - 0.5f) - 1);
        const int Y1 = min(H - 1, (int)ceilf(((float)y + 1.5f) * rh - 0.5f) + 1);
        const int X0 = max(0, (int)floorf(((float)x - 0.5f) * rw - 0.5f) - 1);
        const int X1 = min(W - 1, (int)ceilf(((float)x + 1.5f) * rw - 0.5f) + 1);
        acc = 0.0f;
        for (int Y = Y0; Y <= Y1; ++Y) {
            const UpTap ty = up_tap(Y, sh, h);
            const float wy = (ty.i0 == y ? ty.l0 : 0.0f) + (ty.i1 == y ? ty.l1 : 0.0f);
            if (wy == 0.0f) continue;
            float row = 0.0f;
            for (int X = X0; X <= X1; ++X) {
                const UpTap tx = up_tap(X, sw, w);
                const float wx = (tx.i0 == x ? tx.l0 : 0.0f) + (tx.i1 == x ? tx.l1 : 0.0f);
                if (wx != 0.0f) row = fmaf(wx, __ldg(g + (size_t)Y * W + X), row);
            }
            acc = fmaf(wy, row, acc);
        }
    }
    if (gN) {
        const float inv_m = img_scalars[b * 2], corr = img_scalars[b * 2 + 1];
        acc += smooth_weight * (gN[(size_t)b * h * w + (size_t)y * w + x] * inv_m - corr);
    }
    grad[(size_t)b * h * w + (size_t)y * w + x] = up * acc;
}

}  // namespace

extern "C" {

long long dmh_smooth_fused_workspace_floats(int B, int h, int w) {
    return (long long)B * SF_NB1 + 3LL * B * ceil_div((long long)h * w, SF_THREADS);
}

int dmh_smooth_fused(const float* disp, const float* img, int B, int C, int h, int w, float* ws, float* gN,
                     dmh_stream_t stream) {
    DMH_REQUIRE(disp && img && ws && gN, "dmh_smooth_fused: null pointer");
    DMH_REQUIRE(B > 0 && B <= 65535 && C > 0 && h >= 2 && w >= 2, "dmh_smooth_fused: bad shape");
    const int nb = ceil_div((long long)h * w, SF_THREADS);
    cudaStream_t st = (cudaStream_t)stream;
    DMH_LAUNCH(sf_mean_kernel, dim3(SF_NB1, B), 256, 0, st)(disp, h * w, ws);
    const float inv_nx = (float)(1.0 / ((double)B * h * (w - 1))), inv_ny = (float)(1.0 / ((double)B * (h - 1) * w));
    DMH_LAUNCH(sf_main_kernel, dim3(nb, B), SF_THREADS, 0, st)(disp, img, C, h, w, ws, inv_nx, inv_ny, gN,
                                                             ws + (size_t)B * SF_NB1);
    DMH_CHECK_LAUNCH("dmh_smooth_fused");
    return DMH_OK;
}

int dmh_objective_finish(int S, int B, const float* const* smooth_ws_host, const int* h_host, const int* w_host,
                         const float* const* photo_part_host, const int* photo_n_host,
                         const float* smooth_weight_host, double photo_den, float* img_scalars, float* losses,
                         dmh_stream_t stream) {
    DMH_REQUIRE(S >= 1 && S <= FIN_MAXS && B > 0, "dmh_objective_finish: S=%d outside [1,%d] or B <= 0", S, FIN_MAXS);
    DMH_REQUIRE(smooth_ws_host && h_host && w_host && photo_part_host && photo_n_host && smooth_weight_host &&
                    img_scalars && losses && photo_den > 0.0,
                "dmh_objective_finish: null pointer");
    FinishParams p;
    for (int s = 0; s < S; ++s) {
        DMH_REQUIRE(smooth_ws_host[s] && photo_part_host[s], "dmh_objective_finish: null buffer for scale %d", s);
        p.smooth_ws[s] = smooth_ws_host[s];
        p.photo_part[s] = photo_part_host[s];
        p.photo_n[s] = photo_n_host[s];
        p.h[s] = h_host[s];
        p.w[s] = w_host[s];
        p.smooth_weight[s] = smooth_weight_host[s];
    }
    p.S = S; p.B = B; p.inv_photo_den = 1.0 / photo_den; p.img_scalars = img_scalars; p.losses = losses;
    DMH_LAUNCH(finish_kernel, S * B + 1, 1024, 0, (cudaStream_t)stream)(p);
    DMH_CHECK_LAUNCH("dmh_objective_finish");
    return DMH_OK;
}

int dmh_disp_grad(const float* G_full, const float* gN, const float* img_scalars, float smooth_weight,
                  const float* g_total, const float* g_scale, float inv_S, int B, int h, int w, int H, int W,
                  float* grad_disp, dmh_stream_t stream) {
    DMH_REQUIRE(G_full && grad_disp && (g_total || g_scale), "dmh_disp_grad: null pointer");
    DMH_REQUIRE(!gN || img_scalars, "dmh_disp_grad: gN given without img_scalars");
    DMH_REQUIRE(B > 0 && B <= 65535 && h >= 1 && w >= 1 && H >= h && W >= w, "dmh_disp_grad: bad shape");
    dim3 block(32, 8), grid(ceil_div(w, 32), ceil_div(h, 8), B);
    DMH_LAUNCH(disp_grad_kernel, grid, block, 0, (cudaStream_t)stream)(G_full, gN, img_scalars, smooth_weight, g_total, g_scale,
                                                                    inv_S, h, w, H, W, (float)h / (float)H,
                                                                    (float)w / (float)W, grad_disp);
    DMH_CHECK_LAUNCH("dmh_disp_grad");
    return DMH_OK;
}

}  // extern "C"

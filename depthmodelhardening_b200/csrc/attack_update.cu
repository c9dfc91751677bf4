// Stage 1 -- patch update rules (A6-A8).  All of these touch < 4 MB and are
// launch-latency bound: the point of the kernels is to replace
// ~6 (L-inf) / ~25 (L0) ATen launches and one host sync per iteration by one
// launch each, with the L0 count kept on the device.
//
//   dmh_pgd_linf_step     phy_obj_atk.py:98-100 (same 3 lines: pgd_depth.py:76-78, pgd.py:73-75)
//   dmh_apgd_linf_step    phy_obj_atk_apgd.py:214-222 (APGD step with momentum and double projection)
//   dmh_pgd_l2_step       phy_obj_atk_l2.py:108-120 (gradient normalisation, step, projection onto the L2 ball)
//   dmh_l0_compose_count  phy_obj_atk_l0.py:94-99 + cal_l0 :43-52
//   dmh_l0_adam_step      phy_obj_atk_l0.py:130-138 (mask cost gradient + chain through the
//                         compose clamps + torch.optim.Adam(betas=(0.5,0.9)) update)
//   dmh_l0_finalize       phy_obj_atk_l0.py:143-150
//   dmh_tube_light_patch  light_simulation.py:132-170 + phy_obj_atk_light.py:118-122 (one black-box candidate)
//   dmh_square_linf_candidate  phy_obj_atk_square.py:263-274 (one Square-attack candidate)
//   dmh_keep_best         phy_obj_atk_light.py:148-150 / phy_obj_atk_square.py:281-297 (accept on the device)
//   dmh_topk_select       EXTENSION: exact k-th largest by 4-pass radix
//                         select with warp-aggregated shared-memory histograms
#include "../../include/dmh_b200.h"
#include "dmh_common.cuh"
#include "light_math.cuh"

using namespace dmh;

namespace {

__device__ __forceinline__ float clamp01(float v, float hi) { return fminf(fmaxf(v, 0.0f), hi); }
__device__ __forceinline__ float sgnf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }
// ATen lerp: two branches so that lerp(a, b, 1) == b exactly
__device__ __forceinline__ float lerp_aten(float a, float b, float w) {
    return w < 0.5f ? a + w * (b - a) : b - (b - a) * (1.0f - w);
}

// --------------------------------------------------------------------------- A6
__global__ void pgd_linf_kernel(const float* __restrict__ adv, const float* __restrict__ grad,
                                const float* __restrict__ clean, long long n, float alpha, float eps,
                                float* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = add_rn(adv[i], mul_rn(alpha, sgnf(grad[i])));
    const float c = clean[i];
    const float delta = fminf(fmaxf(sub_rn(a, c), -eps), eps);
    out[i] = fminf(fmaxf(add_rn(c, delta), 0.0f), 1.0f);
}

// --------------------------------------------------------------------------- APGD L-inf step with momentum (next-4)
// phy_obj_atk_apgd.py:214-222 in one launch, every operation rounded as the torch expression rounds it:
//   grad2 = x_adv - x_adv_old
//   z     = clamp(min(max(x_adv + step * sign(grad), x - eps), x + eps), 0, 1)
//   x_new = clamp(min(max(x_adv + (z - x_adv) * a + grad2 * (1 - a), x - eps), x + eps), 0, 1)
// (a = 0.75 after the first iteration, 1 at the first: python floats, exact in fp32.)
__global__ void apgd_linf_kernel(const float* __restrict__ x_adv, const float* __restrict__ x_adv_old,
                                 const float* __restrict__ grad, const float* __restrict__ x0, long long n, float step,
                                 float a, float eps, float* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float xa = x_adv[i], c = x0[i];
    const float lo = sub_rn(c, eps), hi = add_rn(c, eps);
    const float grad2 = sub_rn(xa, x_adv_old[i]);
    float z = add_rn(xa, mul_rn(step, sgnf(grad[i])));
    z = fminf(fmaxf(fminf(fmaxf(z, lo), hi), 0.0f), 1.0f);
    float y = add_rn(add_rn(xa, mul_rn(sub_rn(z, xa), a)), mul_rn(grad2, sub_rn(1.0f, a)));
    out[i] = fminf(fmaxf(fminf(fmaxf(y, lo), hi), 0.0f), 1.0f);
}

// --------------------------------------------------------------------------- L2 PGD update (next-4)
// phy_obj_atk_l2.py:108-120 for the one shared patch, in ONE launch of one CTA (0.94 MB of state: latency-bound):
//   g = grad / (||grad||_2 + 1e-10);  x = adv + alpha * g;  d = x - clean;
//   out = clamp(clean + d * min(eps / ||d||_2, 1), 0, 1)
// Every element-wise operation is rounded as torch rounds it; the two norms are accumulated in double in a fixed
// order (deterministic; torch.norm's fp32 tree differs from it by ~1e-7 relative).
__device__ __forceinline__ double block_sum_d(double v, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    if (wid == 0) v = warp_sum_d(v);
    if (threadIdx.x == 0) red[0] = v;
    __syncthreads();
    v = red[0];
    __syncthreads();
    return v;
}

// --------------------------------------------------------------------------- evaluation metrics (next-4)
// evaluate_depth.py:193-196 + compute_errors (:57-99) for one batch, one launch: both disparity maps go through
// disp_to_depth(|disp|, 0.1, 100)[1] * 5.4 clamped to [1e-3, 80] (fp32, rounded as torch rounds it), then the eight
// (masked) error sums of compute_errors.  The nine sums -- mask total first -- are accumulated in double per thread,
// reduced per block and added to out[9] with double atomics (evaluation statistics, not on the training path;
// order-dependent in the last bits of a double, far below the fp32 the harness prints).
__global__ void __launch_bounds__(256)
depth_errors_kernel(const float* __restrict__ disp_gt, const float* __restrict__ disp_pred,
                    const float* __restrict__ mask, long long n, DepthScale ds, float scale_factor, float min_depth,
                    float max_depth, double* __restrict__ out) {
    __shared__ double red[32];
    double acc[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) acc[q] = 0.0;
    const float t1 = 1.25f, t2 = 1.25f * 1.25f, t3 = 1.25f * 1.25f * 1.25f;   // python 1.25**k, exact in fp32 (k <= 3)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float m = mask ? mask[i] : 1.0f;
        const float gt = fminf(fmaxf(mul_rn(disp_to_depth(fabsf(disp_gt[i]), ds), scale_factor), min_depth), max_depth);
        const float pr = fminf(fmaxf(mul_rn(disp_to_depth(fabsf(disp_pred[i]), ds), scale_factor), min_depth), max_depth);
        const float thr = fmaxf(div_rn(gt, pr), div_rn(pr, gt));
        const float d = sub_rn(gt, pr);
        const float d2 = mul_rn(d, d);
        const float dl = sub_rn(logf(gt), logf(pr));
        acc[0] += (double)m;
        acc[1] += (double)mul_rn(fabsf(d), m);                      // abs_err
        acc[2] += (double)mul_rn(div_rn(fabsf(d), gt), m);          // abs_rel
        acc[3] += (double)mul_rn(div_rn(d2, gt), m);                // sq_rel
        acc[4] += (double)mul_rn(d2, m);                            // rmse^2
        acc[5] += (double)mul_rn(mul_rn(dl, dl), m);                // rmse_log^2
        acc[6] += (double)((thr < t1 ? 1.0f : 0.0f) * m);
        acc[7] += (double)((thr < t2 ? 1.0f : 0.0f) * m);
        acc[8] += (double)((thr < t3 ? 1.0f : 0.0f) * m);
    }
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const double v = block_sum_d(acc[q], red);
        if (threadIdx.x == 0 && v != 0.0) atomicAdd(out + q, v);
    }
}

__global__ void __launch_bounds__(1024)
pgd_l2_kernel(const float* __restrict__ adv, const float* __restrict__ grad, const float* __restrict__ clean,
              long long n, float alpha, float eps, float eps_div, float* __restrict__ out) {
    __shared__ double red[32];
    double ss = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double g = (double)grad[i];
        ss += g * g;
    }
    const float gnorm = add_rn((float)sqrt(block_sum_d(ss, red)), eps_div);
    ss = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float x = add_rn(adv[i], mul_rn(alpha, div_rn(grad[i], gnorm)));
        out[i] = x;                                        // (out may alias adv: element i is read before it is written)
        const double d = (double)sub_rn(x, clean[i]);
        ss += d * d;
    }
    const float dnorm = (float)sqrt(block_sum_d(ss, red));
    const float factor = fminf(div_rn(eps, dnorm), 1.0f);  // eps / 0 = inf -> 1
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float c = clean[i];
        const float d = mul_rn(sub_rn(out[i], c), factor);
        out[i] = fminf(fmaxf(add_rn(c, d), 0.0f), 1.0f);
    }
}

// --------------------------------------------------------------------------- A7
// one thread per PIXEL (3 channels): compose the adversarial patch, count pixels
// whose thresholded pattern is non-zero (warp ballot + one atomic per warp)
__global__ void l0_compose_count_kernel(const float* __restrict__ obj, const float* __restrict__ ppos,
                                        const float* __restrict__ pneg, int C, int npix, float clip_max, float thr,
                                        float* __restrict__ adv, unsigned long long* __restrict__ count) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    bool alive = false;
    if (p < npix) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) {
            const size_t o = (size_t)c * npix + p;
            const float pos = clamp01(mul_rn(ppos[o], clip_max), clip_max);
            const float neg = -clamp01(mul_rn(pneg[o], clip_max), clip_max);
            if (adv) adv[o] = clamp01(add_rn(obj[o], add_rn(pos, neg)), clip_max);
            const float pt = pos < thr ? 0.f : pos;
            const float nt = neg > -thr ? 0.f : neg;
            s = add_rn(s, fabsf(add_rn(pt, nt)));
        }
        alive = s != 0.f;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, alive);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(count, (unsigned long long)__popc(ballot));
}

// --------------------------------------------------------------------------- A8
// state: l0 bookkeeping kept on the device so that no host sync is needed to
// pick mask_weight (phy_obj_atk_l0.py:102-111)
//   counts[0] = current l0, counts[1] = l0 at step 0
__global__ void l0_adam_kernel(const float* __restrict__ obj, const float* __restrict__ g_adv,
                               float* __restrict__ ppos, float* __restrict__ pneg, float* __restrict__ m_pos,
                               float* __restrict__ v_pos, float* __restrict__ m_neg, float* __restrict__ v_neg,
                               int C, int npix, float clip_max, const unsigned long long* __restrict__ counts,
                               float l0_thresh, float mask_weight_init, float step_size, float beta1, float beta2,
                               float adam_eps, float bc2_sqrt, const float* __restrict__ bias_table, int table_len,
                               unsigned* step_state) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (bias_table) {
        // step index on the device (dmh_l0_adam_step_dev): every CTA reads the number of finished steps before any
        // CTA can advance it -- the advance is done by the LAST CTA to arrive at the end of the kernel
        const unsigned done = *reinterpret_cast<volatile unsigned*>(step_state);
        const int idx = min((int)done, table_len - 1);
        step_size = bias_table[2 * idx];
        bc2_sqrt = bias_table[2 * idx + 1];
        __syncthreads();                                 // the whole CTA has read `done`
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(step_state + 1, 1u) == gridDim.x - 1) {
                step_state[1] = 0u;
                __threadfence();
                *reinterpret_cast<volatile unsigned*>(step_state) = done + 1u;
            }
        }
    }
    if (p >= npix) return;
    // mask_weight = 0 once l0/l0_init <= thresh (fp32 division as torch does on the int64 tensors)
    float mask_w = mask_weight_init;
    if (counts) {
        const float ratio = (float)counts[0] / (float)counts[1];
        if (ratio <= l0_thresh) mask_w = 0.f;
    }
    // argmax over channels of tanh(P/10)/(2-1e-7)+0.5 (monotone in P: argmax of P, first max wins)
    int am_p = 0, am_n = 0;
    float best_p = ppos[p], best_n = pneg[p];
    for (int c = 1; c < C; ++c) {
        const float a = ppos[(size_t)c * npix + p], b = pneg[(size_t)c * npix + p];
        if (a > best_p) { best_p = a; am_p = c; }
        if (b > best_n) { best_n = b; am_n = c; }
    }
    const float mask_scale = mask_w / (float)npix;        // mean over the (1,H,W) max map
    const float denom = 2.0f - 1e-7f;
    for (int c = 0; c < C; ++c) {
        const size_t o = (size_t)c * npix + p;
        const float Pp = ppos[o], Pn = pneg[o];
        const float pos_raw = mul_rn(Pp, clip_max), neg_raw = mul_rn(Pn, clip_max);
        const float pos = clamp01(pos_raw, clip_max), neg = -clamp01(neg_raw, clip_max);
        const float pre = add_rn(obj[o], add_rn(pos, neg));
        // clamp passes gradient on the closed interval (torch.clamp backward)
        const float pass_out = (pre >= 0.f && pre <= clip_max) ? 1.f : 0.f;
        const float ga = g_adv ? g_adv[o] * pass_out : 0.f;
        float gp = (pos_raw >= 0.f && pos_raw <= clip_max) ? ga * clip_max : 0.f;
        float gn = (neg_raw >= 0.f && neg_raw <= clip_max) ? -ga * clip_max : 0.f;
        if (mask_scale != 0.f) {
            if (c == am_p) { const float t = tanhf(Pp / 10.0f); gp += mask_scale * (1.0f - t * t) / 10.0f / denom; }
            if (c == am_n) { const float t = tanhf(Pn / 10.0f); gn += mask_scale * (1.0f - t * t) / 10.0f / denom; }
        }
        // torch.optim.Adam (single tensor): m = lerp(m, g, 1-b1); v = b2 v + (1-b2) g^2;
        // p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
        {
            float m = m_pos[o], v = v_pos[o];
            m = lerp_aten(m, gp, 1.0f - beta1);
            v = v * beta2 + (1.0f - beta2) * gp * gp;
            m_pos[o] = m; v_pos[o] = v;
            ppos[o] = Pp - step_size * (m / (sqrtf(v) / bc2_sqrt + adam_eps));
        }
        {
            float m = m_neg[o], v = v_neg[o];
            m = lerp_aten(m, gn, 1.0f - beta1);
            v = v * beta2 + (1.0f - beta2) * gn * gn;
            m_neg[o] = m; v_neg[o] = v;
            pneg[o] = Pn - step_size * (m / (sqrtf(v) / bc2_sqrt + adam_eps));
        }
    }
}

__global__ void l0_finalize_kernel(const float* __restrict__ obj, const float* __restrict__ ppos,
                                   const float* __restrict__ pneg, long long n, float clip_max, float thr,
                                   float* __restrict__ adv, float* __restrict__ pattern) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float pos = clamp01(mul_rn(ppos[i], clip_max), clip_max);
    float neg = -clamp01(mul_rn(pneg[i], clip_max), clip_max);
    if (pos < thr) pos = 0.f;
    if (neg > -thr) neg = 0.f;
    const float pat = add_rn(pos, neg);
    if (pattern) pattern[i] = pat;
    adv[i] = clamp01(add_rn(obj[i], pat), clip_max);
}

// --------------------------------------------------------------------------- top-k (extension)
// magnitude per pixel = max_c max(clamp(P+), clamp(P-)); keep the k largest
// (ties -> lower pixel index).  Single CTA, 4 radix passes over the fp32 bit
// pattern (non-negative floats order like unsigned ints), 256-bin histograms in
// shared memory filled with warp-aggregated atomics (__match_any_sync).
#define TK_THREADS 1024

__device__ __forceinline__ unsigned pixel_key(const float* __restrict__ ppos, const float* __restrict__ pneg, int C,
                                              int npix, int p) {
    float m = 0.f;
    for (int c = 0; c < C; ++c) {
        m = fmaxf(m, clamp01(ppos[(size_t)c * npix + p], 1.0f));
        m = fmaxf(m, clamp01(pneg[(size_t)c * npix + p], 1.0f));
    }
    return __float_as_uint(m);
}

__global__ void __launch_bounds__(TK_THREADS)
topk_select_kernel(float* __restrict__ ppos, float* __restrict__ pneg, int C, int npix, int k,
                   unsigned char* __restrict__ keep_out, unsigned* __restrict__ kth_key_out) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_remaining, s_tie_budget;
    const int tid = threadIdx.x;
    unsigned prefix = 0, prefix_mask = 0;
    unsigned remaining = (unsigned)k;                     // how many still to take among keys matching the prefix
    if (k <= 0 || k >= npix) {
        for (int p = tid; p < npix; p += TK_THREADS) {
            const bool keep = k >= npix;
            if (keep_out) keep_out[p] = keep;
            if (!keep)
                for (int c = 0; c < C; ++c) { ppos[(size_t)c * npix + p] = 0.f; pneg[(size_t)c * npix + p] = 0.f; }
        }
        if (tid == 0 && kth_key_out) *kth_key_out = 0u;
        return;
    }
    for (int pass = 3; pass >= 0; --pass) {
        const int shift = pass * 8;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        for (int base = 0; base < npix; base += TK_THREADS) {
            const int p = base + tid;
            const bool valid = p < npix;
            unsigned key = valid ? pixel_key(ppos, pneg, C, npix, p) : 0u;
            const bool match = valid && ((key & prefix_mask) == prefix);
            const unsigned bin = match ? ((key >> shift) & 255u) : 0xffffffffu;
            // warp-aggregated histogram update: one atomic per distinct bin per warp
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            if (match && (__ffs(peers) - 1) == (tid & 31)) atomicAdd(&hist[bin], (unsigned)__popc(peers));
        }
        __syncthreads();
        if (tid == 0) {
            // walk bins from the largest digit down until `remaining` is covered
            unsigned acc = 0;
            int d = 255;
            for (; d > 0; --d) {
                if (acc + hist[d] >= remaining) break;
                acc += hist[d];
            }
            s_prefix = prefix | ((unsigned)d << shift);
            s_remaining = remaining - acc;
        }
        __syncthreads();
        prefix = s_prefix;
        remaining = s_remaining;
        prefix_mask |= 255u << shift;
        __syncthreads();
    }
    // prefix == k-th largest key; `remaining` of the pixels equal to it are kept (lowest indices first)
    if (tid == 0) { s_tie_budget = remaining; if (kth_key_out) *kth_key_out = prefix; }
    __syncthreads();
    // ordered tie resolution: chunks processed in index order, warp-prefix inside a chunk
    __shared__ unsigned warp_ties[TK_THREADS / 32];
    __shared__ unsigned chunk_base;
    if (tid == 0) chunk_base = 0;
    __syncthreads();
    for (int base = 0; base < npix; base += TK_THREADS) {
        const int p = base + tid;
        const bool valid = p < npix;
        const unsigned key = valid ? pixel_key(ppos, pneg, C, npix, p) : 0u;
        const bool tie = valid && key == prefix;
        const unsigned ballot = __ballot_sync(0xffffffffu, tie);
        const int lane = tid & 31, wid = tid >> 5;
        if (lane == 0) warp_ties[wid] = __popc(ballot);
        __syncthreads();
        unsigned before = chunk_base;
        for (int w = 0; w < wid; ++w) before += warp_ties[w];
        before += __popc(ballot & ((1u << lane) - 1u));
        const bool keep = valid && (key > prefix || (tie && before < s_tie_budget));
        if (valid) {
            if (keep_out) keep_out[p] = keep;
            if (!keep)
                for (int c = 0; c < C; ++c) { ppos[(size_t)c * npix + p] = 0.f; pneg[(size_t)c * npix + p] = 0.f; }
        }
        __syncthreads();
        if (tid == 0) {
            unsigned t = 0;
            for (int w = 0; w < TK_THREADS / 32; ++w) t += warp_ties[w];
            chunk_base += t;
        }
        __syncthreads();
    }
}


// --------------------------------------------------------------------------- black-box candidates (next-4)
// One candidate of the tube-light search (phy_obj_atk_light.py:109-122): the reference fills the (h, w, 3) light
// field in a Python double loop (~0.1 s for the 300 x 260 object), adds it to the 8-bit object image with OpenCV and
// converts back with ToTensor; here one thread per pixel does the same float64 / float32 operations (light_math.cuh)
// and writes the candidate patch as fp32 (k / 255, IEEE division == ToTensor).  base: planar (3, h, w) bytes.
__global__ void __launch_bounds__(256)
tube_light_kernel(const uint8_t* __restrict__ base, int h, int w, TubeLight t, float* __restrict__ patch,
                  uint8_t* __restrict__ lit) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = h * w;
    if (i >= n) return;
    const int y = i / w, x = i - y * w;
    double att = 0.0;
    const int zone = tube_light_zone(t, x, y, &att);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint8_t v = lit_u8(base[c * n + i], t.ca[c], zone, att);
        patch[c * n + i] = div_rn((float)v, 255.0f);
        if (lit) lit[c * n + i] = v;
    }
}

// One candidate of the Square attack's L-inf random search (phy_obj_atk_square.py:263-274): the square
// [vh, vh+s) x [vw, vw+s) of the best patch so far moves by delta[c] = 2 * eps * (+-1) per channel, then the eps-ball
// around the clean patch and [0, 1] are enforced -- three full-size temporaries and five launches in the reference.
__global__ void __launch_bounds__(256)
square_candidate_kernel(const float* __restrict__ x_best, const float* __restrict__ x, int H, int W, int vh, int vw,
                        int s, float d0, float d1, float d2, float eps, float* __restrict__ x_new) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = H * W;
    if (i >= n) return;
    const int py = i / W, px = i - py * W;
    const bool in = py >= vh && py < vh + s && px >= vw && px < vw + s;
    const float d[3] = {d0, d1, d2};
#pragma unroll
    for (int c = 0; c < 3; ++c)
        x_new[c * n + i] = square_linf_candidate(x_best[c * n + i], x[c * n + i], in ? d[c] : 0.0f, eps);
}

// Accepting a candidate without a host round trip: `if cost < best_cost: best_cost, best = cost, candidate`
// (phy_obj_atk_light.py:148-150; phy_obj_atk_square.py:281-297 with its 0/1 blend) -- every thread reads the
// previous best cost from `best_in`, thread 0 writes the new one to `best_out` (the caller ping-pongs the two).
__global__ void __launch_bounds__(256)
keep_best_kernel(const float* __restrict__ cost, const float* __restrict__ best_in, float* __restrict__ best_out,
                 const float* __restrict__ cand, float* __restrict__ best, long long n) {
    const float c = cost[0], b = best_in[0];
    const bool better = c < b;                          // false for NaN, as in Python
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) best_out[0] = better ? c : b;
    if (better && i < n) best[i] = cand[i];
}

}  // namespace

extern "C" {

int dmh_pgd_linf_step(const float* adv, const float* grad, const float* clean, long long n, float alpha, float eps,
                      float* out, dmh_stream_t stream) {
    DMH_REQUIRE(adv && grad && clean && out && n > 0, "dmh_pgd_linf_step: null pointer or n <= 0");
    DMH_LAUNCH(pgd_linf_kernel, ceil_div(n, 256), 256, 0, (cudaStream_t)stream)(adv, grad, clean, n, alpha, eps, out);
    DMH_CHECK_LAUNCH("dmh_pgd_linf_step");
    return DMH_OK;
}

int dmh_apgd_linf_step(const float* x_adv, const float* x_adv_old, const float* grad, const float* x0, long long n,
                       float step, float a, float eps, float* out, dmh_stream_t stream) {
    DMH_REQUIRE(x_adv && x_adv_old && grad && x0 && out && n > 0, "dmh_apgd_linf_step: null pointer or n <= 0");
    DMH_REQUIRE(out != x_adv_old && out != grad && out != x0, "dmh_apgd_linf_step: out may alias x_adv only");
    DMH_LAUNCH(apgd_linf_kernel, ceil_div(n, 256), 256, 0, (cudaStream_t)stream)(x_adv, x_adv_old, grad, x0, n, step, a,
                                                                               eps, out);
    DMH_CHECK_LAUNCH("dmh_apgd_linf_step");
    return DMH_OK;
}

int dmh_tube_light_patch(const uint8_t* base_u8, int h, int w, double k, double b, double norm, double beta,
                         int full_end, int light_end, double ca0, double ca1, double ca2, float* patch, uint8_t* lit_u8,
                         dmh_stream_t stream) {
    DMH_REQUIRE(base_u8 && patch, "dmh_tube_light_patch: null pointer");
    DMH_REQUIRE(h > 0 && w > 0 && (long long)h * w < (1ll << 30), "dmh_tube_light_patch: bad shape %d x %d", h, w);
    DMH_REQUIRE(norm >= 1.0 && full_end >= 0 && light_end >= full_end, "dmh_tube_light_patch: bad beam scalars");
    TubeLight t;
    t.k = k; t.b = b; t.norm = norm; t.beta = beta; t.full_end = (double)full_end; t.light_end = (double)light_end;
    t.ca[0] = ca0; t.ca[1] = ca1; t.ca[2] = ca2;
    DMH_LAUNCH(tube_light_kernel, ceil_div((long long)h * w, 256), 256, 0, (cudaStream_t)stream)(base_u8, h, w, t, patch,
                                                                                                lit_u8);
    DMH_CHECK_LAUNCH("dmh_tube_light_patch");
    return DMH_OK;
}

int dmh_square_linf_candidate(const float* x_best, const float* x, int H, int W, int vh, int vw, int s, float d0,
                              float d1, float d2, float eps, float* x_new, dmh_stream_t stream) {
    DMH_REQUIRE(x_best && x && x_new, "dmh_square_linf_candidate: null pointer");
    DMH_REQUIRE(H > 0 && W > 0 && s >= 0 && vh >= 0 && vw >= 0, "dmh_square_linf_candidate: bad shape / window");
    DMH_REQUIRE(x_new != x, "dmh_square_linf_candidate: x_new may alias x_best only");
    DMH_LAUNCH(square_candidate_kernel, ceil_div((long long)H * W, 256), 256, 0, (cudaStream_t)stream)(
        x_best, x, H, W, vh, vw, s, d0, d1, d2, eps, x_new);
    DMH_CHECK_LAUNCH("dmh_square_linf_candidate");
    return DMH_OK;
}

int dmh_keep_best(const float* cost, const float* best_cost_in, float* best_cost_out, const float* cand, float* best,
                  long long n, dmh_stream_t stream) {
    DMH_REQUIRE(cost && best_cost_in && best_cost_out && cand && best && n > 0, "dmh_keep_best: null pointer or n <= 0");
    DMH_REQUIRE(best_cost_in != best_cost_out && cand != best, "dmh_keep_best: in / out buffers must differ");
    DMH_LAUNCH(keep_best_kernel, ceil_div(n, 256), 256, 0, (cudaStream_t)stream)(cost, best_cost_in, best_cost_out, cand,
                                                                              best, n);
    DMH_CHECK_LAUNCH("dmh_keep_best");
    return DMH_OK;
}

int dmh_depth_errors(const float* disp_gt, const float* disp_pred, const float* mask, long long n, float min_disp_depth,
                     float max_disp_depth, float scale_factor, float min_depth, float max_depth, double* out,
                     dmh_stream_t stream) {
    DMH_REQUIRE(disp_gt && disp_pred && out && n > 0, "dmh_depth_errors: null pointer or n <= 0");
    DepthScale ds;
    ds.min_disp = (float)(1.0 / (double)max_disp_depth);
    ds.range = (float)(1.0 / (double)min_disp_depth - 1.0 / (double)max_disp_depth);
    cudaError_t e = cudaMemsetAsync(out, 0, 9 * sizeof(double), (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("dmh_depth_errors: memset failed: %s", cudaGetErrorString(e)); return DMH_ERR_CUDA; }
    const int blocks = (int)(n / 1024 < 1 ? 1 : (n / 1024 > 1184 ? 1184 : n / 1024));
    DMH_LAUNCH(depth_errors_kernel, blocks, 256, 0, (cudaStream_t)stream)(disp_gt, disp_pred, mask, n, ds, scale_factor,
                                                                         min_depth, max_depth, out);
    DMH_CHECK_LAUNCH("dmh_depth_errors");
    return DMH_OK;
}

int dmh_pgd_l2_step(const float* adv, const float* grad, const float* clean, long long n, float alpha, float eps,
                    float eps_div, float* out, dmh_stream_t stream) {
    DMH_REQUIRE(adv && grad && clean && out && n > 0, "dmh_pgd_l2_step: null pointer or n <= 0");
    DMH_REQUIRE(out != grad && out != clean, "dmh_pgd_l2_step: out may alias adv only");
    DMH_LAUNCH(pgd_l2_kernel, 1, 1024, 0, (cudaStream_t)stream)(adv, grad, clean, n, alpha, eps, eps_div, out);
    DMH_CHECK_LAUNCH("dmh_pgd_l2_step");
    return DMH_OK;
}

int dmh_l0_compose_count(const float* obj, const float* pattern_pos, const float* pattern_neg, int C, int H, int W,
                         float clip_max, float threshold, float* adv, unsigned long long* count,
                         dmh_stream_t stream) {
    DMH_REQUIRE(obj && pattern_pos && pattern_neg && count, "dmh_l0_compose_count: null pointer");
    DMH_REQUIRE(C > 0 && H > 0 && W > 0, "dmh_l0_compose_count: bad shape");
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(unsigned long long), (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("dmh_l0_compose_count: memset failed: %s", cudaGetErrorString(e)); return DMH_ERR_CUDA; }
    const int npix = H * W;
    DMH_LAUNCH(l0_compose_count_kernel, ceil_div(npix, 256), 256, 0, (cudaStream_t)stream)(
        obj, pattern_pos, pattern_neg, C, npix, clip_max, threshold, adv, count);
    DMH_CHECK_LAUNCH("dmh_l0_compose_count");
    return DMH_OK;
}

int dmh_l0_adam_step(const float* obj, const float* grad_adv, float* pattern_pos, float* pattern_neg, float* m_pos,
                     float* v_pos, float* m_neg, float* v_neg, int C, int H, int W, float clip_max,
                     const unsigned long long* counts, float l0_thresh, float mask_weight, float lr, float beta1,
                     float beta2, float adam_eps, int step, dmh_stream_t stream) {
    DMH_REQUIRE(obj && pattern_pos && pattern_neg && m_pos && v_pos && m_neg && v_neg, "dmh_l0_adam_step: null pointer");
    DMH_REQUIRE(C > 0 && H > 0 && W > 0 && step >= 1, "dmh_l0_adam_step: bad shape or step < 1");
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const int npix = H * W;
    DMH_LAUNCH(l0_adam_kernel, ceil_div(npix, 256), 256, 0, (cudaStream_t)stream)(
        obj, grad_adv, pattern_pos, pattern_neg, m_pos, v_pos, m_neg, v_neg, C, npix, clip_max, counts, l0_thresh,
        mask_weight, (float)((double)lr / bc1), beta1, beta2, adam_eps, (float)sqrt(bc2), nullptr, 0, nullptr);
    DMH_CHECK_LAUNCH("dmh_l0_adam_step");
    return DMH_OK;
}

int dmh_l0_adam_bias_table(float lr, float beta1, float beta2, int table_len, float* table_host) {
    DMH_REQUIRE(table_host && table_len >= 1, "dmh_l0_adam_bias_table: null pointer or empty table");
    for (int t = 1; t <= table_len; ++t) {
        // the two scalars dmh_l0_adam_step forms from its host-side step index, same double arithmetic
        const double bc1 = 1.0 - pow((double)beta1, (double)t);
        const double bc2 = 1.0 - pow((double)beta2, (double)t);
        table_host[2 * (t - 1)] = (float)((double)lr / bc1);
        table_host[2 * (t - 1) + 1] = (float)sqrt(bc2);
    }
    return DMH_OK;
}

int dmh_l0_adam_step_dev(const float* obj, const float* grad_adv, float* pattern_pos, float* pattern_neg, float* m_pos,
                         float* v_pos, float* m_neg, float* v_neg, int C, int H, int W, float clip_max,
                         const unsigned long long* counts, float l0_thresh, float mask_weight, float beta1, float beta2,
                         float adam_eps, const float* bias_table, int table_len, unsigned* step_state,
                         dmh_stream_t stream) {
    DMH_REQUIRE(obj && pattern_pos && pattern_neg && m_pos && v_pos && m_neg && v_neg && bias_table && step_state,
                "dmh_l0_adam_step_dev: null pointer");
    DMH_REQUIRE(C > 0 && H > 0 && W > 0 && table_len >= 1, "dmh_l0_adam_step_dev: bad shape or empty table");
    const int npix = H * W;
    DMH_LAUNCH(l0_adam_kernel, ceil_div(npix, 256), 256, 0, (cudaStream_t)stream)(
        obj, grad_adv, pattern_pos, pattern_neg, m_pos, v_pos, m_neg, v_neg, C, npix, clip_max, counts, l0_thresh,
        mask_weight, 0.0f, beta1, beta2, adam_eps, 1.0f, bias_table, table_len, step_state);
    DMH_CHECK_LAUNCH("dmh_l0_adam_step_dev");
    return DMH_OK;
}

int dmh_l0_finalize(const float* obj, const float* pattern_pos, const float* pattern_neg, long long n, float clip_max,
                    float threshold, float* adv, float* pattern, dmh_stream_t stream) {
    DMH_REQUIRE(obj && pattern_pos && pattern_neg && adv && n > 0, "dmh_l0_finalize: null pointer or n <= 0");
    DMH_LAUNCH(l0_finalize_kernel, ceil_div(n, 256), 256, 0, (cudaStream_t)stream)(obj, pattern_pos, pattern_neg, n, clip_max,
                                                                                threshold, adv, pattern);
    DMH_CHECK_LAUNCH("dmh_l0_finalize");
    return DMH_OK;
}

int dmh_topk_select(float* pattern_pos, float* pattern_neg, int C, int H, int W, int k, unsigned char* keep,
                    unsigned* kth_key, dmh_stream_t stream) {
    DMH_REQUIRE(pattern_pos && pattern_neg, "dmh_topk_select: null pointer");
    DMH_REQUIRE(C > 0 && H > 0 && W > 0, "dmh_topk_select: bad shape");
    DMH_LAUNCH(topk_select_kernel, 1, TK_THREADS, 0, (cudaStream_t)stream)(pattern_pos, pattern_neg, C, H * W, k, keep, kth_key);
    DMH_CHECK_LAUNCH("dmh_topk_select");
    return DMH_OK;
}

}  // extern "C"

// Fast path of dmh_photo_scale for ONE source frame (the headline stereo
// configuration, frame_ids [0,'s']; also mono with a single source when no pose
// gradient is requested).  Same contract and results as photo_scale_kernel<1>
// in photo_objective.cu, ~3.4x fewer instructions:
//
//   * 32x32 tile, 256 threads; every thread OWNS 4 interior pixels of one
//     column (rows 4s..4s+3) in the warp phase and in the gradient phase, so
//     the whole backward chain of the gather (bilinear taps -> coordinates ->
//     projection -> depth -> disparity) collapses into 3 scalars per pixel
//     held in registers (no second gather, no coordinate recompute);
//   * SSIM statistics by separable sliding windows: a thread walks down a
//     column of the 34x34 ring and reuses the 3-wide row sums across the 3
//     windows they belong to; value and backward coefficients share one
//     reciprocal; the automask decision gates the coefficients in the same pass
//     (single source frame -> the winner is known immediately);
//   * the 3x3 box sums of the 9 coefficient planes are separable sliding
//     windows too; reflection multiplicities are folded into the edge weights.
//
//   * the source is gathered from a pixel-packed (B,H,W,4) copy written once per step by the identity-loss
//     kernel below: one 128-bit load per bilinear tap, the north-west tap clamped so that the other three sit
//     at fixed offsets (bit-identical to the planar gather);
//   * SSIM in sum form (no divisions by 9), branch-free, channels (0,1) in packed fp32 (explicitly rounded
//     add/mul/fma.rn.f32x2), coefficient planes laid out for 128-bit shared-memory accesses.
//
// Shared memory: tgt 3x36x40 + pred 3x36x36 + coef 9x34x34 floats + gate bytes = 75.8 KB -> 3 CTAs / SM.
// Roofline: HBM by traffic (40 B per target pixel, 419 MB per launch at B=32 -- ncu measures exactly that), but
// instruction-issue bound in practice: ~745 lane-instructions / pixel at ~57 % issue utilisation (profiles/).
//
// Also here: ident_fast_kernel (identity reprojection loss + the packed copy, both tiles by TMA) and the
// two-kernel alternative warp_pred_kernel + photo_fast_kernel<SPLIT> (dmh_photo_scale_split: bit-identical,
// measured slower, opt-in).
//
// Target tile staging: ONE elected thread issues a 4-D TMA load
// (cp.async.bulk.tensor, box 40 x 36 x 3 x 1 at (x0-4, y0-2, 0, b) -- the innermost
// start coordinate must be 16-byte aligned, measured: x0-2 raises an illegal
// instruction -- out-of-image elements zero-filled) that lands on an mbarrier while all warps run the gather
// phase; the 1-px reflection (ReflectionPad2d) of border tiles is patched in
// shared memory afterwards.  Rows that are not 16-byte aligned (W % 4 != 0 or a
// misaligned base) take the plain-load instantiation of the same kernel.
#include <cuda.h>
#include <string.h>

#include "../../include/dmh_b200.h"
#include "dmh_common.cuh"

using namespace dmh;

#include "photo_tile.cuh"

namespace {

// TMA: target tile by cp.async.bulk.tensor; FASTDIV: verified 3-instruction division by W-1 / H-1;
// PK: pixel-packed source (128-bit taps); UP: the disparity map is smaller than the frame (scales 1..3).
// SSIM is always on and the input is a disparity here (no_ssim, depth inputs and the materialised warped images
// take the general kernel).
// SPLIT: the warp was done by warp_pred_kernel (below): the warped tile arrives by a second TMA load from its padded,
// reflect-bordered output and the backward factors D from global memory -- this kernel is then phases B and C only.
template <bool TMA, bool FASTDIV, bool PK, bool UP, bool SPLIT, bool DH = false>
__global__ void __launch_bounds__(FT_THREADS, 3)
photo_fast_kernel(const FastParams p, const __grid_constant__ CUtensorMap tgt_map,
                  const __grid_constant__ CUtensorMap pred_map) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t tgt_bar;
    float* tgt = smem;                       // [3][36][40] (TMA destination: 128-byte aligned)
    float* pred = tgt + 3 * FT_NT;           // [3][N2]
    // gated SSIM coefficients of a ring pixel, laid out for 128-bit accesses: Q1 = (a0, a1, b0, b1),
    // Q2 = (c0, c1, a2, b2), Q3 = c2  (a, b, c of channels 0, 1, 2)
    float4* coefQ1 = reinterpret_cast<float4*>(pred + 3 * FT_N2);  // [N1]
    float4* coefQ2 = coefQ1 + FT_N1;                                // [N1]
    float* coefQ3 = reinterpret_cast<float*>(coefQ2 + FT_N1);       // [N1]
    float* cams = coefQ3 + FT_N1;            // [24]
    float* red = cams + 24;                  // [32]
    uint8_t* gate = reinterpret_cast<uint8_t*>(red + 32);   // [N1]

    const int tid = threadIdx.x;
    const int H = p.H, W = p.W;
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * FT_T, y0 = blockIdx.y * FT_T;
    const int N = H * W;                      // per-item offsets fit 32 bits (checked by the launcher)
    const float w_ssim = 0.85f / 3.0f, w_l1 = 0.15f / 3.0f;

    if (SPLIT) {
    } else if (tid < 12) {
        const int i = tid / 4, j = tid % 4;
        const float* k = p.K + b * 16 + i * 4;
        const float* tt = p.T + b * 16 + j;
        float acc = __ldg(k) * __ldg(tt);
        acc = fmaf(__ldg(k + 1), __ldg(tt + 4), acc);
        acc = fmaf(__ldg(k + 2), __ldg(tt + 8), acc);
        acc = fmaf(__ldg(k + 3), __ldg(tt + 12), acc);
        cams[tid] = acc;
    } else if (tid < 21) {
        const int i = (tid - 12) / 3, j = (tid - 12) % 3;
        cams[tid] = __ldg(p.inv_K + b * 16 + i * 4 + j);
    }
    if (TMA) {
        // ---- target tile by TMA: in flight during the whole gather phase
        if (tid == 0) {
            mbar_init(&tgt_bar, 1);
            mbar_expect_tx(&tgt_bar, (3 * FT_NT + (SPLIT ? 3 * FT_N2 : 0)) * sizeof(float));
            tma_load_4d(tgt, &tgt_map, &tgt_bar, x0 - 2 - FT_TO, y0 - 2, 0, b);
            // warped tile: the padded layout puts image column x at x + 2, so the halo start x0 - 2 is column x0
            if (SPLIT) tma_load_4d(pred, &pred_map, &tgt_bar, x0, y0, 0, b);
        }
    } else if (tid < FT_R2 * 7) {
        // ---- target tile, 2-px reflect halo: 36 columns x 7 row groups = 252 threads, column index maths once
        const int c = tid % FT_R2, rg = tid / FT_R2;
        const float* tp = p.target + (size_t)b * 3 * N + ext_to_img(x0 - 2 + c, W);
        for (int r = rg; r < FT_R2; r += 7) {
            const int o = ext_to_img(y0 - 2 + r, H) * W;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) tgt[ch * FT_NT + r * FT_TP + c + FT_TO] = __ldg(tp + ch * N + o);
        }
    }
    __syncthreads();
    const int oc = tid & 31, os = tid >> 5;
    float D[4][3];
    if (!SPLIT) {
        Camera cam;
#pragma unroll
        for (int i = 0; i < 12; ++i) cam.P[i] = cams[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) cam.iK[i] = cams[12 + i];
        const float* sp = p.src + (size_t)b * (PK ? 4 : 3) * N;
        const float* dp = p.disp.ptr + (size_t)b * (p.disp.h * p.disp.w);

        // ---- phase A: warp.  A thread owns the interior pixels of column tid%32, rows 4*(tid/32)+k, plus one pixel
        // of the halo ring.  Software-pipelined over those 5 pixels: all disparity loads, then the coordinate chains,
        // then the gathers -- the memory latency is paid once per stage instead of once per pixel.
        {
            int hr, hc;
            halo_rc(tid, hr, hc);
            int py[5], pxx[5];
            const int ixo = tile_to_img(x0 + oc, W);
#pragma unroll
            for (int k = 0; k < 4; ++k) { py[k] = tile_to_img(y0 + 4 * os + k, H); pxx[k] = ixo; }
            py[4] = ext_to_img(y0 - 2 + hr, H);
            pxx[4] = ext_to_img(x0 - 2 + hc, W);
            float dv[5];
            if (!UP) {
#pragma unroll
                for (int k = 0; k < 5; ++k) dv[k] = __ldg(dp + py[k] * W + pxx[k]);
            } else {
                const UpTap txo = up_tap(ixo, p.disp.sw, p.disp.w);        // shared by the 4 owned pixels
#pragma unroll
                for (int k = 0; k < 4; ++k) dv[k] = up_sample(dp, p.disp.w, up_tap(py[k], p.disp.sh, p.disp.h), txo);
                dv[4] = up_sample(dp, p.disp.w, up_tap(py[4], p.disp.sh, p.disp.h), up_tap(pxx[4], p.disp.sw, p.disp.w));
            }
            Tap tp[5];
            float gax[5], gay[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) tp[k] = pixel_tap<FASTDIV>(cam, p, pxx[k], py[k], dv[k], k < 4, gax[k], gay[k]);
            // gathers: the taps of TWO pixels are requested before either is consumed
#pragma unroll
            for (int k0 = 0; k0 < 5; k0 += 2) {
                float tv[2][3][4];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (k0 + j < 5) load_taps<PK>(sp, N, W, tp[k0 + j], tv[j]);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int k = k0 + j;
                    if (k >= 5) continue;
                    const Gathered g = combine_taps(tv[j], tp[k], k < 4);
                    const int r = 4 * os + k;
                    const int i2 = (k < 4) ? (r + 2) * FT_R2 + oc + 2 : hr * FT_R2 + hc;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) pred[ch * FT_N2 + i2] = g.v[ch];
                    if (k < 4) {
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) D[k][ch] = g.dix[ch] * gax[k] + g.diy[ch] * gay[k];
                    }
                }
            }
        }
        // the remaining 16 pixels of the halo ring
        if (tid < 272 - FT_THREADS) {
            int r, c;
            halo_rc(tid + FT_THREADS, r, c);
            const int iy = ext_to_img(y0 - 2 + r, H), ix = ext_to_img(x0 - 2 + c, W);
            const float dvh = UP ? up_sample(dp, p.disp.w, up_tap(iy, p.disp.sh, p.disp.h), up_tap(ix, p.disp.sw, p.disp.w))
                                 : __ldg(dp + iy * W + ix);
            float u0, u1;
            const Tap th = pixel_tap<FASTDIV>(cam, p, ix, iy, dvh, false, u0, u1);
            float tvh[3][4];
            load_taps<PK>(sp, N, W, th, tvh);
            const Gathered g = combine_taps(tvh, th, false);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) pred[ch * FT_N2 + r * FT_R2 + c] = g.v[ch];
        }
    }
    // identity losses (+ tie-break noise) of this thread's phase-B pixels: requested before the barrier so
    // that their latency overlaps the other warps' gather.  Ring pixels outside the image are marked by a NaN
    // (rp < NaN is false: they never win, their coefficients are gated to 0 and they add nothing to the loss).
    float idv_pre[FT_ROWS];
    unsigned hflags;
    prefetch_ident<DH>(p, tid, b, x0, y0, idv_pre, hflags);
    __syncthreads();
    if (TMA) {
        mbar_wait(&tgt_bar, 0);
        // ReflectionPad2d(1) at the image border: TMA zero-fills out-of-image elements; patch them from the
        // in-image rows / columns of the same tile (sources are never patched themselves)
        if (x0 < 2 || y0 < 2 || x0 + FT_T + 2 > W || y0 + FT_T + 2 > H) {
            for (int i = tid; i < FT_N2; i += FT_THREADS) {
                const int r = i / FT_R2, c = i - r * FT_R2;
                const int ey = y0 - 2 + r, ex = x0 - 2 + c;
                if (ey < 0 || ey >= H || ex < 0 || ex >= W) {
                    const int sr = ext_to_img(ey, H) - (y0 - 2), sc = ext_to_img(ex, W) - (x0 - 2);
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch)
                        tgt[ch * FT_NT + r * FT_TP + c + FT_TO] = tgt[ch * FT_NT + sr * FT_TP + sc + FT_TO];
                }
            }
            __syncthreads();
        }
    }

    TileSmem sm;
    sm.tgt = tgt; sm.pred = pred; sm.q1 = coefQ1; sm.q2 = coefQ2; sm.q3 = coefQ3; sm.gate = gate;
    float acc4[4] = {0.f, 0.f, 0.f, 0.f};
    const float loss_local = phase_b<DH, UP>(p, sm, tid, b, x0, y0, idv_pre, hflags, acc4);
    __syncthreads();
    if (SPLIT) {
        // d(pred)/d(disp) factors of this thread's 4 pixels, written by warp_pred_kernel (loaded here, not before
        // phase B: 12 registers less across the SSIM phase)
        const float* dg = p.dfac + (size_t)b * 3 * N + x0 + oc;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int y = y0 + 4 * os + k;
            const bool in = y < H && x0 + oc < W;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) D[k][ch] = in ? __ldg(dg + ch * N + y * W) : 0.0f;
        }
    }
    phase_c<DH, UP>(p, sm, tid, b, x0, y0, D, acc4);
    const int blk = (b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (DH) {
        // [4][dh_nblk]: sum reproj*mask_r, sum mask_r, sum proxy*mask_h, sum mask_h
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float s = block_sum(acc4[q], red);
            if (tid == 0) p.loss_partial[q * p.dh_nblk + blk] = s;
        }
    } else {
        const float s = block_sum(loss_local, red);
        if (tid == 0) p.loss_partial[blk] = s;
    }
}


// ---------------------------------------------------------------------------
// Producer / consumer form: a persistent CTA of 12 warps walks over the tiles.  Warps 8-11 (producers) gather the
// warped tile of tile i+1 into the second shared-memory buffer and start its target TMA while warps 0-7 (consumers)
// run phases B and C of tile i: neither the gather latency nor the phase barriers of one tile leave the SM's issue
// slots idle.  Hand-off through mbarriers (full[s]: 128 producer arrivals + the TMA bytes; empty[s]: 256 consumer
// arrivals); consumers synchronise among themselves with named barrier 1, producers with named barrier 2.  The
// backward factors D travel through global memory (written by the producer, read back through L2 by the consumer
// that owns the pixel in phase C).  2 CTAs per SM (108.6 KB each).  Same arithmetic, same bits as the fused kernel.
#define PC_CONS 256
#ifndef PC_PROD
#define PC_PROD 256
#endif
#define PC_THREADS (PC_CONS + PC_PROD)
#define PC_CTAS (PC_PROD == 128 ? 2 : 1)
#define PC_HALVES (1024 / (PC_PROD * 4))          // groups of 4 rows per producer thread
#define PC_NH ((272 + PC_PROD - 1) / PC_PROD)      // halo pixels per producer thread (the last one partial)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <bool FASTDIV, bool PK, bool UP>
__global__ void __launch_bounds__(PC_THREADS, PC_CTAS)
photo_pc_kernel(const FastParams p, const __grid_constant__ CUtensorMap tgt_map, int tiles_x, int tiles_y, int ntiles) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2];
    // buffers s = 0, 1: target tiles at smem + s * 3*FT_NT (TMA destinations, 128-byte aligned), warped tiles after
    float4* q1 = reinterpret_cast<float4*>(smem + 6 * FT_NT + 6 * FT_N2);
    float4* q2 = q1 + FT_N1;
    float* q3 = reinterpret_cast<float*>(q2 + FT_N1);
    float* cams = q3 + FT_N1;                 // [2][24]
    float* red = cams + 48;                   // [32]
    uint8_t* gate = reinterpret_cast<uint8_t*>(red + 32);

    const int tid = threadIdx.x;
    const int H = p.H, W = p.W, N = H * W;
    if (tid == 0) {
        mbar_init(&full_bar[0], PC_PROD + 1); mbar_init(&full_bar[1], PC_PROD + 1);
        mbar_init(&empty_bar[0], PC_CONS); mbar_init(&empty_bar[1], PC_CONS);
    }
    __syncthreads();
    const int per_img = tiles_x * tiles_y;

    if (tid >= PC_CONS) {
        // ================================================================ producers
        const int pt = tid - PC_CONS;
        const int pc_ = pt & 31, pw = pt >> 5;                       // column, row group of the interior
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int s = it & 1, ph = (it >> 1) & 1;
            const int b = t / per_img, rem = t - b * per_img;
            const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
            const int x0 = tx * FT_T, y0 = ty * FT_T;
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (pt == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&full_bar[s], 3 * FT_NT * sizeof(float));
                tma_load_4d(smem + s * 3 * FT_NT, &tgt_map, &full_bar[s], x0 - 2 - FT_TO, y0 - 2, 0, b);
            }
            float* cm = cams + 24 * s;
            if (pt < 12) {
                const int i = pt / 4, j = pt % 4;
                const float* k = p.K + b * 16 + i * 4;
                const float* tt = p.T + b * 16 + j;
                float acc = __ldg(k) * __ldg(tt);
                acc = fmaf(__ldg(k + 1), __ldg(tt + 4), acc);
                acc = fmaf(__ldg(k + 2), __ldg(tt + 8), acc);
                acc = fmaf(__ldg(k + 3), __ldg(tt + 12), acc);
                cm[pt] = acc;
            } else if (pt < 21) {
                const int i = (pt - 12) / 3, j = (pt - 12) % 3;
                cm[pt] = __ldg(p.inv_K + b * 16 + i * 4 + j);
            }
            named_bar(2, PC_PROD);
            Camera cam;
#pragma unroll
            for (int i = 0; i < 12; ++i) cam.P[i] = cm[i];
#pragma unroll
            for (int i = 0; i < 9; ++i) cam.iK[i] = cm[12 + i];
            const float* sp = p.src + (size_t)b * (PK ? 4 : 3) * N;
            const float* dp = p.disp.ptr + (size_t)b * (p.disp.h * p.disp.w);
            float* pred = smem + 6 * FT_NT + s * 3 * FT_N2;
            float* dout = p.dfac_out + (size_t)b * 3 * N;
            // ---- interior: column pc_, rows 8*pw .. 8*pw+7, four at a time (column taps shared)
            const int ixo = tile_to_img(x0 + pc_, W);
#pragma unroll 1
            for (int half = 0; half < PC_HALVES; ++half) {
                const int rb = 4 * PC_HALVES * pw + 4 * half;
                int py[4];
                float dv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) py[k] = tile_to_img(y0 + rb + k, H);
                if (!UP) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) dv[k] = __ldg(dp + py[k] * W + ixo);
                } else {
                    const UpTap txo = up_tap(ixo, p.disp.sw, p.disp.w);
#pragma unroll
                    for (int k = 0; k < 4; ++k) dv[k] = up_sample(dp, p.disp.w, up_tap(py[k], p.disp.sh, p.disp.h), txo);
                }
                Tap tp[4];
                float gax[4], gay[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) tp[k] = pixel_tap<FASTDIV>(cam, p, ixo, py[k], dv[k], true, gax[k], gay[k]);
#pragma unroll
                for (int k0 = 0; k0 < 4; k0 += 2) {
                    float tv[2][3][4];
#pragma unroll
                    for (int j = 0; j < 2; ++j) load_taps<PK>(sp, N, W, tp[k0 + j], tv[j]);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int k = k0 + j, r = rb + k;
                        const Gathered g = combine_taps(tv[j], tp[k], true);
                        const int i2 = (r + 2) * FT_R2 + pc_ + 2;
                        const bool in = y0 + r < H && x0 + pc_ < W;
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            pred[ch * FT_N2 + i2] = g.v[ch];
                            if (in) dout[ch * N + (y0 + r) * W + x0 + pc_] = g.dix[ch] * gax[k] + g.diy[ch] * gay[k];
                        }
                    }
                }
            }
            // ---- halo ring: 272 pixels, up to 3 per thread (the third only for the first 16 threads)
            {
                int hr[PC_NH], hc[PC_NH];
                Tap th[PC_NH];
                const int nh = pt < 272 - (PC_NH - 1) * PC_PROD ? PC_NH : PC_NH - 1;
#pragma unroll
                for (int j = 0; j < PC_NH; ++j) {
                    if (j < nh) {
                        halo_rc(pt + PC_PROD * j, hr[j], hc[j]);
                        const int iy = ext_to_img(y0 - 2 + hr[j], H), ix = ext_to_img(x0 - 2 + hc[j], W);
                        const float dvh = UP ? up_sample(dp, p.disp.w, up_tap(iy, p.disp.sh, p.disp.h),
                                                         up_tap(ix, p.disp.sw, p.disp.w))
                                             : __ldg(dp + iy * W + ix);
                        float u0, u1;
                        th[j] = pixel_tap<FASTDIV>(cam, p, ix, iy, dvh, false, u0, u1);
                    }
                }
#pragma unroll
                for (int j = 0; j < PC_NH; ++j) {
                    if (j < nh) {
                        float tvh[3][4];
                        load_taps<PK>(sp, N, W, th[j], tvh);
                        const Gathered g = combine_taps(tvh, th[j], false);
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) pred[ch * FT_N2 + hr[j] * FT_R2 + hc[j]] = g.v[ch];
                    }
                }
            }
            __threadfence_block();
            mbar_arrive(&full_bar[s]);
        }
    } else {
        // ================================================================ consumers
        const int oc = tid & 31, os = tid >> 5;
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int s = it & 1, ph = (it >> 1) & 1;
            const int b = t / per_img, rem = t - b * per_img;
            const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
            const int x0 = tx * FT_T, y0 = ty * FT_T;
            float idv_pre[FT_ROWS];
            unsigned hflags;
            prefetch_ident<false>(p, tid, b, x0, y0, idv_pre, hflags);
            mbar_wait(&full_bar[s], ph);
            float* tgt = smem + s * 3 * FT_NT;
            if (x0 < 2 || y0 < 2 || x0 + FT_T + 2 > W || y0 + FT_T + 2 > H) {      // ReflectionPad2d(1) of the target
                for (int i = tid; i < FT_N2; i += PC_CONS) {
                    const int r = i / FT_R2, c = i - r * FT_R2;
                    const int ey = y0 - 2 + r, ex = x0 - 2 + c;
                    if (ey < 0 || ey >= H || ex < 0 || ex >= W) {
                        const int sr = ext_to_img(ey, H) - (y0 - 2), sc = ext_to_img(ex, W) - (x0 - 2);
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch)
                            tgt[ch * FT_NT + r * FT_TP + c + FT_TO] = tgt[ch * FT_NT + sr * FT_TP + sc + FT_TO];
                    }
                }
                named_bar(1, PC_CONS);
            }
            TileSmem sm;
            sm.tgt = tgt; sm.pred = smem + 6 * FT_NT + s * 3 * FT_N2; sm.q1 = q1; sm.q2 = q2; sm.q3 = q3; sm.gate = gate;
            float acc4[4];
            const float loss_local = phase_b<false, UP>(p, sm, tid, b, x0, y0, idv_pre, hflags, acc4);
            // backward factors of this thread's 4 pixels (written by the producer of this CTA: read through L2)
            float D[4][3];
            {
                const float* dg = p.dfac_out + (size_t)b * 3 * N + x0 + oc;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int y = y0 + 4 * os + k;
                    const bool in = y < H && x0 + oc < W;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) D[k][ch] = in ? __ldcg(dg + ch * N + y * W) : 0.0f;
                }
            }
            const float wsum = warp_sum(loss_local);
            if ((tid & 31) == 0) red[tid >> 5] = wsum;
            named_bar(1, PC_CONS);                                   // coefficient planes + red[] complete
            phase_c<false, UP>(p, sm, tid, b, x0, y0, D, acc4);
            if (tid == 0) {
                float tot = 0.f;
#pragma unroll
                for (int i = 0; i < PC_CONS / 32; ++i) tot += red[i];
                p.loss_partial[t] = tot;
            }
            mbar_arrive(&empty_bar[s]);                              // this thread is done with buffer s
            named_bar(1, PC_CONS);                                   // nobody still reads the coefficient planes / red[]
        }
    }
}

// ---------------------------------------------------------------------------
// Split design, kernel 1: the warp alone.  One thread per 4 pixels of a column (no halo: every pixel of the frame is
// gathered exactly once, against 1.27x in the fused kernel), low register / no shared-memory footprint, i.e. enough
// resident warps to hide the gather latency.  Writes
//   predp (B,3,H+4,WP): the warped frame with a 2-pixel border, image pixel (x,y) at (x+2,y+2); the border holds the
//          ReflectionPad2d mirror (columns -1 / W = columns 1 / W-2, rows alike; the outer border ring only needs to
//          be finite), so that the loss kernel's TMA box starts on a 16-byte boundary and needs no patching;
//   dfac  (B,3,H,W): d(pred_ch)/d(disp) * grad_scale (the collapsed backward chain of the gather).
template <bool FASTDIV, bool PK, bool UP>
__global__ void __launch_bounds__(FT_THREADS, 4)
warp_pred_kernel(const FastParams p) {
    __shared__ float cams[24];
    const int tid = threadIdx.x;
    const int H = p.H, W = p.W;
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * FT_T, y0 = blockIdx.y * FT_T;
    const int N = H * W;
    if (tid < 12) {
        const int i = tid / 4, j = tid % 4;
        const float* k = p.K + b * 16 + i * 4;
        const float* tt = p.T + b * 16 + j;
        float acc = __ldg(k) * __ldg(tt);
        acc = fmaf(__ldg(k + 1), __ldg(tt + 4), acc);
        acc = fmaf(__ldg(k + 2), __ldg(tt + 8), acc);
        acc = fmaf(__ldg(k + 3), __ldg(tt + 12), acc);
        cams[tid] = acc;
    } else if (tid < 21) {
        const int i = (tid - 12) / 3, j = (tid - 12) % 3;
        cams[tid] = __ldg(p.inv_K + b * 16 + i * 4 + j);
    }
    const int oc = tid & 31, os = tid >> 5;
    const int px = min(x0 + oc, W - 1);
    const float* dp = p.disp.ptr + (size_t)b * (p.disp.h * p.disp.w);
    int py[4];
    float dv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) py[k] = min(y0 + 4 * os + k, H - 1);
    if (!UP) {
#pragma unroll
        for (int k = 0; k < 4; ++k) dv[k] = __ldg(dp + py[k] * W + px);
    } else {
        const UpTap txo = up_tap(px, p.disp.sw, p.disp.w);
#pragma unroll
        for (int k = 0; k < 4; ++k) dv[k] = up_sample(dp, p.disp.w, up_tap(py[k], p.disp.sh, p.disp.h), txo);
    }
    __syncthreads();
    Camera cam;
#pragma unroll
    for (int i = 0; i < 12; ++i) cam.P[i] = cams[i];
#pragma unroll
    for (int i = 0; i < 9; ++i) cam.iK[i] = cams[12 + i];
    const float* sp = p.src + (size_t)b * (PK ? 4 : 3) * N;
    Tap tp[4];
    float gax[4], gay[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) tp[k] = pixel_tap<FASTDIV>(cam, p, px, py[k], dv[k], true, gax[k], gay[k]);
    if (x0 + oc >= W) return;
    const int HP = H + 4, WP = p.WP;
    // mirror column / row of the 2-pixel border this pixel also fills (-1: none)
    const int xp = px + 2;
    const int mxp = px == 1 ? 1 : (px == 2 ? 0 : (px == W - 2 ? W + 2 : (px == W - 3 ? W + 3 : -1)));
    float* pb = p.predp + (size_t)b * 3 * HP * WP;
    float* db = p.dfac_out + (size_t)b * 3 * N + px;
#pragma unroll
    for (int k0 = 0; k0 < 4; k0 += 2) {
        float tv[2][3][4];
#pragma unroll
        for (int j = 0; j < 2; ++j) load_taps<PK>(sp, N, W, tp[k0 + j], tv[j]);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int k = k0 + j;
            const int y = y0 + 4 * os + k;
            if (y >= H) continue;
            const Gathered g = combine_taps(tv[j], tp[k], true);
            const int yp = y + 2;
            const int myp = y == 1 ? 1 : (y == 2 ? 0 : (y == H - 2 ? H + 2 : (y == H - 3 ? H + 3 : -1)));
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float* pc = pb + (size_t)ch * HP * WP;
                pc[yp * WP + xp] = g.v[ch];
                if (mxp >= 0) pc[yp * WP + mxp] = g.v[ch];
                if (myp >= 0) {
                    pc[myp * WP + xp] = g.v[ch];
                    if (mxp >= 0) pc[myp * WP + mxp] = g.v[ch];
                }
                db[ch * N + y * W] = g.dix[ch] * gax[k] + g.diy[ch] * gay[k];
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Identity reprojection loss (M2/trainer.py:608-615): compute_reprojection_loss of the UN-warped
// source against the target.  Scale independent -> computed once per batch (the reference recomputes it
// for every scale).  32x32 tile, 1-px reflect halo, sliding-window SSIM down each column.
#define ID_R 34
#define ID_N (ID_R * ID_R)
struct IdentParams {
    const float* target;
    const float* src[DMH_PHOTO_MAX_FRAMES];
    float* out;                 // (B,F,H,W); nullable when only the packed copy is wanted
    float4* packed;             // nullable (F == 1): source frame 0 re-laid out as (B,H,W,4) for the 128-bit tap gather
    int B, F, H, W, no_ssim;
};

// TMA: both tiles arrive by cp.async.bulk.tensor (box 40 x 34 x 3 at (x0-4, y0-1): 16-byte aligned start column,
// zero fill outside the image, reflection patched in shared memory on border tiles) -- no per-thread load / index
// instructions at all; otherwise plain loads with the same shared-memory layout.
#define ID_P 40                  // row pitch of the tiles
#define ID_O 3                   // column of the tile that holds halo column 0 (image column x0-1)
#define ID_PL (ID_R * ID_P)      // plane stride
struct IdentMaps { CUtensorMap tgt; CUtensorMap src[DMH_PHOTO_MAX_FRAMES]; };

// Shared tail of the identity-loss kernels: packed copy of the source tile (+ optional fp32 copy of the target tile,
// bf16 inputs) and the identity reprojection loss of the tile from the two staged 34 x 34 x 3 tiles.
__device__ __forceinline__ void ident_tile_tail(const IdentParams& p, const float* xs, const float* ys, int tid, int b,
                                                int f, int x0, int y0, float* tgt_f32) {
    const int H = p.H, W = p.W;
    const size_t N = (size_t)H * W;
    __syncthreads();
    const int c = tid & 31, strip = tid >> 5;           // 8 strips of 4 rows
    const int px = x0 + c;
    if (p.packed && px < W) {     // pixel-packed copy of the source tile: a warp writes 512 contiguous bytes per row
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = 4 * strip + k, py = y0 + r;
            if (py < H) {
                const int i = (r + 1) * ID_P + c + 1 + ID_O;
                p.packed[(size_t)b * N + (size_t)py * W + px] = make_float4(xs[i], xs[ID_PL + i], xs[2 * ID_PL + i], 0.0f);
            }
        }
    }
    if (tgt_f32 && x0 + c < W) {   // bf16 inputs: the fp32 target the photometric kernels stage by TMA
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = 4 * strip + k, py = y0 + r;
            if (py < H) {
                const int i = (r + 1) * ID_P + c + 1 + ID_O;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) tgt_f32[((size_t)b * 3 + ch) * N + (size_t)py * W + x0 + c] = ys[ch * ID_PL + i];
            }
        }
    }
    if (!p.out) return;
    // same lane-typed SSIM arithmetic as the fused kernels (channels 0,1 packed, channel 2 scalar): the
    // identity loss and the reprojection loss must come out of identical arithmetic (automask ties)
    Row5T<float2> histP[2];
    Row5T<float> histS[2];
    float2 cenxP = make_float2(0.f, 0.f), cenyP = cenxP;
    float cenxS = 0.f, cenyS = 0.f;
#pragma unroll
    for (int rr = 0; rr < 6; ++rr) {
        const int r2 = 4 * strip + rr;                   // halo-tile row
        const float* x0p = xs + r2 * ID_P + c + ID_O;
        const float* y0p = ys + r2 * ID_P + c + ID_O;
        const float2 xa = make_float2(x0p[0], x0p[ID_PL]), xb = make_float2(x0p[1], x0p[ID_PL + 1]),
                     xc = make_float2(x0p[2], x0p[ID_PL + 2]);
        const float2 ya = make_float2(y0p[0], y0p[ID_PL]), yb = make_float2(y0p[1], y0p[ID_PL + 1]),
                     yc = make_float2(y0p[2], y0p[ID_PL + 2]);
        const Row5T<float2> curP = row5(xa, xb, xc, ya, yb, yc);
        const float* x2p = x0p + 2 * ID_PL;
        const float* y2p = y0p + 2 * ID_PL;
        const Row5T<float> curS = row5(x2p[0], x2p[1], x2p[2], y2p[0], y2p[1], y2p[2]);
        if (rr >= 2) {
            const int py = y0 + 4 * strip + rr - 2;
            if (py < H && px < W) {
                float l1 = fabsf(cenyP.x - cenxP.x);
                l1 += fabsf(cenyP.y - cenxP.y);
                l1 += fabsf(cenyS - cenxS);
                float ss = 0.f;
                if (!p.no_ssim) {
                    float2 passP, rP, nrP;
                    const float2 vP = ssim_value_t(ssim_stats_rows_t(histP[0], histP[1], curP), passP, rP, nrP);
                    float passS, rS, nrS;
                    const float vS = ssim_value_t(ssim_stats_rows_t(histS[0], histS[1], curS), passS, rS, nrS);
                    ss = (vP.x + vP.y) + vS;
                }
                l1 *= (1.0f / 3.0f);
                const float rp = p.no_ssim ? l1 : fmaf(0.85f, ss * (1.0f / 3.0f), 0.15f * l1);
                p.out[((size_t)b * p.F + f) * N + (size_t)py * W + px] = rp;
            }
        }
        histP[0] = histP[1]; histP[1] = curP;
        histS[0] = histS[1]; histS[1] = curS;
        cenxP = xb; cenyP = yb; cenxS = x2p[1]; cenyS = y2p[1];
    }
}

template <bool TMA>
__global__ void __launch_bounds__(256, 4)
ident_fast_kernel(const IdentParams p, const __grid_constant__ IdentMaps maps) {
    __shared__ __align__(128) float xs[3 * ID_PL];
    __shared__ __align__(128) float ys[3 * ID_PL];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const int H = p.H, W = p.W;
    const int b = blockIdx.z / p.F, f = blockIdx.z % p.F;
    const int x0 = blockIdx.x * FT_T, y0 = blockIdx.y * FT_T;
    const size_t N = (size_t)H * W;
    if (TMA) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            mbar_expect_tx(&bar, 2 * 3 * ID_PL * sizeof(float));
            tma_load_4d(xs, &maps.src[f], &bar, x0 - 1 - ID_O, y0 - 1, 0, b);
            tma_load_4d(ys, &maps.tgt, &bar, x0 - 1 - ID_O, y0 - 1, 0, b);
        }
        __syncthreads();                                  // barrier initialised before anyone polls it
        mbar_wait(&bar, 0);
        if (x0 < 1 || y0 < 1 || x0 + FT_T + 1 > W || y0 + FT_T + 1 > H) {
            for (int i = tid; i < ID_N; i += 256) {
                const int r = i / ID_R, c = i - r * ID_R;
                const int ey = y0 - 1 + r, ex = x0 - 1 + c;
                if (ey < 0 || ey >= H || ex < 0 || ex >= W) {
                    const int sr = ext_to_img(ey, H) - (y0 - 1), sc = ext_to_img(ex, W) - (x0 - 1);
                    // a reflected source may itself lie outside the tile when the image is narrower than the halo
                    if (sr >= 0 && sr < ID_R && sc >= 0 && sc < ID_R) {
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            xs[ch * ID_PL + r * ID_P + c + ID_O] = xs[ch * ID_PL + sr * ID_P + sc + ID_O];
                            ys[ch * ID_PL + r * ID_P + c + ID_O] = ys[ch * ID_PL + sr * ID_P + sc + ID_O];
                        }
                    }
                }
            }
        }
    } else {
        // tile + 1-px reflect halo: a warp takes a row (index maths once per row / per lane), 6 loads in flight
        const float* sp = p.src[f] + (size_t)b * 3 * N;
        const float* tp = p.target + (size_t)b * 3 * N;
        const int lane = tid & 31, wid = tid >> 5;
        const int ixa = ext_to_img(x0 - 1 + lane, W);
        const int ixb = ext_to_img(x0 - 1 + 32 + (lane & 1), W);          // columns 32, 33 (lanes 0, 1)
        for (int r = wid; r < ID_R; r += 8) {
            const size_t ro = (size_t)ext_to_img(y0 - 1 + r, H) * W;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                xs[ch * ID_PL + r * ID_P + lane + ID_O] = __ldg(sp + ch * N + ro + ixa);
                ys[ch * ID_PL + r * ID_P + lane + ID_O] = __ldg(tp + ch * N + ro + ixa);
            }
            if (lane < 2) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    xs[ch * ID_PL + r * ID_P + 32 + lane + ID_O] = __ldg(sp + ch * N + ro + ixb);
                    ys[ch * ID_PL + r * ID_P + 32 + lane + ID_O] = __ldg(tp + ch * N + ro + ixb);
                }
            }
        }
    }
    ident_tile_tail(p, xs, ys, tid, b, f, x0, y0, nullptr);
}

// bf16 frames (north_star: "bf16x8 coalesced loads"): the tiles are filled by 128-bit loads of 8 bf16 each -- aligned
// chunks [x0-8, x0+40) of every tile row, zero outside the image like a TMA box, then the same reflection patch --
// and widened to fp32 in shared memory; everything downstream is the fp32 code.  Needs W % 8 == 0 and 16-byte
// aligned frames.  tgt_f32 (B,3,H,W) receives the widened target (what the per-scale kernels read).
__global__ void __launch_bounds__(256, 4)
ident_bf16_kernel(const IdentParams p, const uint16_t* __restrict__ tgt16, const uint16_t* __restrict__ src16,
                  float* __restrict__ tgt_f32) {
    __shared__ __align__(128) float xs[3 * ID_PL];
    __shared__ __align__(128) float ys[3 * ID_PL];
    const int tid = threadIdx.x;
    const int H = p.H, W = p.W;
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * FT_T, y0 = blockIdx.y * FT_T;
    // items: (frame, channel, tile row, chunk of 8 columns); tile column j holds image column x0 - 4 + j
    for (int it = tid; it < 2 * 3 * ID_R * 6; it += 256) {
        const int ck = it % 6, r = (it / 6) % ID_R, ch = (it / (6 * ID_R)) % 3, fr = it / (6 * ID_R * 3);
        const int y = y0 - 1 + r, xc = x0 - 8 + 8 * ck;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (y >= 0 && y < H && xc >= 0 && xc < W) {
            const uint16_t* base = (fr ? tgt16 : src16) + (((size_t)b * 3 + ch) * H + y) * W + xc;
            v = __ldg(reinterpret_cast<const uint4*>(base));
        }
        float* dst = (fr ? ys : xs) + ch * ID_PL + r * ID_P;
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int j = 8 * ck - 4 + e;                 // tile column
            if (j >= 0 && j < ID_P) dst[j] = __uint_as_float((e & 1) ? (w[e >> 1] & 0xffff0000u) : (w[e >> 1] << 16));
        }
    }
    __syncthreads();
    if (x0 < 1 || y0 < 1 || x0 + FT_T + 1 > W || y0 + FT_T + 1 > H) {
        for (int i = tid; i < ID_N; i += 256) {
            const int r = i / ID_R, c = i - r * ID_R;
            const int ey = y0 - 1 + r, ex = x0 - 1 + c;
            if (ey < 0 || ey >= H || ex < 0 || ex >= W) {
                const int sr = ext_to_img(ey, H) - (y0 - 1), sc = ext_to_img(ex, W) - (x0 - 1);
                if (sr >= 0 && sr < ID_R && sc >= 0 && sc < ID_R) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        xs[ch * ID_PL + r * ID_P + c + ID_O] = xs[ch * ID_PL + sr * ID_P + sc + ID_O];
                        ys[ch * ID_PL + r * ID_P + c + ID_O] = ys[ch * ID_PL + sr * ID_P + sc + ID_O];
                    }
                }
            }
        }
    }
    ident_tile_tail(p, xs, ys, tid, b, 0, x0, y0, tgt_f32);
}

size_t fast_smem_bytes() { return sizeof(float) * (3 * FT_NT + 3 * FT_N2 + 9 * FT_N1 + 24 + 32) + FT_N1; }

}  // namespace

namespace dmh {

int photo_fast_tiles(int H, int W) { return ceil_div(W, FT_T) * ceil_div(H, FT_T); }

// floats of the split design's workspace: padded warped frame + backward factors
long long photo_split_workspace_floats(int B, int H, int W) {
    return (long long)B * 3 * ((long long)(H + 4) * ((W + 4 + 3) & ~3) + (long long)H * W);
}

// Called by dmh_photo_scale when F == 1 and no pose gradient is requested.
int launch_photo_fast(const float* target, const float* src, const float* T, const float* disp, int disp_h,
                      int disp_w, const float* K, const float* inv_K, const float* ident, const float* noise, int B, int H, int W, float min_depth,
                      float max_depth, int flags, float grad_scale, float* loss_partial, float* grad_disp,
                      uint8_t* sel, float* warped, float* split_ws, const FastDhArgs* dh, cudaStream_t st) {
    FastParams p;
    p.target = target; p.src = src; p.T = T; p.K = K;
    p.disp.ptr = disp; p.disp.h = disp_h; p.disp.w = disp_w; p.disp.sh = (float)disp_h / (float)H; p.disp.sw = (float)disp_w / (float)W; p.inv_K = inv_K; p.ident = ident;
    p.noise = noise; p.loss_partial = loss_partial; p.grad_disp = grad_disp; p.sel = sel; p.warped = warped;
    p.B = B; p.H = H; p.W = W; p.flags = flags;
    const bool is_depth = (flags & DMH_PHOTO_INPUT_IS_DEPTH) != 0;
    p.ds.min_disp = is_depth ? 0.f : (float)(1.0 / (double)max_depth);
    p.ds.range = is_depth ? 0.f : (float)(1.0 / (double)min_depth - 1.0 / (double)max_depth);
    p.grad_scale = grad_scale;
    p.dfac = nullptr; p.predp = nullptr; p.dfac_out = nullptr; p.WP = 0;
    p.hint_reproj = dh ? dh->hint_reproj : nullptr; p.hint_depth = dh ? dh->hint_depth : nullptr;
    p.hint_valid = dh ? dh->hint_valid : nullptr; p.grad_hint = dh ? dh->grad_hint : nullptr;
    p.dh_nblk = dh ? dh->nblk : 0;
    const size_t smem = fast_smem_bytes();
    static bool configured_dev[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured_dev[dev & 63]) {
        cudaError_t e = cudaSuccess;
#define DMH_FAST_FN(T_, D_, P_, U_) (const void*)photo_fast_kernel<T_, D_, P_, U_, false>
#define DMH_FAST_FN4(P_, U_) DMH_FAST_FN(false, false, P_, U_), DMH_FAST_FN(false, true, P_, U_), \
                             DMH_FAST_FN(true, false, P_, U_), DMH_FAST_FN(true, true, P_, U_)
        const void* fns[16] = {DMH_FAST_FN4(false, false), DMH_FAST_FN4(false, true), DMH_FAST_FN4(true, false),
                               DMH_FAST_FN4(true, true)};
#undef DMH_FAST_FN4
#undef DMH_FAST_FN
        for (int i = 0; i < 16 && e == cudaSuccess; ++i)
            e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute((const void*)photo_fast_kernel<true, false, false, false, true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#define DMH_DH_FN(T_, D_, P_, U_) (const void*)photo_fast_kernel<T_, D_, P_, U_, false, true>
#define DMH_DH_FN4(P_, U_) DMH_DH_FN(false, false, P_, U_), DMH_DH_FN(false, true, P_, U_), \
                           DMH_DH_FN(true, false, P_, U_), DMH_DH_FN(true, true, P_, U_)
        const void* dfns[16] = {DMH_DH_FN4(false, false), DMH_DH_FN4(false, true), DMH_DH_FN4(true, false),
                                DMH_DH_FN4(true, true)};
#undef DMH_DH_FN4
#undef DMH_DH_FN
        for (int i = 0; i < 16 && e == cudaSuccess; ++i)
            e = cudaFuncSetAttribute(dfns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("dmh_photo_scale(fast): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            return DMH_ERR_CUDA;
        }
        configured_dev[dev & 63] = true;
    }
    dim3 grid(ceil_div(W, FT_T), ceil_div(H, FT_T), B);
    // TMA descriptor of the target frames viewed as a (W, H, 3, B) fp32 tensor; needs 16-byte aligned rows
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    bool use_tma = (W % 4 == 0) && ((uintptr_t)target % 16 == 0) && tma_encoder() != nullptr;
    if (use_tma) {
        const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
        const cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * 12};
        const cuuint32_t box[4] = {FT_TP, FT_R2, 3, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult r = tma_encoder()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(target), gdim, gstr,
                                         box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        use_tma = (r == CUDA_SUCCESS);
    }
    // division by W-1 / H-1 in the coordinate chain: the 3-instruction form where it is proven bit-exact
    const bool fastdiv = W > 1 && H > 1 && const_div_exact(W - 1, &p.rcw) && const_div_exact(H - 1, &p.rch);
    const bool packed = (flags & DMH_PHOTO_SRC_PACKED) != 0;
    const bool up = !(disp_h == H && disp_w == W);
    if (split_ws && (flags & DMH_PHOTO_PIPELINED)) {
        // ---- persistent producer / consumer kernel; the workspace holds the backward factors (B,3,H,W)
        if (!use_tma) {
            set_error("dmh_photo_scale_split(pipelined): needs TMA (W %% 4 == 0, 16-byte aligned frames)");
            return DMH_ERR_INVALID;
        }
        p.dfac_out = split_ws;
        p.dfac = split_ws;
        const size_t smem_pc = sizeof(float) * (6 * FT_NT + 6 * FT_N2 + 9 * FT_N1 + 48 + 32) + FT_N1;
        static bool pc_cfg[64] = {false};
        static int pc_sms[64] = {0};
        if (!pc_cfg[dev & 63]) {
            cudaError_t e = cudaSuccess;
            const void* fns[8] = {(const void*)photo_pc_kernel<false, false, false>, (const void*)photo_pc_kernel<false, false, true>,
                                  (const void*)photo_pc_kernel<false, true, false>, (const void*)photo_pc_kernel<false, true, true>,
                                  (const void*)photo_pc_kernel<true, false, false>, (const void*)photo_pc_kernel<true, false, true>,
                                  (const void*)photo_pc_kernel<true, true, false>, (const void*)photo_pc_kernel<true, true, true>};
            for (int i = 0; i < 8 && e == cudaSuccess; ++i)
                e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pc);
            if (e == cudaSuccess) e = cudaDeviceGetAttribute(&pc_sms[dev & 63], cudaDevAttrMultiProcessorCount, dev);
            if (e != cudaSuccess) {
                set_error("dmh_photo_scale_split(pipelined): configuration failed: %s", cudaGetErrorString(e));
                return DMH_ERR_CUDA;
            }
            pc_cfg[dev & 63] = true;
        }
        const int tiles_x = ceil_div(W, FT_T), tiles_y = ceil_div(H, FT_T), ntiles = tiles_x * tiles_y * B;
        const int nblk = ntiles < PC_CTAS * pc_sms[dev & 63] ? ntiles : PC_CTAS * pc_sms[dev & 63];
#define DMH_PC_GO(D_, P_, U_) DMH_LAUNCH((photo_pc_kernel<D_, P_, U_>), nblk, PC_THREADS, smem_pc, st)(p, map, tiles_x, tiles_y, ntiles)
        if (fastdiv) {
            if (packed && up) DMH_PC_GO(true, true, true);
            else if (packed) DMH_PC_GO(true, true, false);
            else if (up) DMH_PC_GO(true, false, true);
            else DMH_PC_GO(true, false, false);
        } else {
            if (packed && up) DMH_PC_GO(false, true, true);
            else if (packed) DMH_PC_GO(false, true, false);
            else if (up) DMH_PC_GO(false, false, true);
            else DMH_PC_GO(false, false, false);
        }
#undef DMH_PC_GO
        return DMH_OK;
    }
    if (split_ws) {
        // ---- split design: warp_pred_kernel (warp only, no halo) + the loss kernel fed by two TMA loads
        if (!use_tma || W < 8 || H < 8) {
            set_error("dmh_photo_scale_split: needs TMA (W %% 4 == 0, 16-byte aligned frames) and W, H >= 8");
            return DMH_ERR_INVALID;
        }
        const int HP = H + 4, WP = (W + 4 + 3) & ~3;
        p.predp = split_ws;
        p.WP = WP;
        p.dfac_out = split_ws + (size_t)B * 3 * HP * WP;
        p.dfac = p.dfac_out;
        CUtensorMap pmap;
        memset(&pmap, 0, sizeof(pmap));
        {
            const cuuint64_t gdim[4] = {(cuuint64_t)(W + 4), (cuuint64_t)HP, 3, (cuuint64_t)B};
            const cuuint64_t gstr[3] = {(cuuint64_t)WP * 4, (cuuint64_t)WP * HP * 4, (cuuint64_t)WP * HP * 12};
            const cuuint32_t box[4] = {FT_R2, FT_R2, 3, 1};
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            const CUresult r = tma_encoder()(&pmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.predp, gdim, gstr, box, estr,
                                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS || (uintptr_t)split_ws % 16 != 0) {
                set_error("dmh_photo_scale_split: tensor map of the warped frame failed (workspace must be 16-byte aligned)");
                return DMH_ERR_INVALID;
            }
        }
#define DMH_WP_GO(D_, P_, U_) DMH_LAUNCH((warp_pred_kernel<D_, P_, U_>), grid, FT_THREADS, 0, st)(p)
        if (fastdiv) {
            if (packed && up) DMH_WP_GO(true, true, true);
            else if (packed) DMH_WP_GO(true, true, false);
            else if (up) DMH_WP_GO(true, false, true);
            else DMH_WP_GO(true, false, false);
        } else {
            if (packed && up) DMH_WP_GO(false, true, true);
            else if (packed) DMH_WP_GO(false, true, false);
            else if (up) DMH_WP_GO(false, false, true);
            else DMH_WP_GO(false, false, false);
        }
#undef DMH_WP_GO
        {
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { set_error("dmh_photo_scale_split: warp launch failed: %s", cudaGetErrorString(e)); return DMH_ERR_CUDA; }
        }
        DMH_LAUNCH((photo_fast_kernel<true, false, false, false, true>), grid, FT_THREADS, smem, st)(p, map, pmap);
        return DMH_OK;
    }
#define DMH_FAST_GO(T_, D_, P_, U_)                                                                           \
    do {                                                                                                      \
        if (dh) DMH_LAUNCH((photo_fast_kernel<T_, D_, P_, U_, false, true>), grid, FT_THREADS, smem, st)(p, map, map); \
        else DMH_LAUNCH((photo_fast_kernel<T_, D_, P_, U_, false>), grid, FT_THREADS, smem, st)(p, map, map);   \
    } while (0)
#define DMH_FAST_GO2(P_, U_)                                           \
    do {                                                               \
        if (use_tma && fastdiv) DMH_FAST_GO(true, true, P_, U_);       \
        else if (use_tma) DMH_FAST_GO(true, false, P_, U_);            \
        else if (fastdiv) DMH_FAST_GO(false, true, P_, U_);            \
        else DMH_FAST_GO(false, false, P_, U_);                        \
    } while (0)
    if (packed && up) DMH_FAST_GO2(true, true);
    else if (packed) DMH_FAST_GO2(true, false);
    else if (up) DMH_FAST_GO2(false, true);
    else DMH_FAST_GO2(false, false);
#undef DMH_FAST_GO2
#undef DMH_FAST_GO
    return DMH_OK;
}


int launch_ident_fast(const float* target, const float* const* src_host, int F, int B, int H, int W, int no_ssim,
                      float* out, float* packed, cudaStream_t st) {
    IdentParams p;
    p.target = target;
    p.packed = reinterpret_cast<float4*>(packed);
    for (int f = 0; f < DMH_PHOTO_MAX_FRAMES; ++f) p.src[f] = f < F ? src_host[f] : nullptr;
    p.out = out; p.B = B; p.F = F; p.H = H; p.W = W; p.no_ssim = no_ssim;
    dim3 grid(ceil_div(W, FT_T), ceil_div(H, FT_T), B * F);
    // TMA descriptors of the frames viewed as (W, H, 3, B) fp32 tensors; rows must be 16-byte aligned
    IdentMaps maps;
    memset(&maps, 0, sizeof(maps));
    bool use_tma = (W % 4 == 0) && ((uintptr_t)target % 16 == 0) && tma_encoder() != nullptr;
    for (int f = 0; f < F && use_tma; ++f) use_tma = (uintptr_t)src_host[f] % 16 == 0;
    if (use_tma) {
        const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
        const cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * 12};
        const cuuint32_t box[4] = {ID_P, ID_R, 3, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        for (int f = -1; f < F && use_tma; ++f) {
            const float* base = f < 0 ? target : src_host[f];
            use_tma = tma_encoder()(f < 0 ? &maps.tgt : &maps.src[f], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                                    const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
    }
    if (use_tma) DMH_LAUNCH(ident_fast_kernel<true>, grid, 256, 0, st)(p, maps);
    else DMH_LAUNCH(ident_fast_kernel<false>, grid, 256, 0, st)(p, maps);
    return DMH_OK;
}

// bf16 frames: identity loss + packed fp32 source + fp32 target in one pass over the bf16 inputs
int launch_ident_bf16(const uint16_t* target, const uint16_t* src, int B, int H, int W, int no_ssim, float* out,
                      float* packed, float* tgt_f32, cudaStream_t st) {
    IdentParams p;
    p.target = nullptr;
    p.packed = reinterpret_cast<float4*>(packed);
    for (int f = 0; f < DMH_PHOTO_MAX_FRAMES; ++f) p.src[f] = nullptr;
    p.out = out; p.B = B; p.F = 1; p.H = H; p.W = W; p.no_ssim = no_ssim;
    dim3 grid(ceil_div(W, FT_T), ceil_div(H, FT_T), B);
    DMH_LAUNCH(ident_bf16_kernel, grid, 256, 0, st)(p, target, src, tgt_f32);
    return DMH_OK;
}

}  // namespace dmh

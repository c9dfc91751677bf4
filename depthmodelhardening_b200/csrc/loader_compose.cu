// Training-batch compositing on the device: the byte work of
// MonoDataset.prep_adv_data + preprocess (DepthNetworks/monodepth2/datasets/mono_dataset.py:186-265, 119-144)
// for a whole collated batch instead of per item inside DataLoader workers.
//
//   dmh_compose_u8   to_tensor(scene) * (1 - m) + obj * m  ->  to_pilimage  (`.mul(255).byte()`: truncation),
//                    optional per-item horizontal flip of the warped patch / mask (mono_dataset.py:222-228)
//   dmh_lanczos_u8   PIL.Image.resize(size, ANTIALIAS) on 8-bit planes -- Pillow's fixed-point Lanczos
//                    (libImaging/Resample.c: 22-bit integer weights, horizontal pass then vertical pass,
//                    each rounded and clipped to 8 bits).  Integer arithmetic: results are bit-exact.
//
// Roofline: HBM by bytes (an 8-bit 1242x375 -> 1024x320 resize moves 1.4 MB + 0.75 MB of intermediate per plane
// triple), in practice bound by the L1 byte gathers of the 9..13 taps; a loader-side step, not on the step's
// critical path (the CPU reference spends ~20 ms per frame on it).
#include "../../include/dmh_b200.h"
#include "dmh_common.cuh"
#include "jitter_math.cuh"

using namespace dmh;

namespace {

#define LZ_PRECISION_BITS 22
#define LZ_UNR 4                   // taps whose loads are in flight together

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= LZ_PRECISION_BITS;                               // arithmetic shift, as Resample.c clip8
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// to_tensor of four pixels: byte / 255 with IEEE division (mono_dataset.py:137-144), fused into the last resize pass
__device__ __forceinline__ float4 bytes_to_float4(unsigned r) {
    return make_float4(div_rn((float)(r & 0xffu), 255.0f), div_rn((float)((r >> 8) & 0xffu), 255.0f),
                       div_rn((float)((r >> 16) & 0xffu), 255.0f), div_rn((float)(r >> 24), 255.0f));
}

// ---- horizontal pass: rows are independent (planes * in_h of them); a thread owns one output column of 4 rows
#define LH_TX 64
#define LH_TY 4
#define LH_ROWS 4
__global__ void __launch_bounds__(LH_TX * LH_TY)
lanczos_h_kernel(const uint8_t* __restrict__ in, int rows, int in_w, int out_w, const int* __restrict__ bounds,
                 const int* __restrict__ kk, int ksize, uint8_t* __restrict__ out) {
    const int xo = blockIdx.x * LH_TX + threadIdx.x;
    const int r0 = (blockIdx.y * LH_TY + threadIdx.y) * LH_ROWS;
    if (xo >= out_w || r0 >= rows) return;
    const int2 bd = __ldg(reinterpret_cast<const int2*>(bounds) + xo);
    const int xmin = bd.x, cnt = bd.y;
    const int* __restrict__ k = kk + xo;                  // weights are laid out (ksize, out): lanes read neighbours
    const uint8_t* p[LH_ROWS];
#pragma unroll
    for (int i = 0; i < LH_ROWS; ++i) p[i] = in + (size_t)min(r0 + i, rows - 1) * in_w + xmin;
    int acc[LH_ROWS];
#pragma unroll
    for (int i = 0; i < LH_ROWS; ++i) acc[i] = 1 << (LZ_PRECISION_BITS - 1);
    // taps in groups of LZ_UNR: the group's loads are issued together (the tap count is a run-time value, so the
    // compiler cannot batch them itself); taps past the span are clamped onto its last pixel with weight 0
    for (int j0 = 0; j0 < cnt; j0 += LZ_UNR) {
        int kj[LZ_UNR];
        unsigned char v[LZ_UNR][LH_ROWS];
#pragma unroll
        for (int u = 0; u < LZ_UNR; ++u) {
            const int j = min(j0 + u, cnt - 1);
            kj[u] = (j0 + u < cnt) ? __ldg(k + (size_t)j * out_w) : 0;
#pragma unroll
            for (int i = 0; i < LH_ROWS; ++i) v[u][i] = __ldg(p[i] + j);
        }
#pragma unroll
        for (int u = 0; u < LZ_UNR; ++u)
#pragma unroll
            for (int i = 0; i < LH_ROWS; ++i) acc[i] += (int)v[u][i] * kj[u];
    }
#pragma unroll
    for (int i = 0; i < LH_ROWS; ++i)
        if (r0 + i < rows) out[(size_t)(r0 + i) * out_w + xo] = clip8(acc[i]);
}

// ---- horizontal pass, dp4a form (spans of <= 4 * NW taps: NW = 3 covers down-scales up to 4/3, NW = 4 up to 2).
// The 22-bit weight k splits exactly into three bytes, k = k2 * 65536 + k1 * 256 + k0 (k0, k1 unsigned, k2 signed),
// so sum_j p_j * k_j is three 8-bit dot products: 3 dp4a per 4 taps instead of 4 byte loads + 4 multiply-adds.
// The pixels of a span come as NW + 1 aligned 32-bit words of the flat byte stream (rows are not word-aligned:
// in_w = 1242), shifted onto tap 0 by a funnel shift.  Taps past the span carry weight 0, so the bytes the words
// bring along (the next pixels of the row, or of the next row) do not matter; word indices are clamped to the buffer.
// ncu on the byte-gather form: 79 lane-instructions per output at 77 % issue utilisation -- instruction-issue bound.
#define LD_ROWS 16
template <int NW>
__global__ void __launch_bounds__(LH_TX * LH_TY)
lanczos_h_dp4a_kernel(const uint8_t* __restrict__ in, int rows, int in_w, int out_w, const int* __restrict__ bounds,
                      const int* __restrict__ kk, unsigned nwords, uint8_t* __restrict__ out) {
    // packed weights of the CTA's 64 output columns, built once per CTA: row ty of the block packs word ty, ty + 4, ...
    __shared__ unsigned sw[3 * NW][LH_TX];
    const int xo = blockIdx.x * LH_TX + threadIdx.x;
    const int r0 = (blockIdx.y * LH_TY + threadIdx.y) * LD_ROWS;
    const int xc = min(xo, out_w - 1);
    const int2 bd = __ldg(reinterpret_cast<const int2*>(bounds) + xc);
    const int xmin = bd.x, cnt = bd.y;
    for (int w = threadIdx.y; w < NW; w += LH_TY) {
        unsigned p0 = 0u, p1 = 0u, p2 = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = 4 * w + i;
            const int k = (j < cnt) ? __ldg(kk + j * out_w + xc) : 0;          // (ksize, out) layout
            p0 |= (unsigned)(k & 0xff) << (8 * i);
            p1 |= (unsigned)((k >> 8) & 0xff) << (8 * i);
            p2 |= (unsigned)((k >> 16) & 0xff) << (8 * i);                      // signed byte (arithmetic shift)
        }
        sw[w][threadIdx.x] = p0;
        sw[NW + w][threadIdx.x] = p1;
        sw[2 * NW + w][threadIdx.x] = p2;
    }
    __syncthreads();
    if (xo >= out_w || r0 >= rows) return;
    unsigned w0[NW], w1[NW], w2[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        w0[w] = sw[w][threadIdx.x];
        w1[w] = sw[NW + w][threadIdx.x];
        w2[w] = sw[2 * NW + w][threadIdx.x];
    }
    const unsigned in_off = (unsigned)((uintptr_t)in & 3);
    const unsigned* __restrict__ base = reinterpret_cast<const unsigned*>(in - in_off);
#pragma unroll 2
    for (int i = 0; i < LD_ROWS; ++i) {
        const int row = min(r0 + i, rows - 1);
        const unsigned a = in_off + (unsigned)row * (unsigned)in_w + (unsigned)xmin;   // byte index of tap 0 (< 2^32)
        const unsigned wi = a >> 2, sh = (a & 3u) * 8u;
        unsigned q[NW + 1];
#pragma unroll
        for (int w = 0; w <= NW; ++w) q[w] = __ldg(base + min(wi + w, nwords - 1));
        unsigned d0 = 0u, d1 = 0u;
        int d2 = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const unsigned v = __funnelshift_r(q[w], q[w + 1], sh);
            d0 = __dp4a(v, w0[w], d0);
            d1 = __dp4a(v, w1[w], d1);
            asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d2) : "r"(v), "r"(w2[w]), "r"(d2));
        }
        const unsigned acc = (1u << (LZ_PRECISION_BITS - 1)) + d0 + (d1 << 8) + ((unsigned)d2 << 16);
        if (r0 + i < rows) out[(size_t)(r0 + i) * out_w + xo] = clip8((int)acc);
    }
}

// ---- vertical pass: a thread owns VEC adjacent columns of one output row; the taps of a row are warp-uniform
template <int VEC>
__global__ void __launch_bounds__(128)
lanczos_v_kernel(const uint8_t* __restrict__ in, int in_h, int w, int out_h, const int* __restrict__ bounds,
                 const int* __restrict__ kk, int ksize, uint8_t* __restrict__ out, float* __restrict__ out_f32) {
    const int x = (blockIdx.x * 128 + threadIdx.x) * VEC;
    const int yo = blockIdx.y;
    if (x >= w) return;
    const int ymin = __ldg(bounds + 2 * yo), cnt = __ldg(bounds + 2 * yo + 1);
    const int* __restrict__ k = kk + yo;                  // (ksize, out) layout; warp-uniform here
    const uint8_t* p = in + ((size_t)blockIdx.z * in_h + ymin) * w + x;
    int acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 1 << (LZ_PRECISION_BITS - 1);
    for (int j0 = 0; j0 < cnt; j0 += LZ_UNR) {
        int kj[LZ_UNR];
        unsigned q[LZ_UNR];
#pragma unroll
        for (int u = 0; u < LZ_UNR; ++u) {
            const int j = min(j0 + u, cnt - 1);
            kj[u] = (j0 + u < cnt) ? __ldg(k + (size_t)j * out_h) : 0;
            if (VEC == 4) q[u] = __ldg(reinterpret_cast<const unsigned*>(p + (size_t)j * w));
            else q[u] = __ldg(p + (size_t)j * w);
        }
#pragma unroll
        for (int u = 0; u < LZ_UNR; ++u) {
            if (VEC == 4) {
                acc[0] += (int)(q[u] & 0xffu) * kj[u];
                acc[1] += (int)((q[u] >> 8) & 0xffu) * kj[u];
                acc[2] += (int)((q[u] >> 16) & 0xffu) * kj[u];
                acc[3] += (int)(q[u] >> 24) * kj[u];
            } else {
                acc[0] += (int)q[u] * kj[u];
            }
        }
    }
    const size_t oi = ((size_t)blockIdx.z * out_h + yo) * w + x;
    uint8_t* o = out + oi;
    if (VEC == 4) {
        const unsigned r = (unsigned)clip8(acc[0]) | ((unsigned)clip8(acc[1]) << 8) | ((unsigned)clip8(acc[2]) << 16) |
                           ((unsigned)clip8(acc[3]) << 24);
        *reinterpret_cast<unsigned*>(o) = r;
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + oi) = bytes_to_float4(r);
    } else {
        o[0] = clip8(acc[0]);
        if (out_f32) out_f32[oi] = div_rn((float)o[0], 255.0f);
    }
}

// ---- vertical pass, row-group form: a thread owns 4 adjacent columns (one 32-bit load per input row) of LV_G
// consecutive output rows, walks the union of their input spans once and feeds every output whose span holds the
// row: each input row is fetched ~LV_G/1.2 times less often than with one output row per thread (the spans of
// neighbouring outputs overlap by 8 of 9 taps).  Span membership and weights are warp-uniform.
#define LV_G 8
#define LV_TS 32                   // rows of the weight table: the union of LV_G spans must fit (checked by the launcher)
__global__ void __launch_bounds__(128)
lanczos_v_group_kernel(const uint8_t* __restrict__ in, int in_h, int w, int out_h, const int* __restrict__ bounds,
                       const int* __restrict__ kk, uint8_t* __restrict__ out, float* __restrict__ out_f32) {
    // sk[t][g]: weight with which input row ylo + t enters output row yo0 + g (0 outside its span): the row loop
    // needs no span test and reads the LV_G weights of a row with two 128-bit broadcast loads
    __shared__ __align__(16) int sk[LV_TS][LV_G];
    __shared__ int s_ylo, s_yhi;
    const int yo0 = blockIdx.y * LV_G;
    const int ylo = __ldg(bounds + 2 * yo0);
    for (int e = threadIdx.x; e < LV_TS * LV_G; e += 128) {
        const int t = e / LV_G, g = e % LV_G;
        int v = 0;
        if (yo0 + g < out_h) {
            const int2 bd = __ldg(reinterpret_cast<const int2*>(bounds) + yo0 + g);
            const int j = t - (bd.x - ylo);
            if (j >= 0 && j < bd.y) v = __ldg(kk + j * out_h + yo0 + g);
        }
        sk[t][g] = v;
    }
    if (threadIdx.x == 0) {
        const int last = min(yo0 + LV_G, out_h) - 1;
        s_ylo = ylo;
        s_yhi = __ldg(bounds + 2 * last) + __ldg(bounds + 2 * last + 1);       // spans are monotone in the output row
    }
    __syncthreads();
    const int x = (blockIdx.x * 128 + threadIdx.x) * 4;
    if (x >= w) return;
    const int n = min(s_yhi - s_ylo, LV_TS);
    int acc[LV_G][4];
#pragma unroll
    for (int g = 0; g < LV_G; ++g)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[g][c] = 1 << (LZ_PRECISION_BITS - 1);
    const int wq = w >> 2;                                 // row pitch in 32-bit words
    const unsigned* p = reinterpret_cast<const unsigned*>(in + ((size_t)blockIdx.z * in_h + ylo) * w + x);
    unsigned qa = __ldg(p), qb = (n > 1) ? __ldg(p + wq) : 0u;
    p += 2 * (size_t)wq;
    for (int t = 0; t < n; ++t) {
        const unsigned cur = qa;
        qa = qb;
        if (t + 2 < n) qb = __ldg(p);                                           // two rows in flight
        p += wq;
        const int b0 = cur & 0xffu, b1 = (cur >> 8) & 0xffu, b2 = (cur >> 16) & 0xffu, b3 = cur >> 24;
        const int4 ka = *reinterpret_cast<const int4*>(&sk[t][0]);
        const int4 kb = *reinterpret_cast<const int4*>(&sk[t][4]);
        const int kv[LV_G] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
#pragma unroll
        for (int g = 0; g < LV_G; ++g) {
            acc[g][0] += b0 * kv[g]; acc[g][1] += b1 * kv[g]; acc[g][2] += b2 * kv[g]; acc[g][3] += b3 * kv[g];
        }
    }
    const size_t oi = ((size_t)blockIdx.z * out_h + yo0) * w + x;
#pragma unroll
    for (int g = 0; g < LV_G; ++g) {
        if (yo0 + g < out_h) {
            const unsigned r = (unsigned)clip8(acc[g][0]) | ((unsigned)clip8(acc[g][1]) << 8) |
                               ((unsigned)clip8(acc[g][2]) << 16) | ((unsigned)clip8(acc[g][3]) << 24);
            *reinterpret_cast<unsigned*>(out + oi + (size_t)g * w) = r;
            if (out_f32) *reinterpret_cast<float4*>(out_f32 + oi + (size_t)g * w) = bytes_to_float4(r);   // to_tensor
        }
    }
}

// ---- composite + quantisation.  One thread per pixel, all channels: the mask is read once.
//   scene (B,C,H,W) u8 or NULL; obj (B,C,H,W) f32; mask (B,1,H,W) f32 or NULL; flip (B) or NULL.
//   scene != NULL: v = scene/255 * (1 - m) + obj * m   (each operation rounded as torch's CPU kernels round it)
//   scene == NULL: v = obj                             (the `color_objmask` image: mask.expand(-1,3,-1,-1))
//   out = (uint8) trunc(v * 255)
__global__ void __launch_bounds__(256)
compose_u8_kernel(const uint8_t* __restrict__ scene, const float* __restrict__ obj, const float* __restrict__ mask,
                  const int* __restrict__ flip, int C, int H, int W, uint8_t* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int b = blockIdx.z;
    if (x >= W) return;
    const bool fl = flip && __ldg(flip + b) != 0;
    const int xs = fl ? W - 1 - x : x;                    // torch.flip(obj_imgs_out, [3]) / torch.flip(masks, [3])
    const size_t N = (size_t)H * W;
    const size_t po = (size_t)y * W + x, so = (size_t)y * W + xs;
    float m = 0.f, om = 1.f;
    if (mask) {
        m = __ldg(mask + (size_t)b * N + so);
        om = sub_rn(1.0f, m);
    }
    for (int c = 0; c < C; ++c) {
        const size_t pl = ((size_t)b * C + c) * N;
        const float o = __ldg(obj + pl + so);
        float v = o;
        if (scene) {
            const float s = div_rn((float)__ldg(scene + pl + po), 255.0f);
            v = add_rn(mul_rn(s, om), mul_rn(o, m));
        }
        const float q = mul_rn(v, 255.0f);
        // `.byte()` of an in-range float truncates; out-of-range inputs (never produced by images in [0,1]) saturate
        out[pl + po] = (uint8_t)min(max((int)q, 0), 255);
    }
}

// --------------------------------------------------------------------------- colour jitter on 8-bit frames (next-2)
// transforms.ColorJitter on the PIL frames of an item (mono_dataset.py:297, 344-350, applied per pyramid level in
// preprocess :140-144): the four steps of torchvision's ColorJitter.forward in the item's drawn order fn_idx --
// 0 brightness, 1 contrast, 2 saturation, 3 hue -- each on 8-bit pixels with Pillow's arithmetic (jitter_math.cuh,
// bit-exact against the oracle / Pillow).  Contrast blends against the rounded MEAN GREY LEVEL of the image as it is
// when the step runs, so the work is two passes: pass 1 applies the steps that precede the contrast step and sums the
// grey levels per image (exact: integer atomics), pass 2 applies all steps.  A step with factor 1 (hue shift 0) is an
// identity in Pillow's arithmetic as well; order entries < 0 are skipped (that step is switched off).
struct JitterItem { int order[4]; float f[3]; int hue_shift; };

__device__ __forceinline__ Rgb8 jitter_steps(Rgb8 p, const JitterItem& it, int upto, uint8_t mean_grey) {
    // applies order[0 .. upto-1]
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k >= upto) break;
        const int op = it.order[k];
        if (op == 0) p = jit_brightness(p, it.f[0]);
        else if (op == 1) p = jit_contrast(p, it.f[1], mean_grey);
        else if (op == 2) p = jit_saturation(p, it.f[2]);
        else if (op == 3) p = jit_hue(p, (uint8_t)it.hue_shift);
    }
    return p;
}

__global__ void __launch_bounds__(256)
jitter_grey_sum_kernel(const uint8_t* __restrict__ in, int npix, const int* __restrict__ order,
                       const float* __restrict__ factors, const int* __restrict__ hue_shift,
                       unsigned long long* __restrict__ sums) {
    __shared__ unsigned int red[8];
    const int b = blockIdx.y;
    JitterItem it;
    int cpos = 4;                                           // position of the contrast step (4: none)
#pragma unroll
    for (int k = 0; k < 4; ++k) { it.order[k] = order[b * 4 + k]; if (it.order[k] == 1 && cpos == 4) cpos = k; }
    it.f[0] = factors[b * 3]; it.f[1] = factors[b * 3 + 1]; it.f[2] = factors[b * 3 + 2];
    it.hue_shift = hue_shift[b];
    if (cpos == 4) return;                                  // no contrast step: the mean is never read
    const uint8_t* base = in + (size_t)b * 3 * npix;
    unsigned int acc = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        Rgb8 p = {base[i], base[npix + i], base[2 * npix + i]};
        p = jitter_steps(p, it, cpos, 0);
        acc += jit_grey(p);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(sums + b, (unsigned long long)t);
    }
}

__global__ void __launch_bounds__(256)
jitter_apply_kernel(const uint8_t* __restrict__ in, int npix, const int* __restrict__ order,
                    const float* __restrict__ factors, const int* __restrict__ hue_shift,
                    const unsigned long long* __restrict__ sums, uint8_t* __restrict__ out_u8,
                    float* __restrict__ out_f32) {
    const int b = blockIdx.y;
    JitterItem it;
#pragma unroll
    for (int k = 0; k < 4; ++k) it.order[k] = order[b * 4 + k];
    it.f[0] = factors[b * 3]; it.f[1] = factors[b * 3 + 1]; it.f[2] = factors[b * 3 + 2];
    it.hue_shift = hue_shift[b];
    // ImageStat.Stat(img.convert("L")).mean[0] is sum / count in double; ImageEnhance.Contrast takes int(mean + 0.5)
    const uint8_t mean_grey = (uint8_t)(int)((double)sums[b] / (double)npix + 0.5);
    const size_t off = (size_t)b * 3 * npix;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        Rgb8 p = {in[off + i], in[off + npix + i], in[off + 2 * npix + i]};
        p = jitter_steps(p, it, 4, mean_grey);
        if (out_u8) { out_u8[off + i] = p.r; out_u8[off + npix + i] = p.g; out_u8[off + 2 * npix + i] = p.b; }
        if (out_f32) {                                      // to_tensor: byte / 255 with IEEE division
            out_f32[off + i] = div_rn((float)p.r, 255.0f);
            out_f32[off + npix + i] = div_rn((float)p.g, 255.0f);
            out_f32[off + 2 * npix + i] = div_rn((float)p.b, 255.0f);
        }
    }
}

}  // namespace


extern "C" {

int dmh_compose_u8(const uint8_t* scene, const float* obj, const float* mask, const int* flip, int B, int C, int H,
                   int W, uint8_t* out, dmh_stream_t stream) {
    DMH_REQUIRE(obj && out, "dmh_compose_u8: null pointer");
    DMH_REQUIRE(!scene || mask, "dmh_compose_u8: a scene needs a mask");
    DMH_REQUIRE(B > 0 && B <= 65535 && C > 0 && H > 0 && H <= 65535 && W > 0, "dmh_compose_u8: bad shape");
    dim3 grid(ceil_div(W, 256), H, B);
    DMH_LAUNCH(compose_u8_kernel, grid, 256, 0, (cudaStream_t)stream)(scene, obj, mask, flip, C, H, W, out);
    DMH_CHECK_LAUNCH("dmh_compose_u8");
    return DMH_OK;
}

int dmh_lanczos_u8(const uint8_t* in, int planes, int in_h, int in_w, int out_h, int out_w, const int* bounds_x,
                   const int* kk_x, int ksize_x, const int* bounds_y, const int* kk_y, int ksize_y, uint8_t* tmp,
                   uint8_t* out, float* out_f32, dmh_stream_t stream) {
    DMH_REQUIRE(in && out, "dmh_lanczos_u8: null pointer");
    DMH_REQUIRE(!out_f32 || ((uintptr_t)out_f32 & 15) == 0, "dmh_lanczos_u8: out_f32 must be 16-byte aligned");
    DMH_REQUIRE(planes > 0 && planes <= 65535 && in_h > 0 && in_w > 0 && out_h > 0 && out_h <= 65535 && out_w > 0,
                "dmh_lanczos_u8: bad shape");
    const bool need_h = out_w != in_w, need_v = out_h != in_h;
    DMH_REQUIRE(!need_h || (bounds_x && kk_x && ksize_x > 0), "dmh_lanczos_u8: horizontal coefficients missing");
    DMH_REQUIRE(!need_v || (bounds_y && kk_y && ksize_y > 0), "dmh_lanczos_u8: vertical coefficients missing");
    DMH_REQUIRE(!(need_h && need_v) || tmp, "dmh_lanczos_u8: two passes need the (planes, in_h, out_w) intermediate");
    cudaStream_t st = (cudaStream_t)stream;
    if (!need_h && !need_v) {
        if (cudaMemcpyAsync(out, in, (size_t)planes * in_h * in_w, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            set_error("dmh_lanczos_u8: copy failed");
            return DMH_ERR_CUDA;
        }
        if (out_f32) return dmh_unpack_u8(out, (long long)planes * in_h * in_w, out_f32, stream);
        return DMH_OK;
    }
    const uint8_t* vin = in;
    if (need_h) {
        uint8_t* hout = need_v ? tmp : out;
        const long long rows = (long long)planes * in_h;
        DMH_REQUIRE(rows < (1ll << 31) && ceil_div(rows, LH_TY * LH_ROWS) <= 65535, "dmh_lanczos_u8: too many rows");
        dim3 block(LH_TX, LH_TY);
        const long long bytes = rows * in_w + 3;                       // + the (<= 3) bytes below an unaligned base
        if (ksize_x <= 16 && bytes < (1ll << 32)) {
            const unsigned nwords = (unsigned)(((uintptr_t)in & 3) + rows * in_w + 3) >> 2;
            dim3 grid(ceil_div(out_w, LH_TX), ceil_div(rows, LH_TY * LD_ROWS));
            if (ksize_x <= 12)
                DMH_LAUNCH(lanczos_h_dp4a_kernel<3>, grid, block, 0, st)(in, (int)rows, in_w, out_w, bounds_x, kk_x, nwords, hout);
            else
                DMH_LAUNCH(lanczos_h_dp4a_kernel<4>, grid, block, 0, st)(in, (int)rows, in_w, out_w, bounds_x, kk_x, nwords, hout);
        } else {
            dim3 grid(ceil_div(out_w, LH_TX), ceil_div(rows, LH_TY * LH_ROWS));
            DMH_LAUNCH(lanczos_h_kernel, grid, block, 0, st)(in, (int)rows, in_w, out_w, bounds_x, kk_x, ksize_x, hout);
        }
        DMH_CHECK_LAUNCH("dmh_lanczos_u8 (horizontal)");
        vin = hout;
    }
    if (need_v) {
        const bool vec = (out_w % 4 == 0) && (((uintptr_t)vin | (uintptr_t)out) % 4 == 0);
        // rows an LV_G-group of outputs can span: (LV_G - 1) centre steps + one full span (+ rounding slack)
        const long long group_span = ((long long)(LV_G - 1) * in_h + out_h - 1) / out_h + ksize_y + 1;
        if (vec && group_span <= LV_TS) {
            dim3 grid(ceil_div(out_w, 128 * 4), ceil_div(out_h, LV_G), planes);
            DMH_LAUNCH(lanczos_v_group_kernel, grid, 128, 0, st)(vin, in_h, out_w, out_h, bounds_y, kk_y, out, out_f32);
        } else if (vec) {
            dim3 grid(ceil_div(out_w, 128 * 4), out_h, planes);
            DMH_LAUNCH(lanczos_v_kernel<4>, grid, 128, 0, st)(vin, in_h, out_w, out_h, bounds_y, kk_y, ksize_y, out, out_f32);
        } else {
            dim3 grid(ceil_div(out_w, 128), out_h, planes);
            DMH_LAUNCH(lanczos_v_kernel<1>, grid, 128, 0, st)(vin, in_h, out_w, out_h, bounds_y, kk_y, ksize_y, out, out_f32);
        }
        DMH_CHECK_LAUNCH("dmh_lanczos_u8 (vertical)");
    } else if (out_f32) {
        // the horizontal pass was the last one: to_tensor as its own (HBM-bound) launch
        return dmh_unpack_u8(out, (long long)planes * out_h * out_w, out_f32, stream);
    }
    return DMH_OK;
}

int dmh_color_jitter_u8(const uint8_t* in, int B, int H, int W, const int* order, const float* factors,
                        const int* hue_shift, unsigned long long* sums, uint8_t* out_u8, float* out_f32,
                        dmh_stream_t stream) {
    DMH_REQUIRE(in && order && factors && hue_shift && sums && (out_u8 || out_f32), "dmh_color_jitter_u8: null pointer");
    DMH_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), "dmh_color_jitter_u8: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(unsigned long long) * (size_t)B, st);
    if (e != cudaSuccess) { set_error("dmh_color_jitter_u8: memset failed: %s", cudaGetErrorString(e)); return DMH_ERR_CUDA; }
    const int npix = H * W;
    const int bx = ceil_div(npix, 256 * 4) < 1 ? 1 : (ceil_div(npix, 256 * 4) > 592 ? 592 : ceil_div(npix, 256 * 4));
    dim3 grid(bx, B);
    DMH_LAUNCH(jitter_grey_sum_kernel, grid, 256, 0, st)(in, npix, order, factors, hue_shift, sums);
    DMH_CHECK_LAUNCH("dmh_color_jitter_u8 (grey sums)");
    DMH_LAUNCH(jitter_apply_kernel, grid, 256, 0, st)(in, npix, order, factors, hue_shift, sums, out_u8, out_f32);
    DMH_CHECK_LAUNCH("dmh_color_jitter_u8 (apply)");
    return DMH_OK;
}

}  // extern "C"

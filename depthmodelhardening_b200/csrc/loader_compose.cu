// Training-batch compositing on the device: the byte work of
// MonoDataset.prep_adv_data + preprocess (DepthNetworks/monodepth2/datasets/mono_dataset.py:186-265, 119-144)
// for a whole collated batch instead of per item inside DataLoader workers.
//
//   dmh_compose_u8   to_tensor(scene) * (1 - m) + obj * m  ->  to_pilimage  (`.mul(255).byte()`: truncation),
//                    optional per-item horizontal flip of the warped patch / mask (mono_dataset.py:226-234)
//   dmh_lanczos_u8   PIL.Image.resize(size, ANTIALIAS) on 8-bit planes -- Pillow's fixed-point Lanczos
//                    (libImaging/Resample.c: 22-bit integer weights, horizontal pass then vertical pass,
//                    each rounded and clipped to 8 bits).  Integer arithmetic: results are bit-exact.
//
// Roofline: HBM by bytes (an 8-bit 1242x375 -> 1024x320 resize moves 1.4 MB + 0.75 MB of intermediate per plane
// triple), in practice bound by the L1 byte gathers of the 9..13 taps; a loader-side step, not on the step's
// critical path (the CPU reference spends ~20 ms per frame on it).
#include "../../include/dmh_b200.h"
#include "dmh_common.cuh"

using namespace dmh;

namespace {

#define LZ_PRECISION_BITS 22
#define LZ_UNR 4                   // taps whose loads are in flight together

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= LZ_PRECISION_BITS;                               // arithmetic shift, as Resample.c clip8
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
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

// ---- vertical pass: a thread owns VEC adjacent columns of one output row; the taps of a row are warp-uniform
template <int VEC>
__global__ void __launch_bounds__(128)
lanczos_v_kernel(const uint8_t* __restrict__ in, int in_h, int w, int out_h, const int* __restrict__ bounds,
                 const int* __restrict__ kk, int ksize, uint8_t* __restrict__ out) {
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
    uint8_t* o = out + ((size_t)blockIdx.z * out_h + yo) * w + x;
    if (VEC == 4) {
        const unsigned r = (unsigned)clip8(acc[0]) | ((unsigned)clip8(acc[1]) << 8) | ((unsigned)clip8(acc[2]) << 16) |
                           ((unsigned)clip8(acc[3]) << 24);
        *reinterpret_cast<unsigned*>(o) = r;
    } else {
        o[0] = clip8(acc[0]);
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
                   uint8_t* out, dmh_stream_t stream) {
    DMH_REQUIRE(in && out, "dmh_lanczos_u8: null pointer");
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
        return DMH_OK;
    }
    const uint8_t* vin = in;
    if (need_h) {
        uint8_t* hout = need_v ? tmp : out;
        const long long rows = (long long)planes * in_h;
        DMH_REQUIRE(rows < (1ll << 31) && ceil_div(rows, LH_TY * LH_ROWS) <= 65535, "dmh_lanczos_u8: too many rows");
        dim3 grid(ceil_div(out_w, LH_TX), ceil_div(rows, LH_TY * LH_ROWS)), block(LH_TX, LH_TY);
        DMH_LAUNCH(lanczos_h_kernel, grid, block, 0, st)(in, (int)rows, in_w, out_w, bounds_x, kk_x, ksize_x, hout);
        DMH_CHECK_LAUNCH("dmh_lanczos_u8 (horizontal)");
        vin = hout;
    }
    if (need_v) {
        const bool vec = (out_w % 4 == 0) && (((uintptr_t)vin | (uintptr_t)out) % 4 == 0);
        if (vec) {
            dim3 grid(ceil_div(out_w, 128 * 4), out_h, planes);
            DMH_LAUNCH(lanczos_v_kernel<4>, grid, 128, 0, st)(vin, in_h, out_w, out_h, bounds_y, kk_y, ksize_y, out);
        } else {
            dim3 grid(ceil_div(out_w, 128), out_h, planes);
            DMH_LAUNCH(lanczos_v_kernel<1>, grid, 128, 0, st)(vin, in_h, out_w, out_h, bounds_y, kk_y, ksize_y, out);
        }
        DMH_CHECK_LAUNCH("dmh_lanczos_u8 (vertical)");
    }
    return DMH_OK;
}

}  // extern "C"

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void k(const __grid_constant__ CUtensorMap map, float* out, int x0, int y0, int b, int bytes) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (MODE >= 1) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        if (MODE == 3)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem_u32(smem)), "l"(&map), "r"(smem_u32(&bar)), "r"(x0), "r"(y0) : "memory");
        if (MODE == 2)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(smem_u32(smem)), "l"(&map), "r"(smem_u32(&bar)), "r"(x0), "r"(y0), "r"(0), "r"(b) : "memory");
    }
    __syncthreads();
    if (MODE >= 2) {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        }
        for (int i = threadIdx.x; i < 3 * 36 * 36; i += blockDim.x) out[i] = smem[i];
    }
}
int main(int argc, char** argv) {
    const int RANK = atoi(argv[1]), BX = atoi(argv[2]), X0 = atoi(argv[3]), BY = atoi(argv[4]);
    const int B = 2, H = 64, W = 96;
    std::vector<float> h(B * 3 * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 3 * 36 * 36 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    printf("entry %d %d %p\n", (int)e, (int)q, f);
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap map; memset(&map, 0, sizeof(map));
    const cuuint64_t gdim[4] = {W, H, 3, B};
    const cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * 12};
    const cuuint32_t box[4] = {(cuuint32_t)BX, (cuuint32_t)BY, (cuuint32_t)(RANK == 4 ? 3 : 1), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = ((Fn)f)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, RANK, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    const int bytes = BX * BY * (RANK == 4 ? 3 : 1) * 4;
    if (RANK == 4) k<2><<<1, 256, 3 * 36 * 36 * 4>>>(map, o, X0, X0, 1, bytes);
    else k<3><<<1, 256, 3 * 36 * 36 * 4>>>(map, o, X0, X0, 1, bytes);
    printf("rank %d box %d x %d x0 %d: %s\n", RANK, BX, BY, X0, cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
    std::vector<float> r2(3 * 36 * 36);
    cudaMemcpy(r2.data(), o, r2.size() * 4, cudaMemcpyDeviceToHost);
    // expect element (ch, r, c) = h[((1*3+ch)*H + (r-2))*W + (c-2)] or 0 if OOB
    int bad = 0;
    for (int ch = 0; ch < 3; ++ch) for (int rr = 0; rr < 36; ++rr) for (int c = 0; c < 36; ++c) {
        int y = rr - 2, x = c - 2;
        float ex = (y < 0 || x < 0) ? 0.f : h[((1 * 3 + ch) * H + y) * W + x];
        if (r2[(ch * 36 + rr) * 36 + c] != ex) ++bad;
    }
    printf("bad %d\n", bad);
    return 0;
}

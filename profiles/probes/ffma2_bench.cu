// throughput probe: scalar FFMA vs packed FFMA2 (independent chains), with and without ALU co-issue
#include <cuda_runtime.h>
#include <stdio.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
    float2 r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    int acc = threadIdx.x;
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { r[i].x = fmaf(r[i].x, a, b); r[i].y = fmaf(r[i].y, a, b); }      // 16 FFMA
            if (MODE == 1) { r[i] = __ffma2_rn(r[i], a2, b2); }                                  // 8 FFMA2 (same flops)
            if (MODE == 2) { r[i].x = fmaf(r[i].x, a, b); r[i].y = fmaf(r[i].y, a, b); acc = (acc ^ (acc << 1)) + i; }
            if (MODE == 3) { r[i] = __ffma2_rn(r[i], a2, b2); acc = (acc ^ (acc << 1)) + i; }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i].x + r[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
}
template <int MODE> float run(float* d, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 256>>>(d, iters, 1.0001f, 0.0001f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(d, iters, 1.0001f, 0.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 4 * 256 * 4);
    const int iters = 20000;
    float t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters), t3 = run<3>(d, iters);
    const double fl = 148.0 * 4 * 256 * iters * 16 * 2;
    printf("FFMA  : %.3f ms  %.1f TFLOP/s\n", t0, fl / t0 / 1e9);
    printf("FFMA2 : %.3f ms  %.1f TFLOP/s\n", t1, fl / t1 / 1e9);
    printf("FFMA  + int: %.3f ms\nFFMA2 + int: %.3f ms\n", t2, t3);
    return 0;
}

// Library-level entry points: thread-local error string, version, build arch.
#include <stdarg.h>

#include <atomic>
#include <map>
#include <mutex>
#include <utility>

#include "../../include/dmh_b200.h"
#include "dmh_common.cuh"

namespace dmh {
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
static std::atomic<long long> g_launches{0};

// ---- exhaustive check of the 3-instruction division by a launch constant (dmh_math.cuh div_const)
__device__ unsigned g_const_div_bad;
__global__ void const_div_check_kernel(float c, float rc) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;          // 2^24 significands: binades [1,2) and [2,4)
    const float a = __uint_as_float(0x3f800000u + i);
    const float q = div_const(a, c, rc), ref = __fdiv_rn(a, c);
    const float qn = div_const(-a, c, rc), refn = __fdiv_rn(-a, c);
    if (__float_as_uint(q) != __float_as_uint(ref) || __float_as_uint(qn) != __float_as_uint(refn))
        atomicAdd(&g_const_div_bad, 1u);
}

bool const_div_exact(int c, float* rc_out) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, bool> cache;               // (device, c) -> verified
    const float rc = (float)(1.0 / (double)c);
    *rc_out = rc;
    if (c < 1) return false;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({dev, c});
    if (it != cache.end()) return it->second;
    // first use of this constant on this device: one small launch + one synchronous 4-byte read-back
    unsigned bad = 0;
    bool ok = cudaMemcpyToSymbol(g_const_div_bad, &bad, sizeof(bad)) == cudaSuccess;
    if (ok) {
        const_div_check_kernel<<<(1u << 24) / 256, 256>>>((float)c, rc);
        bad = 1;
        ok = cudaMemcpyFromSymbol(&bad, g_const_div_bad, sizeof(bad)) == cudaSuccess && bad == 0;
    }
    cudaGetLastError();
    cache[{dev, c}] = ok;
    return ok;
}
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace dmh

extern "C" {
long long dmh_launch_count(void) { return dmh::g_launches.load(std::memory_order_relaxed); }
const char* dmh_last_error(void) { return dmh::g_error; }
int dmh_version(void) { return 1; }
int dmh_build_arch(void) { return 100; }
int dmh_const_div_exact(int c) {
    float rc;
    return dmh::const_div_exact(c, &rc) ? 1 : 0;
}
}

// Library-level entry points: thread-local error string, version, build arch.
#include <stdarg.h>

#include <atomic>

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
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace dmh

extern "C" {
long long dmh_launch_count(void) { return dmh::g_launches.load(std::memory_order_relaxed); }
const char* dmh_last_error(void) { return dmh::g_error; }
int dmh_version(void) { return 1; }
int dmh_build_arch(void) { return 100; }
}

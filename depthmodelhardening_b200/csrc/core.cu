// Library-level entry points: thread-local error string, version, build arch.
#include <stdarg.h>

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
}  // namespace dmh

extern "C" {
const char* dmh_last_error(void) { return dmh::g_error; }
int dmh_version(void) { return 1; }
int dmh_build_arch(void) { return 100; }
}

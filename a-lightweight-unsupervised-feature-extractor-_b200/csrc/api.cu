// Library-wide pieces of the C ABI: error string, version, launch counter.
#include <stdarg.h>

#include "common.cuh"

namespace b200 {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace b200

extern "C" int b200_version(void) { return 100; }
extern "C" const char* b200_last_error(void) { return b200::g_err; }
extern "C" int64_t b200_launch_count(void) { return b200::g_launches.load(); }


// Error reporting, ABI version and launch accounting for libcmrl_b200.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace cmrl {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace cmrl

extern "C" int cmrl_abi_version(void) { return CMRL_B200_ABI_VERSION; }
extern "C" const char* cmrl_last_error(void) { return cmrl::g_err; }
extern "C" long long cmrl_launch_count(void) { return cmrl::g_launches.load(std::memory_order_relaxed); }

// Library-wide runtime bits of libasvgp_sm100a: ABI version and the per-thread error message.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"
#include "../../include/asvgp_b200.h"

namespace asvgp {
static thread_local char g_last_error[512] = {0};

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace asvgp

extern "C" int asvgp_abi_version(void) { return ASVGP_ABI_VERSION; }
extern "C" const char* asvgp_last_error(void) { return asvgp::g_last_error; }
extern "C" int64_t asvgp_launch_count(void) { return (int64_t)asvgp::g_launches.load(std::memory_order_relaxed); }

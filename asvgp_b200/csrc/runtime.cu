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

// Test aid: every SM's shared memory filled with a NaN bit pattern.  Kernels do not inherit zeroed shared memory — they
// inherit whatever the previous kernel on that SM left there — so a read of a slot the kernel itself never wrote is a latent
// bug that only shows when the residue happens to be a NaN (0 * NaN in a "masked" product: r02, asvgp_accum_2d_binned).
// The GPU test suite poisons before every test so that such reads fail deterministically.
namespace asvgp {
__global__ void __launch_bounds__(1024) poison_smem_kernel(int n_words) {
    extern __shared__ unsigned long long poison_smem[];
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) poison_smem[i] = 0x7ff8badc0ffee000ULL;
    __syncthreads();
    if (poison_smem[(threadIdx.x * 31) % n_words] == 0ULL) __trap();      // keeps the stores alive
}
}  // namespace asvgp

extern "C" int asvgp_debug_poison_smem(void* stream) {
    int dev = 0, sms = 148;
    ASVGP_CUDA_OK(cudaGetDevice(&dev));
    ASVGP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int bytes = 227 * 1024;
    static bool allowed = false;
    if (!allowed) {
        ASVGP_CUDA_OK(cudaFuncSetAttribute(asvgp::poison_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        allowed = true;
    }
    // one CTA takes a whole SM's shared memory, so 4 x SM-count CTAs visit every SM (several times)
    asvgp::poison_smem_kernel<<<4 * sms, 1024, bytes, static_cast<cudaStream_t>(stream)>>>(bytes / 8);
    ASVGP_CUDA_OK(cudaGetLastError());
    return asvgp::kOk;
}

extern "C" int asvgp_abi_version(void) { return ASVGP_ABI_VERSION; }
extern "C" const char* asvgp_last_error(void) { return asvgp::g_last_error; }
extern "C" int64_t asvgp_launch_count(void) { return (int64_t)asvgp::g_launches.load(std::memory_order_relaxed); }

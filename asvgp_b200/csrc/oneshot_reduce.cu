// One-shot all-reduce of the packed accumulator over NVLink peer memory (the ONE collective of the data-parallel
// path: [G band | Kuf_y | sum y^2 | N], 400 KB at M = 1e4; NCCL's ring / tree all_reduce costs 18 / 26 / 40 us at 2 / 4 / 8 ranks,
// almost all of it latency).  Every rank holds its partial sums in a SYMMETRIC buffer (torch.distributed._symmetric_memory:
// the same allocation mapped into every peer's address space); ONE kernel per rank
//   1. signals "my buffer is complete" into every peer's signal pad and waits for every peer's signal (system-scope
//      release / acquire, one 32-bit slot per source rank, a monotonically increasing epoch as the value),
//   2. reads all the peers' buffers with 16-byte loads over NVLink and adds them IN RANK ORDER — every rank computes the
//      bit-identical sum, so the replicated factorisations that follow stay bit-identical,
//   3. writes the result to a local (non-symmetric) buffer.
// The symmetric buffers are double-buffered by the caller (asvgp_b200/dist.py): a rank rewrites buffer s % 2 at step s + 2,
// after it has passed the barrier of step s + 1, which every peer enters only after finishing its reads of step s.
// Waits are bounded: after ~2^27 polls a rank gives up, raises its error slot and fills the output with NaN, so that a
// missing peer can never hang the device.
#include <cuda_runtime.h>

#include <algorithm>

#include "common.cuh"
#include "../../include/asvgp_b200.h"

namespace asvgp {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_peer_f64x2(const double* p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_peer_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

constexpr int kMaxRanks = 16;
constexpr long long kPeerSpinLimit = 1LL << 27;

// slot layout inside a rank's signal pad (32-bit words): [pad_offset + src] = epoch of the last completed buffer of rank src
__global__ void __launch_bounds__(256) oneshot_allreduce_kernel(const uint64_t* __restrict__ buffer_ptrs, const uint64_t* __restrict__ pad_ptrs,
                                                                int rank, int world, int64_t offset_doubles, int64_t n, unsigned epoch,
                                                                int pad_offset, double* __restrict__ out, int* __restrict__ status) {
    __shared__ int s_fail;
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    // 1. barrier: block 0 tells every peer, every block waits for every peer
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        unsigned* peer_pad = reinterpret_cast<unsigned*>(pad_ptrs[threadIdx.x]) + pad_offset;
        st_release_sys(peer_pad + rank, epoch);
    }
    if (threadIdx.x < world) {
        const unsigned* my_pad = reinterpret_cast<const unsigned*>(pad_ptrs[rank]) + pad_offset;
        long long spins = 0;
        // (epochs are compared as a signed difference so that the counter may wrap)
        while ((int)(ld_acquire_sys(my_pad + threadIdx.x) - epoch) < 0) {
            if (++spins > kPeerSpinLimit) { s_fail = 1; break; }
        }
    }
    __syncthreads();
    const bool failed = s_fail != 0;
    if (failed && threadIdx.x == 0 && blockIdx.x == 0) atomicExch(status, 1);
    // 2 + 3. sum in rank order
    const double* src[kMaxRanks];
#pragma unroll
    for (int q = 0; q < kMaxRanks; ++q) src[q] = reinterpret_cast<const double*>(buffer_ptrs[q < world ? q : 0]) + offset_doubles;
    const int64_t n2 = n / 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        // all peers' loads are issued before the first add: one NVLink round trip per element, not one per peer
        double2 v[kMaxRanks];
#pragma unroll
        for (int q = 0; q < kMaxRanks; ++q)
            if (q < world) v[q] = ld_peer_f64x2(src[q] + 2 * i);
        double2 acc = make_double2(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < kMaxRanks; ++q)
            if (q < world) { acc.x += v[q].x; acc.y += v[q].y; }
        if (failed) acc = make_double2(nan(""), nan(""));
        *reinterpret_cast<double2*>(out + 2 * i) = acc;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        double acc = 0.0;
        for (int q = 0; q < world; ++q) acc += ld_peer_f64(src[q] + n - 1);
        out[n - 1] = failed ? nan("") : acc;
    }
}

}  // namespace asvgp

using namespace asvgp;

// buffer_ptrs / pad_ptrs: DEVICE arrays of `world` 64-bit addresses (every rank's symmetric buffer / signal pad as mapped into
// this process: _SymmetricMemory.buffer_ptrs_dev / signal_pad_ptrs_dev).  Adds the n doubles at `offset_doubles` of every
// rank's buffer into out[n] (local memory, 16-byte aligned, as the buffers must be).  `epoch` must increase by one per call
// and be the same on every rank; pad_offset selects a block of `world` 32-bit slots in the signal pads.  status[1] (device
// int, zeroed by the caller once): set to 1 if a peer never arrived.
extern "C" int asvgp_allreduce_oneshot(const void* buffer_ptrs, const void* pad_ptrs, int rank, int world, int64_t offset_doubles,
                                       int64_t n, unsigned epoch, int pad_offset, double* out, int* status, void* stream) {
    ASVGP_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "allreduce_oneshot: rank %d of %d", rank, world);
    ASVGP_REQUIRE(n >= 0 && offset_doubles >= 0 && (offset_doubles & 1) == 0, "allreduce_oneshot: n=%lld offset=%lld", (long long)n, (long long)offset_doubles);
    ASVGP_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0, "allreduce_oneshot: out must be 16-byte aligned");
    if (n == 0) return kOk;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // one element pair per thread for the packed 1-D accumulator (25 001 pairs at M = 1e4: 98 CTAs); every CTA polls the barrier
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n / 2 + 255) / 256, 148 * 2));
    oneshot_allreduce_kernel<<<blocks, 256, 0, st>>>(static_cast<const uint64_t*>(buffer_ptrs), static_cast<const uint64_t*>(pad_ptrs), rank,
                                                     world, offset_doubles, n, epoch, pad_offset, out, status); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

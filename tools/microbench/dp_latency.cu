// fp64 latency / throughput probes for B200 (sm_100a): dependent DFMA chain, dependent rsqrt chain, shared-memory
// round trip + barrier, DFMA throughput with 8 warps.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dp_latency dp_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe(double* out, long long* cyc, double seed) {
    __shared__ double sm[256];
    const int tid = threadIdx.x;
    double a = seed + tid * 1e-9, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    // 1. dependent DFMA chain (one warp active at a time is not enforced: all warps run it; latency per op still shows)
    __syncthreads();
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < 1024; ++i) a = fma(a, b, c);
    t1 = clock64();
    if (tid == 0) cyc[0] = t1 - t0;
    // 2. dependent rsqrt chain
    double r = 1.5 + a * 1e-300;
    __syncthreads();
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < 256; ++i) r = rsqrt(r) + 1.25;
    t1 = clock64();
    if (tid == 0) cyc[1] = t1 - t0;
    // 3. smem store -> barrier -> load round trip
    __syncthreads();
    t0 = clock64();
    double v = r;
    for (int i = 0; i < 256; ++i) {
        sm[tid] = v;
        __syncthreads();
        v = sm[(tid + 1) & 255] + 1.0;
        __syncthreads();
    }
    t1 = clock64();
    if (tid == 0) cyc[2] = t1 - t0;
    // 4. DFMA throughput: 16 independent accumulators per thread
    double acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = v + j;
    __syncthreads();
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 256; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fma(acc[j], b, c);
    }
    __syncthreads();
    t1 = clock64();
    if (tid == 0) cyc[3] = t1 - t0;
    double s = a + r + v;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + tid] = s;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 256 * 8 * 148); cudaMalloc(&cyc, 64);
    for (int threads : {32, 256}) {
        probe<<<1, threads>>>(out, cyc, 1.0);
        probe<<<1, threads>>>(out, cyc, 1.0);
        long long h[4]; cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
        printf("threads=%d: dependent DFMA %.1f clk/op | dependent rsqrt(+add) %.1f clk/op | STS+bar+LDS+bar %.1f clk/iter | "
               "16-way independent DFMA: %.2f clk per warp-instruction per SMSP-warp (%lld clk for %d DFMA/thread)\n",
               threads, h[0] / 1024.0, h[1] / 256.0, h[2] / 256.0, h[3] / (256.0 * 16.0), h[3], 256 * 16);
    }
    return 0;
}

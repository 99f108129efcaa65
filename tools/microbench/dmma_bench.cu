// fp64 tensor-core (DMMA, mma.sync m8n8k4) issue rate on one SM, registers only: is it a faster way to do the 64^3 tile
// products of the tile-DAG kernels than the DFMA outer-product loop (57 % of nominal)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void bench(double* out, long long* cyc, int iters) {
    double c[NACC][2];
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-3, b = 0.5 - threadIdx.x * 1e-3;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i], a, b);
    }
    __syncthreads();
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NACC>
void run(int threads) {
    double* out; long long* dc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&dc, 8);
    const int iters = 2000;
    bench<NACC><<<1, threads>>>(out, dc, iters);
    bench<NACC><<<1, threads>>>(out, dc, iters);
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    const double fma = (double)iters * NACC * 256.0 * (threads / 32);
    printf("threads=%4d independent accumulators=%d: %lld cycles, %.1f FMA/clk/SM (DFMA nominal 64), %.2f clk per DMMA per warp; err=%s\n", threads, NACC, c,
           fma / c, (double)c / (iters * NACC), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(dc);
}

int main() {
    run<1>(32); run<8>(32); run<8>(128); run<8>(256); run<8>(512); run<4>(1024);
    return 0;
}

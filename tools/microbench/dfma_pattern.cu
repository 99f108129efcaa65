// DFMA issue rate for the 4x4 outer-product pattern acc[i][j] += a[i]*b[j] (registers only), in the two loop orders,
// with 8 warps on one SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int ORDER>
__global__ void __launch_bounds__(256, 1) k(double* out, long long* cyc, double s) {
    double acc[4][4], a[4], b[4];
    for (int i = 0; i < 4; ++i) { a[i] = s + i + threadIdx.x; b[i] = s * 0.5 + i; for (int j = 0; j < 4; ++j) acc[i][j] = 0.0; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < 1024; ++it) {
        if (ORDER == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        // perturb operands a little so nothing is hoisted (2 extra DADD per 16 DFMA)
        a[it & 3] += 1e-9; b[(it >> 2) & 3] += 1e-9;
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double r = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r += acc[i][j];
    out[threadIdx.x] = r;
}
int main() {
    double* out; long long* dc; cudaMalloc(&out, 2048); cudaMalloc(&dc, 8);
    long long c;
    k<0><<<1, 256>>>(out, dc, 1.0); k<0><<<1, 256>>>(out, dc, 1.0); cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("i-outer: %.2f clk per 16-DFMA block per SMSP-pair (ideal 64)\n", c / 1024.0);
    k<1><<<1, 256>>>(out, dc, 1.0); k<1><<<1, 256>>>(out, dc, 1.0); cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("j-outer: %.2f clk per 16-DFMA block per SMSP-pair (ideal 64)\n", c / 1024.0);
    return 0;
}

// Ceiling of a 1 : 2 read : write stream (the 1-D predictor's traffic: 8 B in, 16 B out per point): out1 = f(x), out2 = g(x)
// with trivial arithmetic, the predictor's access pattern (CTA-contiguous ranges, 16-byte loads, streaming stores).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench/rw12_bench.cu -o /tmp/rw12 && /tmp/rw12
#include <cstdio>
#include <cuda_runtime.h>

template <int U, bool STREAM>
__global__ void __launch_bounds__(256) rw12(const double2* __restrict__ x, double2* __restrict__ a, double2* __restrict__ b, long long n_pairs) {
    const long long per = (n_pairs + gridDim.x - 1) / gridDim.x;
    const long long beg = blockIdx.x * per, end = beg + per < n_pairs ? beg + per : n_pairs;
    for (long long base = beg + threadIdx.x; base < end; base += 256LL * U) {
        double2 v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) { const long long i = base + j * 256LL; v[j] = i < end ? __ldg(x + i) : make_double2(0, 0); }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const long long i = base + j * 256LL;
            if (i >= end) break;
            const double2 p = make_double2(v[j].x * 1.5 + 1.0, v[j].y * 1.5 + 1.0), q = make_double2(v[j].x * v[j].x, v[j].y * v[j].y);
            if (STREAM) { __stcs(a + i, p); __stcs(b + i, q); } else { a[i] = p; b[i] = q; }
        }
    }
}
__global__ void copy_k(const double2* __restrict__ x, double2* __restrict__ a, long long n_pairs) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n_pairs; i += gridDim.x * 256LL) a[i] = __ldg(x + i);
}

int main() {
    const long long n = 100000000, np = n / 2;
    double2 *x, *a, *b;
    cudaMalloc(&x, n * 8); cudaMalloc(&a, n * 8); cudaMalloc(&b, n * 8);
    cudaMemset(x, 0, n * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch, double bytes) {
        for (int i = 0; i < 3; ++i) launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
        printf("%-40s %.3f ms  %.0f GB/s\n", name, ms, bytes / ms / 1e6);
    };
    run("copy 1:1 (16 B/pt)", [&] { copy_k<<<148 * 16, 256>>>(x, a, np); }, 16.0 * n);
    run("cudaMemcpyAsync d2d (16 B/pt)", [&] { cudaMemcpyAsync(a, x, n * 8, cudaMemcpyDeviceToDevice); }, 16.0 * n);
    run("memset (8 B/pt written)", [&] { cudaMemsetAsync(a, 0, n * 8); }, 8.0 * n);
    for (int g : {148 * 2, 148 * 4, 148 * 8, 148 * 16}) {
        char nm[64];
        snprintf(nm, 64, "rw 1:2 U=4 stream grid=%d", g); run(nm, [&] { rw12<4, true><<<g, 256>>>(x, a, b, np); }, 24.0 * n);
        snprintf(nm, 64, "rw 1:2 U=8 stream grid=%d", g); run(nm, [&] { rw12<8, true><<<g, 256>>>(x, a, b, np); }, 24.0 * n);
        snprintf(nm, 64, "rw 1:2 U=4 plain  grid=%d", g); run(nm, [&] { rw12<4, false><<<g, 256>>>(x, a, b, np); }, 24.0 * n);
    }
    return 0;
}

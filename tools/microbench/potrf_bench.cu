// Standalone timing of potrf_regs (the in-register 64 x 64 Cholesky + inverse of the tile-DAG factorisation).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o potrf_bench potrf_bench.cu
//   (-DFRAG times the tensor-core experiment of potrf_frag_experiment.cuh instead)
#include "../../asvgp_b200/csrc/runtime.cu"
#include "../../asvgp_b200/csrc/tiledag_2d.cu"
#ifdef FRAG
#include "potrf_frag_experiment.cuh"
#endif

__global__ void __launch_bounds__(256, 1) bench(const double* A, double* Lout, double* Vout, long long* cyc, int reps) {
    __shared__ __align__(16) double scratch[32 + 8 * 64 + 16 * 64];
#ifdef FRAG
    __shared__ __align__(16) double tile_scratch[64 * 64];
#endif
    __shared__ int s_bad;
    const int tid = threadIdx.x, tm = (tid & 15) * 4, tn = (tid >> 4) * 4;
    double acc[4][4], V[4][4];
    long long best = 1LL << 62;
    for (int r = 0; r < reps; ++r) {
        asvgp::regs_from_tile(acc, A, tm, tn);
        if (tid == 0) s_bad = -1;
        __syncthreads();
        const long long t0 = clock64();
#ifdef FRAG
        asvgp::potrf_frag(acc, V, tm, tn, scratch, tile_scratch, &s_bad);
#else
        asvgp::potrf_regs(acc, V, tm, tn, scratch, scratch + 32, scratch + 32 + 256, &s_bad, nullptr);
#endif
        __syncthreads();
        const long long t1 = clock64();
        if (t1 - t0 < best) best = t1 - t0;
    }
    if (tid == 0) cyc[0] = best;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            if (tm + i < tn + j) { acc[i][j] = 0.0; V[i][j] = 0.0; }
        }
    asvgp::regs_to_tile(acc, Lout, tm, tn);
    asvgp::regs_to_tile(V, Vout, tm, tn);
}

int main() {
    const int n = 64;
    double h[64 * 64], L[64 * 64], V[64 * 64];
    for (int c = 0; c < n; ++c)
        for (int r = 0; r < n; ++r) h[c * n + r] = (r == c ? 70.0 : 0.0) + 1.0 / (1.0 + (r > c ? r - c : c - r));
    double *dA, *dL, *dV; long long* dc;
    cudaMalloc(&dA, sizeof(h)); cudaMalloc(&dL, sizeof(h)); cudaMalloc(&dV, sizeof(h)); cudaMalloc(&dc, 8);
    cudaMemcpy(dA, h, sizeof(h), cudaMemcpyHostToDevice);
    bench<<<1, 256>>>(dA, dL, dV, dc, 20);
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(L, dL, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(V, dV, sizeof(h), cudaMemcpyDeviceToHost);
    // residuals: L L^T - A and V L - I
    double e1 = 0, e2 = 0;
    for (int r = 0; r < n; ++r)
        for (int cc = 0; cc <= r; ++cc) {
            double s = 0, t = 0;
            for (int k = 0; k < n; ++k) { s += L[k * n + r] * L[k * n + cc]; t += V[k * n + r] * L[cc * n + k]; }
            e1 = fmax(e1, fabs(s - h[cc * n + r])); e2 = fmax(e2, fabs(t - (r == cc)));
        }
    printf("potrf_regs: %lld SM cycles (%.2f us at 1.9 GHz)  max|LL^T-A| = %.2e  max|VL-I| = %.2e  err=%s\n", c, c / 1900.0, e1, e2,
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}

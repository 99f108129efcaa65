// Name-for-name counterparts of the reference's Kronecker helpers, for scripts that call them directly:
//
//   asvgp_khatri_rao_csc  <- kronecker.make_kvs_two_sparse / make_kvs_sparse (asvgp/kronecker.py:7-33): row-wise
//                            Khatri-Rao product of two sparse feature matrices, column by column
//   asvgp_kron_dense      <- the dense Kronecker products of utils.bands_to_kron_cholesky (asvgp/utils.py:45-51)
//
// The models never call these (the accumulate kernels of stream_2d.cu fuse the Khatri-Rao product into the moment
// sums and P is never formed densely); they exist so that `asvgp_b200.kronecker` / `asvgp_b200.utils` export what
// `asvgp.kronecker` / `asvgp.utils` export.
#include <cuda_runtime.h>

#include <algorithm>

#include "common.cuh"
#include "../../include/asvgp_b200.h"

namespace asvgp {

// One warp per column n: out[(ia * mB + ib), n] = A[ia, n] * B[ib, n] for every pair of non-zeros of the column, written
// at out_indptr[n] + p * cB + q (p-th non-zero of A's column, q-th of B's) — rows ascending when both inputs are.
__global__ void __launch_bounds__(256) khatri_rao_csc_kernel(const int64_t* __restrict__ ptrA, const int64_t* __restrict__ idxA,
                                                             const double* __restrict__ valA, const int64_t* __restrict__ ptrB,
                                                             const int64_t* __restrict__ idxB, const double* __restrict__ valB,
                                                             int64_t n_cols, int64_t mB, const int64_t* __restrict__ out_ptr,
                                                             int64_t* __restrict__ out_rows, double* __restrict__ out_vals) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t n = warp; n < n_cols; n += n_warps) {
        const int64_t a0 = ptrA[n], cA = ptrA[n + 1] - a0;
        const int64_t b0 = ptrB[n], cB = ptrB[n + 1] - b0;
        const int64_t o0 = out_ptr[n];
        for (int64_t e = lane; e < cA * cB; e += 32) {
            const int64_t p = e / cB, q = e % cB;
            out_rows[o0 + e] = idxA[a0 + p] * mB + idxB[b0 + q];
            out_vals[o0 + e] = valA[a0 + p] * valB[b0 + q];
        }
    }
}

// out[(i1 * mB + i2) * (mA * mB) + (j1 * mB + j2)] = A[i1, j1] * B[i2, j2], A and B dense row-major
__global__ void __launch_bounds__(256) kron_dense_kernel(const double* __restrict__ A, int mA, const double* __restrict__ B,
                                                         int mB, double* __restrict__ out) {
    const int64_t M = (int64_t)mA * mB, total = M * M;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / M, j = t % M;
        out[t] = A[(i / mB) * mA + (j / mB)] * B[(i % mB) * mB + (j % mB)];
    }
}

// Dense lower Cholesky factor of a small SPD matrix (the per-dimension Kuu factors, m x m row-major), one CTA,
// right-looking, in place in `L` (upper triangle zeroed).  info[0] = 0 or the 1-based index of the first non-positive pivot.
__global__ void __launch_bounds__(1024) cholesky_dense_kernel(const double* __restrict__ A, int m, double* __restrict__ L,
                                                              double* __restrict__ info) {
    __shared__ double s_piv;
    __shared__ int s_bad;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) s_bad = 0;
    for (int64_t e = tid; e < (int64_t)m * m; e += nt) L[e] = (e / m >= e % m) ? A[e] : 0.0;
    __syncthreads();
    for (int j = 0; j < m; ++j) {
        if (tid == 0) {
            const double d = L[(int64_t)j * m + j];
            if (!(d > 0.0) && s_bad == 0) s_bad = j + 1;
            s_piv = sqrt(d);
            L[(int64_t)j * m + j] = s_piv;
        }
        __syncthreads();
        const double inv = 1.0 / s_piv;
        for (int i = j + 1 + tid; i < m; i += nt) L[(int64_t)i * m + j] *= inv;
        __syncthreads();
        const int r = m - j - 1;
        for (int64_t e = tid; e < (int64_t)r * r; e += nt) {
            const int i = j + 1 + (int)(e / r), c = j + 1 + (int)(e % r);
            if (c <= i) L[(int64_t)i * m + c] = fma(-L[(int64_t)i * m + j], L[(int64_t)c * m + j], L[(int64_t)i * m + c]);
        }
        __syncthreads();
    }
    if (tid == 0) info[0] = (double)s_bad;
}

}  // namespace asvgp

using namespace asvgp;

extern "C" int asvgp_cholesky_dense(const double* A, int m, double* L, double* info, void* stream) {
    ASVGP_REQUIRE(m > 0 && m <= 4096, "cholesky_dense: m=%d (this helper is for the small per-dimension factors)", m);
    cholesky_dense_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(A, m, L, info); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

extern "C" int asvgp_khatri_rao_csc(const int64_t* indptr_a, const int64_t* indices_a, const double* data_a,
                                    const int64_t* indptr_b, const int64_t* indices_b, const double* data_b,
                                    int64_t n_cols, int64_t m_b, const int64_t* out_indptr, int64_t* out_indices,
                                    double* out_data, void* stream) {
    ASVGP_REQUIRE(n_cols >= 0 && m_b > 0, "khatri_rao_csc: n_cols=%lld m_b=%lld", (long long)n_cols, (long long)m_b);
    if (n_cols == 0) return kOk;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n_cols + 7) / 8, 148 * 8));
    khatri_rao_csc_kernel<<<blocks, 256, 0, st>>>(indptr_a, indices_a, data_a, indptr_b, indices_b, data_b, n_cols, m_b,
                                                  out_indptr, out_indices, out_data); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

extern "C" int asvgp_kron_dense(const double* A, int m_a, const double* B, int m_b, double* out, void* stream) {
    ASVGP_REQUIRE(m_a > 0 && m_b > 0, "kron_dense: m_a=%d m_b=%d", m_a, m_b);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t total = (int64_t)m_a * m_b * m_a * m_b;
    kron_dense_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 16), 256, 0, st>>>(A, m_a, B, m_b, out); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

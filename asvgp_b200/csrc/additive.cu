// Kernels of the additive model GPR_additive (reference asvgp/gpr.py:139-236): f(x) = sum_d f_d(x_d), Kuu block diagonal,
// Kuf = the per-dimension feature matrices stacked.  KufKfu = Kuf Kuf^T has BANDED diagonal blocks (each dimension's own
// Gram: asvgp_accum_1d) and DENSE off-diagonal blocks (basis functions of two different dimensions meet in every point);
// the reference forms it by a sparse product and factorises P = Kuu + KufKfu / sigma2 densely (gpr.py:171-176, 192-195).
//
//   asvgp_accum_cross       <- off-diagonal blocks of `Kuf @ Kuf.T` (gpr.py:174-175), never materialising Kuf
//   asvgp_additive_put_band / _put_cross  <- the dense KufKfu from the per-dimension bands and the cross blocks (gpr.py:175)
//   asvgp_additive_scale + _put_band      <- `Kuu.to_dense() + KufKfu / sigma2` (gpr.py:192, 221)
//   asvgp_additive_terms    <- the contractions the bound's derivatives need from P^-1 (TF reverse mode in the reference)
//   asvgp_predict_additive  <- GPR_additive.predict_f (gpr.py:212-236)
// The dense factorisation itself is asvgp_dense_factor / asvgp_dense_selinv (ndfront_2d.cu).
#include <cuda_runtime.h>

#include <algorithm>

#include "common.cuh"
#include "../../include/asvgp_b200.h"

namespace asvgp {

__device__ __forceinline__ Mesh load_mesh_a(const double* knots, int n_knots) {
    Mesh m;
    m.knots = knots; m.n_knots = n_knots; m.x0 = __ldg(knots); m.inv_delta = 1.0 / (__ldg(knots + 1) - __ldg(knots));
    return m;
}

// C[ia, ib] += sum_n a_ia(xa_n) b_ib(xb_n): (K+1)^2 fp64 REDs per point (the blocks are small and L2-resident; consecutive
// points of a warp that share both knot intervals are combined first by a segmented warp reduction)
template <int K>
__global__ void __launch_bounds__(256) accum_cross_kernel(const double* __restrict__ xa, const double* __restrict__ xb, int64_t stride,
                                                          int64_t n, const double* __restrict__ ka, int nka, const double* __restrict__ kb,
                                                          int nkb, int mb, double* __restrict__ C) {
    const Mesh ma = load_mesh_a(ka, nka), mbm = load_mesh_a(kb, nkb);
    const int lane = threadIdx.x & 31;
    const int64_t n_round = (n + 31) / 32 * 32;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_round; i += (int64_t)gridDim.x * blockDim.x) {
        const bool valid = i < n;
        double wa[K + 1], wb[K + 1];
        int ia = -1, ib = -1;
        if (valid) {
            const double a = __ldg(xa + i * stride), b = __ldg(xb + i * stride);
            ia = locate_interval(ma, a, [](const double* p) { return __ldg(p); });
            ib = locate_interval(mbm, b, [](const double* p) { return __ldg(p); });
            bspline_pieces<K>((a - __ldg(ka + ia)) * ma.inv_delta, wa);
            bspline_pieces<K>((b - __ldg(kb + ib)) * mbm.inv_delta, wb);
        }
        // runs of lanes with the same (ia, ib): the head lane of each run adds the run's sum
        const long long key = valid ? ((long long)ia << 32) | (unsigned)ib : -1 - lane;
        const long long prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = lane == 0 || prev != key;
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const int run_end = (lane == 31) ? 32 : (__ffs(heads >> (lane + 1)) ? lane + __ffs(heads >> (lane + 1)) : 32);
#pragma unroll
        for (int r = 0; r <= K; ++r)
#pragma unroll
            for (int s = 0; s <= K; ++s) {
                double v = valid ? wa[r] * wb[s] : 0.0;
                // segmented inclusive suffix sum inside the run (runs are short in shuffled data, long in ordered data)
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double t = __shfl_down_sync(0xffffffffu, v, o);
                    if (lane + o < run_end) v += t;
                }
                if (head && valid) atomicAdd(C + (int64_t)(ia + r) * mb + ib + s, v);
            }
    }
}

// dense G (M x M, row-major, full symmetric) from the per-dimension lower bands (diagonal blocks) — one launch per dimension —
// and the cross blocks — one launch per pair
__global__ void __launch_bounds__(256) additive_band_block_kernel(const double* __restrict__ band, int m, int K, int off, int M,
                                                                  double scale, bool add, double* __restrict__ out) {
    const int64_t total = (int64_t)m * (2 * K + 1);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / (2 * K + 1)), d = (int)(t % (2 * K + 1)) - K, j = i + d;
        if (j < 0 || j >= m) continue;
        const int ad = d < 0 ? -d : d, c = d < 0 ? j : i;          // band[ad, min(i, j)]
        const double v = scale * __ldg(band + (int64_t)ad * m + c);
        double* o = out + (int64_t)(off + i) * M + off + j;
        *o = add ? __dadd_rn(*o, v) : v;
    }
}
__global__ void __launch_bounds__(256) additive_cross_block_kernel(const double* __restrict__ C, int ma, int mb, int offa, int offb, int M,
                                                                   double* __restrict__ out) {
    const int64_t total = (int64_t)ma * mb;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / mb), j = (int)(t % mb);
        const double v = __ldg(C + t);
        out[(int64_t)(offa + i) * M + offb + j] = v;
        out[(int64_t)(offb + j) * M + offa + i] = v;
    }
}
// P = G / sigma2 (elementwise, the reference's own rounding: a division)
__global__ void __launch_bounds__(256) additive_scale_kernel(const double* __restrict__ G, int64_t total, double sigma2, double* __restrict__ P) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
        P[t] = __ddiv_rn(__ldg(G + t), sigma2);
}

// For dimension d (block offset off, size m, lower bands Kd, dKd) and the dense S = P^-1, x = P^-1 b:
//   out[0] += sum S_dd .* K_d    out[1] += sum S_dd .* dK_d    out[2] += x_d^T K_d x_d    out[3] += x_d^T dK_d x_d
// (full symmetric sums: diagonal once, off-diagonals twice)
__global__ void __launch_bounds__(256) additive_terms_kernel(const double* __restrict__ S, const double* __restrict__ x, int M, int off, int m,
                                                             int K, const double* __restrict__ Kd, const double* __restrict__ dKd,
                                                             double* __restrict__ out) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int64_t total = (int64_t)m * (K + 1);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(t / m), j = (int)(t % m);
        if (j + d >= m) continue;
        const double w = d == 0 ? 1.0 : 2.0;
        const double s = w * __ldg(S + (int64_t)(off + j + d) * M + off + j);
        const double xx = w * __ldg(x + off + j + d) * __ldg(x + off + j);
        const double k = __ldg(Kd + t), dk = __ldg(dKd + t);
        acc[0] = fma(s, k, acc[0]); acc[1] = fma(s, dk, acc[1]);
        acc[2] = fma(xx, k, acc[2]); acc[3] = fma(xx, dk, acc[3]);
    }
    __shared__ double s_red[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) s_red[i][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double v = 0.0;
        for (int wv = 0; wv < 8; ++wv) v += s_red[threadIdx.x][wv];
        atomicAdd(out + threadIdx.x, v);
    }
}
// out[0] += sum S .* G (dense), out[1] += x^T G x
__global__ void __launch_bounds__(256) dense_terms_kernel(const double* __restrict__ S, const double* __restrict__ G, const double* __restrict__ x,
                                                          int M, double* __restrict__ out) {
    double a0 = 0.0, a1 = 0.0;
    const int64_t total = (int64_t)M * M;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const double g = __ldg(G + t);
        a0 = fma(__ldg(S + t), g, a0);
        a1 = fma(__ldg(x + t / M) * __ldg(x + t % M), g, a1);
    }
    __shared__ double s_red[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = a0; s_red[1][threadIdx.x >> 5] = a1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double v = 0.0;
        for (int wv = 0; wv < 8; ++wv) v += s_red[threadIdx.x][wv];
        atomicAdd(out + threadIdx.x, v);
    }
}

// meta[4 d .. 4 d + 3] = { offset of the dimension's knots in `meshes`, number of knots, offset of its basis functions, m_d }
template <int K>
__global__ void __launch_bounds__(128) predict_additive_kernel(const double* __restrict__ X, int64_t n, int D, const double* __restrict__ meshes,
                                                               const int* __restrict__ meta, int M, const double* __restrict__ alpha,
                                                               const double* __restrict__ Pinv, const double* __restrict__ Sall,
                                                               double prior_var, double* __restrict__ mean, double* __restrict__ var) {
    constexpr int kMaxD = 8;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double w[kMaxD][K + 1];
        int row0[kMaxD];
        double mu = 0.0, v = prior_var;
        for (int d = 0; d < D; ++d) {
            const double* knots = meshes + meta[4 * d];
            const Mesh mesh = load_mesh_a(knots, meta[4 * d + 1]);
            const double xv = __ldg(X + i * D + d);
            const int idx = locate_interval(mesh, xv, [](const double* p) { return __ldg(p); });
            bspline_pieces<K>((xv - __ldg(knots + idx)) * mesh.inv_delta, w[d]);
            const int off = meta[4 * d + 2], m = meta[4 * d + 3];
            row0[d] = off + idx;
            const double* Sd = Sall + (int64_t)(K + 1) * off;            // lower band of K_d^-1, (K+1) x m
            double q = 0.0;
#pragma unroll
            for (int r = 0; r <= K; ++r) {
                mu = fma(w[d][r], __ldg(alpha + row0[d] + r), mu);
#pragma unroll
                for (int s = 0; s <= r; ++s)
                    q = fma((r == s ? 1.0 : 2.0) * w[d][r] * w[d][s], __ldg(Sd + (int64_t)(r - s) * m + idx + s), q);
            }
            v -= q;
        }
        for (int d = 0; d < D; ++d)
            for (int e = 0; e < D; ++e)
#pragma unroll
                for (int r = 0; r <= K; ++r)
#pragma unroll
                    for (int s = 0; s <= K; ++s)
                        v = fma(w[d][r] * w[e][s], __ldg(Pinv + (int64_t)(row0[d] + r) * M + row0[e] + s), v);
        mean[i] = mu;
        var[i] = v;
    }
}

}  // namespace asvgp

using namespace asvgp;

#define ASVGP_DISPATCH_ORDER(order, CALL)                         \
    switch (order) {                                              \
        case 1: { constexpr int K = 1; CALL; } break;             \
        case 2: { constexpr int K = 2; CALL; } break;             \
        case 3: { constexpr int K = 3; CALL; } break;             \
        case 4: { constexpr int K = 4; CALL; } break;             \
        case 5: { constexpr int K = 5; CALL; } break;             \
        case 6: { constexpr int K = 6; CALL; } break;             \
        default:                                                  \
            set_last_error("spline order %d not in 1..6", order); \
            return kBadArgument;                                  \
    }

extern "C" int asvgp_accum_cross(const double* xa, const double* xb, int64_t stride, int64_t n, const double* mesh_a, int n_knots_a,
                                 const double* mesh_b, int n_knots_b, int order, double* C, void* stream) {
    ASVGP_REQUIRE(n >= 0 && stride >= 1 && n_knots_a >= 2 && n_knots_b >= 2, "accum_cross: n=%lld stride=%lld", (long long)n, (long long)stride);
    if (n == 0) return kOk;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int mb = n_knots_b + order - 1;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 148 * 8));
    ASVGP_DISPATCH_ORDER(order, (accum_cross_kernel<K><<<blocks, 256, 0, st>>>(xa, xb, stride, n, mesh_a, n_knots_a, mesh_b, n_knots_b, mb, C))); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

extern "C" int asvgp_additive_put_band(const double* band, int m, int order, int offset, int M, double scale, int add, double* out, void* stream) {
    ASVGP_REQUIRE(m > 0 && order >= 1 && offset >= 0 && offset + m <= M, "additive_put_band: m=%d offset=%d M=%d", m, offset, M);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t total = (int64_t)m * (2 * order + 1);
    additive_band_block_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 8), 256, 0, st>>>(band, m, order, offset, M, scale, add != 0, out); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

extern "C" int asvgp_additive_put_cross(const double* C, int m_a, int m_b, int offset_a, int offset_b, int M, double* out, void* stream) {
    ASVGP_REQUIRE(m_a > 0 && m_b > 0 && offset_a + m_a <= M && offset_b + m_b <= M, "additive_put_cross: offsets %d,%d M=%d", offset_a, offset_b, M);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t total = (int64_t)m_a * m_b;
    additive_cross_block_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 8), 256, 0, st>>>(C, m_a, m_b, offset_a, offset_b, M, out); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

extern "C" int asvgp_additive_scale(const double* G, int M, double sigma2, double* P, void* stream) {
    ASVGP_REQUIRE(M > 0 && sigma2 > 0.0, "additive_scale: M=%d sigma2=%g", M, sigma2);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t total = (int64_t)M * M;
    additive_scale_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 16), 256, 0, st>>>(G, total, sigma2, P); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

extern "C" int asvgp_additive_terms(const double* Pinv, const double* x, int M, int offset, int m, int order, const double* Kd,
                                    const double* dKd, double* out4, void* stream) {
    ASVGP_REQUIRE(m > 0 && offset >= 0 && offset + m <= M, "additive_terms: m=%d offset=%d M=%d", m, offset, M);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ASVGP_CUDA_OK(cudaMemsetAsync(out4, 0, 4 * sizeof(double), st));
    const int64_t total = (int64_t)m * (order + 1);
    additive_terms_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148), 256, 0, st>>>(Pinv, x, M, offset, m, order, Kd, dKd, out4); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

extern "C" int asvgp_dense_terms(const double* Pinv, const double* G, const double* x, int M, double* out2, void* stream) {
    ASVGP_REQUIRE(M > 0, "dense_terms: M=%d", M);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ASVGP_CUDA_OK(cudaMemsetAsync(out2, 0, 2 * sizeof(double), st));
    const int64_t total = (int64_t)M * M;
    dense_terms_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 4), 256, 0, st>>>(Pinv, G, x, M, out2); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

extern "C" int asvgp_predict_additive(const double* Xnew, int64_t n, int D, const double* meshes, const int* meta, int M, int order,
                                      const double* alpha, const double* Pinv, const double* S_all, double prior_var, double* mean,
                                      double* var, void* stream) {
    ASVGP_REQUIRE(n >= 0 && D >= 1 && D <= 8 && M > 0, "predict_additive: n=%lld D=%d (at most 8 dimensions)", (long long)n, D);
    if (n == 0) return kOk;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 127) / 128, 148 * 8));
    ASVGP_DISPATCH_ORDER(order, (predict_additive_kernel<K><<<blocks, 128, 0, st>>>(Xnew, n, D, meshes, meta, M, alpha, Pinv, S_all, prior_var, mean, var))); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

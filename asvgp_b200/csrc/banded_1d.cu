// Latency-bound banded kernels of the 1-D model: Kuu assembly, collapsed ELBO + hyper-parameter gradients,
// posterior weights for the predictor.
//
//   asvgp_kuu_assemble   <- SplineFeatures1D.make_Kuu                    (reference asvgp/inducing_features.py:12-44)
//   asvgp_elbo_grad_1d   <- GPR_1d.elbo + its TF-autodiff gradient       (reference asvgp/gpr.py:49-89, example.py:31-32)
//   asvgp_posterior_1d   <- the factorisations/solves of GPR_1d.predict_f (reference asvgp/gpr.py:96-108)
//
// Design.  The work is O(M k^2) flops — nothing — but a length-M dependency chain of
// sqrt/div/FMA.  Each "chain" (one SPD banded matrix) runs in ONE CTA through the partitioned engine of
// band_engine.cuh: P lanes eliminate P chunks in lock-step, lane 0 eliminates the (P-1)k separator system, the lanes
// sweep back.  Independent chains run in different CTAs of the same launch:
//   ELBO+grad : [Kuu, d/dl] with Takahashi + trace,  [P, d/dl] forward only,  [P, d/dsigma2] forward only
//   posterior : [Kuu] Takahashi,  [P] solve + Takahashi
// Derivatives ride along as Dual<1> tangents; the variance derivative follows analytically from the sigma2 one
// because Kuu is proportional to 1/variance (see elbo_finalize_kernel).
#include <cuda_runtime.h>

#include <algorithm>

#include "band_engine.cuh"
#include "../../include/asvgp_b200.h"

namespace asvgp {

constexpr int kChainThreads = 128;          // >= max chunk count P
constexpr int kMaxTerms = 12;

// ------------------------------------------------------------------------------------------------------------------
// Kuu assembly
// ------------------------------------------------------------------------------------------------------------------
struct KuuTerms {
    int n_terms;
    double coef[kMaxTerms];
    double dcoef[kMaxTerms];
};

__global__ void __launch_bounds__(256) kuu_assemble_kernel(const double* __restrict__ tables, KuuTerms terms,
                                                           int64_t band_elems, double* __restrict__ Kuu,
                                                           double* __restrict__ dKuu) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < band_elems;
         i += (int64_t)gridDim.x * blockDim.x) {
        // The reference's own operation order (inducing_features.py:17-44: every term `coefficient * table` rounded, then
        // summed left to right) instead of a fused multiply-add chain: at l / delta ~ 18 the bound moves by a few 1e-10
        // relative per ulp of Kuu (cond(Kuu) ~ 6e4 against 1 / sigma2), so the assembly should round where the
        // reference's does.
        double v = 0.0, dv = 0.0;
        for (int t = 0; t < terms.n_terms; ++t) {
            const double s = __ldg(tables + (int64_t)t * band_elems + i);
            v = __dadd_rn(v, __dmul_rn(terms.coef[t], s));
            dv = __dadd_rn(dv, __dmul_rn(terms.dcoef[t], s));
        }
        Kuu[i] = v;
        if (dKuu != nullptr) dKuu[i] = dv;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// matrix / right-hand-side functors fed to the engine
// ------------------------------------------------------------------------------------------------------------------
// All accessors are branch-free (clamped index + select) so that the loads of one window row issue back to back
// and can be hoisted two columns ahead of their use.
template <class T> struct BandMat;      // A = Kuu + beta*G  with tangent  use_dK*dKuu + tb*G
template <> struct BandMat<Dual<1>> {
    const double* Kuu; const double* dKuu; const double* G;
    double beta, tb; int use_dK; int M;
    __device__ __forceinline__ Dual<1> operator()(int d, int j) const {
        const bool ok = (j >= 0) & (j + d < M);
        const size_t i = ok ? (size_t)d * M + j : 0;
        const double kv = __ldg(Kuu + i);
        const double dk = use_dK ? __ldg(dKuu + i) : 0.0;
        const double g = (G != nullptr) ? __ldg(G + i) : 0.0;
        Dual<1> r;
        r.v = ok ? fma(beta, g, kv) : 0.0;
        r.d[0] = ok ? fma(tb, g, dk) : 0.0;
        return r;
    }
};
template <> struct BandMat<double> {
    const double* Kuu; const double* dKuu; const double* G;
    double beta, tb; int use_dK; int M;
    __device__ __forceinline__ double operator()(int d, int j) const {
        const bool ok = (j >= 0) & (j + d < M);
        const size_t i = ok ? (size_t)d * M + j : 0;
        const double kv = __ldg(Kuu + i);
        const double g = (G != nullptr) ? __ldg(G + i) : 0.0;
        return ok ? fma(beta, g, kv) : 0.0;
    }
};
template <class T> struct VecRhs {      // use == 0: zero right-hand side (b must still be a readable pointer)
    const double* b; int M; int use;
    __device__ __forceinline__ T operator()(int j) const {
        const bool ok = (use != 0) & (j >= 0) & (j < M);
        const double v = __ldg(b + (ok ? j : 0));
        return make_scalar<T>(ok ? v : 0.0, 0.0);
    }
};

// ------------------------------------------------------------------------------------------------------------------
// lane-interleaved row tables
// ------------------------------------------------------------------------------------------------------------------
// The P lanes of a chain sweep P different chunks in lock-step, so reading the band directly costs 32 separate sectors
// per load instruction (the sweeps were LSU-bound: per-column time GREW with P).  A fully parallel pre-pass therefore
// tabulates, for every chain, what lane p needs at row rho of its chunk —
//     rows[(rho * (K+1) + bidx) * P + p] = A[g0(p) + rho, g0(p) + rho - K + bidx],     rhs[rho * P + p] = b[g0(p) + rho]
// — and the sweeps read those tables with fully coalesced loads.  Every access A(d, col) of the engine maps to
// rho = col + d - g0, bidx = K - d.
template <class T, int K>
struct RowsMat {
    const T* rows; int g0, P, p, n_rho;
    __device__ __forceinline__ T operator()(int d, int col) const {
        const int rho = col + d - g0;
        const bool ok = (rho >= 0) & (rho < n_rho);
        const T v = rows[ok ? ((size_t)rho * (K + 1) + (K - d)) * P + p : (size_t)p];      // branch-free
        return ok ? v : zero_of<T>();
    }
};
template <class T>
struct RowsRhs {
    const T* r; int g0, P, p, n_rho;
    __device__ __forceinline__ T operator()(int j) const {
        const int rho = j - g0;
        const bool ok = (rho >= 0) & (rho < n_rho);
        const T v = r[ok ? (size_t)rho * P + p : (size_t)p];
        return ok ? v : zero_of<T>();
    }
};
__host__ __device__ inline int chain_n_rho(const ChunkLayout& lay, int K) { return lay.max_size() + K + 4; }
template <class T, int K>
__host__ __device__ inline size_t chain_rows_count(const ChunkLayout& lay) {
    return (size_t)chain_n_rho(lay, K) * (K + 2) * lay.P;            // (K+1) window entries + the right-hand side
}

template <class T> struct ChainSpec { BandMat<T> A; VecRhs<T> rhs; };

// Sink of the ELBO's Kuu chain: trace(Kuu^-1 G) = sum band(Kuu^-1) .* band(G) (off-diagonals twice, reference
// gpr.py:60-70) is accumulated entry by entry as the Takahashi recursion produces band(Kuu^-1), against a row table of
// G in the lanes' interleaved layout — the inverse band is never stored and there is no separate trace pass.
template <class T, int K>
struct TraceSink {
    const T* gtab; int g0, P, p, n_rho;
    T acc;
    __device__ __forceinline__ void operator()(int d, int col, const T& v) {
        const int rho = col + d - g0;
        const double g = (d == 0 ? 1.0 : 2.0) * gtab[((size_t)rho * (K + 1) + (K - d)) * P + p].v;
        acc.v = fma(v.v, g, acc.v);
        acc.d[0] = fma(v.d[0], g, acc.d[0]);
    }
};

template <class T, int K, int NCHAINS>
__global__ void __launch_bounds__(256) chain_rows_kernel(ChunkLayout lay, ChainSpec<T> s0, ChainSpec<T> s1, ChainSpec<T> s2, ChainSpec<T> s3,
                                                         T* __restrict__ out) {
    const int n_rho = chain_n_rho(lay, K), P = lay.P;
    const size_t per_chain = chain_rows_count<T, K>(lay);
    const size_t total = per_chain * NCHAINS;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int chain = (int)(t / per_chain);
        const size_t e = t % per_chain;
        const ChainSpec<T>& sp = chain == 0 ? s0 : (chain == 1 ? s1 : (chain == 2 ? s2 : s3));
        const int p = (int)(e % P);
        const size_t q = e / P;
        const int g0 = P > 1 ? lay.start(p) : 0;
        if (q < (size_t)n_rho * (K + 1)) {
            const int bidx = (int)(q % (K + 1)), rho = (int)(q / (K + 1));
            const int row = g0 + rho, col = row - K + bidx;
            out[t] = sp.A(K - bidx, col);
        } else {
            const int rho = (int)(q - (size_t)n_rho * (K + 1));
            out[t] = sp.rhs(g0 + rho);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// workspace carving (host) — one ChainWork per chain out of a caller-provided buffer
// ------------------------------------------------------------------------------------------------------------------
template <class T, int K>
struct ChainPlan {
    static size_t bytes(const ChunkLayout& lay) {
        const size_t n = ColumnStore<T, K, true>::count(lay.max_size(), lay.P) * sizeof(T);
        return (n + 255) & ~(size_t)255;
    }
    static ColumnStore<T, K, true> carve(const ChunkLayout& lay, char* base) {
        return ColumnStore<T, K, true>{reinterpret_cast<T*>(base), lay.max_size(), lay.P};
    }
};

// Shared-memory budget: the separator system, its factor and the per-chunk Schur pieces live in dynamic shared
// memory (they are touched by the single-thread phase, where global-memory latency would be fully exposed).
constexpr size_t kChainSmemLimit = 200 * 1024;

// One CTA = one chain.  `Mat`/`Rhs` feed the matrix; logdet/quad totals go to `tot`.
template <class T, int K, bool STORE, bool SOLVE, bool SELINV, class Mat, class Rhs, class Sink>
__device__ __forceinline__ void run_chain(const ChunkLayout& lay, const ColumnStore<T, K, true>& cols, char* smem,
                                          Mat A, Rhs rhs, T* x_out, Sink& sink, ChainTotals<T, K>* tot,
                                          long long* clk) {
    const int p = threadIdx.x;
    ChainWork<T, K> w;
    w.cols = cols;
    ChainSmall<T, K>::carve(lay.P, smem, w);
    if (p == 0 && clk) clk[0] = clock64();
    if (p < lay.P) chain_phase1<T, K, STORE>(lay, p, A, rhs, w);
    __syncthreads();
    if (p == 0 && clk) clk[1] = clock64();
    // separator system: block cyclic reduction, one lane per separator, two barriers per level (band_engine.cuh)
    const int n = lay.P - 1;
    if (p < n) cr_assemble<T, K>(lay, p, w);
    __syncthreads();
    for (int s = 1; s < n; s *= 2) {
        if (p < n && (p & (2 * s - 1)) == s) cr_eliminate<T, K>(n, s, p, lay.M, w);
        __syncthreads();
        if (p < n && (p & (2 * s - 1)) == 0) cr_update<T, K>(n, s, p, w);
        __syncthreads();
    }
    if (p == 0) {
        if (n > 0) cr_eliminate<T, K>(n, 0, 0, lay.M, w);
        *tot = chain_totals<T, K>(lay, w);
        if ((SOLVE || SELINV) && n > 0) cr_back<T, K, SOLVE, SELINV>(n, 0, 0, w);
    }
    __syncthreads();
    if ((SOLVE || SELINV) && n > 0) {
        for (int s = cr_top_stride(n); s >= 1; s /= 2) {
            if (p < n && (p & (2 * s - 1)) == s) cr_back<T, K, SOLVE, SELINV>(n, s, p, w);
            __syncthreads();
        }
        if (p < n) cr_export<T, K, SOLVE, SELINV>(lay, p, w);
        __syncthreads();
    }
    if (p == 0 && clk) clk[2] = clock64();
    if ((SOLVE || SELINV) && p < lay.P) chain_phase3<T, K, SOLVE, SELINV>(lay, p, w, x_out, sink);
    __syncthreads();
    if (p == 0 && clk) clk[3] = clock64();
}

// ------------------------------------------------------------------------------------------------------------------
// ELBO + gradient
// ------------------------------------------------------------------------------------------------------------------
template <int K>
struct ElboArgs {
    ChunkLayout lay;
    ColumnStore<Dual<1>, K, true> cols[3];
    const double* Kuu; const double* dKuu; const double* G; const double* b;
    double sigma2;
    double* partial;        // [3 chains][16]: logdet, dlogdet, quad, dquad, trace, dtrace, info, -, clocks[4]
    const Dual<1>* rows;    // lane-interleaved row tables of the three chains (chain_rows_kernel)
};

template <int K>
__global__ void __launch_bounds__(kChainThreads) elbo_chains_kernel(ElboArgs<K> a) {
    using T = Dual<1>;
    extern __shared__ __align__(16) char smem[];
    const int chain = blockIdx.x, p = threadIdx.x;
    const ChunkLayout lay = a.lay;
    const int M = lay.M;
    __shared__ ChainTotals<T, K> tot;
    __shared__ double s_red[2][kChainThreads / 32];
    __shared__ long long clk[4];
    double* out = a.partial + chain * 16;

    const int n_rho = chain_n_rho(lay, K), g0 = lay.P > 1 ? lay.start(p < lay.P ? p : 0) : 0, pp = p < lay.P ? p : 0;
    const T* tab = a.rows + (size_t)chain * chain_rows_count<T, K>(lay);
    const RowsMat<T, K> A{tab, g0, lay.P, pp, n_rho};
    const RowsRhs<T> rhs{tab + (size_t)n_rho * (K + 1) * lay.P, g0, lay.P, pp, n_rho};
    BandSink<T> no_sink{nullptr, M};
    if (chain == 0) {
        TraceSink<T, K> sink{a.rows + 3 * chain_rows_count<T, K>(lay), g0, lay.P, pp, n_rho, zero_of<T>()};
        run_chain<T, K, true, false, true>(lay, a.cols[0], smem, A, rhs, static_cast<T*>(nullptr), sink, &tot, clk);
        double tr = sink.acc.v, dtr = sink.acc.d[0];           // this lane's share of trace(Kuu^-1 G) and of its d/dl
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            tr += __shfl_xor_sync(0xffffffffu, tr, o);
            dtr += __shfl_xor_sync(0xffffffffu, dtr, o);
        }
        if ((p & 31) == 0) { s_red[0][p >> 5] = tr; s_red[1][p >> 5] = dtr; }
        __syncthreads();
        if (p == 0) {
            tr = 0.0; dtr = 0.0;
            for (int i = 0; i < kChainThreads / 32; ++i) { tr += s_red[0][i]; dtr += s_red[1][i]; }
            out[4] = tr; out[5] = dtr;
        }
    } else {
        // chain 1: P with tangent d/dl;  chain 2: P with tangent d/dsigma2 (tables built by launch_elbo)
        run_chain<T, K, false, false, false>(lay, a.cols[chain], smem, A, rhs, static_cast<T*>(nullptr), no_sink, &tot, clk);
    }
    if (p == 0) {
        out[0] = tot.logdet.v; out[1] = tot.logdet.d[0];
        out[2] = tot.quad.v;   out[3] = tot.quad.d[0];
        out[6] = (double)tot.info;
        const long long t_end = clock64();
        out[8] = (double)(clk[1] - clk[0]); out[9] = (double)(clk[2] - clk[1]);
        out[10] = (double)(clk[3] - clk[2]); out[11] = (double)(t_end - clk[3]);
    }
}

// Collapsed bound of reference gpr.py:81-87 and its derivatives.  With Q = b^T P^-1 b (so that
// sum(c^2) = Q / sigma2^2, gpr.py:75,85), tr = trace(Kuu^-1 G):
//   ELBO = -N/2 log(2 pi s2) - 1/2 log|P| + 1/2 log|Kuu| - yy/(2 s2) + Q/(2 s2^2) - N v/(2 s2) + tr/(2 s2)
// Kuu = Kt(l)/v  =>  P = (Kt + (v/s2) G)/v, hence d/dv of log|P| and Q follow from d/ds2:
//   dlog|P|/dv = -M/v - (s2/v) dlog|P|/ds2,   dQ/dv = Q/v - (s2/v) dQ/ds2,   dlog|Kuu|/dv = -M/v,   dtr/dv = tr/v.
__global__ void elbo_finalize_kernel(const double* __restrict__ partial, const double* __restrict__ scal, int M,
                                     double variance, double sigma2, double* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double* cK = partial;          // Kuu chain (d/dl)
    const double* cL = partial + 16;     // P chain (d/dl)
    const double* cS = partial + 32;     // P chain (d/dsigma2)
    const double yy = scal[0], N = scal[1];
    const double v = variance, s2 = sigma2;
    const double logdetK = cK[0], dlogdetK_dl = cK[1], tr = cK[4], dtr_dl = cK[5];
    const double logdetP = cL[0], dlogdetP_dl = cL[1], Q = cL[2], dQ_dl = cL[3];
    const double dlogdetP_ds = cS[1], dQ_ds = cS[3];
    const double two_pi = 6.283185307179586476925286766559;
    const double elbo = -0.5 * N * log(two_pi * s2) - 0.5 * logdetP + 0.5 * logdetK - 0.5 * yy / s2
                        + 0.5 * Q / (s2 * s2) - 0.5 * N * v / s2 + 0.5 * tr / s2;
    const double d_l = -0.5 * dlogdetP_dl + 0.5 * dlogdetK_dl + 0.5 * dQ_dl / (s2 * s2) + 0.5 * dtr_dl / s2;
    const double d_s = -0.5 * N / s2 - 0.5 * dlogdetP_ds + 0.5 * yy / (s2 * s2) + 0.5 * dQ_ds / (s2 * s2)
                       - Q / (s2 * s2 * s2) + 0.5 * N * v / (s2 * s2) - 0.5 * tr / (s2 * s2);
    const double dlogdetP_dv = -(double)M / v - (s2 / v) * dlogdetP_ds;
    const double dQ_dv = Q / v - (s2 / v) * dQ_ds;
    const double d_v = -0.5 * dlogdetP_dv - 0.5 * (double)M / v + 0.5 * dQ_dv / (s2 * s2) - 0.5 * N / s2
                       + 0.5 * tr / (v * s2);
    out[0] = elbo; out[1] = d_v; out[2] = d_l; out[3] = d_s;
    out[4] = logdetK; out[5] = logdetP; out[6] = Q; out[7] = tr;
    double info = cK[6];
    if (info == 0.0) info = cL[6];
    if (info == 0.0) info = cS[6];
    out[8] = info;
    // diagnostics: SM cycles of the Kuu chain's phases (chunk sweep, separator system, back sweep, trace)
    out[9] = cK[8]; out[10] = cK[9]; out[11] = cK[10]; out[12] = cK[11];
    out[13] = cL[8]; out[14] = cL[9];
    out[15] = dtr_dl;             // d trace(Kuu^-1 G) / d lengthscale (multi-output bound: trace terms count once)
}

// ------------------------------------------------------------------------------------------------------------------
// posterior weights
// ------------------------------------------------------------------------------------------------------------------
template <int K>
struct PosteriorArgs {
    ChunkLayout lay;
    ColumnStore<double, K, true> cols[2];
    const double* Kuu; const double* G; const double* b;
    double sigma2;
    double* sigK; double* sigP; double* x;       // (K+1) x M, (K+1) x M, M
    double* info;                                // [2]
    const double* rows;                          // lane-interleaved row tables of the two chains
};

template <int K>
__global__ void __launch_bounds__(kChainThreads) posterior_chains_kernel(PosteriorArgs<K> a) {
    using T = double;
    extern __shared__ __align__(16) char smem[];
    const int chain = blockIdx.x, p = threadIdx.x;
    const ChunkLayout lay = a.lay;
    const int M = lay.M;
    __shared__ ChainTotals<T, K> tot;
    (void)M;
    const int n_rho = chain_n_rho(lay, K), g0 = lay.P > 1 ? lay.start(p < lay.P ? p : 0) : 0, pp = p < lay.P ? p : 0;
    const T* tab = a.rows + (size_t)chain * chain_rows_count<T, K>(lay);
    const RowsMat<T, K> A{tab, g0, lay.P, pp, n_rho};
    const RowsRhs<T> rhs{tab + (size_t)n_rho * (K + 1) * lay.P, g0, lay.P, pp, n_rho};
    BandSink<T> sinkK{a.sigK, lay.M}, sinkP{a.sigP, lay.M};
    if (chain == 0) run_chain<T, K, true, false, true>(lay, a.cols[0], smem, A, rhs, static_cast<T*>(nullptr), sinkK, &tot, nullptr);
    else run_chain<T, K, true, true, true>(lay, a.cols[1], smem, A, rhs, a.x, sinkP, &tot, nullptr);
    if (p == 0) a.info[chain] = (double)tot.info;
}

__global__ void __launch_bounds__(256) posterior_combine_kernel(const double* __restrict__ sigK,
                                                                const double* __restrict__ sigP,
                                                                const double* __restrict__ x, double inv_s2, int M,
                                                                int band_elems, double* __restrict__ alpha,
                                                                double* __restrict__ S) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < band_elems; i += gridDim.x * blockDim.x) {
        S[i] = sigP[i] - sigK[i];
        if (i < M) alpha[i] = x[i] * inv_s2;
    }
}


// ------------------------------------------------------------------------------------------------------------------
// band(A^-1) and log|A| of one SPD band matrix with one tangent — used for the per-dimension factors K1, K2 of the
// Kronecker model (log|K1 (x) K2| = m2 log|K1| + m1 log|K2| and trace((K1 (x) K2)^-1 G) only need the bands of the
// factor inverses, reference gpr.py:288-289,307)
// ------------------------------------------------------------------------------------------------------------------
template <int K>
struct BandInvArgs {
    ChunkLayout lay;
    ColumnStore<Dual<1>, K, true> cols;
    const double* A; const double* dA;
    Dual<1>* sig;             // scratch (K+1) x M duals
    double* sig_val; double* sig_tan;
    double* scal;             // [4]: log|A|, d log|A|, info, -
    const Dual<1>* rows;      // lane-interleaved row table
};

template <int K>
__global__ void __launch_bounds__(kChainThreads) band_inverse_kernel(BandInvArgs<K> a) {
    using T = Dual<1>;
    extern __shared__ __align__(16) char smem[];
    const int p = threadIdx.x;
    const ChunkLayout lay = a.lay;
    const int M = lay.M;
    __shared__ ChainTotals<T, K> tot;
    const int n_rho = chain_n_rho(lay, K), g0 = lay.P > 1 ? lay.start(p < lay.P ? p : 0) : 0, pp = p < lay.P ? p : 0;
    const RowsMat<T, K> A{a.rows, g0, lay.P, pp, n_rho};
    const RowsRhs<T> rhs{a.rows + (size_t)n_rho * (K + 1) * lay.P, g0, lay.P, pp, n_rho};
    BandSink<T> sink{a.sig, lay.M};
    run_chain<T, K, true, false, true>(lay, a.cols, smem, A, rhs, static_cast<T*>(nullptr), sink, &tot, nullptr);
    for (int i = p; i < (K + 1) * M; i += kChainThreads) {
        const T s = a.sig[i];
        a.sig_val[i] = s.v;
        a.sig_tan[i] = s.d[0];
    }
    if (p == 0) {
        a.scal[0] = tot.logdet.v; a.scal[1] = tot.logdet.d[0]; a.scal[2] = (double)tot.info; a.scal[3] = 0.0;
    }
}

template <int K>
static int launch_band_inverse(const ChunkLayout& lay, const double* A, const double* dA, double* sig_val,
                               double* sig_tan, double* scal, char* work, cudaStream_t st) {
    BandInvArgs<K> a;
    a.lay = lay;
    char* p = work;
    a.cols = ChainPlan<Dual<1>, K>::carve(lay, p);
    p += ChainPlan<Dual<1>, K>::bytes(lay);
    a.sig = reinterpret_cast<Dual<1>*>(p);
    a.A = A; a.dA = dA; a.sig_val = sig_val; a.sig_tan = sig_tan; a.scal = scal;
    p += ((size_t)(K + 1) * lay.M * sizeof(Dual<1>) + 255) & ~(size_t)255;
    Dual<1>* rows = reinterpret_cast<Dual<1>*>(p);
    a.rows = rows;
    {
        using T = Dual<1>;
        ChainSpec<T> s0{BandMat<T>{A, dA, nullptr, 0.0, 0.0, 1, lay.M}, VecRhs<T>{A, lay.M, 0}};
        const size_t total = chain_rows_count<T, K>(lay);
        chain_rows_kernel<T, K, 1><<<(int)((total + 255) / 256), 256, 0, st>>>(lay, s0, s0, s0, s0, rows); ASVGP_LAUNCHED();
        ASVGP_CUDA_OK(cudaGetLastError());
    }
    ASVGP_CUDA_OK(cudaMemsetAsync(a.sig, 0, (size_t)(K + 1) * lay.M * sizeof(Dual<1>), st));
    const size_t smem = ChainSmall<Dual<1>, K>::bytes(lay.P);
    ASVGP_CUDA_OK(cudaFuncSetAttribute(band_inverse_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    band_inverse_kernel<K><<<1, kChainThreads, smem, st>>>(a); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

template <int K>
static size_t elbo_work_bytes(const ChunkLayout& lay) {
    return 3 * ChainPlan<Dual<1>, K>::bytes(lay) + (((size_t)(K + 1) * lay.M * sizeof(Dual<1>) + 255) & ~(size_t)255)
           + 512 + 4 * chain_rows_count<Dual<1>, K>(lay) * sizeof(Dual<1>) + 256;
}
template <int K>
static size_t posterior_work_bytes(const ChunkLayout& lay) {
    return 2 * ChainPlan<double, K>::bytes(lay) + 3 * ((((size_t)(K + 1) * lay.M * sizeof(double)) + 255) & ~(size_t)255)
           + 256 + (((size_t)lay.M * sizeof(double) + 255) & ~(size_t)255) + 2 * chain_rows_count<double, K>(lay) * sizeof(double) + 256;
}

template <int K>
static int launch_elbo(const ChunkLayout& lay, const double* Kuu, const double* dKuu, const double* acc,
                       double variance, double sigma2, double* out, char* work, cudaStream_t st) {
    ElboArgs<K> a;
    a.lay = lay;
    char* p = work;
    for (int c = 0; c < 3; ++c) { a.cols[c] = ChainPlan<Dual<1>, K>::carve(lay, p); p += ChainPlan<Dual<1>, K>::bytes(lay); }
    a.partial = reinterpret_cast<double*>(p);
    p += 512;
    const int M = lay.M;
    a.Kuu = Kuu; a.dKuu = dKuu; a.G = acc; a.b = acc + (size_t)(K + 1) * M;
    a.sigma2 = sigma2;
    {
        using T = Dual<1>;
        T* rows = reinterpret_cast<T*>(p);
        a.rows = rows;
        const double inv_s2 = 1.0 / sigma2;
        // chain 0: Kuu with tangent d/dl;  chain 1: P = Kuu + G/s2 with d/dl;  chain 2: P with d/dsigma2 = -G/s2^2
        ChainSpec<T> s0{BandMat<T>{Kuu, dKuu, nullptr, 0.0, 0.0, 1, M}, VecRhs<T>{Kuu, M, 0}};
        ChainSpec<T> s1{BandMat<T>{Kuu, dKuu, a.G, inv_s2, 0.0, 1, M}, VecRhs<T>{a.b, M, 1}};
        ChainSpec<T> s2{BandMat<T>{Kuu, dKuu, a.G, inv_s2, -inv_s2 * inv_s2, 0, M}, VecRhs<T>{a.b, M, 1}};
        ChainSpec<T> s3{BandMat<T>{a.G, a.G, nullptr, 0.0, 0.0, 0, M}, VecRhs<T>{Kuu, M, 0}};      // G itself (for the trace)
        const size_t total = 4 * chain_rows_count<T, K>(lay);
        chain_rows_kernel<T, K, 4><<<(int)std::min<size_t>((total + 255) / 256, 148 * 8), 256, 0, st>>>(lay, s0, s1, s2, s3, rows); ASVGP_LAUNCHED();
        ASVGP_CUDA_OK(cudaGetLastError());
    }
    const size_t smem = ChainSmall<Dual<1>, K>::bytes(lay.P);
    ASVGP_CUDA_OK(cudaFuncSetAttribute(elbo_chains_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    elbo_chains_kernel<K><<<3, kChainThreads, smem, st>>>(a); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    elbo_finalize_kernel<<<1, 32, 0, st>>>(a.partial, acc + (size_t)(K + 2) * M, M, variance, sigma2, out); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

template <int K>
static int launch_posterior(const ChunkLayout& lay, const double* Kuu, const double* acc, double sigma2,
                            double* alpha, double* S, double* info, char* work, cudaStream_t st) {
    PosteriorArgs<K> a;
    a.lay = lay;
    char* p = work;
    for (int c = 0; c < 2; ++c) { a.cols[c] = ChainPlan<double, K>::carve(lay, p); p += ChainPlan<double, K>::bytes(lay); }
    const int M = lay.M;
    const size_t band_bytes = (((size_t)(K + 1) * M * sizeof(double)) + 255) & ~(size_t)255;
    a.sigK = reinterpret_cast<double*>(p); p += band_bytes;
    a.sigP = reinterpret_cast<double*>(p); p += band_bytes;
    a.x = reinterpret_cast<double*>(p);
    p += ((size_t)M * sizeof(double) + 255) & ~(size_t)255;
    a.Kuu = Kuu; a.G = acc; a.b = acc + (size_t)(K + 1) * M;
    a.sigma2 = sigma2;
    a.info = info;
    {
        using T = double;
        T* rows = reinterpret_cast<T*>(p);
        a.rows = rows;
        ChainSpec<T> s0{BandMat<T>{Kuu, nullptr, nullptr, 0.0, 0.0, 0, M}, VecRhs<T>{Kuu, M, 0}};
        ChainSpec<T> s1{BandMat<T>{Kuu, nullptr, a.G, 1.0 / sigma2, 0.0, 0, M}, VecRhs<T>{a.b, M, 1}};
        const size_t total = 2 * chain_rows_count<T, K>(lay);
        chain_rows_kernel<T, K, 2><<<(int)std::min<size_t>((total + 255) / 256, 148 * 8), 256, 0, st>>>(lay, s0, s1, s1, s1, rows); ASVGP_LAUNCHED();
        ASVGP_CUDA_OK(cudaGetLastError());
    }
    ASVGP_CUDA_OK(cudaMemsetAsync(a.sigK, 0, 2 * band_bytes, st));
    const size_t smem = ChainSmall<double, K>::bytes(lay.P);
    ASVGP_CUDA_OK(cudaFuncSetAttribute(posterior_chains_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    posterior_chains_kernel<K><<<2, kChainThreads, smem, st>>>(a); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    const int band_elems = (K + 1) * M;
    posterior_combine_kernel<<<(band_elems + 255) / 256, 256, 0, st>>>(a.sigK, a.sigP, a.x, 1.0 / sigma2, M, band_elems,
                                                                      alpha, S); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

template <int K>
static int max_chunks_for_smem() {
    int P = kChainThreads;
    while (P > 1 && ChainSmall<Dual<1>, K>::bytes(P) > kChainSmemLimit) --P;
    return P;
}

static ChunkLayout pick_layout(int M, int K, int chunks) {
    int P = chunks > 0 ? chunks : default_chunks(M, K);
    int cap = kChainThreads;
    switch (K) {
        case 1: cap = max_chunks_for_smem<1>(); break;
        case 2: cap = max_chunks_for_smem<2>(); break;
        case 3: cap = max_chunks_for_smem<3>(); break;
        case 4: cap = max_chunks_for_smem<4>(); break;
        case 5: cap = max_chunks_for_smem<5>(); break;
        case 6: cap = max_chunks_for_smem<6>(); break;
    }
    if (P > cap) P = cap;
    return make_layout(M, K, P);
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

extern "C" int asvgp_kuu_assemble(const double* tables, int n_terms, const double* h_coef, const double* h_dcoef,
                                  int M, int order, double* Kuu, double* dKuu, void* stream) {
    ASVGP_REQUIRE(n_terms >= 1 && n_terms <= kMaxTerms, "kuu_assemble: n_terms=%d not in 1..%d", n_terms, kMaxTerms);
    ASVGP_REQUIRE(M > 0 && order >= 1 && order <= kMaxOrder, "kuu_assemble: M=%d order=%d", M, order);
    KuuTerms t;
    t.n_terms = n_terms;
    for (int i = 0; i < n_terms; ++i) { t.coef[i] = h_coef[i]; t.dcoef[i] = h_dcoef ? h_dcoef[i] : 0.0; }
    const int64_t elems = (int64_t)(order + 1) * M;
    const int blocks = (int)((elems + 255) / 256);
    kuu_assemble_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(tables, t, elems, Kuu, dKuu); ASVGP_LAUNCHED();
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

extern "C" int64_t asvgp_workspace_bytes_1d(int M, int order, int chunks) {
    if (M <= 0 || order < 1 || order > kMaxOrder) return -1;
    const ChunkLayout lay = pick_layout(M, order, chunks);
    size_t e = 0, p = 0;
    switch (order) {
        case 1: e = elbo_work_bytes<1>(lay); p = posterior_work_bytes<1>(lay); break;
        case 2: e = elbo_work_bytes<2>(lay); p = posterior_work_bytes<2>(lay); break;
        case 3: e = elbo_work_bytes<3>(lay); p = posterior_work_bytes<3>(lay); break;
        case 4: e = elbo_work_bytes<4>(lay); p = posterior_work_bytes<4>(lay); break;
        case 5: e = elbo_work_bytes<5>(lay); p = posterior_work_bytes<5>(lay); break;
        case 6: e = elbo_work_bytes<6>(lay); p = posterior_work_bytes<6>(lay); break;
    }
    return (int64_t)(e > p ? e : p);
}

extern "C" int asvgp_elbo_grad_1d(const double* Kuu, const double* dKuu, const double* acc, int M, int order,
                                  double variance, double sigma2, int chunks, double* out, void* work,
                                  int64_t work_bytes, void* stream) {
    ASVGP_REQUIRE(M > 2 * order && order >= 1 && order <= kMaxOrder, "elbo_grad_1d: M=%d order=%d", M, order);
    ASVGP_REQUIRE(variance > 0.0 && sigma2 > 0.0, "elbo_grad_1d: variance=%g sigma2=%g must be positive", variance, sigma2);
    ASVGP_REQUIRE(work_bytes >= asvgp_workspace_bytes_1d(M, order, chunks), "elbo_grad_1d: workspace too small");
    const ChunkLayout lay = pick_layout(M, order, chunks);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ASVGP_DISPATCH_ORDER(order, { if (int rc = launch_elbo<K>(lay, Kuu, dKuu, acc, variance, sigma2, out,
                                                              static_cast<char*>(work), st)) return rc; });
    return kOk;
}

extern "C" int asvgp_posterior_1d(const double* Kuu, const double* acc, int M, int order, double sigma2, int chunks,
                                  double* alpha, double* S_band, double* info, void* work, int64_t work_bytes,
                                  void* stream) {
    ASVGP_REQUIRE(M > 2 * order && order >= 1 && order <= kMaxOrder, "posterior_1d: M=%d order=%d", M, order);
    ASVGP_REQUIRE(sigma2 > 0.0, "posterior_1d: sigma2=%g must be positive", sigma2);
    ASVGP_REQUIRE(work_bytes >= asvgp_workspace_bytes_1d(M, order, chunks), "posterior_1d: workspace too small");
    const ChunkLayout lay = pick_layout(M, order, chunks);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ASVGP_DISPATCH_ORDER(order, { if (int rc = launch_posterior<K>(lay, Kuu, acc, sigma2, alpha, S_band, info,
                                                                   static_cast<char*>(work), st)) return rc; });
    return kOk;
}

extern "C" int asvgp_band_inverse_1d(const double* A, const double* dA, int M, int order, int chunks, double* sig_val,
                                     double* sig_tan, double* scal, void* work, int64_t work_bytes, void* stream) {
    ASVGP_REQUIRE(M > 2 * order && order >= 1 && order <= kMaxOrder, "band_inverse_1d: M=%d order=%d", M, order);
    ASVGP_REQUIRE(work_bytes >= asvgp_workspace_bytes_1d(M, order, chunks), "band_inverse_1d: workspace too small");
    const ChunkLayout lay = pick_layout(M, order, chunks);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ASVGP_DISPATCH_ORDER(order, { if (int rc = launch_band_inverse<K>(lay, A, dA, sig_val, sig_tan, scal,
                                                                      static_cast<char*>(work), st)) return rc; });
    return kOk;
}

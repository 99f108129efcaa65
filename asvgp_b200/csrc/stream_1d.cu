// Streaming (HBM-bound) 1-D kernels: Kuf evaluation, fused Gram/projection accumulation, posterior predictor.
//
//   asvgp_basis_eval_1d  <- SplineBasis.evaluate_basis        (reference asvgp/basis.py:51-80)
//   asvgp_accum_1d       <- GPR_1d.__init__ precompute        (reference asvgp/gpr.py:39-44, utils.py:24-30)
//   asvgp_predict_1d     <- GPR_1d.predict_f                  (reference asvgp/gpr.py:91-136)
//
// accum_1d design.  Kuf is never materialised.  Each CTA streams one contiguous slice of (x, y)
// with 128-bit read-only loads, U=4 loads of x and of y in flight per thread.  Every warp keeps, in registers, the
// (k+1)(k+2)/2 + (k+1) partial sums of w w^T and w*y for ONE knot interval `cur` (warp-uniform).  While the
// points a warp reads stay inside that interval (time-series / raster order: ~N/M consecutive points do) the inner
// loop is pure register FMAs.  When a warp meets a point of another interval it butterfly-reduces its partial sums
// with warp shuffles and issues one fp64 RED per band entry, then switches interval.  Points that belong to neither
// the old nor the new interval (only possible for unsorted input) are added with per-point REDs, so any input order
// gives the right answer; sorted input costs ~(k+1)(k+4)/2 REDs per warp per interval crossing, i.e. nothing.
#include <cuda_runtime.h>

#include <algorithm>

#include "common.cuh"
#include "../../include/asvgp_b200.h"

namespace asvgp {

// x0 and delta = mesh[1]-mesh[0] (reference basis.py:18) are read from the device-resident mesh by the kernel
// itself, so no entry point has to synchronise with the host to learn them.
__device__ __forceinline__ Mesh load_mesh(const double* knots, int n_knots) {
    Mesh m;
    m.knots = knots;
    m.n_knots = n_knots;
    m.x0 = __ldg(knots);
    m.inv_delta = 1.0 / (__ldg(knots + 1) - m.x0);
    return m;
}

__device__ __forceinline__ int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }

struct LdgLoader {
    __device__ __forceinline__ double operator()(const double* p) const { return __ldg(p); }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------------------------
// basis evaluation (API parity with evaluate_basis; materialises the (k+1) n non-zeros)
// ------------------------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256) basis_eval_1d_kernel(const double* __restrict__ x, int64_t n, const double* __restrict__ knots, int n_knots,
                                                            int dx, const double* __restrict__ coef,
                                                            int64_t* __restrict__ idx_out,
                                                            double* __restrict__ vals) {
    __shared__ double s_coef[(K + 1) * (K + 1)];
    if (dx > 0) {
        for (int i = threadIdx.x; i < (K + 1) * (K + 1); i += blockDim.x) s_coef[i] = coef[i];
        __syncthreads();
    }
    const Mesh mesh = load_mesh(knots, n_knots);
    double scale = 1.0;
    for (int i = 0; i < dx; ++i) scale *= mesh.inv_delta;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double xi = x[i];
        const int idx = locate_interval(mesh, xi, LdgLoader());
        const double t = (xi - __ldg(mesh.knots + idx)) * mesh.inv_delta;
        double w[K + 1];
        if (dx == 0) bspline_pieces<K>(t, w);
        else bspline_pieces_coef<K>(t, s_coef, scale, w);
        idx_out[i] = idx;
#pragma unroll
        for (int r = 0; r <= K; ++r) vals[(int64_t)r * n + i] = w[r];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// fused accumulation
// ------------------------------------------------------------------------------------------------------------------
constexpr int kAccumThreads = 256;
constexpr int kAccumUnroll = 8;

template <int K>
struct WarpAccum {
    static constexpr int kPairs = Counts<K>::kPairs;
    static constexpr int kAcc = Counts<K>::kAcc;
    static constexpr int NG = 2 * K;          // Gram moments sum tau^j, j = 1..2K (j = 0 is the point count)
    static constexpr int NY = K + 1;          // projection moments sum y tau^j, j = 0..K
    double mg[NG], my[NY];
    int cnt;          // points added by this lane
    int cur;          // interval the register sums belong to (warp-uniform), -1 = none
    double u;         // mesh[cur]
    double lo, hi;    // x in (lo, hi]  <=>  locate_interval(x) == cur
    bool dirty;       // warp-uniform: the sums hold something

    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < NG; ++i) mg[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NY; ++i) my[i] = 0.0;
        cnt = 0;
        dirty = false;
    }
    __device__ __forceinline__ void set_interval(const Mesh& mesh, int idx) {
        cur = idx;
        u = __ldg(mesh.knots + idx);
        lo = (idx == 0) ? -INFINITY : u;
        hi = (idx == mesh.n_knots - 2) ? INFINITY : __ldg(mesh.knots + idx + 1);
    }
    __device__ __forceinline__ bool inside(double x) const { return x > lo && x <= hi; }

    // centred monomial moments of the point (see MomentCoef in common.cuh): 5K+2 fp64 instructions
    __device__ __forceinline__ void add(const Mesh& mesh, double x, double y) {
        const double tau = fma(x - u, mesh.inv_delta, -0.5);
        double p = tau;
        mg[0] += p;
        my[0] += y;
#pragma unroll
        for (int j = 2; j <= NG; ++j) {
            if (j - 1 <= K) my[j - 1] = fma(y, p, my[j - 1]);        // y tau^(j-1)
            p *= tau;
            mg[j - 1] += p;
        }
        ++cnt;
    }
    // shuffle-reduce the moments over the warp, convert them to band entries and add those to the band: one RED per
    // entry per warp, issued by lane `entry`
    __device__ __forceinline__ void flush(double* __restrict__ G, double* __restrict__ b, int M, int lane) {
        if (dirty) {
            double Mg[NG + 1], My[NY];
            Mg[0] = (double)__reduce_add_sync(0xffffffffu, cnt);
#pragma unroll
            for (int j = 0; j < NG; ++j) Mg[j + 1] = warp_sum(mg[j]);
#pragma unroll
            for (int j = 0; j < NY; ++j) My[j] = warp_sum(my[j]);
            const MomentCoef<K>& mc = g_moment_coef<K>;
#pragma unroll
            for (int e0 = 0; e0 < kAcc; e0 += 32) {               // kAcc = 35 for K = 6: a second round for three entries
                const int e = e0 + lane;
                if (e < kPairs) {
                    double v = 0.0;
#pragma unroll
                    for (int j = 0; j <= NG; ++j) v = fma(mc.cg[e][j], Mg[j], v);
                    const int r = mc.rr[e], s = mc.ss[e];
                    atomicAdd(G + (int64_t)(r - s) * M + cur + s, v);
                } else if (e < kAcc) {
                    const int r = e - kPairs;
                    double v = 0.0;
#pragma unroll
                    for (int j = 0; j < NY; ++j) v = fma(mc.cb[r][j], My[j], v);
                    atomicAdd(b + cur + r, v);
                }
            }
        }
        clear();
    }
};

// one point straight to global memory (unsorted-input slow path)
template <int K>
__device__ __forceinline__ void scatter_point(const Mesh& mesh, int idx, double x, double y,
                                              double* __restrict__ G, double* __restrict__ b, int M) {
    const double t = (x - __ldg(mesh.knots + idx)) * mesh.inv_delta;
    double w[K + 1];
    bspline_pieces<K>(t, w);
#pragma unroll
    for (int r = 0; r <= K; ++r) {
#pragma unroll
        for (int s = 0; s <= r; ++s) atomicAdd(G + (int64_t)(r - s) * M + idx + s, w[r] * w[s]);
        atomicAdd(b + idx + r, w[r] * y);
    }
}

// VEC = 2: x and y are 16-byte aligned and read as double2; VEC = 1: scalar loads (misaligned views).
template <int K, int VEC>
__global__ void __launch_bounds__(kAccumThreads, 2)
accum_1d_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n,
                const double* __restrict__ knots, int n_knots, int M,
                double* __restrict__ G, double* __restrict__ b, double* __restrict__ scal) {
    constexpr int U = kAccumUnroll;
    const Mesh mesh = load_mesh(knots, n_knots);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    constexpr int kWarps = kAccumThreads / 32;

    // slots of VEC points; CTA c owns a contiguous range of whole warp-tiles (32*U slots)
    const int64_t n_slots = (n + VEC - 1) / VEC;
    const int64_t tile = 32 * U;
    const int64_t n_tiles = (n_slots + tile - 1) / tile;
    const int64_t tiles_per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int64_t t_begin = blockIdx.x * tiles_per_cta;
    const int64_t t_end = imin64(t_begin + tiles_per_cta, n_tiles);

    WarpAccum<K> wa;
    wa.clear();
    wa.cur = -1;
    wa.u = 0.0;
    wa.lo = INFINITY;
    wa.hi = -INFINITY;
    double yy = 0.0;

    // within the CTA's range warp w takes the w-th contiguous share, so that (for ordered input) each warp meets as
    // few interval crossings as possible
    const int64_t my_tiles = t_end > t_begin ? t_end - t_begin : 0;
    const int64_t per_warp = (my_tiles + kWarps - 1) / kWarps;
    const int64_t w_begin = t_begin + warp * per_warp;
    const int64_t w_end = imin64(w_begin + per_warp, t_end);

    for (int64_t tl = w_begin; tl < w_end; ++tl) {
        double xs[U][VEC], ys[U][VEC];
        const int64_t slot0 = tl * tile + lane;
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int64_t slot = slot0 + j * 32;
            const int64_t p = slot * VEC;
            if (VEC == 2) {
                if (p + 1 < n) {
                    const double2 xv = __ldg(reinterpret_cast<const double2*>(x + p));
                    const double2 yv = __ldg(reinterpret_cast<const double2*>(y + p));
                    xs[j][0] = xv.x; xs[j][VEC - 1] = xv.y;
                    ys[j][0] = yv.x; ys[j][VEC - 1] = yv.y;
                } else {
                    xs[j][0] = (p < n) ? __ldg(x + p) : 0.0;
                    ys[j][0] = (p < n) ? __ldg(y + p) : 0.0;
                    xs[j][VEC - 1] = xs[j][0];
                    ys[j][VEC - 1] = 0.0;
                }
            } else {
                xs[j][0] = (p < n) ? __ldg(x + p) : 0.0;
                ys[j][0] = (p < n) ? __ldg(y + p) : 0.0;
            }
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int64_t p = (slot0 + j * 32) * VEC;
            bool valid[VEC], in[VEC];
            bool all_in = true;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                valid[e] = (p + e) < n;
                in[e] = wa.inside(xs[j][e]);
                all_in = all_in && (in[e] || !valid[e]);
                if (valid[e]) yy = fma(ys[j][e], ys[j][e], yy);
            }
            if (__all_sync(0xffffffffu, all_in)) {
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                    if (valid[e]) wa.add(mesh, xs[j][e], ys[j][e]);
                wa.dirty = true;
            } else {
                // interval crossing (or unsorted input)
                int where[VEC];
                bool pending = false, added_old = false;
                int cand = -1;
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    where[e] = -1;
                    if (valid[e]) {
                        if (in[e]) { wa.add(mesh, xs[j][e], ys[j][e]); added_old = true; }
                        else {
                            where[e] = locate_interval(mesh, xs[j][e], LdgLoader());
                            pending = true;
                            cand = where[e];
                        }
                    }
                }
                // (cur == -1 before the first point: nothing accumulated yet, nothing to flush)
                wa.dirty = wa.dirty || __any_sync(0xffffffffu, added_old);
                wa.flush(G, b, M, lane);
                const unsigned pend = __ballot_sync(0xffffffffu, pending);   // non-zero here
                const int src = 31 - __clz(pend);
                wa.set_interval(mesh, __shfl_sync(0xffffffffu, cand, src));
                bool added = false;
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    if (where[e] >= 0) {
                        if (where[e] == wa.cur) { wa.add(mesh, xs[j][e], ys[j][e]); added = true; }
                        else scatter_point<K>(mesh, where[e], xs[j][e], ys[j][e], G, b, M);
                    }
                }
                wa.dirty = __any_sync(0xffffffffu, added);
            }
        }
    }
    wa.flush(G, b, M, lane);

    __shared__ double s_yy[kWarps];
    yy = warp_sum(yy);
    if (lane == 0) s_yy[warp] = yy;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) tot += s_yy[w];
        atomicAdd(scal, tot);
        if (blockIdx.x == 0) atomicAdd(scal + 1, (double)n);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// predictor
// ------------------------------------------------------------------------------------------------------------------
// Inside one knot interval the posterior mean is a polynomial of degree K and the variance one of degree 2K in
// tau = t - 1/2 (the same conversion tables as the accumulate, MomentCoef<K>): every lane caches the K+1 + 2K+1 coefficients
// of the interval it is in, so that a point of a time-ordered test set costs two Horner evaluations (3K+2 fp64
// instructions) and no gathers; lanes take consecutive pairs of points, loads and stores are fully coalesced 128-bit
// accesses.  A lane whose points jump between intervals (unordered test sets) evaluates pieces and window entries
// directly instead of refreshing its cache every time.  24 B of traffic per point.
template <int K>
__global__ void __launch_bounds__(256) predict_1d_kernel(const double* __restrict__ xs, int64_t n, const double* __restrict__ knots,
                                                         int n_knots, int M,
                                                         const double* __restrict__ alpha,
                                                         const double* __restrict__ S, double variance,
                                                         double* __restrict__ mean, double* __restrict__ var) {
    constexpr int U = 4;                 // pairs of points in flight per thread (8 measured no faster)
    const Mesh mesh = load_mesh(knots, n_knots);
    const MomentCoef<K>& mc = g_moment_coef<K>;
    const bool vec = ((reinterpret_cast<uintptr_t>(xs) | reinterpret_cast<uintptr_t>(mean) | reinterpret_cast<uintptr_t>(var)) & 15u) == 0;

    double cm[K + 1], cv[2 * K + 1];     // mean / variance polynomials (ascending powers of tau) of the cached interval
    int cur = -1, since_refresh = 1 << 20;
    double u = 0.0, lo = INFINITY, hi = -INFINITY;

    auto eval = [&](double x, double& mu, double& vv) {
        if (x > lo && x <= hi) {
            const double tau = fma(x - u, mesh.inv_delta, -0.5);
            double m = cm[K];
#pragma unroll
            for (int j = K - 1; j >= 0; --j) m = fma(m, tau, cm[j]);
            double v = cv[2 * K];
#pragma unroll
            for (int j = 2 * K - 1; j >= 0; --j) v = fma(v, tau, cv[j]);
            mu = m;
            vv = v;
            ++since_refresh;
            return;
        }
        const int idx = locate_interval(mesh, x, LdgLoader());
        const double ui = __ldg(mesh.knots + idx);
        if (since_refresh >= 8) {
            // the lane has been sitting in one interval: adopt the new one
            cur = idx;
            u = ui;
            lo = (idx == 0) ? -INFINITY : ui;
            hi = (idx == mesh.n_knots - 2) ? INFINITY : __ldg(mesh.knots + idx + 1);
#pragma unroll
            for (int j = 0; j <= K; ++j) cm[j] = 0.0;
#pragma unroll
            for (int j = 0; j <= 2 * K; ++j) cv[j] = 0.0;
#pragma unroll
            for (int r = 0; r <= K; ++r) {
                const double ar = __ldg(alpha + idx + r);
#pragma unroll
                for (int j = 0; j <= K; ++j) cm[j] = fma(mc.cb[r][j], ar, cm[j]);
#pragma unroll
                for (int q = 0; q <= r; ++q) {
                    const double sv = (q == r ? 1.0 : 2.0) * __ldg(S + (int64_t)(r - q) * M + idx + q);
#pragma unroll
                    for (int j = 0; j <= 2 * K; ++j) cv[j] = fma(mc.cg[tri_index(r, q)][j], sv, cv[j]);
                }
            }
            cv[0] += variance;
            since_refresh = 0;
            const double tau = fma(x - u, mesh.inv_delta, -0.5);
            double m = cm[K];
#pragma unroll
            for (int j = K - 1; j >= 0; --j) m = fma(m, tau, cm[j]);
            double v = cv[2 * K];
#pragma unroll
            for (int j = 2 * K - 1; j >= 0; --j) v = fma(v, tau, cv[j]);
            mu = m;
            vv = v;
            return;
        }
        // unordered points: direct evaluation from the pieces and the window entries
        since_refresh = 0;
        const double t = (x - ui) * mesh.inv_delta;
        double w[K + 1];
        bspline_pieces<K>(t, w);
        double m = 0.0, q = 0.0;
#pragma unroll
        for (int r = 0; r <= K; ++r) {
            m = fma(w[r], __ldg(alpha + idx + r), m);
            double row = 0.5 * w[r] * __ldg(S + idx + r);                       // diagonal counted once
#pragma unroll
            for (int sidx = 0; sidx < r; ++sidx) row = fma(w[sidx], __ldg(S + (int64_t)(r - sidx) * M + idx + sidx), row);
            q = fma(w[r], row, q);
        }
        mu = m;
        vv = variance + 2.0 * q;
    };

    if (vec) {
        // every CTA walks its own contiguous range of pairs in tiles of 256 * U, so that for an ordered test set a lane's
        // successive points are 2 * 256 * U apart and stay in one knot interval for many tiles
        const int64_t n_pairs = n >> 1;
        const int64_t per_cta = (n_pairs + gridDim.x - 1) / gridDim.x;
        const int64_t c_begin = blockIdx.x * per_cta, c_end = imin64(c_begin + per_cta, n_pairs);
        const double2* __restrict__ x2 = reinterpret_cast<const double2*>(xs);
        double2* __restrict__ m2 = reinterpret_cast<double2*>(mean);
        double2* __restrict__ v2 = reinterpret_cast<double2*>(var);
        for (int64_t base = c_begin + threadIdx.x; base < c_end; base += (int64_t)blockDim.x * U) {
            double2 xv[U];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int64_t i = base + j * (int64_t)blockDim.x;
                xv[j] = (i < c_end) ? __ldg(x2 + i) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int64_t i = base + j * (int64_t)blockDim.x;
                if (i >= c_end) break;
                double2 mo, vo;
                eval(xv[j].x, mo.x, vo.x);
                eval(xv[j].y, mo.y, vo.y);
                m2[i] = mo;
                v2[i] = vo;
            }
        }
        if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) eval(__ldg(xs + n - 1), mean[n - 1], var[n - 1]);
    } else {
        const int64_t stride = (int64_t)gridDim.x * blockDim.x;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) eval(__ldg(xs + i), mean[i], var[i]);
    }
}

static int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) cached = 148;
    }
    return cached;
}

}  // namespace asvgp

using namespace asvgp;

#define ASVGP_DISPATCH_ORDER(order, CALL)          \
    switch (order) {                               \
        case 1: { constexpr int K = 1; CALL; } break; \
        case 2: { constexpr int K = 2; CALL; } break; \
        case 3: { constexpr int K = 3; CALL; } break; \
        case 4: { constexpr int K = 4; CALL; } break; \
        case 5: { constexpr int K = 5; CALL; } break; \
        case 6: { constexpr int K = 6; CALL; } break; \
        default:                                   \
            set_last_error("spline order %d not in 1..6", order); \
            return kBadArgument;                   \
    }

extern "C" int asvgp_basis_eval_1d(const double* x, int64_t n, const double* mesh, int n_knots, int order, int dx,
                                   const double* coef, int64_t* idx, double* vals, void* stream) {
    ASVGP_REQUIRE(n >= 0 && n_knots >= 2, "basis_eval_1d: n=%lld n_knots=%d", (long long)n, n_knots);
    ASVGP_REQUIRE(dx >= 0 && dx <= 3, "basis_eval_1d: dx=%d not in 0..3 (reference basis.py:61-70)", dx);
    ASVGP_REQUIRE(dx == 0 || coef != nullptr, "basis_eval_1d: dx>0 needs the piece coefficient table");
    if (n == 0) return kOk;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8);
    ASVGP_DISPATCH_ORDER(order, (basis_eval_1d_kernel<K><<<blocks, 256, 0, st>>>(x, n, mesh, n_knots, dx, coef, idx, vals)));
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

extern "C" int asvgp_accum_1d(const double* x, const double* y, int64_t n, const double* mesh, int n_knots, int order,
                              double* acc, void* stream) {
    ASVGP_REQUIRE(n >= 0 && n_knots >= 2, "accum_1d: n=%lld n_knots=%d", (long long)n, n_knots);
    ASVGP_REQUIRE(order >= 1 && order <= kMaxOrder, "accum_1d: spline order %d not in 1..6", order);
    if (n == 0) return kOk;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int M = n_knots + order - 1;
    double* G = acc;
    double* b = acc + (int64_t)(order + 1) * M;
    double* scal = b + M;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0;
    const int64_t per_tile = 32 * kAccumUnroll * (vec ? 2 : 1);
    const int64_t n_tiles = (n + per_tile - 1) / per_tile;
    const int blocks = (int)std::min<int64_t>((n_tiles + 7) / 8, (int64_t)sm_count() * 2);
    if (vec) {
        ASVGP_DISPATCH_ORDER(order, (accum_1d_kernel<K, 2><<<blocks, kAccumThreads, 0, st>>>(x, y, n, mesh, n_knots, M, G, b, scal)));
    } else {
        ASVGP_DISPATCH_ORDER(order, (accum_1d_kernel<K, 1><<<blocks, kAccumThreads, 0, st>>>(x, y, n, mesh, n_knots, M, G, b, scal)));
    }
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

extern "C" int asvgp_predict_1d(const double* xnew, int64_t n, const double* mesh, int n_knots, int order,
                                const double* alpha, const double* S_band, double variance, double* mean, double* var,
                                void* stream) {
    ASVGP_REQUIRE(n >= 0 && n_knots >= 2, "predict_1d: n=%lld n_knots=%d", (long long)n, n_knots);
    if (n == 0) return kOk;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int M = n_knots + order - 1;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 256 * 8 - 1) / (256 * 8), (int64_t)sm_count() * 4));
    ASVGP_DISPATCH_ORDER(order, (predict_1d_kernel<K><<<blocks, 256, 0, st>>>(xnew, n, mesh, n_knots, M, alpha, S_band, variance, mean, var)));
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

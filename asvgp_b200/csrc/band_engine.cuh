// Partitioned banded SPD engine: Cholesky / log-determinant / solve / Takahashi selected inverse of a symmetric
// positive definite matrix of bandwidth K, with the length-M dependency chain cut into P independent chunks.
//
// Stands for the banded_matrices ops the reference calls once per optimiser step (reference gpr.py:56-75):
// cholesky_band, the log-det from its first band row, solve_triang_mat, inverse_from_cholesky_band — and for the
// CHOLMOD factorisations/solves of predict_f (gpr.py:96-108).  Everything the ELBO, its gradients and the
// predictor need (log-dets, b^T P^-1 b, alpha = P^-1 b, band(A^-1)) is invariant to the elimination order, so the
// order used here differs from the reference's left-to-right sweep (SURVEY §7 hard part 1):
//
//   indices:  [ I_0 | S_0 | I_1 | S_1 | ... | S_{P-2} | I_{P-1} ]     S_q = K-wide separators
//
//   phase 1 (one thread per chunk, lock-step):  right-looking elimination of the columns of I_p.  The band rows
//            that reach into S_p are a natural continuation of the band; the coupling to S_{p-1} is carried as K
//            extra "spike" rows.  Leaves the Schur complement pieces D_p, E_p and the reduced right-hand side.
//   phase 2 (one thread):  the reduced system on the separators is itself SPD banded, of size (P-1)K and bandwidth
//            2K-1, and goes through the same serial routine (no spikes).
//   phase 3 (one thread per chunk):  back-substitution and/or the Takahashi recursion inside each chunk, seeded
//            with the separator values from phase 2.
//
// The serial dependency chain is ~M/P + (2K-1)/K*(P-1)K columns instead of M.  Every routine is templated on the
// scalar (double or Dual<1>) so the same sweep yields the hyper-parameter derivative, and is ASVGP_HD so that
// tests/host_harness.cpp can run it on the CPU against dense numpy (test infrastructure; the product only
// launches it from banded_1d.cu).
#pragma once
#include "common.cuh"
#include "dual.cuh"

namespace asvgp {

// ---- index layout ----------------------------------------------------------------------------------------------
struct ChunkLayout {
    int M, K, P;
    int base, rem;     // interior sizes: base+1 for p < rem, base otherwise
    ASVGP_HD int size(int p) const { return base + (p < rem ? 1 : 0); }
    ASVGP_HD int start(int p) const { return p * (base + K) + (p < rem ? p : rem); }   // first interior index
    ASVGP_HD int sep_start(int q) const { return start(q) + size(q); }                  // first index of S_q
    ASVGP_HD int max_size() const { return base + (rem > 0 ? 1 : 0); }
    ASVGP_HD int n_reduced() const { return (P - 1) * K; }
};

inline ChunkLayout make_layout(int M, int K, int P) {
    ChunkLayout L;
    L.M = M; L.K = K;
    // every interior needs at least K+1 columns so that no band entry couples two different separators
    while (P > 1 && (M - (P - 1) * K) / P < K + 1) --P;
    L.P = P;
    const int interior = M - (P - 1) * K;
    L.base = interior / P;
    L.rem = interior % P;
    return L;
}

// Chunk count: with the separator system solved by block cyclic reduction (log2 P levels) the chunk sweeps (M/P
// columns) dominate, so as many lanes as the CTA has; the callers cap it by what fits in shared memory.
inline int default_chunks(int M, int K) {
    // (measured at M = 200, K = 3, the per-dimension factor of the 2-D bench: 112 us with one chunk, 40 us with 12)
    if (M < 16 * (K + 1)) return 1;
    int P = M / (4 * (K + 1));          // keep chunks at least a few bandwidths long
    if (P > 128) P = 128;
    if (P < 1) P = 1;
    return P;
}

// ---- per-column record kept for the backward passes ----------------------------------------------------------------
// Stored "column-step major" so that the P lanes working in lock-step touch consecutive addresses:
//   rec[(field * n_steps + j) * P + p],  fields: 0 = 1/pivot, 1..K = L[j+a, j], [K+1..2K = spike X[rho, j],] last = y_j
template <class T, int K, bool SPIKE = true>
struct ColumnStore {
    T* rec;
    int n_steps, P;
    static constexpr int kFields = SPIKE ? 2 * K + 2 : K + 2;
    static constexpr int kY = kFields - 1;
    ASVGP_HD T& at(int field, int j, int p) const { return rec[((size_t)field * n_steps + j) * P + p]; }
    ASVGP_HD static size_t count(int n_steps, int P) { return (size_t)kFields * n_steps * P; }
};

template <class T, int K>
struct ChunkSchur {          // what phase 1 leaves behind for chunk p
    T logdet, quad;
    T Dend[K][K];            // lower triangle: updated block (S_p, S_p), original entries included
    T rend[K];               // updated right-hand side on S_p
    T E[K][K];               // E[a][rho] = coupling (S_p row a, S_{p-1} col rho)
    T SLL[K][K];             // lower triangle: sum_j X[rho,j] X[rho',j]   (to subtract from block (S_{p-1}, S_{p-1}))
    T rhsL[K];               // sum_j X[rho,j] y_j                          (to subtract from rhs on S_{p-1})
    int info;                // 0 or 1 + global index of the first non-positive pivot
};

// ---- phase 1 / phase 2 forward: right-looking elimination of n columns starting at global column g0 -----------
// A(d, j) -> T : entry A[j+d, j] of the lower band (must return 0 when j+d >= size or j >= size or j < 0)
// rhs(j)  -> T : right-hand side (0 outside)
// SPIKE: carry the K spike rows coupling to the K indices just before g0.
template <class T, int K>
struct WindowRow {           // what enters the active window after eliminating column g: row g+1+K and its rhs
    T a[K + 1];
    T r;
};

template <class T, int K, class MatFn, class RhsFn>
ASVGP_HD void fetch_row(MatFn& A, RhsFn& rhs, int g, WindowRow<T, K>& row) {
#pragma unroll
    for (int b = 0; b <= K; ++b) row.a[b] = A(K - b, g + 1 + b);
    row.r = rhs(g + 1 + K);
}

template <class T, int K, bool SPIKE, bool STORE, class MatFn, class RhsFn, class Store>
ASVGP_HD void eliminate_columns(int g0, int n, int n_steps, MatFn A, RhsFn rhs, const Store& store,
                                int p, ChunkSchur<T, K>& out) {
    T W[K + 1][K + 1];       // lower triangle of the active window, rows/cols g .. g+K
    T r[K + 1];
    T c[SPIKE ? K : 1][K + 1];
#pragma unroll
    for (int a = 0; a <= K; ++a) {
#pragma unroll
        for (int b = 0; b <= a; ++b) W[a][b] = A(a - b, g0 + b);
        r[a] = rhs(g0 + a);
    }
    if (SPIKE) {
#pragma unroll
        for (int rho = 0; rho < K; ++rho)
#pragma unroll
            for (int a = 0; a <= K; ++a)
                c[rho][a] = (a <= rho) ? A(a + K - rho, g0 - K + rho) : zero_of<T>();
    }
    LogAccum<T> logdet;
    logdet.init();
    T quad = zero_of<T>();
    T SLL[K][K], rhsL[K];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        rhsL[a] = zero_of<T>();
#pragma unroll
        for (int b = 0; b < K; ++b) SLL[a][b] = zero_of<T>();
    }
    int info = 0;

    // The rows entering the window are fetched two columns ahead of their use so that their memory latency hides
    // behind the rsqrt/FMA chain of the columns in between (the accessors are branch-free, so the K+2 loads of a
    // row issue back to back).
    WindowRow<T, K> rowA, rowB;
    fetch_row<T, K>(A, rhs, g0, rowA);
    fetch_row<T, K>(A, rhs, g0 + 1, rowB);

    auto step = [&](int j, const WindowRow<T, K>& row) {
        const int g = g0 + j;
        if (!(value_of(W[0][0]) > 0.0) && info == 0) info = g + 1;
        const T ip = rsqrt_of(W[0][0]);
        logdet.add(W[0][0], ip);
        T l[K + 1];
#pragma unroll
        for (int a = 1; a <= K; ++a) l[a] = W[a][0] * ip;
        const T yj = r[0] * ip;
        quad += yj * yj;
        T X[SPIKE ? K : 1];
        if (SPIKE) {
#pragma unroll
            for (int rho = 0; rho < K; ++rho) {
                X[rho] = c[rho][0] * ip;
                rhsL[rho] += X[rho] * yj;
#pragma unroll
                for (int r2 = 0; r2 <= rho; ++r2) SLL[rho][r2] += X[rho] * X[r2];
            }
        }
        if (STORE) {
            store.at(0, j, p) = ip;
#pragma unroll
            for (int a = 1; a <= K; ++a) store.at(a, j, p) = l[a];
            if (SPIKE) {
#pragma unroll
                for (int rho = 0; rho < K; ++rho) store.at(K + 1 + rho, j, p) = X[rho];
            }
            store.at(Store::kY, j, p) = yj;
        }
        // trailing update and window shift
#pragma unroll
        for (int a = 1; a <= K; ++a) {
#pragma unroll
            for (int b = 1; b <= a; ++b) W[a - 1][b - 1] = W[a][b] - l[a] * l[b];
            r[a - 1] = r[a] - l[a] * yj;
            if (SPIKE) {
#pragma unroll
                for (int rho = 0; rho < K; ++rho) c[rho][a - 1] = c[rho][a] - l[a] * X[rho];
            }
        }
#pragma unroll
        for (int b = 0; b <= K; ++b) W[K][b] = row.a[b];
        r[K] = row.r;
        if (SPIKE) {
#pragma unroll
            for (int rho = 0; rho < K; ++rho) c[rho][K] = zero_of<T>();
        }
    };

    for (int j = 0; j < n_steps; j += 2) {
        WindowRow<T, K> nextA, nextB;
        fetch_row<T, K>(A, rhs, g0 + j + 2, nextA);
        if (j < n) step(j, rowA);
        fetch_row<T, K>(A, rhs, g0 + j + 3, nextB);
        if (j + 1 < n) step(j + 1, rowB);
        rowA = nextA;
        rowB = nextB;
    }
    out.logdet = logdet.result();
    out.quad = quad;
    out.info = info;
#pragma unroll
    for (int a = 0; a < K; ++a) {
        out.rend[a] = r[a];
        out.rhsL[a] = rhsL[a];
#pragma unroll
        for (int b = 0; b < K; ++b) {
            out.Dend[a][b] = (b <= a) ? W[a][b] : zero_of<T>();
            out.SLL[a][b] = SLL[a][b];
            out.E[a][b] = SPIKE ? c[b][a] : zero_of<T>();
        }
    }
}

// ---- serial back-substitution L^T x = y on stored columns (used for the reduced system) -----------------------------
template <class T, int K, class Store>
ASVGP_HD void backsolve_serial(int n, const Store& store, T* x) {
    T xw[K + 1];
#pragma unroll
    for (int a = 0; a <= K; ++a) xw[a] = zero_of<T>();
    for (int j = n - 1; j >= 0; --j) {
        T acc = store.at(Store::kY, j, 0);
#pragma unroll
        for (int a = 1; a <= K; ++a) acc -= store.at(a, j, 0) * xw[a];
        const T xj = acc * store.at(0, j, 0);
#pragma unroll
        for (int a = K; a >= 2; --a) xw[a] = xw[a - 1];
        xw[1] = xj;
        x[j] = xj;
    }
}

// ---- serial Takahashi recursion on stored columns: lower band of (L L^T)^-1, sig[d * n + j] ---------------------------
// Where the entries of the inverse band go: sink(d, col, v) receives entry (row col + d, column col).
template <class T>
struct BandSink {                 // lower band (K+1) x M, row-major
    T* sig; int M;
    ASVGP_HD void operator()(int d, int col, const T& v) const { sig[(size_t)d * M + col] = v; }
};

template <class T, int K, class Store, class Sink>
ASVGP_HD void selinv_serial(int n, const Store& store, Sink& sink) {
    T Z[K][K];     // Sigma[g+1+a, g+1+b] of the columns already done (symmetric, full storage)
#pragma unroll
    for (int a = 0; a < K; ++a)
#pragma unroll
        for (int b = 0; b < K; ++b) Z[a][b] = zero_of<T>();
    for (int j = n - 1; j >= 0; --j) {
        const T ip = store.at(0, j, 0);
        T l[K], w[K];
#pragma unroll
        for (int a = 0; a < K; ++a) l[a] = store.at(a + 1, j, 0);
        T dot = zero_of<T>();
#pragma unroll
        for (int a = 0; a < K; ++a) {
            T acc = zero_of<T>();
#pragma unroll
            for (int b = 0; b < K; ++b) acc += Z[a][b] * l[b];
            w[a] = acc;
            dot += l[a] * acc;
        }
        const T ip2 = ip * ip;
        const T sjj = ip2 + dot * ip2;
        sink(0, j, sjj);
        T col[K];
#pragma unroll
        for (int a = 0; a < K; ++a) {
            col[a] = -(w[a] * ip);
            if (j + 1 + a < n) sink(a + 1, j, col[a]);
        }
        // shift: new neighbour set is {j, j+1, .., j+K-1}
#pragma unroll
        for (int a = K - 1; a >= 1; --a)
#pragma unroll
            for (int b = K - 1; b >= 1; --b) Z[a][b] = Z[a - 1][b - 1];
        Z[0][0] = sjj;
#pragma unroll
        for (int a = 1; a < K; ++a) { Z[a][0] = col[a - 1]; Z[0][a] = col[a - 1]; }
    }
}

// ---- phase 3: backward sweep inside chunk p (solve and/or selected inverse) --------------------------------------------
// x_red / sig_red: solution and selected-inverse band ((2K) x n_red, row-major) of the reduced system (may be null
// when the corresponding output is not requested).  x_out[M]; sig_out[(K+1) x M] lower band of A^-1.
template <class T, int K, bool SOLVE, bool SELINV, class Sink>
ASVGP_HD void chunk_backward(const ChunkLayout& lay, int p, int n_steps, const ColumnStore<T, K, true>& store,
                             const T* x_red, const T* sig_red, T* x_out, Sink& sink) {
    constexpr int KR = 2 * K - 1;
    const int n = lay.size(p), s = lay.start(p), M = lay.M, nred = lay.n_reduced();
    const bool has_left = p > 0, has_right = p < lay.P - 1;
    (void)KR;
    T xw[K + 1], xL[K];
    T Z[2 * K][2 * K];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        xw[a + 1] = (SOLVE && has_right) ? x_red[p * K + a] : zero_of<T>();
        xL[a] = (SOLVE && has_left) ? x_red[(p - 1) * K + a] : zero_of<T>();
    }
    xw[0] = zero_of<T>();
    if (SELINV) {
#pragma unroll
        for (int a = 0; a < 2 * K; ++a)
#pragma unroll
            for (int b = 0; b < 2 * K; ++b) Z[a][b] = zero_of<T>();
#pragma unroll
        for (int a = 0; a < K; ++a) {
#pragma unroll
            for (int b = 0; b <= a; ++b) {
                if (has_right) { Z[a][b] = sig_red[(size_t)(a - b) * nred + p * K + b]; Z[b][a] = Z[a][b]; }
                if (has_left) {
                    Z[K + a][K + b] = sig_red[(size_t)(a - b) * nred + (p - 1) * K + b];
                    Z[K + b][K + a] = Z[K + a][K + b];
                }
            }
#pragma unroll
            for (int rho = 0; rho < K; ++rho) {
                if (has_left && has_right) {
                    Z[a][K + rho] = sig_red[(size_t)(K + a - rho) * nred + (p - 1) * K + rho];
                    Z[K + rho][a] = Z[a][K + rho];
                }
            }
        }
    }
    // column records are fetched two columns ahead of their use (they sit in L2: written by phase 1)
    struct Rec { T ip; T lv[2 * K]; T y; };
    auto fetch = [&](int j, Rec& rec) {
        const int jj = j < 0 ? 0 : (j >= n_steps ? n_steps - 1 : j);     // clamped: out-of-range records are never used
        rec.ip = store.at(0, jj, p);
#pragma unroll
        for (int a = 0; a < K; ++a) {
            rec.lv[a] = store.at(a + 1, jj, p);
            rec.lv[K + a] = store.at(K + 1 + a, jj, p);
        }
        rec.y = store.at(ColumnStore<T, K, true>::kY, jj, p);
    };
    auto step = [&](int j, const Rec& rec) {
        const int g = s + j;
        const T ip = rec.ip;
        if (SOLVE) {
            T acc = rec.y;
#pragma unroll
            for (int a = 0; a < K; ++a) { acc -= rec.lv[a] * xw[a + 1]; acc -= rec.lv[K + a] * xL[a]; }
            const T xj = acc * ip;
#pragma unroll
            for (int a = K; a >= 2; --a) xw[a] = xw[a - 1];
            xw[1] = xj;
            x_out[g] = xj;
        }
        if (SELINV) {
            T w[2 * K];
            T dot = zero_of<T>();
#pragma unroll
            for (int a = 0; a < 2 * K; ++a) {
                T acc = zero_of<T>();
#pragma unroll
                for (int b = 0; b < 2 * K; ++b) acc += Z[a][b] * rec.lv[b];
                w[a] = acc;
                dot += rec.lv[a] * acc;
            }
            const T ip2 = ip * ip;
            const T sjj = ip2 + dot * ip2;
            sink(0, g, sjj);
            T col[2 * K];
#pragma unroll
            for (int a = 0; a < 2 * K; ++a) col[a] = -(w[a] * ip);
#pragma unroll
            for (int a = 0; a < K; ++a) {
                if (g + 1 + a < M) sink(a + 1, g, col[a]);                              // rows below, same band
                // (row g, column S_{p-1}[rho]) lies inside the band iff j <= rho
                if (has_left && j <= a) sink(j + K - a, s - K + a, col[K + a]);
            }
            // shift the band part of Z; the separator part stays
#pragma unroll
            for (int a = K - 1; a >= 1; --a) {
#pragma unroll
                for (int b = K - 1; b >= 1; --b) Z[a][b] = Z[a - 1][b - 1];
#pragma unroll
                for (int rho = 0; rho < K; ++rho) { Z[a][K + rho] = Z[a - 1][K + rho]; Z[K + rho][a] = Z[a][K + rho]; }
            }
            Z[0][0] = sjj;
#pragma unroll
            for (int a = 1; a < K; ++a) { Z[a][0] = col[a - 1]; Z[0][a] = col[a - 1]; }
#pragma unroll
            for (int rho = 0; rho < K; ++rho) { Z[0][K + rho] = col[K + rho]; Z[K + rho][0] = col[K + rho]; }
        }
    };
    Rec recA, recB;
    const int top = ((n_steps + 1) & ~1) - 1;        // odd index >= n_steps - 1: steps are taken in pairs (j, j-1)
    fetch(top, recA);
    fetch(top - 1, recB);
    for (int j = top; j >= 0; j -= 2) {
        Rec nextA, nextB;
        fetch(j - 2, nextA);
        if (j < n) step(j, recA);
        fetch(j - 3, nextB);
        if (j - 1 < n && j - 1 >= 0) step(j - 1, recB);
        recA = nextA;
        recB = nextB;
    }
}

}  // namespace asvgp

// =====================================================================================================================
// Phase drivers shared by the CUDA kernels (banded_1d.cu) and the CPU test harness (tests/host_harness.cpp)
// =====================================================================================================================
namespace asvgp {

// ---- the separator system: block cyclic reduction -----------------------------------------------------------------------
// The (P-1) separators form an SPD block-tridiagonal system with K x K blocks.  A serial sweep over its (P-1) K columns
// was the longest phase of a chain; block cyclic reduction needs ceil(log2(P-1)) levels instead, every level handled by
// one lane per separator: level l (stride s = 2^l) eliminates the nodes i = s (2j+1), whose neighbours a = i - s and
// b = i + s stay; it is a Cholesky factorisation in nested-dissection order, so log-det, forward substitution, back
// substitution and the Takahashi recursion (Sigma_ai = -Sigma_aa Y_a - Sigma_ab Y_b, ..., run over the levels in reverse)
// all carry over block by block, and the tangents ride along because everything is templated on the scalar.
template <class T, int B>
struct CrNode {
    T D[B][B];    // diagonal block (lower used) -> its Cholesky factor L_i (lower) -> Sigma_ii (full symmetric)
    T E[B][B];    // coupling with the previous node still present: rows = this node, columns = that node
    T Wa[B][B];   // factor block of this node's column in the rows of neighbour a:  W_a = R_{a,i} L_i^-T
    T Wb[B][B];   // ... in the rows of neighbour b
    T Sa[B][B];   // Sigma_{a,i}
    T Sb[B][B];   // Sigma_{b,i}
    T r[B];       // right-hand side -> y_i = L_i^-1 (...) -> x_i
    T ip[B];      // 1 / L_i[j][j]
    T logdet, quad;
    int info;
};

// Address of the same shared-memory object in CTA `rank` of the thread-block cluster (distributed shared memory)
#if defined(__CUDACC__)
template <class P> __device__ __forceinline__ P* cluster_peer(P* p, int rank) {
    unsigned long long out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((unsigned long long)p), "r"(rank));
    return reinterpret_cast<P*>(out);
}
#endif

template <class T, int K>
struct ChainWork {
    ColumnStore<T, K, true> cols;              // chunk columns (global memory): kFields x n_steps x P
    // everything below is small and lives in shared memory inside the kernels (host memory in the test harness)
    ChunkSchur<T, K>* schur;                   // [P]
    CrNode<T, K>* nodes;                       // [P-1] separator system
    T* x_red;                                  // n_red            } written by cr_export when the Schur pieces are dead:
    T* sig_red;                                // (2K) x n_red     } they alias the schur array (global scratch when clustered)
    int per_cta = 0;                           // > 0: the arrays are dealt over the CTAs of a cluster, per_cta entries each
    ASVGP_HD ChunkSchur<T, K>& sch(int p) const {
#if defined(__CUDA_ARCH__)
        if (per_cta > 0) return *cluster_peer(schur + p % per_cta, p / per_cta);
#endif
        return schur[p];
    }
    ASVGP_HD CrNode<T, K>& node(int i) const {
#if defined(__CUDA_ARCH__)
        if (per_cta > 0) return *cluster_peer(nodes + i % per_cta, i / per_cta);
#endif
        return nodes[i];
    }
};

// bytes of the small (shared-memory) part of a ChainWork and its carving; the same code sizes the host harness
template <class T, int K>
struct ChainSmall {
    ASVGP_HD static size_t align16(size_t n) { return (n + 15) & ~(size_t)15; }
    ASVGP_HD static size_t head_bytes(int P) {
        const size_t nred = (size_t)(P - 1) * K + 1;
        const size_t a = align16((size_t)P * sizeof(ChunkSchur<T, K>));
        const size_t b = align16((size_t)(2 * K) * nred * sizeof(T)) + align16(nred * sizeof(T));
        return a > b ? a : b;
    }
    ASVGP_HD static size_t bytes(int P) { return head_bytes(P) + align16((size_t)(P > 1 ? P - 1 : 1) * sizeof(CrNode<T, K>)); }
    ASVGP_HD static void carve(int P, char* base, ChainWork<T, K>& w) {
        const size_t nred = (size_t)(P - 1) * K + 1;
        w.schur = reinterpret_cast<ChunkSchur<T, K>*>(base);
        w.sig_red = reinterpret_cast<T*>(base);
        w.x_red = reinterpret_cast<T*>(base + align16((size_t)(2 * K) * nred * sizeof(T)));
        w.nodes = reinterpret_cast<CrNode<T, K>*>(base + head_bytes(P));
    }
};

template <class T, int K>
struct ChainTotals { T logdet, quad; int info; };

template <class T, int K, bool STORE, class MatFn, class RhsFn>
ASVGP_HD void chain_phase1(const ChunkLayout& lay, int p, MatFn A, RhsFn rhs, const ChainWork<T, K>& w) {
    const int n_steps = lay.max_size();
    if (lay.P == 1) eliminate_columns<T, K, false, STORE>(0, lay.M, n_steps, A, rhs, w.cols, 0, w.sch(0));
    else eliminate_columns<T, K, true, STORE>(lay.start(p), lay.size(p), n_steps, A, rhs, w.cols, p, w.sch(p));
}

// Executed by lane q < P-1: node q of the separator system from the Schur pieces of the chunks on either side of S_q.
template <class T, int K>
ASVGP_HD void cr_assemble(const ChunkLayout& lay, int q, const ChainWork<T, K>& w) {
    const ChunkSchur<T, K>& left = w.sch(q);        // chunk q ends at S_q
    const ChunkSchur<T, K>& right = w.sch(q + 1);   // chunk q+1 starts after S_q
    CrNode<T, K>& nd = w.node(q);
#pragma unroll
    for (int a = 0; a < K; ++a) {
        nd.r[a] = left.rend[a] - right.rhsL[a];
#pragma unroll
        for (int b = 0; b < K; ++b) {
            nd.D[a][b] = (b <= a) ? left.Dend[a][b] - right.SLL[a][b] : zero_of<T>();
            nd.E[a][b] = left.E[a][b];                // rows of S_q, columns of S_{q-1} (zero for q = 0)
        }
    }
    nd.logdet = zero_of<T>();
    nd.quad = zero_of<T>();
    nd.info = 0;
}

// Executed by the lane of node i when it is eliminated at stride s (neighbours a = i - s, b = i + s when they exist):
// Cholesky of D_i, y_i, its share of log-det and quadratic form, and the factor blocks W_a, W_b.
template <class T, int K>
ASVGP_HD void cr_eliminate(int n, int s, int i, int g_offset, const ChainWork<T, K>& w) {
    CrNode<T, K>& nd = w.node(i);
    const bool has_a = s > 0 && i - s >= 0, has_b = s > 0 && i + s < n;
    // work on register copies: the node lives in shared memory, where every dependent access costs a round trip
    T D[K][K], r[K], ipv[K], E[K][K], Eb[K][K];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        r[a] = nd.r[a];
#pragma unroll
        for (int b = 0; b < K; ++b) {
            D[a][b] = (b <= a) ? nd.D[a][b] : zero_of<T>();
            E[a][b] = has_a ? nd.E[a][b] : zero_of<T>();
            Eb[a][b] = has_b ? w.node(i + s).E[a][b] : zero_of<T>();
        }
    }
    LogAccum<T> logdet;
    logdet.init();
    T quad = zero_of<T>();
    int info = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        T d = D[j][j];
#pragma unroll
        for (int q = 0; q < j; ++q) d -= D[j][q] * D[j][q];
        if (!(value_of(d) > 0.0) && info == 0) info = g_offset + i * K + j + 1;
        const T ip = rsqrt_of(d);
        logdet.add(d, ip);
        ipv[j] = ip;
        D[j][j] = d * ip;
#pragma unroll
        for (int rr = j + 1; rr < K; ++rr) {
            T v = D[rr][j];
#pragma unroll
            for (int q = 0; q < j; ++q) v -= D[rr][q] * D[j][q];
            D[rr][j] = v * ip;
        }
        T y = r[j];
#pragma unroll
        for (int q = 0; q < j; ++q) y -= D[j][q] * r[q];
        y = y * ip;
        r[j] = y;
        quad += y * y;
    }
    T Wa[K][K], Wb[K][K];
#pragma unroll
    for (int x = 0; x < K; ++x)
#pragma unroll
        for (int j = 0; j < K; ++j) {
            T va = E[j][x], vb = Eb[x][j];            // W_a = E_i^T L^-T,  W_b = E_b L^-T (row x solves L z = ...)
#pragma unroll
            for (int q = 0; q < j; ++q) { va -= D[j][q] * Wa[x][q]; vb -= D[j][q] * Wb[x][q]; }
            Wa[x][j] = va * ipv[j];
            Wb[x][j] = vb * ipv[j];
        }
#pragma unroll
    for (int a = 0; a < K; ++a) {
        nd.r[a] = r[a];
        nd.ip[a] = ipv[a];
#pragma unroll
        for (int b = 0; b < K; ++b) {
            if (b <= a) nd.D[a][b] = D[a][b];
            if (has_a) nd.Wa[a][b] = Wa[a][b];
            if (has_b) nd.Wb[a][b] = Wb[a][b];
        }
    }
    nd.logdet = logdet.result();
    nd.quad = quad;
    if (info != 0 && nd.info == 0) nd.info = info;
}

// Executed by the lane of node c that STAYS at stride s (c is a multiple of 2s): Schur updates from the eliminated
// neighbours c + s (c is their `a`) and c - s (c is their `b`), and the new coupling with c - 2s.
template <class T, int K>
ASVGP_HD void cr_update(int n, int s, int c, const ChainWork<T, K>& w) {
    CrNode<T, K>& nd = w.node(c);
    const bool has_p = c + s < n, has_m = c - s >= 0, has_a = c - 2 * s >= 0;
    T D[K][K], r[K], Wp[K][K], yp[K], Wm[K][K], Wma[K][K], ym[K];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        r[a] = nd.r[a];
        yp[a] = has_p ? w.node(c + s).r[a] : zero_of<T>();
        ym[a] = has_m ? w.node(c - s).r[a] : zero_of<T>();
#pragma unroll
        for (int b = 0; b < K; ++b) {
            D[a][b] = (b <= a) ? nd.D[a][b] : zero_of<T>();
            Wp[a][b] = has_p ? w.node(c + s).Wa[a][b] : zero_of<T>();       // c is the `a` of node c + s
            Wm[a][b] = has_m ? w.node(c - s).Wb[a][b] : zero_of<T>();       // c is the `b` of node c - s
            Wma[a][b] = (has_m && has_a) ? w.node(c - s).Wa[a][b] : zero_of<T>();
        }
    }
#pragma unroll
    for (int x = 0; x < K; ++x) {
        T rr = r[x];
#pragma unroll
        for (int j = 0; j < K; ++j) { rr -= Wp[x][j] * yp[j]; rr -= Wm[x][j] * ym[j]; }
        nd.r[x] = rr;
#pragma unroll
        for (int y = 0; y <= x; ++y) {
            T v = D[x][y];
#pragma unroll
            for (int j = 0; j < K; ++j) { v -= Wp[x][j] * Wp[y][j]; v -= Wm[x][j] * Wm[y][j]; }
            nd.D[x][y] = v;
        }
        if (has_m) {
#pragma unroll
            for (int y = 0; y < K; ++y) {             // new coupling R'_{c, c-2s} = -W_b W_a^T
                T v = zero_of<T>();
#pragma unroll
                for (int j = 0; j < K; ++j) v -= Wm[x][j] * Wma[y][j];
                nd.E[x][y] = v;
            }
        }
    }
}

// Back substitution and Takahashi recursion for node i eliminated at stride s (s = 0: the root, no neighbours).
template <class T, int K, bool SOLVE, bool SELINV>
ASVGP_HD void cr_back(int n, int s, int i, const ChainWork<T, K>& w) {
    CrNode<T, K>& nd = w.node(i);
    const bool has_a = s > 0 && i - s >= 0, has_b = s > 0 && i + s < n;
    const CrNode<T, K>& na = w.node(has_a ? i - s : i);
    const CrNode<T, K>& nb = w.node(has_b ? i + s : i);
    // register copies of everything that is read (the node lives in shared memory)
    T L[K][K], ipv[K], Wa[K][K], Wb[K][K];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        ipv[a] = nd.ip[a];
#pragma unroll
        for (int b = 0; b < K; ++b) {
            L[a][b] = (b <= a) ? nd.D[a][b] : zero_of<T>();
            Wa[a][b] = has_a ? nd.Wa[a][b] : zero_of<T>();
            Wb[a][b] = has_b ? nd.Wb[a][b] : zero_of<T>();
        }
    }
    if (SOLVE) {
        T t[K], xa[K], xb[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            t[j] = nd.r[j];
            xa[j] = has_a ? na.r[j] : zero_of<T>();
            xb[j] = has_b ? nb.r[j] : zero_of<T>();
        }
#pragma unroll
        for (int j = 0; j < K; ++j) {
            T v = t[j];
#pragma unroll
            for (int x = 0; x < K; ++x) { v -= Wa[x][j] * xa[x]; v -= Wb[x][j] * xb[x]; }
            t[j] = v;
        }
#pragma unroll
        for (int j = K - 1; j >= 0; --j) {
            T v = t[j];
#pragma unroll
            for (int q = j + 1; q < K; ++q) v -= L[q][j] * t[q];
            t[j] = v * ipv[j];
        }
#pragma unroll
        for (int j = 0; j < K; ++j) nd.r[j] = t[j];
    }
    if (SELINV) {
        // Sigma_aa, Sigma_bb and Sigma_ba (a and b are neighbours one level up, where exactly one of them was eliminated)
        const bool both = has_a && has_b;
        const bool a_is_odd = both && (((i - s) / (2 * s)) & 1);
        T Saa[K][K], Sbb[K][K], Sba[K][K];
#pragma unroll
        for (int x = 0; x < K; ++x)
#pragma unroll
            for (int y = 0; y < K; ++y) {
                Saa[x][y] = has_a ? na.D[x][y] : zero_of<T>();
                Sbb[x][y] = has_b ? nb.D[x][y] : zero_of<T>();
                Sba[x][y] = both ? (a_is_odd ? na.Sb[x][y] : nb.Sa[y][x]) : zero_of<T>();      // rows of b, columns of a
            }
        // L^-1 (lower), column by column
        T Li[K][K];
#pragma unroll
        for (int c = 0; c < K; ++c)
#pragma unroll
            for (int r = 0; r < K; ++r) {
                if (r < c) { Li[r][c] = zero_of<T>(); continue; }
                T v = (r == c) ? make_scalar<T>(1.0, 0.0) : zero_of<T>();
#pragma unroll
                for (int q = c; q < r; ++q) v -= L[r][q] * Li[q][c];
                Li[r][c] = v * ipv[r];
            }
        T Ya[K][K], Yb[K][K];                         // Y_k = W_k L^-1
#pragma unroll
        for (int x = 0; x < K; ++x)
#pragma unroll
            for (int c = 0; c < K; ++c) {
                T va = zero_of<T>(), vb = zero_of<T>();
#pragma unroll
                for (int j = c; j < K; ++j) { va += Wa[x][j] * Li[j][c]; vb += Wb[x][j] * Li[j][c]; }
                Ya[x][c] = va;
                Yb[x][c] = vb;
            }
        T Sai[K][K], Sbi[K][K];                       // Sigma_ai = -Sigma_aa Y_a - Sigma_ab Y_b,  Sigma_bi = -Sigma_ba Y_a - Sigma_bb Y_b
#pragma unroll
        for (int x = 0; x < K; ++x)
#pragma unroll
            for (int c = 0; c < K; ++c) {
                T va = zero_of<T>(), vb = zero_of<T>();
#pragma unroll
                for (int y = 0; y < K; ++y) {
                    va -= Saa[x][y] * Ya[y][c];
                    va -= Sba[y][x] * Yb[y][c];
                    vb -= Sba[x][y] * Ya[y][c];
                    vb -= Sbb[x][y] * Yb[y][c];
                }
                Sai[x][c] = va;
                Sbi[x][c] = vb;
            }
#pragma unroll
        for (int c = 0; c < K; ++c)
#pragma unroll
            for (int c2 = 0; c2 < K; ++c2) {
                T v = zero_of<T>();                   // (L^-T L^-1)[c][c2] - (Y_a^T Sigma_ai)[c][c2] - (Y_b^T Sigma_bi)[c][c2]
#pragma unroll
                for (int r = 0; r < K; ++r) {
                    if (r >= c && r >= c2) v += Li[r][c] * Li[r][c2];
                }
#pragma unroll
                for (int x = 0; x < K; ++x) { v -= Ya[x][c] * Sai[x][c2]; v -= Yb[x][c] * Sbi[x][c2]; }
                nd.D[c][c2] = v;
                if (has_a) nd.Sa[c][c2] = Sai[c][c2];
                if (has_b) nd.Sb[c][c2] = Sbi[c][c2];
            }
    }
}

// Executed by lane q < P-1 once the recursion is complete: separator solution and the block-tridiagonal part of the
// inverse in the band layout phase 3 reads (x_red, sig_red alias the Schur pieces, which are dead by now).
template <class T, int K, bool SOLVE, bool SELINV>
ASVGP_HD void cr_export(const ChunkLayout& lay, int q, const ChainWork<T, K>& w) {
    const int nred = lay.n_reduced();
    const CrNode<T, K>& nd = w.node(q);
    if (SOLVE) {
#pragma unroll
        for (int a = 0; a < K; ++a) w.x_red[q * K + a] = nd.r[a];
    }
    if (SELINV) {
#pragma unroll
        for (int a = 0; a < K; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) w.sig_red[(size_t)(a - b) * nred + q * K + b] = nd.D[a][b];
        if (q >= 1) {
            // Sigma_{q,q-1}: the odd one of (q-1, q) was eliminated at stride 1 with the other as its neighbour
            const bool q_odd = q & 1;
            const CrNode<T, K>& prev = w.node(q - 1);
#pragma unroll
            for (int a = 0; a < K; ++a)
#pragma unroll
                for (int rho = 0; rho < K; ++rho)
                    w.sig_red[(size_t)(K + a - rho) * nred + (q - 1) * K + rho] = q_odd ? nd.Sa[rho][a] : prev.Sb[a][rho];
        }
    }
}

// Sums of the per-chunk and per-separator log-determinants and quadratic forms (one lane).
template <class T, int K>
ASVGP_HD ChainTotals<T, K> chain_totals(const ChunkLayout& lay, const ChainWork<T, K>& w) {
    ChainTotals<T, K> tot;
    tot.logdet = zero_of<T>();
    tot.quad = zero_of<T>();
    tot.info = 0;
    for (int p = 0; p < lay.P; ++p) {
        tot.logdet += w.sch(p).logdet;
        tot.quad += w.sch(p).quad;
        if (w.sch(p).info != 0 && (tot.info == 0 || w.sch(p).info < tot.info)) tot.info = w.sch(p).info;
    }
    for (int q = 0; q + 1 < lay.P; ++q) {
        tot.logdet += w.node(q).logdet;
        tot.quad += w.node(q).quad;
        if (w.node(q).info != 0 && tot.info == 0) tot.info = w.node(q).info;   // failure inside the separator system
    }
    return tot;
}

// The level schedule of the separator recursion, shared by the kernels and the host harness:
//   forward : for (s = 1; s < n; s *= 2) { nodes i % 2s == s: cr_eliminate(i, s);  barrier;  nodes c % 2s == 0: cr_update(c, s);  barrier }
//             node 0: cr_eliminate(0, s = 0);  chain_totals (needs the Schur pieces: before cr_export)
//   backward: node 0: cr_back(0, s = 0);  barrier;  for (s = top; s >= 1; s /= 2) { nodes i % 2s == s: cr_back(i, s);  barrier }
//             nodes q: cr_export(q)
ASVGP_HD int cr_top_stride(int n) {
    int s = 1;
    while (2 * s < n) s *= 2;
    return n > 1 ? s : 0;
}

template <class T, int K, bool SOLVE, bool SELINV, class Sink>
ASVGP_HD void chain_phase3(const ChunkLayout& lay, int p, const ChainWork<T, K>& w, T* x_out, Sink& sink) {
    const int n_steps = lay.max_size();
    if (lay.P == 1) {
        if (SOLVE) backsolve_serial<T, K>(lay.M, w.cols, x_out);
        if (SELINV) selinv_serial<T, K>(lay.M, w.cols, sink);
        return;
    }
    const int nred = lay.n_reduced();
    if (p < lay.P - 1) {                 // separator S_p: copy from the reduced solution
        const int g = lay.sep_start(p);
        for (int a = 0; a < K; ++a) {
            if (SOLVE) x_out[g + a] = w.x_red[p * K + a];
            if (SELINV)
                for (int b = 0; b <= a; ++b)
                    sink(a - b, g + b, w.sig_red[(size_t)(a - b) * nred + p * K + b]);
        }
    }
    chunk_backward<T, K, SOLVE, SELINV>(lay, p, n_steps, w.cols, w.x_red, w.sig_red, x_out, sink);
}

}  // namespace asvgp

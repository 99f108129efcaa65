// Block-band kernels of the 2-D (Kronecker) model: P = K1 (x) K2 + G / sigma^2 is block banded with scalar
// bandwidth w = k (m2 + 1) (reference gpr.py:262); the reference factorises it as a DENSE m1 m2 x m1 m2 matrix
// (tf.linalg.cholesky, gpr.py:293) which is infeasible at 200 x 200 (12.8 GB, 2e13 flop — SURVEY §6).
//
//   asvgp_kron_assemble  <- utils.bands_to_kron_cholesky's Kronecker product + `Kuu + KufKfu / sigma2` (gpr.py:287-292)
//   asvgp_kron_factor    <- tf.linalg.cholesky(P), log-det, triangular_solve(L_P, Kuf_y)              (gpr.py:293-295)
//   asvgp_kron_selinv    <- what TF reverse mode / cholesky_solve extract from P^-1 (gpr.py:307, 319-326): the entries of
//                           P^-1 on the stencil pattern and alpha-like solves
//
// Storage: LAPACK-style lower band, column-major, ab[(i-j) + j*ld] = A[i,j] with ld = w + NB so that every NB-wide
// block column, including its triangular tail, is addressable as a dense column-major window with leading
// dimension ld-1 (element (i0+r, j0+c) = window[r + c*(ld-1)]).  Entries between the true band and the padding
// are zero and stay zero.  The column count is padded to a multiple of NB with unit diagonal.
//
// Factorisation: right-looking blocked Cholesky, three launches per block column (POTRF in shared memory,
// row-parallel TRSM + right-hand-side update, 64x64-tiled SYRK on the fp64 pipe).  Selected inverse: blocked
// Takahashi recursion backwards over the block columns, Sigma_WJ = -Sigma_WW (L_WJ L_JJ^-1),
// Sigma_JJ = L_JJ^-T L_JJ^-1 - Sigma_JW L_WJ L_JJ^-1, two launches per block column.  No tensor cores: B200's fp64
// tensor rate equals its fp64 FMA rate, so DMMA would buy nothing.
#include <cuda_runtime.h>

#include <algorithm>

#include "common.cuh"
#include "../../include/asvgp_b200.h"

namespace asvgp {

constexpr int NB = 64;                 // block-column width
constexpr int TILE = 64;               // GEMM tile (TILE x TILE outputs per CTA, 4x4 per thread)
constexpr int BK = 16;                 // k-slab staged through shared memory

struct BandGeom {
    int m1, m2, K;
    int M;          // m1 * m2
    int Mpad;       // padded to a multiple of NB
    int w;          // scalar bandwidth K * m2 + K
    int ld;         // w + NB
    __host__ __device__ int64_t lda() const { return ld - 1; }
};

static BandGeom make_geom(int m1, int m2, int K) {
    BandGeom g;
    g.m1 = m1; g.m2 = m2; g.K = K;
    g.M = m1 * m2;
    g.Mpad = ((g.M + NB - 1) / NB) * NB;
    g.w = K * m2 + K;
    g.ld = g.w + NB;
    return g;
}

__device__ __forceinline__ double* win(double* ab, const BandGeom& g, int i0, int j0) {
    return ab + (i0 - j0) + (int64_t)j0 * g.ld;       // dense window origin (i0 >= j0)
}

// ------------------------------------------------------------------------------------------------------------------
// assembly
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bb_assemble_kernel(BandGeom g, const double* __restrict__ K1,
                                                          const double* __restrict__ K2,
                                                          const double* __restrict__ Gs, double inv_s2,
                                                          double* __restrict__ ab) {
    const int NS = 2 * g.K + 1;
    const int n_e = (g.K + 1) * NS;
    const int64_t total = (int64_t)g.Mpad * n_e;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = t / n_e;
        const int e = (int)(t % n_e);
        const int d1 = e / NS, d2 = e % NS - g.K;
        if (j >= g.M) {                                     // padding columns: unit diagonal
            if (d1 == 0 && d2 == 0) ab[j * g.ld] = 1.0;
            continue;
        }
        if (d1 == 0 && d2 < 0) continue;
        const int j1 = (int)(j / g.m2), j2 = (int)(j % g.m2);
        const int i1 = j1 + d1, i2 = j2 + d2;
        if (i1 >= g.m1 || i2 < 0 || i2 >= g.m2) continue;
        const int a2 = d2 < 0 ? -d2 : d2, c2 = d2 < 0 ? i2 : j2;
        const double kv = K1[(int64_t)d1 * g.m1 + j1] * K2[(int64_t)a2 * g.m2 + c2];
        const double gv = Gs[(int64_t)e * g.M + j];
        ab[(int64_t)(d1 * g.m2 + d2) + j * g.ld] = fma(inv_s2, gv, kv);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// POTRF of the NB x NB diagonal block in shared memory (+ forward substitution of the right-hand side, log-det)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bb_potrf_kernel(BandGeom g, double* __restrict__ ab, int J0,
                                                       double* __restrict__ rhs, double* __restrict__ scal) {
    __shared__ double sA[NB][NB + 1];
    __shared__ int s_info;
    const int tid = threadIdx.x;
    double* A = win(ab, g, J0, J0);
    const int64_t lda = g.lda();
    if (tid == 0) s_info = 0;
    for (int t = tid; t < NB * NB; t += blockDim.x) {
        const int r = t % NB, c = t / NB;
        sA[r][c] = (r >= c) ? A[r + c * lda] : 0.0;
    }
    __syncthreads();
    for (int k = 0; k < NB; ++k) {
        const double akk = sA[k][k];
        if (tid == 0 && !(akk > 0.0) && s_info == 0) s_info = J0 + k + 1;
        const double ip = rsqrt(akk);
        __syncthreads();
        if (tid == k) sA[k][k] = akk * ip;
        else if (tid > k && tid < NB) sA[tid][k] *= ip;
        __syncthreads();
        const int rem = NB - k - 1;
        for (int t = tid; t < rem * rem; t += blockDim.x) {
            const int r = k + 1 + t % rem, c = k + 1 + t / rem;
            if (r >= c) sA[r][c] -= sA[r][k] * sA[c][k];
        }
        __syncthreads();
    }
    for (int t = tid; t < NB * NB; t += blockDim.x) {
        const int r = t % NB, c = t / NB;
        if (r >= c) A[r + c * lda] = sA[r][c];
    }
    // y_J = L11^-1 b_J (warp 0, column-oriented forward substitution) and this block's share of log|P|, ||y||^2
    if (tid < 32) {
        double y0 = rhs[J0 + tid], y1 = rhs[J0 + 32 + tid];
        double logd = log(sA[tid][tid]) + log(sA[tid + 32][tid + 32]);
        for (int k = 0; k < NB; ++k) {
            double yk = __shfl_sync(0xffffffffu, k < 32 ? y0 : y1, k & 31) / sA[k][k];
            if (tid == (k & 31)) { if (k < 32) y0 = yk; else y1 = yk; }
            if (tid > k) y0 -= sA[tid][k] * yk;
            if (tid + 32 > k) y1 -= sA[tid + 32][k] * yk;
        }
        rhs[J0 + tid] = y0;
        rhs[J0 + 32 + tid] = y1;
        double q = y0 * y0 + y1 * y1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            logd += __shfl_xor_sync(0xffffffffu, logd, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (tid == 0) {
            scal[0] += 2.0 * logd;           // single CTA, stream-ordered: no atomics needed
            scal[1] += q;
            if (s_info != 0 && scal[2] == 0.0) scal[2] = (double)s_info;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// TRSM: rows below the diagonal block, one thread per row:  L21[r,:] = A21[r,:] L11^-T ; rhs[r] -= L21[r,:] . y_J
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bb_trsm_kernel(BandGeom g, double* __restrict__ ab, int J0, int n_rows,
                                                      double* __restrict__ rhs) {
    __shared__ double sL[NB][NB + 1];
    __shared__ double sy[NB], sdinv[NB];
    const int tid = threadIdx.x;
    const int64_t lda = g.lda();
    const double* L11 = win(ab, g, J0, J0);
    for (int t = tid; t < NB * NB; t += blockDim.x) {
        const int r = t % NB, c = t / NB;
        sL[r][c] = (r >= c) ? L11[r + c * lda] : 0.0;
    }
    if (tid < NB) { sy[tid] = rhs[J0 + tid]; sdinv[tid] = 1.0 / L11[tid + tid * lda]; }
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + tid;
    if (r >= n_rows) return;
    double* row = win(ab, g, J0 + NB, J0) + r;        // element (J0+NB+r, J0+c) = row[c*lda]
    double x[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) x[c] = row[c * lda];
    double dot = 0.0;
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        double v[4] = {x[c], 0.0, 0.0, 0.0};          // four partial sums: the dot product is not one long chain
#pragma unroll
        for (int c2 = 0; c2 < c; ++c2) v[c2 & 3] = fma(-x[c2], sL[c][c2], v[c2 & 3]);
        const double xc = ((v[0] + v[1]) + (v[2] + v[3])) * sdinv[c];
        x[c] = xc;
        dot = fma(xc, sy[c], dot);
    }
#pragma unroll
    for (int c = 0; c < NB; ++c) row[c * lda] = x[c];
    rhs[J0 + NB + r] -= dot;
}

// ------------------------------------------------------------------------------------------------------------------
// fp64 tile GEMM building block: acc[4][4] (+)= sum_k A(m,k) B(n,k) over a k-range, operands given by strides
// ------------------------------------------------------------------------------------------------------------------
struct Operand {
    const double* p;
    int64_t s_mn;      // stride along the tile's m (or n) index
    int64_t s_k;       // stride along k
    int valid_mn;      // rows/cols of the tile that exist (others read as zero)
};

__device__ __forceinline__ void load_slab(const Operand& op, int k0, int kvalid, double (*dst)[TILE + 1], int tid) {
    // dst[k][mn]; choose the thread mapping so that the unit-stride direction is the fast one
    if (op.s_mn == 1) {
        for (int t = tid; t < TILE * BK; t += 256) {
            const int mn = t % TILE, k = t / TILE;
            dst[k][mn] = (mn < op.valid_mn && k < kvalid) ? op.p[mn + (int64_t)(k0 + k) * op.s_k] : 0.0;
        }
    } else {
        for (int t = tid; t < TILE * BK; t += 256) {
            const int k = t % BK, mn = t / BK;
            dst[k][mn] = (mn < op.valid_mn && k < kvalid) ? op.p[(int64_t)mn * op.s_mn + (int64_t)(k0 + k) * op.s_k] : 0.0;
        }
    }
}

__device__ __forceinline__ void tile_gemm(double (&acc)[4][4], const Operand& A, const Operand& B, int kdim,
                                          double (*sA)[TILE + 1], double (*sB)[TILE + 1], int tid) {
    const int tm = (tid % 16) * 4, tn = (tid / 16) * 4;
    for (int k0 = 0; k0 < kdim; k0 += BK) {
        const int kvalid = min(BK, kdim - k0);
        __syncthreads();
        load_slab(A, k0, kvalid, sA, tid);
        load_slab(B, k0, kvalid, sB, tid);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sA[k][tm + i]; b[i] = sB[k][tn + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
    }
}

// SYRK on the trailing window: A22[r, c] -= sum_k L21[r,k] L21[c,k], lower triangle, TILE x TILE tiles
__global__ void __launch_bounds__(256) bb_syrk_kernel(BandGeom g, double* __restrict__ ab, int J0, int n_rows) {
    __shared__ double sA[BK][TILE + 1], sB[BK][TILE + 1];
    // linear tile index -> (ti >= tj)
    int ti = 0, t = blockIdx.x;
    while (t > ti) { t -= ti + 1; ++ti; }
    const int tj = t;
    const int tid = threadIdx.x;
    const int64_t lda = g.lda();
    const double* L21 = win(ab, g, J0 + NB, J0);
    Operand A{L21 + ti * TILE, 1, lda, min(TILE, n_rows - ti * TILE)};
    Operand B{L21 + tj * TILE, 1, lda, min(TILE, n_rows - tj * TILE)};
    double acc[4][4] = {};
    tile_gemm(acc, A, B, NB, sA, sB, tid);
    const int r0 = J0 + NB + ti * TILE, c0 = J0 + NB + tj * TILE;
    const int tm = (tid % 16) * 4, tn = (tid / 16) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = r0 + tm + i, c = c0 + tn + j;
            if (r >= c && tm + i < A.valid_mn && tn + j < B.valid_mn && r - c < g.ld)
                ab[(r - c) + (int64_t)c * g.ld] -= acc[i][j];
        }
}

// ------------------------------------------------------------------------------------------------------------------
// batched inverse of the triangular diagonal blocks (after the factorisation, off the critical path)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NB) bb_trinv_kernel(BandGeom g, const double* __restrict__ ab,
                                                      double* __restrict__ Linv) {
    __shared__ double sL[NB][NB + 1];
    const int blk = blockIdx.x, J0 = blk * NB, c = threadIdx.x;
    const int64_t lda = g.lda();
    const double* L11 = ab + (int64_t)J0 * g.ld;
    for (int t = c; t < NB * NB; t += NB) {
        const int r = t % NB, cc = t / NB;
        sL[r][cc] = (r >= cc) ? L11[r + cc * lda] : 0.0;
    }
    __syncthreads();
    // thread c: column c of L^-1 by forward substitution
    double x[NB];
#pragma unroll
    for (int r = 0; r < NB; ++r) {
        double v = (r == c) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < r; ++k) v = fma(-sL[r][k], x[k], v);
        x[r] = (r >= c) ? v / sL[r][r] : 0.0;
    }
    double* out = Linv + (int64_t)blk * NB * NB;         // column-major NB x NB
#pragma unroll
    for (int r = 0; r < NB; ++r) out[r + c * NB] = x[r];
}

// ------------------------------------------------------------------------------------------------------------------
// selected inverse, step 1 of a block column: Y = L21 Linv (one thread per row), Sigma_JJ <- Linv^T Linv,
// back-substitution x_J = Linv^T (y_J - L21^T x_W)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bb_sel_prep_kernel(BandGeom g, const double* __restrict__ Lb,
                                                          const double* __restrict__ Linv_all, int J0, int n_rows,
                                                          double* __restrict__ Y, int ldy, double* __restrict__ Sig,
                                                          double* __restrict__ x) {
    __shared__ double sI[NB][NB + 1];        // Linv (lower)
    __shared__ double sv[NB];
    const int tid = threadIdx.x;
    const int64_t lda = g.lda();
    const double* Linv = Linv_all + (int64_t)(J0 / NB) * NB * NB;
    for (int t = tid; t < NB * NB; t += blockDim.x) sI[t % NB][t / NB] = Linv[t];
    __syncthreads();
    const int n_row_blocks = (n_rows + 127) / 128;
    if ((int)blockIdx.x < n_row_blocks) {
        const int r = blockIdx.x * 128 + tid;
        if (r < n_rows) {
            const double* row = Lb + (J0 + NB - J0) + (int64_t)J0 * g.ld + r;     // element (J0+NB+r, J0+c)
            double l[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) l[c] = row[c * lda];
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                double v = 0.0;
#pragma unroll
                for (int c2 = c; c2 < NB; ++c2) v = fma(l[c2], sI[c2][c], v);
                Y[r + (int64_t)c * ldy] = v;
            }
        }
    } else if ((int)blockIdx.x == n_row_blocks) {
        // Sigma_JJ (lower) <- Linv^T Linv
        double* S = Sig + (int64_t)J0 * g.ld;
        for (int t = tid; t < NB * NB; t += blockDim.x) {
            const int r = t % NB, c = t / NB;
            if (r < c) continue;
            double v = 0.0;
            for (int k = r; k < NB; ++k) v = fma(sI[k][r], sI[k][c], v);
            S[r + c * lda] = v;
        }
    } else {
        // x_J = Linv^T (y_J - L21^T x_W)
        if (tid < NB) {
            const double* col = Lb + NB + (int64_t)(J0 + tid) * g.ld - tid;      // element (J0+NB+r, J0+tid) = col[r]
            double v = x[J0 + tid];
            for (int r = 0; r < n_rows; ++r) v = fma(-col[r], x[J0 + NB + r], v);
            sv[tid] = v;
        }
        __syncthreads();
        if (tid < NB) {
            double v = 0.0;
            for (int k = tid; k < NB; ++k) v = fma(sI[k][tid], sv[k], v);
            x[J0 + tid] = v;
        }
    }
}

// step 2: T_I = -(Sigma_WW Y)[I,:]  -> Sigma band (rows I of block column J);  Sigma_JJ -= T_I^T Y_I (atomics)
__global__ void __launch_bounds__(256) bb_sel_symm_kernel(BandGeom g, double* __restrict__ Sig, int J0, int n_rows,
                                                          const double* __restrict__ Y, int ldy) {
    extern __shared__ __align__(16) double symm_smem[];
    double (*sA)[TILE + 1] = reinterpret_cast<double (*)[TILE + 1]>(symm_smem);
    double (*sB)[TILE + 1] = reinterpret_cast<double (*)[TILE + 1]>(symm_smem + BK * (TILE + 1));
    double (*sT)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(symm_smem + 2 * BK * (TILE + 1));
    const int tid = threadIdx.x;
    const int ti = blockIdx.x;                 // row tile of W
    const int W0 = J0 + NB;
    const int64_t lda = g.lda();
    const int rows_here = min(TILE, n_rows - ti * TILE);
    double acc[4][4] = {};
    const int n_kt = (n_rows + TILE - 1) / TILE;
    for (int kt = 0; kt < n_kt; ++kt) {
        const int kvalid = min(TILE, n_rows - kt * TILE);
        Operand B{Y + kt * TILE, ldy, 1, NB};                       // B(n, k) = Y[kt*TILE + k, n]
        if (kt < ti) {
            // block (ti, kt) is stored (lower): Sigma[W0 + ti*T + m, W0 + kt*T + k]
            const double* p = Sig + ((ti - kt) * TILE) + (int64_t)(W0 + kt * TILE) * g.ld;
            Operand A{p, 1, lda, rows_here};
            tile_gemm(acc, A, B, kvalid, sA, sB, tid);
        } else if (kt > ti) {
            // block (ti, kt) = block (kt, ti)^T: Sigma[W0 + kt*T + k, W0 + ti*T + m]
            const double* p = Sig + ((kt - ti) * TILE) + (int64_t)(W0 + ti * TILE) * g.ld;
            Operand A{p, lda, 1, rows_here};
            tile_gemm(acc, A, B, kvalid, sA, sB, tid);
        } else {
            // diagonal block: symmetric read of the stored lower triangle
            const double* p = Sig + (int64_t)(W0 + ti * TILE) * g.ld;
            const int tm = (tid % 16) * 4, tn = (tid / 16) * 4;
            for (int k0 = 0; k0 < kvalid; k0 += BK) {
                const int kv = min(BK, kvalid - k0);
                __syncthreads();
                for (int t = tid; t < TILE * BK; t += 256) {
                    const int m = t % TILE, k = t / TILE, kk = k0 + k;
                    double v = 0.0;
                    if (m < rows_here && k < kv) v = (m >= kk) ? p[m + kk * lda] : p[kk + m * lda];
                    sA[k][m] = v;
                }
                load_slab(B, k0, kv, sB, tid);
                __syncthreads();
#pragma unroll
                for (int k = 0; k < BK; ++k) {
                    double a[4], b[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { a[i] = sA[k][tm + i]; b[i] = sB[k][tn + i]; }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
                }
            }
        }
    }
    // write T_I = -acc into the Sigma band and keep it in shared memory for the Sigma_JJ contribution
    const int tm = (tid % 16) * 4, tn = (tid / 16) * 4;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = tm + i, c = tn + j;
            const double v = -acc[i][j];
            sT[m][c] = (m < rows_here) ? v : 0.0;
            if (m < rows_here) {
                const int r = W0 + ti * TILE + m, col = J0 + c;
                if (r - col < g.ld) Sig[(r - col) + (int64_t)col * g.ld] = v;
            }
        }
    __syncthreads();
    // Sigma_JJ[a, b] -= sum_m T_I[m, a] Y_I[m, b]   (lower triangle a >= b)
    double* S = Sig + (int64_t)J0 * g.ld;
    for (int t = tid; t < NB * NB; t += 256) {
        const int a = t % NB, b = t / NB;
        if (a < b) continue;
        double v = 0.0;
        for (int m = 0; m < rows_here; ++m) v = fma(sT[m][a], Y[ti * TILE + m + (int64_t)b * ldy], v);
        atomicAdd(S + a + b * lda, -v);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// stencil extraction and the scalar contractions needed by the gradients
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bb_extract_stencil_kernel(BandGeom g, const double* __restrict__ Sig,
                                                                 double* __restrict__ out) {
    const int NS = 2 * g.K + 1, n_e = (g.K + 1) * NS;
    const int64_t total = (int64_t)g.M * n_e;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(t / g.M);
        const int64_t j = t % g.M;
        const int d1 = e / NS, d2 = e % NS - g.K;
        const int j1 = (int)(j / g.m2), j2 = (int)(j % g.m2);
        const int i1 = j1 + d1, i2 = j2 + d2;
        double v = 0.0;
        if (!(d1 == 0 && d2 < 0) && i1 < g.m1 && i2 >= 0 && i2 < g.m2) v = Sig[(int64_t)(d1 * g.m2 + d2) + j * g.ld];
        out[t] = v;
    }
}

// For the stencil operators  Op in { G,  dK1 (x) K2,  K1 (x) dK2,  K1 (x) K2 }  and a symmetric stencil field S:
//   out[o]     = sum_{i,j} S[i,j] Op[i,j]   (full symmetric sum = diagonal once + off-diagonals twice)
//   out[4 + o] = x^T Op x
// plus the Kronecker trace terms  out[8..10] = sum G[(i),(j)] T1[i1,j1] T2[i2,j2] for (T1,T2) in
//   (S1,S2), (dS1,S2), (S1,dS2)   (reference gpr.py:307 trace(cholesky_solve(L_Kuu, KufKfu)) and its derivatives).
struct StencilTerms {
    const double *SigP, *Gs, *x;
    const double *K1, *dK1, *K2, *dK2;        // lower bands (K+1) x m
    const double *S1, *dS1, *S2, *dS2;        // lower bands of K1^-1, d(K1^-1), K2^-1, d(K2^-1)
};

__device__ __forceinline__ double band_sym(const double* B, int m, int i, int j) {
    const int d = i - j;
    return d >= 0 ? B[(int64_t)d * m + j] : B[(int64_t)(-d) * m + i];
}

__global__ void __launch_bounds__(256) bb_terms_kernel(BandGeom g, StencilTerms a, double* __restrict__ out) {
    const int NS = 2 * g.K + 1, n_e = (g.K + 1) * NS;
    const int64_t total = (int64_t)g.M * n_e;
    double acc[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) acc[i] = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(t / g.M);
        const int64_t j = t % g.M;
        const int d1 = e / NS, d2 = e % NS - g.K;
        if (d1 == 0 && d2 < 0) continue;
        const int j1 = (int)(j / g.m2), j2 = (int)(j % g.m2);
        const int i1 = j1 + d1, i2 = j2 + d2;
        if (i1 >= g.m1 || i2 < 0 || i2 >= g.m2) continue;
        const int64_t i = (int64_t)i1 * g.m2 + i2;
        const double wgt = (d1 == 0 && d2 == 0) ? 1.0 : 2.0;
        const double s = a.SigP[t], gv = a.Gs[t];
        const double k1 = a.K1[(int64_t)d1 * g.m1 + j1], dk1 = a.dK1[(int64_t)d1 * g.m1 + j1];
        const double k2 = band_sym(a.K2, g.m2, i2, j2), dk2 = band_sym(a.dK2, g.m2, i2, j2);
        const double xx = wgt * a.x[i] * a.x[j], ws = wgt * s;
        const double op[4] = {gv, dk1 * k2, k1 * dk2, k1 * k2};
#pragma unroll
        for (int o = 0; o < 4; ++o) { acc[o] = fma(ws, op[o], acc[o]); acc[4 + o] = fma(xx, op[o], acc[4 + o]); }
        const double s1 = a.S1[(int64_t)d1 * g.m1 + j1], ds1 = a.dS1[(int64_t)d1 * g.m1 + j1];
        const double s2 = band_sym(a.S2, g.m2, i2, j2), ds2 = band_sym(a.dS2, g.m2, i2, j2);
        const double wg = wgt * gv;
        acc[8] = fma(wg, s1 * s2, acc[8]);
        acc[9] = fma(wg, ds1 * s2, acc[9]);
        acc[10] = fma(wg, s1 * ds2, acc[10]);
    }
    __shared__ double s_red[11][8];
#pragma unroll
    for (int i = 0; i < 11; ++i) {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) s_red[i][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 11) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += s_red[threadIdx.x][w];
        atomicAdd(out + threadIdx.x, v);
    }
}

}  // namespace asvgp

using namespace asvgp;

extern "C" int64_t asvgp_kron_band_doubles(int m1, int m2, int order) {
    if (m1 <= 0 || m2 <= 0 || order < 1 || order > kMaxOrder) return -1;
    const BandGeom g = make_geom(m1, m2, order);
    return (int64_t)g.Mpad * g.ld;
}

extern "C" int64_t asvgp_kron_work_doubles(int m1, int m2, int order) {
    if (m1 <= 0 || m2 <= 0 || order < 1 || order > kMaxOrder) return -1;
    const BandGeom g = make_geom(m1, m2, order);
    // Linv for every block + Y scratch (ld x NB) + padded rhs/x vector
    return (int64_t)(g.Mpad / NB) * NB * NB + (int64_t)g.ld * NB + g.Mpad + g.ld + 64;
}

// Assembles P into `band` (asvgp_kron_band_doubles doubles) and factorises it in place.
// rhs_io[Mpad + ld]: in = Kuf_y (zero padded), out = y = L^-1 Kuf_y.   scal[3] = log|P|, ||y||^2, info.
extern "C" int asvgp_kron_factor(const double* K1, const double* K2, const double* Gs, int m1, int m2, int order,
                                 double sigma2, double* band, double* rhs_io, double* scal, void* stream) {
    ASVGP_REQUIRE(m1 > 0 && m2 > 0 && order >= 1 && order <= kMaxOrder, "kron_factor: m=%d,%d order=%d", m1, m2, order);
    ASVGP_REQUIRE(sigma2 > 0.0, "kron_factor: sigma2=%g", sigma2);
    const BandGeom g = make_geom(m1, m2, order);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ASVGP_CUDA_OK(cudaMemsetAsync(band, 0, (size_t)g.Mpad * g.ld * sizeof(double), st));
    ASVGP_CUDA_OK(cudaMemsetAsync(scal, 0, 3 * sizeof(double), st));
    const int64_t total = (int64_t)g.Mpad * (order + 1) * (2 * order + 1);
    bb_assemble_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 16), 256, 0, st>>>(g, K1, K2, Gs, 1.0 / sigma2, band);
    ASVGP_CUDA_OK(cudaGetLastError());
    for (int J0 = 0; J0 < g.Mpad; J0 += NB) {
        bb_potrf_kernel<<<1, 256, 0, st>>>(g, band, J0, rhs_io, scal);
        const int n_rows = std::min(g.w, g.Mpad - (J0 + NB));
        if (n_rows > 0) {
            bb_trsm_kernel<<<(n_rows + 127) / 128, 128, 0, st>>>(g, band, J0, n_rows, rhs_io);
            const int nt = (n_rows + TILE - 1) / TILE;
            bb_syrk_kernel<<<nt * (nt + 1) / 2, 256, 0, st>>>(g, band, J0, n_rows);
        }
    }
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

// From the factor: sigma_stencil[(order+1)(2 order+1) x M] = entries of P^-1 on the stencil, x_io: in y = L^-1 b,
// out x = P^-1 b.  sig_band: scratch of asvgp_kron_band_doubles doubles; work: asvgp_kron_work_doubles doubles.
extern "C" int asvgp_kron_selinv(const double* band, int m1, int m2, int order, double* sig_band, double* x_io,
                                 double* sigma_stencil, double* work, void* stream) {
    ASVGP_REQUIRE(m1 > 0 && m2 > 0 && order >= 1 && order <= kMaxOrder, "kron_selinv: m=%d,%d order=%d", m1, m2, order);
    const BandGeom g = make_geom(m1, m2, order);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* Linv = work;
    double* Y = work + (int64_t)(g.Mpad / NB) * NB * NB;
    const int ldy = g.ld;
    const size_t symm_bytes = (size_t)(2 * BK * (TILE + 1) + TILE * (NB + 1)) * sizeof(double);
    ASVGP_CUDA_OK(cudaFuncSetAttribute(bb_sel_symm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)symm_bytes));
    ASVGP_CUDA_OK(cudaMemsetAsync(sig_band, 0, (size_t)g.Mpad * g.ld * sizeof(double), st));
    bb_trinv_kernel<<<g.Mpad / NB, NB, 0, st>>>(g, band, Linv);
    ASVGP_CUDA_OK(cudaGetLastError());
    for (int J0 = g.Mpad - NB; J0 >= 0; J0 -= NB) {
        const int n_rows = std::min(g.w, g.Mpad - (J0 + NB));
        const int n_row_blocks = (n_rows + 127) / 128;
        bb_sel_prep_kernel<<<n_row_blocks + 2, 128, 0, st>>>(g, band, Linv, J0, n_rows, Y, ldy, sig_band, x_io);
        if (n_rows > 0)
            bb_sel_symm_kernel<<<(n_rows + TILE - 1) / TILE, 256, symm_bytes, st>>>(g, sig_band, J0, n_rows, Y, ldy);
    }
    ASVGP_CUDA_OK(cudaGetLastError());
    const int64_t total = (int64_t)g.M * (order + 1) * (2 * order + 1);
    bb_extract_stencil_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 16), 256, 0, st>>>(g, sig_band, sigma_stencil);
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

// out[11] (device, zeroed here): see bb_terms_kernel.
extern "C" int asvgp_kron_terms(const double* SigP, const double* Gs, const double* x, const double* K1,
                                const double* dK1, const double* K2, const double* dK2, const double* S1,
                                const double* dS1, const double* S2, const double* dS2, int m1, int m2, int order,
                                double* out, void* stream) {
    ASVGP_REQUIRE(m1 > 0 && m2 > 0 && order >= 1 && order <= kMaxOrder, "kron_terms: m=%d,%d order=%d", m1, m2, order);
    const BandGeom g = make_geom(m1, m2, order);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ASVGP_CUDA_OK(cudaMemsetAsync(out, 0, 11 * sizeof(double), st));
    StencilTerms a{SigP, Gs, x, K1, dK1, K2, dK2, S1, dS1, S2, dS2};
    const int64_t total = (int64_t)g.M * (order + 1) * (2 * order + 1);
    bb_terms_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 4), 256, 0, st>>>(g, a, out);
    ASVGP_CUDA_OK(cudaGetLastError());
    return kOk;
}

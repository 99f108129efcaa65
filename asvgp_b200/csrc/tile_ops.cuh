// Device building blocks shared by the tile kernels of the 2-D (Kronecker) model (tiledag_2d.cu: block-band tile DAG;
// ndfront_2d.cu: nested-dissection fronts): 64 x 64 fp64 tiles staged by TMA bulk copies, tile products on the fp64 tensor
// cores (mma.sync.m8n8k4.f64), the in-register Cholesky + inverse of a diagonal tile, release/acquire flag waits.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>

#include "async_copy.cuh"
#include "common.cuh"

namespace asvgp {

constexpr int NB = 64;
constexpr int TILE = NB * NB;                   // doubles per tile
constexpr int TILE_BYTES = TILE * 8;
constexpr int kTdThreads = 256;                 // 16 x 16 threads, 4 x 4 outputs each
constexpr long long kSpinLimit = 1LL << 24;     // polls before a wait gives up and raises the abort flag (~ seconds)
constexpr int kColStat = 20;              // per block column: log-det share, ||y||^2 share, 6 + 4 factorisation stamps, 4 selected-inverse stamps
constexpr int kNoBadPivot = 0x7f7f7f7f;         // what cudaMemset(0x7f) leaves in the "first bad pivot" slot

__device__ __forceinline__ void tma_load_tile_(double* dst_smem, const double* src_gmem, uint64_t* bar) {
    tma_load_bulk(dst_smem, src_gmem, (uint32_t)TILE_BYTES, bar);
}

// ---- tile products on the fp64 tensor cores --------------------------------------------------------------------------
// mma.sync.m8n8k4.f64 sustains the full fp64 rate of the SM (64 FMA/clk, tools/microbench/dmma_bench.cu) where the DFMA
// outer-product loop of tile_mma reaches 42 %.  A fragment load takes element (row m = lane/4, k = lane%4) of an operand
// stored k-major; with 64-double columns the four k of a row share a bank, so the operands of the bulk products are staged
// with a column stride of LDT = 68 doubles (68 mod 16 = 4: the 16 lanes of a half-warp hit 16 different banks): one TMA
// bulk copy per column (64 x 512 B, spread over the lanes of warp 0) instead of one per tile, same mbarrier, same byte count.
constexpr int LDT = NB + 4;
constexpr int PTILE = NB * LDT;                 // doubles of a staged (padded) operand tile
__device__ __forceinline__ void tma_load_tile_padded(double* dst_smem, const double* src_gmem, uint64_t* bar, int lane) {
    for (int c = lane; c < NB; c += 32) tma_load_bulk(dst_smem + c * LDT, src_gmem + c * NB, (uint32_t)(NB * 8), bar);
}
// unpadded tile (as the TMA delivered it) -> padded copy, all threads of the CTA
__device__ __forceinline__ void repack_padded(double* __restrict__ dst, const double* __restrict__ src, int tid) {
#pragma unroll
    for (int idx = tid; idx < TILE / 2; idx += kTdThreads) {
        const int c = idx >> 5, r = (idx & 31) * 2;
        *reinterpret_cast<double2*>(dst + c * LDT + r) = *reinterpret_cast<const double2*>(src + c * NB + r);
    }
}
// Same for a DIAGONAL tile of the selected inverse, which is symmetric only up to rounding (it is a sum of REDs of
// termwise unsymmetric products): the copy handed to the tensor cores is 0.5 (S + S^T).  The Takahashi recursion
// amplifies an unsymmetric rounding component by about 2x per block column when Kuu dominates P (l / delta ~ 18 at
// 200 x 200: 1e-16 -> 1e+60 over 625 block columns), so symmetry is enforced wherever such a tile is consumed.
// Thread (c = tid % 64, g = tid / 64) walks rows (c + 16 g + j) % 64: both reads and the write are bank-conflict free
// (banks r % 16, c % 16 and (5 c + j) % 16 over the 16 lanes of a half-warp).
__device__ __forceinline__ void repack_padded_sym(double* __restrict__ dst, const double* __restrict__ src, int tid) {
    const int c = tid & (NB - 1), g = tid >> 6;
#pragma unroll
    for (int j = 0; j < NB / (kTdThreads / NB); ++j) {
        const int r = (c + g * (NB / (kTdThreads / NB)) + j) & (NB - 1);
        dst[c * LDT + r] = 0.5 * (src[c * NB + r] + src[r * NB + c]);
    }
}
__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// cf += A B^T for operands staged k-major with stride LDT (element (m, k) at [k * LDT + m]); warp w owns rows 8w..8w+7 of
// the 64 x 64 result as eight 8 x 8 fragments (lane: row lane/4, columns 2 (lane%4) + {0, 1} of each)
__device__ __forceinline__ void dmma_tile(double (&cf)[8][2], const double* __restrict__ A, const double* __restrict__ B, int warp, int lane) {
    const double* pa = A + (lane & 3) * LDT + warp * 8 + (lane >> 2);
    const double* pb = B + (lane & 3) * LDT + (lane >> 2);
#pragma unroll 4
    for (int k0 = 0; k0 < NB; k0 += 4) {
        const double a = pa[k0 * LDT];
#pragma unroll
        for (int cb = 0; cb < 8; ++cb) dmma_m8n8k4(cf[cb], a, pb[k0 * LDT + cb * 8]);
    }
}
// acc -= (the product held as fragments), through a column-major 64 x 64 scratch tile in shared memory
__device__ __forceinline__ void frags_subtract(double (&acc)[4][4], const double (&cf)[8][2], double* scratch, int warp, int lane,
                                               int tm, int tn) {
    const int row = warp * 8 + (lane >> 2), col = (lane & 3) * 2;
#pragma unroll
    for (int cb = 0; cb < 8; ++cb) {
        scratch[(cb * 8 + col) * NB + row] = cf[cb][0];
        scratch[(cb * 8 + col + 1) * NB + row] = cf[cb][1];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double2 a = *reinterpret_cast<const double2*>(scratch + (tn + j) * NB + tm);
        const double2 b = *reinterpret_cast<const double2*>(scratch + (tn + j) * NB + tm + 2);
        acc[0][j] -= a.x; acc[1][j] -= a.y; acc[2][j] -= b.x; acc[3][j] -= b.y;
    }
    __syncthreads();
}

// Spin until *flag >= want; gives up (and makes everybody give up) after kSpinLimit polls so that a logic error can
// never hang the device.
__device__ __forceinline__ void wait_flag(const int* flag, int want, int* abort_flag) {
    long long spins = 0;
    while (ld_acquire(flag) < want) {
        if ((++spins & 1023) == 0) {
            if (ld_acquire(abort_flag) != 0) return;
            if (spins > kSpinLimit) { atomicExch(abort_flag, 1); return; }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 64 x 64 x 64 tile product on the fp64 pipe: acc[i][j] (+/-)= sum_k A(tm+i, k) B(tn+j, k),
// A(m, k) at A[k*LDA + m], B(n, k) at B[k*LDB + n] (shared memory)
// ------------------------------------------------------------------------------------------------------------------
template <int LD>
__device__ __forceinline__ void load4(const double* p, double (&v)[4]) {
    if (LD % 2 == 0) {
        const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
        v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3];
    }
}
template <int LDA, int LDB, bool SUB>
__device__ __forceinline__ void tile_mma(double (&acc)[4][4], const double* __restrict__ A,
                                         const double* __restrict__ B, int tm, int tn) {
#pragma unroll 4
    for (int k = 0; k < NB; ++k) {
        double a[4], b[4];
        load4<LDA>(A + k * LDA + tm, a);
        load4<LDB>(B + k * LDB + tn, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double ai = SUB ? -a[i] : a[i];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(ai, b[j], acc[i][j]);
        }
    }
}

// stage s of a two-stage ring laid out [A0 | B0 | A1 | B1] (pointer arithmetic instead of an indexed pointer array,
// which would live in local memory)
struct StageBufs {
    double* base;
    __device__ __forceinline__ double* operator[](int s) const { return base + s * 2 * TILE; }
};
struct Phases {                // parity bits of the two mbarriers
    uint32_t bits = 0u;
    __device__ __forceinline__ uint32_t get(int s) const { return (bits >> s) & 1u; }
    __device__ __forceinline__ void flip(int s) { bits ^= 1u << s; }
};

// registers <-> column-major tile (element (r, c) at [c*64 + r]); thread owns rows tm..tm+3, columns tn..tn+3
__device__ __forceinline__ void regs_from_tile(double (&acc)[4][4], const double* __restrict__ t, int tm, int tn) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double2 a = *reinterpret_cast<const double2*>(t + (tn + j) * NB + tm);
        const double2 b = *reinterpret_cast<const double2*>(t + (tn + j) * NB + tm + 2);
        acc[0][j] = a.x; acc[1][j] = a.y; acc[2][j] = b.x; acc[3][j] = b.y;
    }
}
__device__ __forceinline__ void regs_to_tile(const double (&acc)[4][4], double* __restrict__ t, int tm, int tn) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        *reinterpret_cast<double2*>(t + (tn + j) * NB + tm) = make_double2(acc[0][j], acc[1][j]);
        *reinterpret_cast<double2*>(t + (tn + j) * NB + tm + 2) = make_double2(acc[2][j], acc[3][j]);
    }
}
template <int LD>
__device__ __forceinline__ void regs_to_tile_ld(const double (&acc)[4][4], double* __restrict__ t, int tm, int tn) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        *reinterpret_cast<double2*>(t + (tn + j) * LD + tm) = make_double2(acc[0][j], acc[1][j]);
        *reinterpret_cast<double2*>(t + (tn + j) * LD + tm + 2) = make_double2(acc[2][j], acc[3][j]);
    }
}
// transposed store: element (r, c) at [r*64 + c]
__device__ __forceinline__ void regs_to_tile_t(const double (&acc)[4][4], double* __restrict__ t, int tm, int tn) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        *reinterpret_cast<double2*>(t + (tm + i) * NB + tn) = make_double2(acc[i][0], acc[i][1]);
        *reinterpret_cast<double2*>(t + (tm + i) * NB + tn + 2) = make_double2(acc[i][2], acc[i][3]);
    }
}

template <int LD>
__device__ __forceinline__ void regs_to_tile_t_ld(const double (&acc)[4][4], double* __restrict__ t, int tm, int tn) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        *reinterpret_cast<double2*>(t + (tm + i) * LD + tn) = make_double2(acc[i][0], acc[i][1]);
        *reinterpret_cast<double2*>(t + (tm + i) * LD + tn + 2) = make_double2(acc[i][2], acc[i][3]);
    }
}

// out[r] = sum over the 16 column groups of part[q][r]; part is [16][64] in shared memory (deterministic reduction)
__device__ __forceinline__ double reduce16(const double* part, int r) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) s += part[q * NB + r];
    return s;
}

__device__ __forceinline__ long long clock_mem() {          // clock64 that the compiler cannot move across memory ops / barriers
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
}

// Cholesky of the symmetric 64 x 64 tile held in registers (4 x 4 per thread), right-looking in 16 block steps of four
// columns, carrying V = L^-1 along (V starts as the identity and receives the same row operations).  Per block step:
//   the thread that owns the 4 x 4 diagonal block factorises it and inverts its factor in registers   -> s11 (barrier)
//   the 16 threads that own the block column form their rows of the panel  L = A L11^-T               -> spanel
//   the 16 threads that own the block row of V form the final rows           W = L11^-1 V             -> swrow (barrier)
//   every thread applies the rank-4 update to its 4 x 4 block of A (columns to the right) or of V (rows below).
// Two barriers per FOUR columns instead of one per column, and the scalar sqrt chain runs in one thread's registers.
// On exit acc holds L (lower incl. diagonal; entries above the diagonal are zero) and V holds L^-1 (lower).
__device__ __forceinline__ void potrf_regs(double (&acc)[4][4], double (&V)[4][4], int tm, int tn, double* s11,
                                           double* spanel, double* swrow, int* s_bad, double* prof) {
    long long t_diag = 0, t_b1 = 0, t_b2 = 0, t_upd = 0;        // thread 0's clock64 per phase (diagnostics)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) V[i][j] = (tm + i == tn + j) ? 1.0 : 0.0;
#pragma unroll 1
    for (int k0 = 0; k0 < NB; k0 += 4) {
        const long long c0 = clock_mem();
        if (tm == k0 && tn == k0) {
            // ---- 4 x 4 diagonal block: Cholesky in registers; s11 <- { l (row-major 4 x 4, lower), 1 / l_cc } -----------------
            double l[4][4], r[4];
            bool bad = false;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                double d = acc[c][c];
#pragma unroll
                for (int q = 0; q < c; ++q) d = fma(-l[c][q], l[c][q], d);
                bad = bad || !(d > 0.0);
                if (bad && *s_bad < 0) *s_bad = k0 + c;
                r[c] = rsqrt(d);
                l[c][c] = d * r[c];
#pragma unroll
                for (int i = c + 1; i < 4; ++i) {
                    double v = acc[i][c];
#pragma unroll
                    for (int q = 0; q < c; ++q) v = fma(-l[i][q], l[c][q], v);
                    l[i][c] = v * r[c];
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                s11[16 + i] = r[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s11[i * 4 + j] = (j <= i) ? l[i][j] : 0.0;
                    acc[i][j] = (j <= i) ? l[i][j] : 0.0;
                }
            }
        }
        const long long c1 = clock_mem();
        __syncthreads();
        const long long c2 = clock_mem();
        if (tn == k0 || tm == k0) {
            double l[4][4], r[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                r[i] = s11[16 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) l[i][j] = s11[i * 4 + j];
            }
            if (tn == k0) {
                // ---- panel rows below the block: X L11^T = A by forward substitution, x_c = (a_c - sum_{q<c} x_q l[c][q]) / l[c][c]
                if (tm > k0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            double v = acc[i][c];
#pragma unroll
                            for (int q = 0; q < c; ++q) v = fma(-acc[i][q], l[c][q], v);
                            acc[i][c] = v * r[c];
                        }
                    }
                } else if (tm < k0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[i][c] = 0.0;
                }
                // spanel is stored TRANSPOSED, [c][row]: the 16 owners write (and everybody later reads) consecutive 32-byte
                // chunks, free of bank conflicts
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    *reinterpret_cast<double2*>(spanel + c * NB + tm) = make_double2(acc[0][c], acc[1][c]);
                    *reinterpret_cast<double2*>(spanel + c * NB + tm + 2) = make_double2(acc[2][c], acc[3][c]);
                }
            }
            if (tm == k0) {
                // ---- final rows k0..k0+3 of L^-1: L11 W = V_block by forward substitution -----------------------------------------
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        double v = V[i][j];
#pragma unroll
                        for (int q = 0; q < i; ++q) v = fma(-l[i][q], V[q][j], v);
                        V[i][j] = v * r[i];
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    *reinterpret_cast<double2*>(swrow + i * NB + tn) = make_double2(V[i][0], V[i][1]);
                    *reinterpret_cast<double2*>(swrow + i * NB + tn + 2) = make_double2(V[i][2], V[i][3]);
                }
            }
        }
        __syncthreads();
        const long long c3 = clock_mem();
        if (tn > k0) {
            // ---- rank-4 update of A: acc[i][j] -= sum_c L[tm+i][k0+c] L[tn+j][k0+c]   (rows above the block carry zeros) -------
            if (tm + 3 >= tn) {                      // blocks strictly above the diagonal are never read
                double pr[4][4], pc[4][4];           // [c][i]
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double2 a0 = *reinterpret_cast<const double2*>(spanel + c * NB + tm);
                    const double2 a1 = *reinterpret_cast<const double2*>(spanel + c * NB + tm + 2);
                    pr[c][0] = a0.x; pr[c][1] = a0.y; pr[c][2] = a1.x; pr[c][3] = a1.y;
                    const double2 b0 = *reinterpret_cast<const double2*>(spanel + c * NB + tn);
                    const double2 b1 = *reinterpret_cast<const double2*>(spanel + c * NB + tn + 2);
                    pc[c][0] = b0.x; pc[c][1] = b0.y; pc[c][2] = b1.x; pc[c][3] = b1.y;
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fma(-pr[c][i], pc[c][j], acc[i][j]);
            }
        } else if (tm > k0) {
            // ---- rank-4 update of V (columns <= k0+3, rows below the block): V[i][j] -= sum_c L[tm+i][k0+c] W[c][tn+j] --------
            double pr[4][4], w[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double2 a0 = *reinterpret_cast<const double2*>(spanel + c * NB + tm);
                const double2 a1 = *reinterpret_cast<const double2*>(spanel + c * NB + tm + 2);
                pr[c][0] = a0.x; pr[c][1] = a0.y; pr[c][2] = a1.x; pr[c][3] = a1.y;
                const double2 w0 = *reinterpret_cast<const double2*>(swrow + c * NB + tn);
                const double2 w1 = *reinterpret_cast<const double2*>(swrow + c * NB + tn + 2);
                w[c][0] = w0.x; w[c][1] = w0.y; w[c][2] = w1.x; w[c][3] = w1.y;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) V[i][j] = fma(-pr[c][i], w[c][j], V[i][j]);
        }
#ifdef ASVGP_DIAG_POTRF
        __syncthreads();                 // diagnostic build: barrier-to-barrier phase times
#endif
        const long long c4 = clock_mem();
        if (k0 == 0) t_diag = c1 - c0;
        t_b1 += c2 - c1; t_b2 += c3 - c2; t_upd += c4 - c3;
    }
    if (threadIdx.x == 0 && prof != nullptr) {
        prof[0] = (double)t_diag; prof[1] = (double)t_b1; prof[2] = (double)t_b2; prof[3] = (double)t_upd;
    }
}

// persistent grid: one CTA per SM (as many as can be co-resident), never more than there are tasks
template <class Kernel>
static int persistent_grid(Kernel kernel, size_t smem, int n_tasks, int* grid) {
    int dev = 0, sms = 0, per_sm = 0;
    ASVGP_CUDA_OK(cudaGetDevice(&dev));
    ASVGP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ASVGP_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ASVGP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTdThreads, smem));
    if (per_sm < 1) {
        set_last_error("tile-DAG kernel does not fit on an SM (%zu bytes of shared memory)", smem);
        return kCudaError;
    }
    *grid = std::max(1, std::min(sms, n_tasks));
    return kOk;
}

constexpr size_t kTdSmem = (4 * (size_t)TILE + 2 * (size_t)PTILE) * sizeof(double);   // two landing stages of (A, B) + one padded pair

}  // namespace asvgp

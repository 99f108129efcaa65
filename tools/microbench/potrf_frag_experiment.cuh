// EXPERIMENT (not part of the library): POTRF + inverse of a 64 x 64 tile with the rank-4 updates of each block step on the
// fp64 tensor cores, the tile living in DMMA fragment layout.  Numerically right (max|LL^T-A| = 8.5e-14 in potrf_bench -DFRAG)
// but 93.9 k cycles against 34.5 k for potrf_regs.  Ablation (cycles per tile): the updates themselves 7.6 k (they were
// ~15 k), the single-lane 4 x 4 Cholesky 11.8 k, turning rows k0..k0+3 of V into rows of L^-1 with shuffles inside ONE warp
// 46 k (!), everything else (panel substitution, two barriers, 20 broadcast LDS per thread, layout conversions) 28 k.
// To be competitive the W rows must be spread over all warps again and the skeleton slimmed; kept here as the starting point.
// Include after tiledag_2d.cu.
#pragma once
namespace asvgp {
// ---- POTRF + inverse of a 64 x 64 tile with the rank-4 updates on the tensor cores ---------------------------------------
// Same contract as potrf_regs (in: the tile in the 4 x 4-per-thread layout; out: L and L^-1 in that layout, junk above the
// diagonal), but inside the tile lives in DMMA fragment layout — warp w holds rows 8w..8w+7, lane (m = lane/4, q = lane%4)
// columns 8cb + 2q + {0, 1} of row 8w + m for cb = 0..7 — so that the two rank-4 updates of a block step,
//   A[i][j] -= sum_c X[i][c] X[j][c]   and   V[i][j] -= sum_c X[i][c] W[c][j]      (X = panel, W = final rows of L^-1),
// are one mma.sync.m8n8k4 per 8 x 8 block (<= 9 per warp per step) instead of 16 LDS.128 + 64 DFMA per thread: the update
// phase was ~930 of the ~2100 cycles of a block step.  A block step: the 8 lanes holding the 4 x 4 diagonal block put it in
// shared memory and one of them factorises it (s11 <- l, 1/l_cc); after barrier 1 the two lanes holding a row's four panel
// entries do the forward substitution (one shuffle hands x0, x1 to the second lane) and warp k0/8 turns rows k0..k0+3 of V
// into rows of L^-1 (four shuffles per value fetch the rows, the substitution is local); panel and W rows go to shared
// memory with a stride of 68 doubles (conflict-free fragment loads); after barrier 2 the updates.  `work`: 608 doubles,
// `scratch`: one 64 x 64 tile (layout conversions at entry and exit).
__device__ __forceinline__ void tile_to_frags(const double* __restrict__ t, double (&f)[8][2], int warp, int lane) {
    const int row = warp * 8 + (lane >> 2), col = (lane & 3) * 2;
#pragma unroll
    for (int cb = 0; cb < 8; ++cb) {
        f[cb][0] = t[(cb * 8 + col) * NB + row];
        f[cb][1] = t[(cb * 8 + col + 1) * NB + row];
    }
}
__device__ __forceinline__ void frags_to_tile(const double (&f)[8][2], double* __restrict__ t, int warp, int lane) {
    const int row = warp * 8 + (lane >> 2), col = (lane & 3) * 2;
#pragma unroll
    for (int cb = 0; cb < 8; ++cb) {
        t[(cb * 8 + col) * NB + row] = f[cb][0];
        t[(cb * 8 + col + 1) * NB + row] = f[cb][1];
    }
}

__device__ __forceinline__ void potrf_frag(double (&acc)[4][4], double (&V)[4][4], int tm, int tn, double* work, double* scratch,
                                           int* s_bad) {
    constexpr int LDS_ = NB + 4;
    double* s11 = work;                 // [16] l (row-major, lower) + [4] 1 / l_cc
    double* sraw = work + 32;           // [16] the diagonal block as it stands
    double* spanel = work + 64;         // [4][68]  X[row][c] at [c * 68 + row]
    double* swrow = spanel + 4 * LDS_;  // [4][68]  W[c][col] at [c * 68 + col]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, fm = lane >> 2, fq = lane & 3;
    const int row = warp * 8 + fm;
    double a[8][2], v[8][2];
    regs_to_tile(acc, scratch, tm, tn);
    __syncthreads();
    tile_to_frags(scratch, a, warp, lane);
#pragma unroll
    for (int cb = 0; cb < 8; ++cb) {
        v[cb][0] = (row == cb * 8 + fq * 2) ? 1.0 : 0.0;
        v[cb][1] = (row == cb * 8 + fq * 2 + 1) ? 1.0 : 0.0;
    }
#pragma unroll 1
    for (int k0 = 0; k0 < NB; k0 += 4) {
        const int cb0 = k0 >> 3, h = (k0 >> 2) & 1;          // column block and half of it; warp cb0 holds rows k0..k0+3
        const bool my_cols = (fq >> 1) == h;                 // this lane holds two of the four panel columns of its row
        double a0 = 0.0, a1 = 0.0;                           // ... namely these entries
#pragma unroll
        for (int cb = 0; cb < 8; ++cb)
            if (cb == cb0) { a0 = a[cb][0]; a1 = a[cb][1]; }
        // ---- phase 1: the diagonal block -------------------------------------------------------------------------------------
        if (warp == cb0) {
            const int i = fm - 4 * h;
            if (i >= 0 && i < 4 && my_cols) *reinterpret_cast<double2*>(sraw + i * 4 + 2 * (fq & 1)) = make_double2(a0, a1);
            __syncwarp();
#ifndef ASVGP_ABL_CHOL
            if (lane == 0) {
                double l[4][4], r[4];
                bool bad = false;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    double d = sraw[c * 4 + c];
#pragma unroll
                    for (int q = 0; q < c; ++q) d = fma(-l[c][q], l[c][q], d);
                    bad = bad || !(d > 0.0);
                    if (bad && *s_bad < 0) *s_bad = k0 + c;
                    r[c] = rsqrt(d);
                    l[c][c] = d * r[c];
#pragma unroll
                    for (int i2 = c + 1; i2 < 4; ++i2) {
                        double t = sraw[i2 * 4 + c];
#pragma unroll
                        for (int q = 0; q < c; ++q) t = fma(-l[i2][q], l[c][q], t);
                        l[i2][c] = t * r[c];
                    }
                }
#pragma unroll
                for (int i2 = 0; i2 < 4; ++i2) {
                    s11[16 + i2] = r[i2];
#pragma unroll
                    for (int j = 0; j < 4; ++j) s11[i2 * 4 + j] = (j <= i2) ? l[i2][j] : 0.0;
                }
            }
#endif
        }
        __syncthreads();
        // ---- phase 2: panel rows and the final rows of L^-1 --------------------------------------------------------------------
        double l[4][4], r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r[i] = s11[16 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) l[i][j] = s11[i * 4 + j];
        }
        {
            // forward substitution X L11^T = A for this lane's row: lane 2h of the row has columns 0, 1, lane 2h+1 columns 2, 3
            const double xa = a0 * r[0];
            const double xb = fma(-xa, l[1][0], a1) * r[1];
            const int src = (lane & ~3) | (2 * h);
            const double X0 = __shfl_sync(0xffffffffu, xa, src), X1 = __shfl_sync(0xffffffffu, xb, src);
            const double xc = fma(-X1, l[2][1], fma(-X0, l[2][0], a0)) * r[2];
            const double xd = fma(-xc, l[3][2], fma(-X1, l[3][1], fma(-X0, l[3][0], a1))) * r[3];
            if (my_cols) {
                const bool second = (fq & 1) != 0;
                double n0 = 0.0, n1 = 0.0;                      // rows above the block carry zeros
                if (row > k0 + 3) {
                    n0 = second ? xc : xa;
                    n1 = second ? xd : xb;
                } else if (row >= k0) {                         // rows of the block itself: L11 (selects, not a dynamic register index)
                    const int i = row - k0;
                    double li[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) li[j] = i == 0 ? l[0][j] : (i == 1 ? l[1][j] : (i == 2 ? l[2][j] : l[3][j]));
                    n0 = second ? li[2] : li[0];
                    n1 = second ? li[3] : li[1];
                }
                spanel[(second ? 2 : 0) * LDS_ + row] = n0;
                spanel[(second ? 3 : 1) * LDS_ + row] = n1;
                if (row >= k0) {
#pragma unroll
                    for (int cb = 0; cb < 8; ++cb)
                        if (cb == cb0) { a[cb][0] = n0; a[cb][1] = n1; }
                }
            }
        }
#ifndef ASVGP_ABL_W
        if (warp == cb0) {
            // rows k0..k0+3 of V -> rows of L^-1: W = L11^-1 V_rows, column by column (this lane's columns of every block <= cb0)
            const int base = (4 * h) << 2 | fq;
            const int i_own = fm - 4 * h;
#pragma unroll
            for (int cb = 0; cb < 8; ++cb) {
                if (cb <= cb0) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double t0 = __shfl_sync(0xffffffffu, v[cb][e], base), t1 = __shfl_sync(0xffffffffu, v[cb][e], base + 4);
                        const double t2 = __shfl_sync(0xffffffffu, v[cb][e], base + 8), t3 = __shfl_sync(0xffffffffu, v[cb][e], base + 12);
                        const double w0 = t0 * r[0];
                        const double w1 = fma(-l[1][0], w0, t1) * r[1];
                        const double w2 = fma(-l[2][1], w1, fma(-l[2][0], w0, t2)) * r[2];
                        const double w3 = fma(-l[3][2], w2, fma(-l[3][1], w1, fma(-l[3][0], w0, t3))) * r[3];
                        if (i_own >= 0 && i_own < 4) v[cb][e] = i_own == 0 ? w0 : (i_own == 1 ? w1 : (i_own == 2 ? w2 : w3));
                        if (fm == 0) {
                            const int col = cb * 8 + fq * 2 + e;
                            swrow[0 * LDS_ + col] = w0; swrow[1 * LDS_ + col] = w1; swrow[2 * LDS_ + col] = w2; swrow[3 * LDS_ + col] = w3;
                        }
                    }
                }
            }
        }
#endif
        __syncthreads();
        // ---- phase 3: rank-4 updates on the tensor cores (rows below the block only) --------------------------------------------
#ifndef ASVGP_ABL_UPD
        if (warp >= cb0) {
            const double am = (row > k0 + 3) ? -spanel[fq * LDS_ + row] : 0.0;          // A fragment: (m = fm, k = fq)
#pragma unroll
            for (int cb = 0; cb < 8; ++cb) {
                if (cb > cb0) {
                    dmma_m8n8k4(a[cb], am, spanel[fq * LDS_ + cb * 8 + fm]);
                } else if (cb == cb0 && h == 0) {                                       // columns k0+4..k0+7 of the block only
                    double t[2] = {0.0, 0.0};
                    dmma_m8n8k4(t, am, spanel[fq * LDS_ + cb * 8 + fm]);
                    if (fq >= 2) { a[cb][0] += t[0]; a[cb][1] += t[1]; }
                }
                if (cb <= cb0) dmma_m8n8k4(v[cb], am, swrow[fq * LDS_ + cb * 8 + fm]);
            }
        }
#endif
    }
    __syncthreads();
    frags_to_tile(a, scratch, warp, lane);
    __syncthreads();
    regs_from_tile(acc, scratch, tm, tn);
    __syncthreads();
    frags_to_tile(v, scratch, warp, lane);
    __syncthreads();
    regs_from_tile(V, scratch, tm, tn);
    __syncthreads();
}

}  // namespace asvgp

"""Extended-precision (x87 80-bit long double, 64-bit mantissa) evaluation of the Kronecker bound — a checker OF THE
ORACLE, used once at the 200 x 200 fixture where the fp64 LAPACK oracle and the GPU differ by a few 1e-10 relative and
the question is which one carries the rounding error.  TEST INFRASTRUCTURE; minutes of CPU.

    python -m oracle.extended_check        # prints the bound of tests/golden/scale_kron_c4's case term by term
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

LD = np.longdouble


def cholesky_band_ld(band):
    """Lower band (w+1, M) of the Cholesky factor in long double, right-looking, one rank-1 update of the w x w window
    per column (strided parallelogram view of the band; w scratch rows above it absorb the unused upper triangle)."""
    w, M = band.shape[0] - 1, band.shape[1]
    ab = np.zeros((2 * w + 1, M + w), dtype=LD)
    ab[w:, :M] = band
    ab[w, M:] = 1.0
    s = ab.strides[1]
    rs = ab.strides[0]
    for j in range(M):
        d = np.sqrt(ab[w, j])
        ab[w:, j] /= d
        l = ab[w + 1:, j].copy()                   # rows j+1 .. j+w
        V = np.lib.stride_tricks.as_strided(ab[w:, j + 1:], shape=(w, w), strides=(rs, s - rs))   # V[a, b] = A[j+1+a, j+1+b]
        V -= np.multiply.outer(l, l)
    return ab[w:, :M]


def solve_lower_band_ld(L, b):
    w, M = L.shape[0] - 1, L.shape[1]
    y = np.zeros(M + w, dtype=LD)
    y[:M] = b
    for j in range(M):
        y[j] /= L[0, j]
        y[j + 1: j + w + 1] -= L[1:, j] * y[j]
    return y[:M]


def dense_from_band_ld(band):
    k, m = band.shape[0] - 1, band.shape[1]
    A = np.zeros((m, m), dtype=LD)
    for d in range(k + 1):
        i = np.arange(m - d)
        A[i + d, i] = band[d, : m - d]
        A[i, i + d] = band[d, : m - d]
    return A


def chol_dense_ld(A):
    A = A.copy()
    m = A.shape[0]
    for j in range(m):
        A[j, j] = np.sqrt(A[j, j])
        A[j + 1:, j] /= A[j, j]
        A[j + 1:, j + 1:] -= np.multiply.outer(A[j + 1:, j], A[j + 1:, j])
    return np.tril(A)


def inv_spd_dense_ld(A):
    L = chol_dense_ld(A)
    m = A.shape[0]
    Li = np.zeros_like(L)
    for c in range(m):                              # forward substitution, column by column
        e = np.zeros(m, dtype=LD); e[c] = 1
        for j in range(c, m):
            e[j] /= L[j, j]
            e[j + 1:] -= L[j + 1:, j] * e[j]
        Li[:, c] = e
    return Li.T @ Li, 2 * np.sum(np.log(np.diag(L)))


def main():
    import scale_cases as SC
    from oracle import asvgp_oracle as O

    c = SC.C4
    k, ms = c["order"], list(c["m"])
    X, y = SC.case_2d(c["raster"], c["seed"])
    meshes, deltas = zip(*[O.make_mesh(SC.DOM_2D[i][0], SC.DOM_2D[i][1], ms[i], k) for i in range(2)])
    G, b, yy = O.precompute_kron(meshes, deltas, k, ms, X, y)
    n = X.shape[0]
    T = [O.static_bands(k, m, d) for m, d in zip(ms, deltas)]
    (v1, l1), (v2, l2) = c["hypers"]
    s2 = c["sigma2"]
    Ks = [O.make_Kuu("Matern32", l1, v1, T[0]), O.make_Kuu("Matern32", l2, v2, T[1])]
    t0 = time.time()
    Pb = O.kron_band(Ks, G, s2, k, ms)             # fp64 assembly (one rounding per entry, like every implementation)
    L = cholesky_band_ld(Pb.astype(LD))
    logdetP = 2 * np.sum(np.log(L[0]))
    cvec = solve_lower_band_ld(L, b[:, 0].astype(LD)) / LD(s2)
    quad = np.sum(cvec * cvec)
    print("band cholesky in long double: %.0f s" % (time.time() - t0))
    m1, m2 = ms
    Kinv, ldK = zip(*[inv_spd_dense_ld(dense_from_band_ld(kb.astype(LD))) for kb in Ks])
    logdetK = m2 * ldK[0] + m1 * ldK[1]
    Gc = G.tocoo()
    tr = np.sum(Gc.data.astype(LD) * Kinv[0][Gc.row // m2, Gc.col // m2] * Kinv[1][Gc.row % m2, Gc.col % m2])
    yy_ld = np.sum(np.square(y.astype(LD)))
    elbo = (-LD(0.5) * n * np.log(2 * LD(np.pi) * LD(s2)) - LD(0.5) * logdetP + LD(0.5) * logdetK - LD(0.5) * yy_ld / LD(s2)
            + LD(0.5) * quad - LD(0.5) * n * LD(v1) * LD(v2) / LD(s2) + LD(0.5) * tr / LD(s2))
    print("long double: elbo %.12f  logdetP %.12f  logdetK %.12f  quad*s2^2 %.12f  trace %.12f"
          % (elbo, logdetP, logdetK, quad * LD(s2) ** 2, tr))
    e64, terms = None, None
    e64 = O.elbo_kron_banded(Ks, G, b, yy, n, [v1, v2], s2, k, ms)
    print("fp64 LAPACK oracle: elbo %.12f   (difference %.3e relative)" % (e64, abs(e64 - float(elbo)) / abs(float(elbo))))


if __name__ == "__main__":
    main()

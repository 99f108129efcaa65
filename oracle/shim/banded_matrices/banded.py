"""Stand-ins for the eight `banded_matrices.banded` ops the reference calls (gpr.py:56-75,185,215;
utils.py:8,37,40-55; kronecker.py:37).  TEST INFRASTRUCTURE ONLY — forward values, no gradients.

Layout (validated end-to-end by the reference notebook's stored ELBO): a matrix with
lower/upper bandwidth (l, u) is a dense (l+u+1) x n array with band[u + i - j, j] = A[i, j]
(LAPACK/SciPy `ab`), unused tail entries zero; symmetric matrices travel as lower bands (l=k, u=0).
"""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

from tensorflow._core import _t


def _to_sparse(band, l, u):
    band = np.asarray(band, dtype=np.float64)
    n = band.shape[1]
    A = sp.lil_matrix((n, n))
    for r in range(l + u + 1):
        off = r - u  # i - j
        if off >= 0:
            cols = np.arange(0, n - off)
        else:
            cols = np.arange(-off, n)
        A[cols + off, cols] = band[r, cols]
    return A.tocsr()


def unpack_banded_matrix_to_dense(band, lower_bandwidth, upper_bandwidth):
    return _t(_to_sparse(band, int(lower_bandwidth), int(upper_bandwidth)).toarray())


def pack_dense_matrix_to_banded(dense, lower_bandwidth, upper_bandwidth):
    dense = np.asarray(dense)
    l, u = int(lower_bandwidth), int(upper_bandwidth)
    n = dense.shape[0]
    band = np.zeros((l + u + 1, n))
    for r in range(l + u + 1):
        off = r - u
        d = np.diagonal(dense, offset=-off)
        if off >= 0:
            band[r, : n - off] = d
        else:
            band[r, -off:] = d
    return _t(band)


def transpose_band(band, lower_bandwidth, upper_bandwidth):
    l, u = int(lower_bandwidth), int(upper_bandwidth)
    A = _to_sparse(band, l, u).T.toarray()
    return pack_dense_matrix_to_banded(A, u, l)


def symmetrise_band(lower, lower_bandwidth):
    k = int(lower_bandwidth)
    up = np.asarray(transpose_band(lower, k, 0))
    return _t(np.concatenate([up[:-1, :], np.asarray(lower)], axis=0))


def cholesky_band(band):
    return _t(sla.cholesky_banded(np.asarray(band, dtype=np.float64), lower=True))


def inverse_from_cholesky_band(L_band):
    """Lower band of (L L^T)^-1 (the Takahashi / sparse-inverse subset), via a dense inverse."""
    L_band = np.asarray(L_band)
    k = L_band.shape[0] - 1
    L = np.asarray(unpack_banded_matrix_to_dense(L_band, k, 0))
    Linv = sla.solve_triangular(L, np.eye(L.shape[0]), lower=True)
    return pack_dense_matrix_to_banded(Linv.T @ Linv, k, 0)


def product_band_band(A, B, left_lower_bandwidth, left_upper_bandwidth, right_lower_bandwidth,
                      right_upper_bandwidth, result_lower_bandwidth, result_upper_bandwidth):
    P = (_to_sparse(A, left_lower_bandwidth, left_upper_bandwidth)
         @ _to_sparse(B, right_lower_bandwidth, right_upper_bandwidth)).toarray()
    return pack_dense_matrix_to_banded(P, result_lower_bandwidth, result_upper_bandwidth)


def solve_triang_mat(L_band, rhs):
    L_band = np.asarray(L_band)
    k = L_band.shape[0] - 1
    return _t(sla.solve_banded((k, 0), L_band, np.asarray(rhs)))

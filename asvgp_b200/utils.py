"""Band helpers with the names of reference asvgp/utils.py:7-57 (host-side conveniences, numpy).

Layout everywhere: lower band (k+1) x n with band[d, j] = A[j+d, j] and d trailing zeros in row d."""
import numpy as np
import scipy.sparse as sparse


def symmetrise_banded(K_lower):
    """(k+1) x n lower band -> (2k+1) x n full band, row r = diagonal offset r-k (reference utils.py:7-9)."""
    K_lower = np.asarray(K_lower)
    k, n = K_lower.shape[0] - 1, K_lower.shape[1]
    upper = np.zeros((k, n))
    for d in range(1, k + 1):
        upper[k - d, d:] = K_lower[d, : n - d]
    return np.concatenate([upper, K_lower], axis=0)


def sparse_to_band(K_sparse, bandwidth):
    """Lower band of a (symmetric) sparse matrix (reference utils.py:24-30)."""
    n = K_sparse.shape[0]
    band = np.zeros((bandwidth + 1, n))
    for d in range(bandwidth + 1):
        band[d, : n - d] = K_sparse.diagonal(k=-d)
    return band


def band_to_sparse(K_lower):
    """Lower-triangular sparse matrix from a lower band (reference utils.py:32-33; CHOLMOD reads only the lower
    triangle)."""
    K_lower = np.asarray(K_lower)
    return sparse.spdiags(K_lower, np.arange(0, -(K_lower.shape[0]), -1), K_lower.shape[1], K_lower.shape[1])


def band_to_dense(K_lower):
    """Dense symmetric matrix from a lower band."""
    K_lower = np.asarray(K_lower)
    k, n = K_lower.shape[0] - 1, K_lower.shape[1]
    A = np.zeros((n, n))
    for d in range(k + 1):
        i = np.arange(n - d)
        A[i + d, i] = K_lower[d, : n - d]
        A[i, i + d] = K_lower[d, : n - d]
    return A


def bands_to_sparse(K_bands, mat_bandwidth):
    """Sparse Kronecker product of two banded factors (reference utils.py:53-57, d = 2 only)."""
    Ks = [sparse.csc_matrix(band_to_dense(k)) for k in K_bands]
    return sparse.kron(Ks[0], Ks[1])

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


def bands_to_kron_cholesky(K_bands, mat_bandwidth=None):
    """(K1 (x) K2 dense, L1 (x) L2 dense) from two lower bands (reference utils.py:45-51; `mat_bandwidth` is accepted
    for signature parity — the reference uses it only to unpack the bands).  Dense (m1 m2)^2 outputs: small models only,
    as in the reference; `GPR_kron` itself never forms them.  GPU: asvgp_cholesky_dense + asvgp_kron_dense."""
    import torch

    from . import _lib, ops

    dense = [ops.to_device(band_to_dense(np.asarray(_np(k)))) for k in K_bands]
    if len(dense) != 2:
        raise NotImplementedError("d = 2 only (as reference utils.py:57)")
    facs = []
    for A in dense:
        L = torch.empty_like(A)
        info = torch.zeros(1, dtype=torch.float64, device=A.device)
        _lib.call("asvgp_cholesky_dense", ops._p(A), A.shape[0], ops._p(L), ops._p(info), ops._stream())
        if info.item() != 0:
            raise np.linalg.LinAlgError("Cholesky failed: non-positive pivot %d" % int(info.item()))
        facs.append(L)
    m1, m2 = dense[0].shape[0], dense[1].shape[0]
    out = []
    for a, b in ((dense[0], dense[1]), (facs[0], facs[1])):
        o = torch.empty((m1 * m2, m1 * m2), dtype=torch.float64, device=a.device)
        _lib.call("asvgp_kron_dense", ops._p(a), m1, ops._p(b), m2, ops._p(o), ops._stream())
        out.append(o.cpu().numpy())
    return out[0], out[1]


def _np(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else a


def bands_to_sparse(K_bands, mat_bandwidth):
    """Sparse Kronecker product of two banded factors (reference utils.py:53-57, d = 2 only)."""
    Ks = [sparse.csc_matrix(band_to_dense(k)) for k in K_bands]
    return sparse.kron(Ks[0], Ks[1])


def stencil_to_sparse(Gs, m1, m2, order):
    """Full symmetric sparse matrix from the stencil layout of the Kronecker kernels: Gs[e, j] with
    e = d1*(2k+1) + (d2+k) holds A[(j1+d1, j2+d2), (j1, j2)], j = j1*m2 + j2 (include/asvgp_b200.h)."""
    Gs = np.asarray(Gs)
    k, M = order, m1 * m2
    j = np.arange(M)
    j1, j2 = j // m2, j % m2
    rows, cols, vals = [], [], []
    for d1 in range(k + 1):
        for d2 in range(-k, k + 1):
            if d1 == 0 and d2 < 0:
                continue
            ok = (j1 + d1 < m1) & (j2 + d2 >= 0) & (j2 + d2 < m2)
            i = (j1 + d1) * m2 + (j2 + d2)
            v = Gs[d1 * (2 * k + 1) + d2 + k]
            rows.append(i[ok]); cols.append(j[ok]); vals.append(v[ok])
            if d1 or d2:
                rows.append(j[ok]); cols.append(i[ok]); vals.append(v[ok])
    A = sparse.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(M, M))
    return A.tocsr()


def sparse_to_stencil(A, m1, m2, order):
    """Inverse of stencil_to_sparse: the lower part of a block-banded sparse matrix in stencil layout."""
    A = sparse.csr_matrix(A)
    k, M = order, m1 * m2
    out = np.zeros(((k + 1) * (2 * k + 1), M))
    j = np.arange(M)
    j1, j2 = j // m2, j % m2
    for d1 in range(k + 1):
        for d2 in range(-k, k + 1):
            if d1 == 0 and d2 < 0:
                continue
            ok = (j1 + d1 < m1) & (j2 + d2 >= 0) & (j2 + d2 < m2)
            i = (j1 + d1) * m2 + (j2 + d2)
            out[d1 * (2 * k + 1) + d2 + k, j[ok]] = np.asarray(A[i[ok], j[ok]]).ravel()
    return out

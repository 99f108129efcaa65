"""`asvgp_b200.kronecker` — the names of reference asvgp/kronecker.py:7-40, for scripts that call them directly.

The models never form the Khatri-Rao feature matrix (`GPR_kron` fuses it into the accumulate kernels,
asvgp_accum_2d); these helpers exist so that `from asvgp import kronecker as kron` switches unchanged.  The products
run on the GPU (asvgp_khatri_rao_csc); SciPy matrices are only the container the reference's return type asks for."""
from functools import reduce
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sparse
import torch

from . import _lib, ops


def make_kvs_two_sparse(A, B):
    """Row-wise Khatri-Rao product: out[ia * B.shape[0] + ib, n] = A[ia, n] * B[ib, n]  (reference kronecker.py:7-27,
    there as sparse_repeats(A) .multiply. sparse_tile(B))."""
    A, B = sparse.csc_matrix(A), sparse.csc_matrix(B)
    if A.shape[1] != B.shape[1]:
        raise ValueError("both feature matrices must have one column per datapoint")
    A.sort_indices(); B.sort_indices()
    n, mB = A.shape[1], B.shape[0]
    counts = np.diff(A.indptr).astype(np.int64) * np.diff(B.indptr).astype(np.int64)
    out_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    nnz = int(out_ptr[-1])
    dev = ops.device()
    i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(dev)      # noqa: E731
    f64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)    # noqa: E731
    pa, ia, va, pb, ib, vb, po = i64(A.indptr), i64(A.indices), f64(A.data), i64(B.indptr), i64(B.indices), f64(B.data), i64(out_ptr)
    rows = torch.empty(max(nnz, 1), dtype=torch.int64, device=dev)
    vals = torch.empty(max(nnz, 1), dtype=torch.float64, device=dev)
    _lib.call("asvgp_khatri_rao_csc", ops._p(pa), ops._p(ia), ops._p(va), ops._p(pb), ops._p(ib), ops._p(vb), n, mB,
              ops._p(po), ops._p(rows), ops._p(vals), ops._stream())
    out = sparse.csc_matrix((vals[:nnz].cpu().numpy(), rows[:nnz].cpu().numpy(), out_ptr), shape=(A.shape[0] * mB, n))
    return out.tocsr()


def make_kvs_sparse(A_list):
    """Khatri-Rao product of a list of per-dimension Kuf matrices, first dimension slowest (reference kronecker.py:29-33)."""
    return reduce(make_kvs_two_sparse, A_list)


def sparse_repeats(A, repeats):
    """Rows of A repeated `repeats` times each: out[i * repeats + r, n] = A[i, n]  (reference kronecker.py:7-15)."""
    return make_kvs_two_sparse(A, sparse.csc_matrix(np.ones((repeats, A.shape[1]))))


def sparse_tile(A, repeats):
    """A stacked `repeats` times: out[r * A.shape[0] + i, n] = A[i, n]  (reference kronecker.py:17-25)."""
    return make_kvs_two_sparse(sparse.csc_matrix(np.ones((repeats, A.shape[1]))), A)


def kron_log_determinant(Kuu, M, d):
    """log|K_1 (x) ... (x) K_d| = sum_i (M^d / M) log|K_i| for d banded factors of equal size M (what reference
    kronecker.py:35-40 means to compute; as written there it multiplies by a list and fails — SURVEY §2 row 7).
    Kuu: list of lower bands (k+1, M), numpy or CUDA tensors."""
    total = 0.0
    for band in Kuu:
        band = ops.to_device(band)
        k, m = band.shape[0] - 1, band.shape[1]
        assert m == M and len(Kuu) == d
        _, _, scal = ops.band_inverse_1d(band, None, SimpleNamespace(order=k, m=m))   # only order and m are read
        s = scal.cpu().numpy()
        if s[2] != 0:
            raise np.linalg.LinAlgError("banded Cholesky failed: non-positive pivot %d" % int(s[2]))
        total += (float(M) ** (d - 1)) * s[0]
    return total

"""`sksparse.cholmod.cholesky(A_csc, ordering_method="natural")` stand-in (dense LAPACK underneath).
Like CHOLMOD it reads only the LOWER triangle of A (the reference relies on that: utils.py:32-33 builds
lower-only matrices).  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp


class Factor:
    def __init__(self, A):
        A = sp.csc_matrix(A)
        low = sp.tril(A).toarray()
        full = low + low.T - np.diag(np.diag(low))
        self._L = np.linalg.cholesky(full)

    @staticmethod
    def _dense(B):
        return (B.toarray(), True) if sp.issparse(B) else (np.asarray(B), False)

    def solve_L(self, B, use_LDLt_decomposition=True):
        assert not use_LDLt_decomposition
        b, was_sparse = self._dense(B)
        x = sla.solve_triangular(self._L, b, lower=True)
        return sp.csc_matrix(x) if was_sparse else x

    def __call__(self, B):
        b, was_sparse = self._dense(B)
        x = sla.cho_solve((self._L, True), b)
        return sp.csc_matrix(x) if was_sparse else x


def cholesky(A, ordering_method="natural"):
    assert ordering_method == "natural"
    return Factor(A)

"""tf.linalg.* stand-ins used by reference utils.py:45-57 and gpr.py:185-334 (TEST INFRASTRUCTURE ONLY)."""
import numpy as np
import scipy.linalg as sla

from ._core import _t


def diag(diagonal, k=0):
    return _t(np.diag(np.asarray(diagonal), k=k))


def diag_part(x, k=0):
    return _t(np.diagonal(np.asarray(x), offset=k).copy())


def cholesky(x):
    return _t(np.linalg.cholesky(np.asarray(x)))


def trace(x):
    return _t(np.trace(np.asarray(x)))


def triangular_solve(matrix, rhs, lower=True):
    return _t(sla.solve_triangular(np.asarray(matrix), np.asarray(rhs), lower=lower))


def cholesky_solve(chol, rhs):
    return _t(sla.cho_solve((np.asarray(chol), True), np.asarray(rhs)))


class LinearOperatorFullMatrix:
    def __init__(self, matrix):
        self.matrix = np.asarray(matrix)

    def to_dense(self):
        return _t(self.matrix)


class LinearOperatorKronecker:
    def __init__(self, operators):
        self.operators = operators

    def to_dense(self):
        out = np.ones((1, 1))
        for op in self.operators:
            out = np.kron(out, np.asarray(op.to_dense()))
        return _t(out)


class LinearOperatorBlockDiag:
    def __init__(self, operators):
        self.operators = operators

    def to_dense(self):
        return _t(sla.block_diag(*[np.asarray(op.to_dense()) for op in self.operators]))

    def solve(self, rhs):
        return _t(np.linalg.solve(np.asarray(self.to_dense()), np.asarray(rhs)))

    def log_abs_determinant(self):
        return _t(np.linalg.slogdet(np.asarray(self.to_dense()))[1])

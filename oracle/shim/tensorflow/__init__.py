"""numpy stand-in for the ~45 `tf.*` symbols that /root/reference/asvgp/*.py touches.

TEST INFRASTRUCTURE ONLY.  This package exists so that the *unmodified* reference files
(`/root/reference/asvgp/{basis,inducing_features,gpr,utils,kronecker}.py`) can be imported in
the build container to generate golden vectors (see `oracle/make_golden.py`).  Nothing under
`asvgp_b200/` imports it.  It provides forward values only (no autodiff).

Symbol inventory: `grep -oE "tf\\.[A-Za-z_.]+" /root/reference/asvgp/*.py`.
"""
import numpy as np

from ._core import Tensor, _t
from . import linalg, math, nn  # noqa: F401,E402

float64 = np.float64
float32 = np.float32
int64 = np.int64
int32 = np.int32


def _dt(dtype):
    return None if dtype is None else np.dtype(dtype)


def cast(x, dtype=None):
    return _t(np.asarray(x).astype(_dt(dtype)))


def constant(x, dtype=None):
    return _t(np.array(np.asarray(x), dtype=_dt(dtype)))


def linspace(start, stop, num):
    """Emulates TF dtype inference (reference basis.py:17): Python floats become float32 tensors and
    the arithmetic runs in float32; Python ints give float64 [upstream-memory, SURVEY Q1]."""
    if isinstance(start, (int, np.integer)) and isinstance(stop, (int, np.integer)):
        return _t(np.linspace(float(start), float(stop), int(num), dtype=np.float64))
    f = np.float32
    a, b = f(start), f(stop)
    n = int(num)
    step = f(f(b - a) / f(n - 1))
    out = np.empty(n, dtype=f)
    out[0] = a
    out[-1] = b
    j = np.arange(1, n - 1, dtype=f)
    out[1:-1] = (a + step * j).astype(f)
    return _t(out)


def repeat(x, repeats, axis=None):
    return _t(np.repeat(np.asarray(x), int(repeats), axis=axis))


def zeros(shape, dtype=np.float32):
    if isinstance(shape, (int, np.integer)):
        shape = (int(shape),)
    return _t(np.zeros(tuple(int(s) for s in shape), dtype=_dt(dtype)))


def ones(shape, dtype=np.float32):
    if isinstance(shape, (int, np.integer)):
        shape = (int(shape),)
    return _t(np.ones(tuple(int(s) for s in np.atleast_1d(shape)), dtype=_dt(dtype)))


def concat(values, axis=0):
    vals = [np.atleast_1d(np.asarray(v)) for v in values]
    return _t(np.concatenate(vals, axis=axis))


def stack(values, axis=0):
    return _t(np.stack([np.asarray(v) for v in values], axis=axis))


def searchsorted(sorted_sequence, values, side="left"):
    return _t(np.searchsorted(np.asarray(sorted_sequence), np.asarray(values), side=side).astype(np.int64))


def gather(params, indices):
    return _t(np.asarray(params)[np.asarray(indices)])


def tile(x, multiples):
    return _t(np.tile(np.asarray(x), tuple(int(m) for m in np.atleast_1d(np.asarray(multiples)))))


def range(*args):  # noqa: A001
    return _t(np.arange(*[int(a) for a in args], dtype=np.int64))


def reshape(x, shape):
    return _t(np.reshape(np.asarray(x), shape))


def transpose(x):
    return _t(np.asarray(x).T)


def reduce_sum(x, axis=None):
    return _t(np.sum(np.asarray(x), axis=axis))


def square(x):
    return _t(np.square(np.asarray(x)))


def shape(x):
    return _t(np.array(np.shape(x), dtype=np.int64))


def size(x):
    return _t(np.array(np.size(x), dtype=np.int64))


def multiply(a, b):
    return _t(np.multiply(np.asarray(a), np.asarray(b)))


def add(a, b):
    return _t(np.add(np.asarray(a), np.asarray(b)))


def matmul(a, b):
    return _t(np.asarray(a) @ np.asarray(b))


def expand_dims(x, axis):
    return _t(np.expand_dims(np.asarray(x), axis))


def reverse(x, axis):
    return _t(np.flip(np.asarray(x), axis=tuple(axis)))


def scatter_nd(indices, updates, shape):
    out = np.zeros(tuple(int(s) for s in np.asarray(shape)), dtype=np.asarray(updates).dtype)
    idx = np.asarray(indices)
    np.add.at(out, tuple(idx[:, i] for i in np.arange(idx.shape[1])), np.asarray(updates))
    return _t(out)

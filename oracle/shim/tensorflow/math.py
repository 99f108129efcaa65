"""tf.math.* stand-ins (TEST INFRASTRUCTURE ONLY)."""
import numpy as np

from ._core import _t


def cumsum(x, axis=0):
    return _t(np.cumsum(np.asarray(x), axis=axis))


def reduce_sum(x, axis=None):
    return _t(np.sum(np.asarray(x), axis=axis))


def log(x):
    return _t(np.log(np.asarray(x)))


def square(x):
    return _t(np.square(np.asarray(x)))

"""tf.nn.* stand-ins (TEST INFRASTRUCTURE ONLY)."""
import numpy as np

from ._core import _t


def relu(x):
    return _t(np.maximum(np.asarray(x), 0))

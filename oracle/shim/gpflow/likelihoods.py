"""gpflow.likelihoods.Gaussian stand-in (TEST ONLY)."""
import numpy as np

from tensorflow._core import _t


class Gaussian:
    def __init__(self, variance=1.0):
        self.variance = _t(np.float64(variance))

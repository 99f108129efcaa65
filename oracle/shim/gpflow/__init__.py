"""Minimal stand-in for the eight GPflow 2.5.2 names the reference touches:
kernels.Matern12/32/52, likelihoods.Gaussian, mean_functions.Zero, models.GPModel,
models.InternalDataTrainingLossMixin, optimizers.Scipy.  TEST INFRASTRUCTURE ONLY.
Parameters are plain float64 0-d arrays (with `.numpy()`); GPflow's defaults are all 1.0."""
from . import kernels, likelihoods, mean_functions, models  # noqa: F401

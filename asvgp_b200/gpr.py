"""Models `GPR_1d` and `GPR_kron` — same constructor, attributes and methods as reference asvgp/gpr.py:18-136 and
:239-359, computed by the sm_100a kernels of libasvgp_sm100a (no CPU fallback).

What differs from the reference is only *how*:
  * the O(N) precompute never materialises Kuf (fused accumulate kernel, asvgp_accum_1d);
  * the per-step banded algebra (2 Choleskys, Takahashi inverse, triangular solve, gpr.py:55-75) runs as partitioned
    sweeps on the GPU and returns the hyper-parameter gradients with the ELBO (the reference gets them from TF
    reverse mode when `gpflow.optimizers.Scipy` asks, example.py:31-32);
  * prediction is O(n*) instead of O(n* M): only the (k+1) window entries of alpha and of
    band(P^-1) - band(Kuu^-1) matter (gpr.py:103-118);
  * with torch.distributed initialised the data passed in are this rank's shard and the packed accumulator is
    all-reduced once.
"""
import numpy as np
import torch

from . import dist as _dist
from . import ops
from .inducing_features import SplineFeatures1D
from .kernels import Gaussian, Parameter, hyper_value, kernel_kind


class _ModelBase:
    """The slice of gpflow.models.GPModel + InternalDataTrainingLossMixin the reference's scripts use."""

    def maximum_log_likelihood_objective(self):
        return self.elbo()

    def training_loss(self):
        return -self.maximum_log_likelihood_objective()

    def training_loss_and_gradients(self):
        """(-ELBO, d(-ELBO)/d unconstrained variables) in the order of `trainable_variables`."""
        elbo, grads = self.elbo_and_grad()
        g = [-(grads[id(p)] * p.dvalue_dunconstrained()) for p in self.trainable_variables]
        return -elbo, np.array(g, dtype=np.float64)

    def predict_log_density(self, data):
        """log N(y | mean, var + sigma2) per test point (used by reference electricity.py:138)."""
        X, y = data
        mean, var = self.predict_f(X)
        mean, var = np.asarray(_to_numpy(mean)), np.asarray(_to_numpy(var))
        s2 = var + hyper_value(self.likelihood.variance)
        y = np.asarray(_to_numpy(y)).reshape(mean.shape)
        return (-0.5 * np.log(2 * np.pi * s2) - 0.5 * (y - mean) ** 2 / s2).sum(-1)


def _to_numpy(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a


class GPR_1d(_ModelBase):
    """Collapsed-bound sparse GP regression in 1-D with B-spline inducing features (reference gpr.py:18-136)."""

    def __init__(self, data, kernel, basis, distributed="auto", check_inputs=True, chunks=0):
        # Check inputs (reference gpr.py:22-26)
        self._kind = kernel_kind(kernel)
        X, y = data
        assert X.shape[1] == 1
        if y.ndim == 1:
            y = y.reshape(-1, 1)
        if y.shape[1] != 1:
            raise NotImplementedError("multi-output y (D > 1) is not implemented yet")
        self.X, self.y = X, y
        on_device = isinstance(X, torch.Tensor) and X.is_cuda
        if check_inputs and X.shape[0]:
            lo, hi = (torch.aminmax(X) if isinstance(X, torch.Tensor) else (X.min(), X.max()))
            assert float(lo) > basis.a
            assert float(hi) < basis.b

        # Init model (reference gpr.py:29-34)
        self.kernel = kernel
        self.likelihood = Gaussian()
        self.mean_function = None
        self.num_latent_gps = 1
        self.basis = basis
        self.inducing_features = SplineFeatures1D(kernel, basis)
        self.bandwidth = self.basis.order
        self._chunks = chunks

        # Precompute static quantities (reference gpr.py:39-44): one fused pass over this rank's points,
        # then (multi-GPU) one all-reduce of the packed buffer.
        if on_device:
            self._acc = ops.accum_1d(X.reshape(-1), y.reshape(-1), basis)
        else:                       # host data: streamed through pinned staging buffers, never fully resident
            self._acc = ops.accum_1d_host(X, y, basis)
        self._distributed = _dist.is_distributed(distributed)
        if self._distributed:
            _dist.allreduce_packed(self._acc)
        self._G, self._b, self._scal = ops.split_accum_1d(self._acc, basis)
        self._host = None
        self._out = torch.empty(16, dtype=torch.float64, device=self._acc.device)

    # -- the reference's cached attributes (host copies, fetched lazily) --------------------------------------
    def _host_stats(self):
        if self._host is None:
            self._host = (self._G.cpu().numpy(), self._b.cpu().numpy().reshape(-1, 1), self._scal.cpu().numpy())
        return self._host

    @property
    def KufKfu(self):
        return self._host_stats()[0]

    @property
    def Kuf_y(self):
        return self._host_stats()[1]

    @property
    def tr_yTy(self):
        return float(self._host_stats()[2][0])

    @property
    def num_data(self):
        """Global number of datapoints (all shards)."""
        return int(self._host_stats()[2][1])

    @property
    def KufKfu_sparse(self):
        from . import utils

        full = utils.band_to_sparse(self.KufKfu)
        return full + full.T - __import__("scipy.sparse", fromlist=["diags"]).diags(self.KufKfu[0])

    @property
    def trainable_variables(self):
        return [self.kernel.variance, self.kernel.lengthscales, self.likelihood.variance]

    # -- objective ---------------------------------------------------------------------------------------------
    def _launch_elbo(self):
        var = hyper_value(self.kernel.variance)
        s2 = hyper_value(self.likelihood.variance)
        Kuu, dKuu = self.inducing_features.make_Kuu_device(self.kernel, want_grad=True)
        ops.elbo_grad_1d(Kuu, dKuu, self._acc, self.basis, var, s2, chunks=self._chunks, out=self._out)
        return self._out

    def elbo_and_grad(self):
        """ELBO (reference gpr.py:49-89) and {id(param): dELBO/dparam} for variance, lengthscales, sigma^2."""
        out = self._launch_elbo().cpu().numpy()
        if out[8] != 0:
            raise np.linalg.LinAlgError("banded Cholesky failed: non-positive pivot %d" % int(out[8]))
        grads = {id(self.kernel.variance): out[1], id(self.kernel.lengthscales): out[2],
                 id(self.likelihood.variance): out[3]}
        self.last_terms = dict(log_det_Kuu=out[4], log_det_P=out[5], quad=out[6], trace=out[7])
        return float(out[0]), grads

    def elbo(self):
        """Variational bound on the log marginal likelihood (reference gpr.py:49-89)."""
        return np.float64(self.elbo_and_grad()[0])

    # -- prediction ----------------------------------------------------------------------------------------------
    def posterior_weights(self):
        """(alpha, S) on the device: alpha = P^-1 Kuf_y / sigma2, S = band(P^-1) - band(Kuu^-1)."""
        s2 = hyper_value(self.likelihood.variance)
        Kuu, _ = self.inducing_features.make_Kuu_device(self.kernel, want_grad=False)
        alpha, S, info = ops.posterior_1d(Kuu, self._acc, self.basis, s2, chunks=self._chunks)
        return alpha, S, info

    def predict_f(self, Xnew, full_cov=False, full_output_cov=False, batch=False):
        """Posterior mean and variance at Xnew, each (n*, 1) (reference gpr.py:91-136).  numpy in -> numpy out,
        CUDA tensor in -> CUDA tensors out.  `batch` is accepted for compatibility; no chunking is needed (and the
        reference's silent drop of the last n* mod 10000 points, SURVEY Q6, is not reproduced)."""
        assert not full_output_cov
        if full_cov:
            raise NotImplementedError
        alpha, S, info = self.posterior_weights()
        xs = ops.to_device(Xnew).reshape(-1)
        mean, var = ops.predict_1d(xs, self.basis, alpha, S, hyper_value(self.kernel.variance))
        if info.any().item():
            raise np.linalg.LinAlgError("banded Cholesky failed in predict_f")
        mean, var = mean.view(-1, 1), var.view(-1, 1)
        if isinstance(Xnew, torch.Tensor) and Xnew.is_cuda:
            return mean, var
        return mean.cpu().numpy(), var.cpu().numpy()

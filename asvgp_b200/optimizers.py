"""`Scipy` — the slice of `gpflow.optimizers.Scipy` the reference uses (experiments/snelson/example.py:31-32):
L-BFGS-B over the unconstrained (softplus-transformed) variables, loss and gradient from one GPU evaluation."""
import numpy as np
import scipy.optimize


def _same_on_every_rank(u):
    """Data-parallel runs: every rank drives its own L-BFGS on (up to rounding) the same loss and gradient; rank 0's iterate
    is broadcast before each evaluation so that the replicas can never drift apart by an ulp and take different line-search
    branches (the factorisation uses fp64 REDs, whose summation order is not fixed)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return u
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.as_tensor(np.asarray(u, dtype=np.float64)).to(dev)
    dist.broadcast(t, src=0)
    return t.cpu().numpy()


class Scipy:
    def minimize(self, closure, variables, method="L-BFGS-B", **scipy_kwargs):
        model = getattr(closure, "__self__", None)
        if model is None or not hasattr(model, "training_loss_and_gradients"):
            raise TypeError("closure must be the bound `training_loss` of an asvgp_b200 model")
        variables = list(variables)
        model_vars = model.trainable_variables
        index = [next(i for i, p in enumerate(model_vars) if p is v) for v in variables]

        def fun(u):
            u = _same_on_every_rank(u)
            for v, ui in zip(variables, u):
                v.unconstrained = float(ui)
            loss, grad = model.training_loss_and_gradients()
            return loss, grad[index]

        u0 = np.array([v.unconstrained for v in variables], dtype=np.float64)
        res = scipy.optimize.minimize(fun, u0, jac=True, method=method, **scipy_kwargs)
        for v, ui in zip(variables, res.x):
            v.unconstrained = float(ui)
        return res

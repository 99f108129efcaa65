"""`Scipy` — the slice of `gpflow.optimizers.Scipy` the reference uses (experiments/snelson/example.py:31-32):
L-BFGS-B over the unconstrained (softplus-transformed) variables, loss and gradient from one GPU evaluation."""
import numpy as np
import scipy.optimize


class Scipy:
    def minimize(self, closure, variables, method="L-BFGS-B", **scipy_kwargs):
        model = getattr(closure, "__self__", None)
        if model is None or not hasattr(model, "training_loss_and_gradients"):
            raise TypeError("closure must be the bound `training_loss` of an asvgp_b200 model")
        variables = list(variables)
        model_vars = model.trainable_variables
        index = [next(i for i, p in enumerate(model_vars) if p is v) for v in variables]

        def fun(u):
            for v, ui in zip(variables, u):
                v.unconstrained = float(ui)
            loss, grad = model.training_loss_and_gradients()
            return loss, grad[index]

        u0 = np.array([v.unconstrained for v in variables], dtype=np.float64)
        res = scipy.optimize.minimize(fun, u0, jac=True, method=method, **scipy_kwargs)
        for v, ui in zip(variables, res.x):
            v.unconstrained = float(ui)
        return res

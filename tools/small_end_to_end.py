"""Small end-to-end run of every kernel family (1-D and 2-D models, every accumulate path, predictors) in a few seconds:
a quick "does everything still launch and finish" check on a GPU box.  (Written for compute-sanitizer memcheck, which is
closed on this pool; bounds are covered by the ragged / boundary-size parity tests instead.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from asvgp_b200 import basis as B, kernels as Kn, ops
from asvgp_b200.gpr import GPR_1d, GPR_kron

rng = np.random.default_rng(0)
# 1-D, sorted and shuffled, odd n, P > 1 chains
m, n = 400, 70_001
x = np.sort(rng.uniform(0, m, n)); y = np.sin(x / 9) + 0.1 * rng.standard_normal(n)
for xs in (x, rng.permutation(x)):
    mdl = GPR_1d((xs.reshape(-1, 1), y.reshape(-1, 1)), Kn.Matern52(), B.B3Spline(-1, m + 1, m))
    mdl.likelihood.variance.assign(0.3)
    print("1-D elbo", mdl.elbo_and_grad()[0])
    print("1-D predict", [a.shape for a in mdl.predict_f(np.linspace(1, m - 1, 1001).reshape(-1, 1))])
# 2-D: separable raster, curvilinear raster (odd row length), x1 runs, scattered
g1, g2 = np.sort(rng.uniform(0.02, 0.98, 90)), np.sort(rng.uniform(0.02, 1.98, 131))
Xs = np.stack(np.meshgrid(g1, g2, indexing="ij"), -1)
Xc = Xs.copy(); Xc[:, :, 1] += 0.001 * np.sin(np.arange(90))[:, None]
cases = {"separable": Xs.reshape(-1, 2), "curvilinear": Xc.reshape(-1, 2),
         "runs": Xs.reshape(-1, 2)[rng.uniform(size=90 * 131) > 0.1], "scattered": rng.permutation(Xs.reshape(-1, 2))}
for name, X in cases.items():
    yy = np.sin(5 * X[:, 0]) * np.cos(3 * X[:, 1]) + 0.05 * rng.standard_normal(X.shape[0])
    mk = GPR_kron((X, yy.reshape(-1, 1)), [Kn.Matern32(0.8, 0.3), Kn.Matern32(1.2, 0.5)], [B.B3Spline(0, 1, 20), B.B3Spline(0, 2, 24)])
    mk.likelihood.variance.assign(0.05)
    e, g = mk.elbo_and_grad()
    mu, var = mk.predict_f(X[:500])
    mu2, var2 = mk.predict_f(Xs.reshape(-1, 2))
    print("2-D", name, e, mu.shape, mu2.shape)
torch.cuda.synchronize()
print("small end-to-end run complete")

import os, sys, cProfile, pstats, time
sys.path.insert(0, os.getcwd())
import torch
from asvgp_b200 import basis as B, kernels as Kn, ops
from asvgp_b200.inducing_features import SplineFeatures1D
n, m = 100_000_000, 10000
b = B.B3Spline(-1, m + 1, m)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.sort(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * m).values.clamp_(1e-9, m - 1e-9)
y = torch.sin(x / 5)
kern = Kn.Matern52(variance=1.0, lengthscales=1.0)
feats = SplineFeatures1D(kern, b)
acc = torch.zeros(ops.accum_size_1d(b), dtype=torch.float64, device="cuda")
out = torch.empty(16, dtype=torch.float64, device="cuda")
def step():
    acc.zero_()
    Kuu, dKuu = feats.make_Kuu_device(kern, want_grad=True)
    kuu = ops.kuu_chain_1d(Kuu, dKuu, b, gate=True)
    ops.accum_1d(x, y, b, acc=acc)
    ops.elbo_grad_1d(Kuu, dKuu, acc, b, 1.0, 0.1, out=out, kuu=kuu)
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue per step: %.1f us" % ((t1 - t0) / 300 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(300): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)

"""Per-step device time of the first steps after a synchronize (1-D bench step): is there a ramp?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
for rep in range(2):
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
    evs[0].record()
    for i in range(40):
        step(); evs[i + 1].record()
    torch.cuda.synchronize()
    print("rep %d per-step us:" % rep, " ".join("%.0f" % (evs[i].elapsed_time(evs[i + 1]) * 1e3) for i in range(40)))

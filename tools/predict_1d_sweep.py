"""CUDA-event time of asvgp_predict_1d on 1e8 sorted points (M = 1e4, k = 3) — used with ASVGP_PRED_VARIANT / ASVGP_PRED_MULT
while tuning the kernel's launch shape."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asvgp_b200 import basis as B, ops
m, n = 10000, 100_000_000
b = B.B3Spline(-1, m + 1, m)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.sort(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * m).values.clamp_(1e-9, m - 1e-9)
alpha = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
S = torch.randn((4, m), dtype=torch.float64, device="cuda", generator=g) * 0.01
mean = torch.empty(n, dtype=torch.float64, device="cuda"); var = torch.empty_like(mean)
for _ in range(3): ops.predict_1d(x, b, alpha, S, 1.0, mean=mean, var=var)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.predict_1d(x, b, alpha, S, 1.0, mean=mean, var=var)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("variant %s mult %s: %.3f ms  %.0f GB/s  checksum %.6e" % (os.environ.get("ASVGP_PRED_VARIANT", "0"), os.environ.get("ASVGP_PRED_MULT", "4"), ms, 24e8 / ms / 1e6, float(mean[::1000003].sum() + var[::1000003].sum())))

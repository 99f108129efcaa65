"""CUDA-event time of asvgp_predict_2d_apply on a 1e4 x 1e4 raster (200 x 200 features) — launch-shape tuning."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asvgp_b200 import basis as B, ops
m, n1, n2 = 200, 10000, 10000
bases = [B.B3Spline(-80, -25, m), B.B3Spline(15, 55, m)]
x1 = torch.linspace(-75, -30, n1 + 2, dtype=torch.float64, device="cuda")[1:-1]
x2 = torch.linspace(20, 50, n2, dtype=torch.float64, device="cuda")
X = torch.stack([x1[:, None].expand(n1, n2), x2[None, :].expand(n1, n2)], -1).reshape(-1, 2).contiguous()
g = torch.Generator(device="cuda").manual_seed(1)
alpha = torch.randn(m * m, dtype=torch.float64, device="cuda", generator=g)
SigP = torch.randn((28, m * m), dtype=torch.float64, device="cuda", generator=g) * 1e-3
S1 = torch.randn((4, m), dtype=torch.float64, device="cuda", generator=g) * 1e-2
S2 = torch.randn((4, m), dtype=torch.float64, device="cuda", generator=g) * 1e-2
table = ops.predict_2d_prepare(bases, alpha, SigP, S1, S2)
mean = torch.empty(n1 * n2, dtype=torch.float64, device="cuda"); var = torch.empty_like(mean)
for _ in range(3): ops.predict_2d_apply(X, bases, table, 1.0, raster_row_len=n2, mean=mean, var=var)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.predict_2d_apply(X, bases, table, 1.0, raster_row_len=n2, mean=mean, var=var)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("pred2d mult %s: %.3f ms  %.0f GB/s  checksum %.6e" % (os.environ.get("ASVGP_PRED2D_MULT", "2"), ms, 32e8 / ms / 1e6, float(mean[::1000003].sum() + var[::1000003].sum())))

"""Runs only the binned 2-D accumulate on a shuffled raster (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvgp_b200 import basis as B, ops
n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
bases = [B.B3Spline(-80, -25, 200), B.B3Spline(15, 55, 200)]
x1 = torch.linspace(-75, -30, n1, dtype=torch.float64, device="cuda")
x2 = torch.linspace(20, 50, n1, dtype=torch.float64, device="cuda")
X = torch.stack([x1[:, None].expand(n1, n1), x2[None, :].expand(n1, n1)], -1).reshape(-1, 2)
g = torch.Generator(device="cuda"); g.manual_seed(1)
perm = torch.randperm(n1 * n1, device="cuda", generator=g)
X = X[perm].contiguous()
y = torch.sin(X[:, 0] / 4) * torch.cos(X[:, 1] / 3)
del perm
acc = torch.zeros(ops.accum_size_2d(bases), dtype=torch.float64, device="cuda")
cm = ops.moment_table_2d(bases)
for _ in range(2):
    ops.accum_2d(X, y, bases, cm, ops.split_accum_2d(acc, bases)[2], binned=True)
torch.cuda.synchronize()
print("done")

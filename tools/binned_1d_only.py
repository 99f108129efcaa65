"""Runs only the binned 1-D accumulate on shuffled input (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvgp_b200 import basis as B, ops
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
m = 10_000
b = B.B3Spline(-1, m + 1, m)
g = torch.Generator(device="cuda"); g.manual_seed(1997)
x = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * m
y = torch.sin(x / 37)
acc = torch.zeros(ops.accum_size_1d(b), dtype=torch.float64, device="cuda")
for _ in range(2):
    ops.accum_1d(x, y, b, acc, binned=True)
torch.cuda.synchronize()
print("done")

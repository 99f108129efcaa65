"""Runs only the 2-D accumulate at the C4 shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvgp_b200 import basis as B, ops
n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 200
k = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cls = getattr(B, "B%dSpline" % k)
bases = [cls(-80, -25, m), cls(15, 55, m)]
x1 = torch.linspace(-75, -30, n1, dtype=torch.float64, device="cuda")
x2 = torch.linspace(20, 50, n1, dtype=torch.float64, device="cuda")
X = torch.stack([x1[:, None].expand(n1, n1), x2[None, :].expand(n1, n1)], -1).reshape(-1, 2).contiguous()
y = torch.sin(X[:, 0] / 4) * torch.cos(X[:, 1] / 3)
acc = torch.zeros(ops.accum_size_2d(bases), dtype=torch.float64, device="cuda")
cm = ops.moment_table_2d(bases)
for _ in range(4):
    ops.accum_2d(X, y, bases, cm, ops.split_accum_2d(acc, bases)[2])
torch.cuda.synchronize()
print("done")

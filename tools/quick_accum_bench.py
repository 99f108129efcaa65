"""Scratch timing of accum_1d at the C3 shape (N=1e8 sorted, M=1e4, k=3); prints GB/s vs MEASURED_PEAKS."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvgp_b200 import basis as B, ops

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
k = 3
basis = B.B3Spline(-1, m + 1, m)
gen = torch.Generator(device="cuda").manual_seed(1997)
x = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) * m
y = torch.sin(x / 37.0) + 0.3 * torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
for order in ("sorted", "random"):
    if order == "sorted":
        xs, perm = torch.sort(x)
        ys = y[perm]
        del perm
    else:
        xs, ys = x, y
    acc = torch.zeros(ops.accum_size_1d(basis), dtype=torch.float64, device="cuda")
    for _ in range(3):
        acc.zero_(); ops.accum_1d(xs, ys, basis, acc=acc)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        acc.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.accum_1d(xs, ys, basis, acc=acc); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(json.dumps({"order": order, "n": n, "m": m, "ms": ms, "min_ms": min(ts), "GBps": 16 * n / ms / 1e6, "pts_per_s": n / ms * 1e3}))

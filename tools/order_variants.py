"""Accumulate timings for the secondary orderings of SURVEY §8(d): 1-D random order (C3 ii) and 2-D shuffled raster (C4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvgp_b200 import basis as B, ops

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
only_1d = len(sys.argv) > 2 and sys.argv[2] == "1d"
g = torch.Generator(device="cuda"); g.manual_seed(1997)
m = 10_000
b = B.B3Spline(-1, m + 1, m)
x = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * m
y = torch.sin(x / 37)
acc = torch.zeros(ops.accum_size_1d(b), dtype=torch.float64, device="cuda")
t_rand = timeit(lambda: ops.accum_1d(x, y, b, acc))
t_binned = timeit(lambda: ops.accum_1d(x, y, b, acc, binned=True))
t_auto = timeit(lambda: ops.accum_1d(x, y, b, acc, binned="auto"))
xs, order = torch.sort(x)
ys = y[order]
t_sorted = timeit(lambda: ops.accum_1d(xs, ys, b, acc))
t_sort = timeit(lambda: torch.sort(x), 2)
print("1-D n=%d: random order %.3f ms streaming, %.3f ms binned (%.3f ms with the order probe); sorted %.3f ms; torch.sort alone %.3f ms"
      % (n, t_rand, t_binned, t_auto, t_sorted, t_sort))
if only_1d:
    sys.exit(0)
del xs, ys, order, x, y

n1 = int(round(n ** 0.5))
bases = [B.B3Spline(-80, -25, 200), B.B3Spline(15, 55, 200)]
x1 = torch.linspace(-75, -30, n1, dtype=torch.float64, device="cuda")
x2 = torch.linspace(20, 50, n1, dtype=torch.float64, device="cuda")
X = torch.stack([x1[:, None].expand(n1, n1), x2[None, :].expand(n1, n1)], -1).reshape(-1, 2).contiguous()
yy = torch.sin(X[:, 0] / 4) * torch.cos(X[:, 1] / 3)
acc2 = torch.zeros(ops.accum_size_2d(bases), dtype=torch.float64, device="cuda")
cm = ops.moment_table_2d(bases)
mom = ops.split_accum_2d(acc2, bases)[2]
t_raster = timeit(lambda: ops.accum_2d(X, yy, bases, cm, mom))
perm = torch.randperm(n1 * n1, device="cuda", generator=g)
Xp = X[perm].contiguous(); yp = yy[perm].contiguous()
del X, yy, perm
t_shuf = timeit(lambda: ops.accum_2d(Xp, yp, bases, cm, mom), 2)
t_bin = timeit(lambda: ops.accum_2d(Xp, yp, bases, cm, mom, binned=True), 3)
t_auto = timeit(lambda: ops.accum_2d(Xp, yp, bases, cm, mom, binned="auto"), 3)
print("2-D n=%d: raster %.3f ms; shuffled %.3f ms streaming, %.3f ms binned (%.3f ms with the order probe)"
      % (n1 * n1, t_raster, t_shuf, t_bin, t_auto))

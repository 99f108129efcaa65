"""CUDA-event times of asvgp_accum_1d (1e8 sorted points, M = 1e4) and asvgp_accum_2d_raster (1e4 x 1e4 raster, 200 x 200) —
launch-shape tuning with ASVGP_ACC1D_MULT / ASVGP_ACC2D_TPW."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asvgp_b200 import basis as B, ops
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
which = sys.argv[1]
if which == "1d":
    m, n = 10000, 100_000_000
    b = B.B3Spline(-1, m + 1, m)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.sort(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * m).values.clamp_(1e-9, m - 1e-9)
    y = torch.sin(x)
    acc = torch.zeros(ops.accum_size_1d(b), dtype=torch.float64, device="cuda")
    ms = timeit(lambda: ops.accum_1d(x, y, b, acc=acc))
    print("acc1d mult %s: %.4f ms  %.0f GB/s  checksum %.10e" % (os.environ.get("ASVGP_ACC1D_MULT", "2"), ms, 16e8 / ms / 1e6, float(acc[:40000].sum()) / 13))
else:
    m, n1, n2 = 200, 10000, 10000
    bases = [B.B3Spline(-80, -25, m), B.B3Spline(15, 55, m)]
    x1 = torch.linspace(-75, -30, n1 + 2, dtype=torch.float64, device="cuda")[1:-1]
    x2 = torch.linspace(20, 50, n2, dtype=torch.float64, device="cuda")
    X = torch.stack([x1[:, None].expand(n1, n2), x2[None, :].expand(n1, n2)], -1).reshape(-1, 2).contiguous()
    y = torch.sin(X[:, 0]) * torch.cos(X[:, 1])
    acc = torch.zeros(ops.accum_size_2d(bases), dtype=torch.float64, device="cuda")
    cm = ops.moment_table_2d(bases)
    scal = ops.split_accum_2d(acc, bases)[2]
    ms = timeit(lambda: ops.accum_2d(X, y, bases, cm, scal, raster_row_len=n2))
    print("acc2d tpw %s: %.4f ms  %.0f GB/s  checksum %.10e" % (os.environ.get("ASVGP_ACC2D_TPW", "4"), ms, 24e8 / ms / 1e6, float(cm.sum()) / 13))

import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from asvgp_b200 import basis as B, ops
k, m1, m2, n = 3, 400, 12, 150000
rng = np.random.default_rng(n + 7 * k)
bases = [B.B3Spline(-80, -25, m1), B.B3Spline(15, 55, m2)]
X = np.stack([rng.uniform(-79.5, -25.5, n), rng.uniform(15.5, 54.5, n)], 1)
y = np.sin(X[:, 0] / 4.0) * np.cos(X[:, 1] / 3.0) + 0.05 * rng.standard_normal(n)
Xd, yd = ops.to_device(X), ops.to_device(y)
ref = None
bad = 0
from asvgp_b200 import _lib
nbytes = _lib.load().asvgp_accum_2d_binned_work_bytes(n)
for it in range(400):
    # poison the block the work buffer will be carved from (the suite's other tests leave arbitrary bytes there)
    junk = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device="cuda") if it % 2 else torch.full((nbytes,), 255, dtype=torch.uint8, device="cuda")
    del junk
    acc = torch.zeros(ops.accum_size_2d(bases), dtype=torch.float64, device="cuda")
    cm = ops.moment_table_2d(bases)
    ops.accum_2d(Xd, yd, bases, cm, ops.split_accum_2d(acc, bases)[2], binned=True)
    cmc = cm.clone()
    ops.expand_moments_2d(cm, bases, acc)
    a = acc.clone()
    if it % 3 == 0:   # churn the allocator / other kernels in between
        junk = torch.randn(1 << 22, device="cuda")
    if ref is None:
        ref, refcm = a, cmc
        continue
    d = (a - ref).abs().max().item(); dc = (cmc - refcm).abs().max().item()
    if d > 1e-9 * ref.abs().max().item() or dc > 1e-9 * refcm.abs().max().item():
        bad += 1
        print("iter", it, "acc diff", d, "moment diff", dc, "scal", a[-2:].tolist(), ref[-2:].tolist())
print("bad", bad, "of 399")

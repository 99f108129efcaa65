"""Phase timing of the 2-D Kronecker path at the C4 shape (N = n1*n2 raster, M = m x m, order k): accumulate, expand,
per-dimension band inverses, block-band factor, selected inverse, contractions, predictor."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvgp_b200 import basis as B, kernels as Kn, ops
from asvgp_b200.gpr import GPR_kron

n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 200
k = int(sys.argv[3]) if len(sys.argv) > 3 else 3
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
n2 = n1
cls = getattr(B, "B%dSpline" % k)
bases = [cls(-80, -25, m), cls(15, 55, m)]
x1 = torch.linspace(-75, -30, n1, dtype=torch.float64, device="cuda")
x2 = torch.linspace(20, 50, n2, dtype=torch.float64, device="cuda")
X = torch.stack([x1[:, None].expand(n1, n2), x2[None, :].expand(n1, n2)], -1).reshape(-1, 2).contiguous()
gen = torch.Generator(device="cuda").manual_seed(3)
y = torch.sin(X[:, 0] / 4) * torch.cos(X[:, 1] / 3) + 0.05 * torch.randn(X.shape[0], dtype=torch.float64, device="cuda", generator=gen)
n = X.shape[0]


def timed(fn, reps=reps):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


out = {"n": n, "m": m, "k": k}
acc = torch.zeros(ops.accum_size_2d(bases), dtype=torch.float64, device="cuda")
cm = ops.moment_table_2d(bases)
scal = ops.split_accum_2d(acc, bases)[2]
out["accum_ms"] = timed(lambda: ops.accum_2d(X, y, bases, cm, scal))
out["accum_GBps"] = 24 * n / out["accum_ms"] / 1e6
cm.zero_(); acc.zero_()
ops.accum_2d(X, y, bases, cm, scal)
out["expand_ms"] = timed(lambda: (acc[:-2].zero_(), ops.expand_moments_2d(cm, bases, acc)))
kerns = [Kn.Matern32(variance=1.0, lengthscales=5.0), Kn.Matern32(variance=1.0, lengthscales=4.0)]
model = GPR_kron.__new__(GPR_kron)
from asvgp_b200.inducing_features import SplineFeatures1D
model.kernels, model.bases, model.order = kerns, bases, k
model.inducing_features = [SplineFeatures1D(kerns[i], bases[i]) for i in range(2)]
out["factors_ms"] = timed(lambda: model._factors(True))
Ks, dKs, Ss, dSs, scals = model._factors(True)
ws = ops.kron_workspace(m, m, k, "band")
s2 = 0.01
out["kron_factor_ms"] = timed(lambda: ops.kron_factor(Ks[0], Ks[1], acc, bases, s2, ws))
M = m * m
w = k * (m + 1)
out["kron_factor_GFLOPs"] = M * w * w / out["kron_factor_ms"] / 1e6


def sel():
    ops.kron_factor(Ks[0], Ks[1], acc, bases, s2, ws)
    ops.kron_selinv(bases, ws)


out["factor_plus_selinv_ms"] = timed(sel)
SigP, x = ops.kron_selinv(bases, ws)
out["terms_ms"] = timed(lambda: ops.kron_terms(SigP, acc, x, Ks[0], dKs[0], Ks[1], dKs[1], Ss[0], dSs[0], Ss[1], dSs[1], bases, ws.terms))
alpha = x / s2
out["predict_ms"] = timed(lambda: ops.predict_2d(X, bases, alpha, SigP, Ss[0], Ss[1], 1.0))
out["predict_GBps"] = 32 * n / out["predict_ms"] / 1e6
print(json.dumps(out))

"""Critical-path breakdown of the tile-DAG factorisation from the %globaltimer stamps it leaves in the band buffer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from asvgp_b200 import basis as B, kernels as Kn, ops, _lib
from asvgp_b200.inducing_features import SplineFeatures1D

m = int(sys.argv[1]) if len(sys.argv) > 1 else 200
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cls = getattr(B, "B%dSpline" % k)
bases = [cls(-80, -25, m), cls(15, 55, m)]
n1 = 2000
x1 = torch.linspace(-75, -30, n1, dtype=torch.float64, device="cuda")
x2 = torch.linspace(20, 50, n1, dtype=torch.float64, device="cuda")
X = torch.stack([x1[:, None].expand(n1, n1), x2[None, :].expand(n1, n1)], -1).reshape(-1, 2).contiguous()
y = torch.sin(X[:, 0] / 4) * torch.cos(X[:, 1] / 3)
acc = torch.zeros(ops.accum_size_2d(bases), dtype=torch.float64, device="cuda")
cm = ops.moment_table_2d(bases)
ops.accum_2d(X, y, bases, cm, ops.split_accum_2d(acc, bases)[2]); ops.expand_moments_2d(cm, bases, acc)
kerns = [Kn.Matern32(variance=1.0, lengthscales=5.0), Kn.Matern32(variance=1.0, lengthscales=4.0)]
Ks = [SplineFeatures1D(kerns[i], bases[i]).make_Kuu_device(kerns[i])[0] for i in range(2)]
ws = ops.kron_workspace(m, m, k, "band")
for _ in range(3):
    ops.kron_factor(Ks[0], Ks[1], acc, bases, 0.01, ws)
torch.cuda.synchronize()
off = _lib.load().asvgp_kronband_colstat_offset(m, m, k)
nb = -(-m * m // 64)
NS = 20
st = ws.band[off: off + NS * nb].view(nb, NS).cpu().numpy()
t = st[:, 2:8]
prof = st[:, 8:12]
mid = slice(nb // 4, 3 * nb // 4)
def us(a): return float(np.median(a[mid])) / 1e3
print("block columns", nb, " total chain %.2f ms" % ((t[-1, 3] - t[0, 0]) / 1e6))
print("per block column (median, us): period %.2f" % us(np.diff(t[:, 3])))
print("  diag: begin->operands landed %.2f | potrf+inverse %.2f | publish %.2f" % (us(t[:, 1] - t[:, 0]), us(t[:, 2] - t[:, 1]), us(t[:, 3] - t[:, 2])))
print("  d=1 : diag published -> inverse seen %.2f | trsm tile done %.2f" % (us(t[:, 4] - t[:, 3]), us(t[:, 5] - t[:, 4])))
print("  next diag potrf start - d=1 tile published: %.2f" % us(t[1:, 1] - t[:-1, 5]))
print("  potrf (thread 0, SM cycles, median): 4x4 diag block %d | wait barrier 1 (sum of 16) %d | panel+barrier 2 %d | rank-4 updates %d" % tuple(np.median(prof[mid], axis=0)))

# selected inverse: stamps of the first sub-diagonal tile (R = C + 1) of every block column (columns run downwards)
ops.kron_selinv(bases, ws)
torch.cuda.synchronize()
st = ws.band[off: off + NS * nb].view(nb, NS).cpu().numpy()
s = st[:-1, 12:16]                      # the last block column has no sub-diagonal tile
w = st[:-1, 16:19]
print("selected inverse, per block column (median, us): period %.2f" % us(-np.diff(s[:, 3])))
print("  d=1 tile: Sigma(R,R) landed -> product + conversion done %.2f | published %.2f | own Y^T landed, contribution added, counted %.2f"
      % (us(s[:, 1] - s[:, 0]), us(s[:, 2] - s[:, 1]), us(s[:, 3] - s[:, 2])))
print("  contribution of tile (C+1, C) counted -> Sigma(C, C) seen landed by tile (C, C-1): %.2f" % us(s[:-1, 0] - s[1:, 3]))
print("  tile (C, C-1): starts waiting for Sigma(C,C) %.2f before the last contribution is counted | sees it complete %.2f after | copies issued +%.2f | landed +%.2f"
      % (us(s[1:, 3] - w[:-1, 0]), us(w[:-1, 1] - s[1:, 3]), us(w[:-1, 2] - w[:-1, 1]), us(s[:-1, 0] - w[:-1, 2])))

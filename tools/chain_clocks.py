"""Prints the SM-cycle breakdown of the ELBO chains kernel (diagnostic slots out[9..14])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asvgp_b200 import basis as B, kernels as Kn, ops
from asvgp_b200.inducing_features import SplineFeatures1D

m = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
n = 2_000_000
basis = B.B3Spline(-1, m + 1, m)
kern = Kn.Matern52()
x = torch.sort(torch.rand(n, dtype=torch.float64, device="cuda") * m).values.clamp_(1e-9, m - 1e-9)
y = torch.sin(x / 37.0)
acc = ops.accum_1d(x, y, basis)
feats = SplineFeatures1D(kern, basis)
Kuu, dKuu = feats.make_Kuu_device(kern)
for chunks in (0, 16, 24, 32, 47, 64, 96, 128):
    for _ in range(3):
        out = ops.elbo_grad_1d(Kuu, dKuu, acc, basis, 1.0, 0.1, chunks=chunks)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = ops.elbo_grad_1d(Kuu, dKuu, acc, basis, 1.0, 0.1, chunks=chunks)
    e1.record(); torch.cuda.synchronize()
    o = out.cpu().numpy()
    print("chunks=%3d  %.1f us/call  elbo=%.6f info=%d  Kuu-chain cycles: sweep=%d sep=%d back=%d trace=%d | P-chain: sweep=%d sep=%d"
          % (chunks, e0.elapsed_time(e1) * 100, o[0], o[8], o[9], o[10], o[11], o[12], o[13], o[14]))

"""Time of asvgp_band_inverse_1d at m = 200 (the per-dimension factor of the 2-D bench) for several chunk counts."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asvgp_b200 import basis as B, kernels as Kn, ops
from asvgp_b200.inducing_features import SplineFeatures1D
m = int(sys.argv[1]) if len(sys.argv) > 1 else 200
b = B.B3Spline(-80, -25, m)
K, dK = SplineFeatures1D(Kn.Matern32(lengthscales=5.0), b).make_Kuu_device(Kn.Matern32(lengthscales=5.0))
ref = None
for chunks in (1, 2, 4, 6, 8, 12, 16, 25):
    S, dS, sc = ops.band_inverse_1d(K, dK, b, chunks=chunks)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.band_inverse_1d(K, dK, b, chunks=chunks)
    e1.record(); torch.cuda.synchronize()
    if ref is None: ref = S.clone()
    print("chunks %2d: %.1f us  max rel diff vs chunks=1 %.2e" % (chunks, e0.elapsed_time(e1) / 20 * 1e3, float((S - ref).abs().max() / ref.abs().max())))

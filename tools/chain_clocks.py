"""SM-cycle split of the Kuu chain of asvgp_elbo_grad_1d (out[9..12]: chunk sweep, separator system, back sweep, trace
reduction) and CUDA-event time of the call, at the bench's 1-D size."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asvgp_b200 import basis as B, kernels as Kn, ops
from asvgp_b200.inducing_features import SplineFeatures1D
m = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 0
b = B.B3Spline(-1, m + 1, m)
kern = Kn.Matern52(variance=1.0, lengthscales=1.0)
n = 2_000_000
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.sort(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * m).values.clamp_(1e-9, m - 1e-9)
y = torch.sin(x / 5)
acc = ops.accum_1d(x, y, b)
feats = SplineFeatures1D(kern, b)
Kuu, dKuu = feats.make_Kuu_device(kern)
out = torch.empty(16, dtype=torch.float64, device="cuda")
for _ in range(3):
    ops.elbo_grad_1d(Kuu, dKuu, acc, b, 1.0, 0.1, chunks=chunks, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.elbo_grad_1d(Kuu, dKuu, acc, b, 1.0, 0.1, chunks=chunks, out=out)
e1.record(); torch.cuda.synchronize()
kuu = ops.kuu_chain_1d(Kuu, dKuu, b, chunks=chunks)
torch.cuda.current_stream().wait_event(kuu.event)
torch.cuda.synchronize()
t = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
t[0].record()
for _ in range(20):
    h = ops.kuu_chain_1d(Kuu, dKuu, b, chunks=chunks)
    torch.cuda.current_stream().wait_event(h.event)
t[1].record()
for _ in range(20):
    ops.elbo_grad_1d(Kuu, dKuu, acc, b, 1.0, 0.1, chunks=chunks, out=out, kuu=kuu)
t[2].record()
for _ in range(20):
    ops.elbo_grad_1d_single_stream(Kuu, dKuu, acc, b, 1.0, 0.1, chunks=chunks, out=out)
t[3].record(); torch.cuda.synchronize()
print("  Kuu chain alone %.1f us; P chains + bound alone %.1f us; single-stream call %.1f us"
      % (t[0].elapsed_time(t[1]) / 20 * 1e3, t[1].elapsed_time(t[2]) / 20 * 1e3, t[2].elapsed_time(t[3]) / 20 * 1e3))
o = out.cpu().numpy()
print("m %d chunks %d: %.1f us per call; Kuu chain cycles: sweep %.0f separators %.0f back %.0f tail %.0f; P chain: sweep %.0f separators %.0f; elbo %.6f"
      % (m, chunks, e0.elapsed_time(e1) / 20 * 1e3, o[9], o[10], o[11], o[12], o[13], o[14], o[0]))

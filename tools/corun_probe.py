"""Does ANY kernel running beside asvgp_accum_1d slow it?  (a) alone, (b) torch.cuda._sleep (one spinning thread, no memory
traffic) on a side stream, (c) the same on a high-priority stream, (d) the Kuu chain (4-CTA cluster) beside it."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asvgp_b200 import basis as B, kernels as Kn, ops
from asvgp_b200.inducing_features import SplineFeatures1D
n, m = 100_000_000, 10000
b = B.B3Spline(-1, m + 1, m)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.sort(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * m).values.clamp_(1e-9, m - 1e-9)
y = torch.sin(x / 5)
acc = torch.zeros(ops.accum_size_1d(b), dtype=torch.float64, device="cuda")
kern = Kn.Matern52(variance=1.0, lengthscales=1.0)
Kuu, dKuu = SplineFeatures1D(kern, b).make_Kuu_device(kern, want_grad=True)
main = torch.cuda.current_stream()
def run(name, side_fn):
    ts = []
    for it in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        side_fn()
        e0.record(main)
        ops.accum_1d(x, y, b, acc=acc)
        e1.record(main)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%-40s accumulate %.4f ms (min %.4f)" % (name, sorted(ts)[len(ts) // 2], min(ts)))
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)
def sleeper(stream, cycles):
    def f():
        stream.wait_stream(main)
        with torch.cuda.stream(stream):
            torch.cuda._sleep(cycles)
        ev = torch.cuda.Event(); ev.record(stream)
        # let the sleeper get going first
        torch.cuda._sleep(20000)
    return f
run("alone", lambda: None)
run("beside _sleep(1e6 cycles), prio 0", sleeper(s0, 1_000_000))
run("beside _sleep(1e6 cycles), prio -1", sleeper(s1, 1_000_000))
run("beside the Kuu chain (gate, 512 chunks)", lambda: ops.kuu_chain_1d(Kuu, dKuu, b, gate=True))
run("beside the Kuu chain (gate, 1024 chunks)", lambda: ops.kuu_chain_1d(Kuu, dKuu, b, chunks=1024, gate=True))
run("alone again", lambda: None)

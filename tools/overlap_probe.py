"""Timeline of one bench step (1-D, N = 1e8, M = 1e4): when the side-stream Kuu chain starts/ends relative to the accumulate.
usage: overlap_probe.py [gate 0/1] [priority 0/-1]"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asvgp_b200 import basis as B, kernels as Kn, ops, _lib
from asvgp_b200.inducing_features import SplineFeatures1D
gate = int(sys.argv[1]) if len(sys.argv) > 1 else 1
prio = int(sys.argv[2]) if len(sys.argv) > 2 else -1
n, m, k = 100_000_000, 10000, 3
b = B.B3Spline(-1, m + 1, m)
kern = Kn.Matern52(variance=1.0, lengthscales=1.0)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.sort(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * m).values.clamp_(1e-9, m - 1e-9)
y = torch.sin(x / 5)
feats = SplineFeatures1D(kern, b)
acc = torch.zeros(ops.accum_size_1d(b), dtype=torch.float64, device="cuda")
out = torch.empty(16, dtype=torch.float64, device="cuda")
side = torch.cuda.Stream(priority=prio)
main = torch.cuda.current_stream()
state = torch.empty(_lib.load().asvgp_kuu_state_doubles(m, k), dtype=torch.float64, device="cuda")
ws_a, ws_b = ops.workspace_1d(m, k, 0, slot="kuu"), ops.workspace_1d(m, k, 0)
P = lambda t: ctypes.c_void_p(t.data_ptr())
E = lambda: torch.cuda.Event(enable_timing=True)
rows = []
evs = []
sync_each = os.environ.get("PROBE_SYNC", "1") != "0"
for it in range(12):
    ev = {name: E() for name in ("t0", "asm", "side_begin", "gate", "chain_done", "acc_begin", "acc_end", "end")}
    acc.zero_()
    ev["t0"].record(main)
    Kuu, dKuu = feats.make_Kuu_device(kern, want_grad=True)
    ev["asm"].record(main)
    side.wait_stream(main)
    ev["side_begin"].record(side)
    ev["gate"].record(side)
    _lib.call("asvgp_kuu_chain_1d", P(Kuu), P(dKuu), m, k, 0, P(state), P(ws_a), ws_a.numel(), ctypes.c_void_p(ev["gate"].cuda_event),
              ctypes.c_void_p(side.cuda_stream))
    ev["chain_done"].record(side)
    if gate:
        main.wait_event(ev["gate"])
    ev["acc_begin"].record(main)
    ops.accum_1d(x, y, b, acc=acc)
    ev["acc_end"].record(main)
    _lib.call("asvgp_elbo_grad_1d_prepared", P(state), P(Kuu), P(dKuu), P(acc), m, k, 1.0, 0.1, 0, P(out), P(ws_b), ws_b.numel(),
              ctypes.c_void_p(ev["chain_done"].cuda_event), ctypes.c_void_p(main.cuda_stream))
    ev["end"].record(main)
    if sync_each:
        torch.cuda.synchronize()
    evs.append(ev)
torch.cuda.synchronize()
rows = [{name: ev["t0"].elapsed_time(e) * 1e3 for name, e in ev.items()} for ev in evs]
for r in rows[6:]:
    print("gate %d prio %d: " % (gate, prio) + "  ".join("%s %.0f" % (k_, v) for k_, v in r.items()) + "  (us after t0)")

"""torchrun --nproc-per-node N tools/oneshot_check.py : asvgp_allreduce_oneshot against torch.distributed.all_reduce (same
sums, bit-identical across ranks) and their device times for the 1-D packed accumulator (M = 1e4, k = 3: 50 002 doubles) and the
2-D one (200 x 200: 1.16 M doubles)."""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asvgp_b200 import dist as D

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
for n in (50_002, 1_160_002):
    red = D.oneshot_reducer(n)
    if red is None:
        if rank == 0: print("symmetric memory unavailable:", D._REDUCERS.get(((n, id(None), local), "why")))
        break
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    ok = True
    for step in range(6):
        part = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
        buf = red.buffer(); buf.copy_(part)
        red.reduce(out)
        ref = part.clone(); dist.all_reduce(ref)
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        err = float((out - ref).abs().max() / ref.abs().max())
        ok = ok and same and err < 1e-14
    red.check()
    # timing
    ev = lambda: torch.cuda.Event(enable_timing=True)
    def timeit(fn, reps=200):
        for _ in range(10): fn()
        dist.barrier(); torch.cuda.synchronize()
        a, b = ev(), ev(); a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps * 1e3], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()
    t_one = timeit(lambda: red.reduce(out))
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    t_nccl = timeit(lambda: dist.all_reduce(x))
    if rank == 0:
        print("n=%d world=%d: one-shot %s (bit-identical across ranks, = NCCL sum to 1e-14); one-shot %.1f us, NCCL all_reduce %.1f us"
              % (n, world, "OK" if ok else "MISMATCH", t_one, t_nccl))
dist.destroy_process_group()

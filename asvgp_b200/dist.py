"""Data-parallel plumbing: N shards over ranks (one process per GPU), ONE sum-allreduce of the packed accumulator
[G band | b | sum y^2 | count] over NCCL/NVLink, factorisation replicated per rank.  Prediction shards
the test points and needs no collective.  Works with any torch.distributed backend (the CPU tests use gloo)."""
import torch
import torch.distributed as dist


def is_distributed(flag="auto"):
    if flag is False:
        return False
    ok = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if flag is True and not ok:
        raise RuntimeError("distributed=True but torch.distributed is not initialised with world_size > 1")
    return ok


def shard_bounds(n, rank, world, align=2):
    """Contiguous [lo, hi) slice of n points for `rank`; boundaries are multiples of `align` points so that every
    shard of a 16-byte aligned array stays 16-byte aligned (the accumulate kernel's 128-bit load path)."""
    per = -(-n // world)
    per = -(-per // align) * align
    lo = min(rank * per, n)
    hi = min(lo + per, n)
    return lo, hi


def allreduce_packed(acc, group=None):
    """In-place SUM all-reduce of the packed accumulator; returns it."""
    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


class OneShotReducer:
    """The packed accumulator's all-reduce as ONE kernel over NVLink peer memory (asvgp_allreduce_oneshot) instead of NCCL's
    ring / tree: two symmetric buffers (torch.distributed._symmetric_memory) that the accumulate kernels write into directly,
    alternating per step; `reduce(out)` adds all ranks' current buffers, in rank order, into the local tensor `out`.

        red = OneShotReducer(n)            # collective: every rank, same n
        buf = red.buffer(); buf.zero_(); accumulate into buf ...; red.reduce(out)

    Raises RuntimeError at construction when symmetric memory is unavailable (no NVLink peer access, gloo, one rank); callers
    fall back to `allreduce_packed` (NCCL) — a different collective, the same sums."""

    def __init__(self, n, group=None, device=None):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib

        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            raise RuntimeError("OneShotReducer needs an initialised process group with more than one rank")
        if dist.get_backend(group) != "nccl":
            raise RuntimeError("OneShotReducer needs CUDA peers (nccl backend)")
        self._lib, self._ctypes = _lib, ctypes
        self.n = int(n)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        grp = group if group is not None else dist.group.WORLD
        n_alloc = self.n + (self.n & 1)
        self._bufs = [symm_mem.empty(n_alloc, dtype=torch.float64, device=self.device) for _ in range(2)]
        self._hdl = [symm_mem.rendezvous(b, grp.group_name) for b in self._bufs]
        self.rank, self.world = self._hdl[0].rank, self._hdl[0].world_size
        if self.world > 16:
            raise RuntimeError("OneShotReducer supports at most 16 ranks")
        for b in self._bufs:
            b.zero_()
        # a block of `world` 32-bit slots at the end of the signal pad (torch's own barriers use its head)
        self._pad_offset = [int(h.signal_pad_size) // 4 - 32 for h in self._hdl]
        self._epoch = [0, 0]
        self._turn = 0
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier(group)

    def buffer(self):
        """This step's symmetric buffer [n] (a view): accumulate this rank's partial sums into it."""
        return self._bufs[self._turn][: self.n]

    def reduce(self, out):
        """out[n] (local CUDA tensor) <- sum over ranks of their current buffers; flips to the other buffer."""
        t = self._turn
        h = self._hdl[t]
        self._epoch[t] += 1
        c = self._ctypes
        self._lib.call("asvgp_allreduce_oneshot", c.c_void_p(int(h.buffer_ptrs_dev)), c.c_void_p(int(h.signal_pad_ptrs_dev)), self.rank,
                       self.world, 0, self.n, c.c_uint(self._epoch[t] & 0xFFFFFFFF), self._pad_offset[t], c.c_void_p(out.data_ptr()),
                       c.c_void_p(self.status.data_ptr()), c.c_void_p(torch.cuda.current_stream().cuda_stream))
        self._turn ^= 1
        return out

    def check(self):
        """Host-synchronising: raises if a peer failed to arrive in any reduce so far."""
        if int(self.status.item()) != 0:
            raise RuntimeError("one-shot all-reduce: a peer never signalled (timed out); results are NaN")


_REDUCERS = {}


def oneshot_reducer(n, group=None):
    """Cached OneShotReducer for buffers of n doubles, or None when symmetric memory cannot be used here."""
    key = (n, id(group), torch.cuda.current_device() if torch.cuda.is_available() else -1)
    if key not in _REDUCERS:
        try:
            _REDUCERS[key] = OneShotReducer(n, group)
        except Exception as exc:                      # no peer access / unsupported backend: NCCL all_reduce does the job
            _REDUCERS[key] = None
            _REDUCERS[(key, "why")] = repr(exc)
    return _REDUCERS[key]


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1

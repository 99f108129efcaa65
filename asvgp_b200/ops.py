"""Thin Python faces of the C-ABI entry points.  torch is used for device memory and streams only; every
operator below runs a hand-written sm_100a kernel through ctypes (no torch math on the hot path, no CPU
fallback).  Inputs may be numpy arrays, torch tensors or any object exporting `__dlpack__`."""
import ctypes

import numpy as np
import torch

from . import _lib

F64 = torch.float64


def device():
    if not torch.cuda.is_available():
        raise _lib.AsvgpNativeError("asvgp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_device(a, dtype=F64):
    """numpy / torch / DLPack object -> contiguous CUDA tensor of `dtype` on the current device."""
    if isinstance(a, torch.Tensor):
        t = a
    elif isinstance(a, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(a))
    elif hasattr(a, "__dlpack__"):
        t = torch.from_dlpack(a)
    else:
        t = torch.as_tensor(np.asarray(a))
    return t.to(device=device(), dtype=dtype).contiguous()


def _p(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def device_mesh(basis):
    """The basis' knot mesh as a CUDA tensor (cached on the basis object per device)."""
    dev = device()
    cache = basis.__dict__.setdefault("_dev_mesh", {})
    if dev not in cache:
        cache[dev] = torch.from_numpy(np.ascontiguousarray(basis.mesh, dtype=np.float64)).to(dev)
    return cache[dev]


def basis_eval_1d(X, basis, dx=0):
    """(idx[n] int64, vals[(k+1), n]) as numpy arrays — rows idx+r of Kuf (reference basis.py:51-80)."""
    x = to_device(X).reshape(-1)
    n, k = x.numel(), basis.order
    mesh = device_mesh(basis)
    idx = torch.empty(n, dtype=torch.int64, device=x.device)
    vals = torch.empty((k + 1, n), dtype=F64, device=x.device)
    coef = to_device(basis.piece_coefficients(dx)) if dx > 0 else None
    _lib.call("asvgp_basis_eval_1d", _p(x), n, _p(mesh), mesh.numel(), k, int(dx), _p(coef), _p(idx), _p(vals),
              _stream())
    return idx.cpu().numpy(), vals.cpu().numpy()


def accum_size_1d(basis):
    return (basis.order + 2) * basis.m + 2


def accum_1d(x, y, basis, acc=None):
    """Adds sum_n w_n w_n^T (lower band), sum_n w_n y_n, sum y^2 and the count of the points (x, y) into the
    packed accumulator [G_band | b | sum(y^2) | count] (allocated zeroed if None).  Reference gpr.py:39-44."""
    x = to_device(x).reshape(-1)
    y = to_device(y).reshape(-1)
    if x.numel() != y.numel():
        raise ValueError("x and y must have the same number of points")
    mesh = device_mesh(basis)
    if acc is None:
        acc = torch.zeros(accum_size_1d(basis), dtype=F64, device=x.device)
    _lib.call("asvgp_accum_1d", _p(x), _p(y), x.numel(), _p(mesh), mesh.numel(), basis.order, _p(acc), _stream())
    return acc


def split_accum_1d(acc, basis):
    """Views (G_band (k+1, M), b (M,), scal (2,)) into the packed accumulator."""
    k, m = basis.order, basis.m
    return acc[: (k + 1) * m].view(k + 1, m), acc[(k + 1) * m: (k + 2) * m], acc[(k + 2) * m:]


def predict_1d(xnew, basis, alpha, S_band, variance, mean=None, var=None):
    """Posterior mean/variance at xnew from alpha = P^-1 b / sigma2 and S = band(P^-1) - band(Kuu^-1)
    (reference gpr.py:91-136)."""
    x = to_device(xnew).reshape(-1)
    n = x.numel()
    mesh = device_mesh(basis)
    alpha = to_device(alpha).reshape(-1)
    S_band = to_device(S_band)
    if mean is None:
        mean = torch.empty(n, dtype=F64, device=x.device)
    if var is None:
        var = torch.empty(n, dtype=F64, device=x.device)
    _lib.call("asvgp_predict_1d", _p(x), n, _p(mesh), mesh.numel(), basis.order, _p(alpha), _p(S_band),
              float(variance), _p(mean), _p(var), _stream())
    return mean, var


# ---- banded (latency-bound) operators ----------------------------------------------------------------------------------
def workspace_1d(m, order, chunks=0):
    """Scratch tensor for elbo_grad_1d / posterior_1d (cached per (device, m, order, chunks))."""
    dev = device()
    key = (dev, m, order, chunks)
    ws = _WORKSPACES.get(key)
    if ws is None:
        nbytes = _lib.load().asvgp_workspace_bytes_1d(m, order, chunks)
        if nbytes < 0:
            raise _lib.AsvgpNativeError("asvgp_workspace_bytes_1d rejected m=%d order=%d" % (m, order))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _WORKSPACES[key] = ws
    return ws


_WORKSPACES = {}


def device_tables(basis, names):
    """The named static bands of the basis stacked as one CUDA tensor [len(names), k+1, m] (cached)."""
    dev = device()
    cache = basis.__dict__.setdefault("_dev_tables", {})
    key = (dev, tuple(names))
    if key not in cache:
        stack = np.stack([np.asarray(getattr(basis, n), dtype=np.float64) for n in names])
        cache[key] = torch.from_numpy(np.ascontiguousarray(stack)).to(dev)
    return cache[key]


def kuu_assemble(basis, names, coef, dcoef=None, want_grad=True):
    """Kuu = sum coef[n] * basis.<n> (and its lengthscale derivative) as CUDA bands (reference
    inducing_features.py:12-44)."""
    tables = device_tables(basis, names)
    k, m = basis.order, basis.m
    Kuu = torch.empty((k + 1, m), dtype=F64, device=tables.device)
    dKuu = torch.empty_like(Kuu) if want_grad else None
    c = (ctypes.c_double * len(names))(*[float(v) for v in coef])
    dc = (ctypes.c_double * len(names))(*[float(v) for v in (dcoef if dcoef is not None else [0.0] * len(names))])
    _lib.call("asvgp_kuu_assemble", _p(tables), len(names), ctypes.cast(c, ctypes.c_void_p),
              ctypes.cast(dc, ctypes.c_void_p), m, k, _p(Kuu), _p(dKuu), _stream())
    return Kuu, dKuu


def elbo_grad_1d(Kuu, dKuu, acc, basis, variance, sigma2, chunks=0, out=None):
    """Launches the ELBO+gradient kernels; returns the 16-slot device result (see include/asvgp_b200.h)."""
    k, m = basis.order, basis.m
    ws = workspace_1d(m, k, chunks)
    if out is None:
        out = torch.empty(16, dtype=F64, device=acc.device)
    _lib.call("asvgp_elbo_grad_1d", _p(Kuu), _p(dKuu), _p(acc), m, k, float(variance), float(sigma2), int(chunks),
              _p(out), _p(ws), ws.numel(), _stream())
    return out


def posterior_1d(Kuu, acc, basis, sigma2, chunks=0):
    """alpha = P^-1 b / sigma2 and S = band(P^-1) - band(Kuu^-1) on the device (reference gpr.py:96-108)."""
    k, m = basis.order, basis.m
    ws = workspace_1d(m, k, chunks)
    alpha = torch.empty(m, dtype=F64, device=acc.device)
    S = torch.empty((k + 1, m), dtype=F64, device=acc.device)
    info = torch.zeros(2, dtype=F64, device=acc.device)
    _lib.call("asvgp_posterior_1d", _p(Kuu), _p(acc), m, k, float(sigma2), int(chunks), _p(alpha), _p(S), _p(info),
              _p(ws), ws.numel(), _stream())
    return alpha, S, info


# ---- host-resident data: stream it through pinned staging buffers ------------------------------------------------------
_STAGING = {}


def _staging(dev, chunk):
    key = (dev, chunk)
    st = _STAGING.get(key)
    if st is None:
        st = dict(
            host=[torch.empty((2, chunk), dtype=F64).pin_memory() for _ in range(2)],
            dev=[torch.empty((2, chunk), dtype=F64, device=dev) for _ in range(2)],
            copy_stream=torch.cuda.Stream(device=dev),
            filled=[torch.cuda.Event() for _ in range(2)],
            drained=[torch.cuda.Event() for _ in range(2)],
            staged=[torch.cuda.Event() for _ in range(2)],
        )
        _STAGING[key] = st
    return st


def accum_1d_host(x, y, basis, acc=None, chunk=1 << 23):
    """accum_1d for HOST arrays (numpy or CPU tensors): the points are streamed to the GPU in `chunk`-point pieces,
    host->pinned staging (unless already pinned), async H2D on a copy stream, double-buffered against the accumulate
    kernel — the device never holds more than two chunks."""
    dev = device()
    xt = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x, dtype=F64).reshape(-1)
    yt = torch.as_tensor(np.asarray(y) if not isinstance(y, torch.Tensor) else y, dtype=F64).reshape(-1)
    n = xt.numel()
    if yt.numel() != n:
        raise ValueError("x and y must have the same number of points")
    mesh = device_mesh(basis)
    if acc is None:
        acc = torch.zeros(accum_size_1d(basis), dtype=F64, device=dev)
    if n == 0:
        return acc
    chunk = min(chunk, n + (n & 1))
    st = _staging(dev, chunk)
    pinned = xt.is_pinned() and yt.is_pinned()
    main = torch.cuda.current_stream()
    cs = st["copy_stream"]
    n_chunks = -(-n // chunk)
    for c in range(n_chunks):
        lo, hi = c * chunk, min((c + 1) * chunk, n)
        s = c & 1
        dbuf = st["dev"][s]
        if c >= 2:
            cs.wait_event(st["drained"][s])            # kernel that read dbuf two chunks ago has finished
        with torch.cuda.stream(cs):
            if pinned:
                dbuf[0, : hi - lo].copy_(xt[lo:hi], non_blocking=True)
                dbuf[1, : hi - lo].copy_(yt[lo:hi], non_blocking=True)
            else:
                hbuf = st["host"][s]
                if c >= 2:
                    st["staged"][s].synchronize()      # previous H2D out of this pinned buffer has completed
                hbuf[0, : hi - lo].copy_(xt[lo:hi])
                hbuf[1, : hi - lo].copy_(yt[lo:hi])
                dbuf[:, : hi - lo].copy_(hbuf[:, : hi - lo], non_blocking=True)
                st["staged"][s].record(cs)
            st["filled"][s].record(cs)
        main.wait_event(st["filled"][s])
        _lib.call("asvgp_accum_1d", _p(dbuf[0]), _p(dbuf[1]), hi - lo, _p(mesh), mesh.numel(), basis.order, _p(acc),
                  ctypes.c_void_p(main.cuda_stream))
        st["drained"][s].record(main)
    return acc

"""GPU parity of the streaming 1-D kernels (through the C ABI) against the golden vectors produced by the
reference-under-shim and against the CPU oracle.  Tolerances: rel 1e-10 on G, Kuf_y, sum y^2;
1e-9 absolute on predictive mean / variance."""
import numpy as np
import pytest

from oracle import asvgp_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _basis(k, a, b, m, **kw):
    from asvgp_b200 import basis as B

    return getattr(B, "B%dSpline" % k)(a, b, m, **kw)


def _close(got, want, rtol=RTOL):
    scale = np.abs(want).max()
    np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * scale)


def _accum(x, y, basis):
    from asvgp_b200 import ops

    acc = ops.accum_1d(x, y, basis)
    return [t.cpu().numpy() for t in ops.split_accum_1d(acc, basis)]


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("tag,a,b", [("f32", -3.5, 10.5), ("f64", -1, 41)])
def test_basis_eval_matches_reference(cuda, golden, k, tag, a, b):
    g = golden("basis_eval")
    key = "k%d_%s" % (k, tag)
    basis = _basis(k, a, b, 40)
    np.testing.assert_array_equal(basis.mesh, g[key + "_mesh"])
    x = g[key + "_x"]
    for dx in range(4):
        name = key + "_dx%d" % dx
        if name not in g.files:
            continue
        got = basis.evaluate_basis(x.reshape(-1, 1), dx=dx).toarray()
        _close(got, g[name], rtol=1e-12)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_accum_matches_reference_unsorted(cuda, golden, k):
    """Golden inputs are in random order -> exercises the interval-switch and per-point RED paths."""
    g = golden("synth_1d")
    key = "k%d" % k
    m = int(g[key + "_m"])
    basis = _basis(k, -1, m + 1, m)
    G, b, scal = _accum(g[key + "_x"], g[key + "_y"], basis)
    _close(G, g[key + "_G"])
    _close(b, g[key + "_Kuf_y"].ravel())
    assert abs(scal[0] - float(g[key + "_tr_yTy"])) <= RTOL * float(g[key + "_tr_yTy"])
    assert scal[1] == g[key + "_x"].shape[0]


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_accum_binned_matches_reference(cuda, golden, k):
    """The bucket-partition path on the reference's own (random-order) inputs."""
    from asvgp_b200 import ops

    g = golden("synth_1d")
    key = "k%d" % k
    m = int(g[key + "_m"])
    basis = _basis(k, -1, m + 1, m)
    acc = ops.accum_1d(g[key + "_x"], g[key + "_y"], basis, binned=True)
    G, b, scal = [t.cpu().numpy() for t in ops.split_accum_1d(acc, basis)]
    _close(G, g[key + "_G"])
    _close(b, g[key + "_Kuf_y"].ravel())
    assert abs(scal[0] - float(g[key + "_tr_yTy"])) <= RTOL * float(g[key + "_tr_yTy"])
    assert scal[1] == g[key + "_x"].shape[0]


@pytest.mark.parametrize("n,m,k,dist", [(1, 30, 3, "uniform"), (4095, 30, 3, "uniform"), (4097, 700, 2, "uniform"),
                                        (300_001, 3000, 3, "uniform"), (250_000, 10_000, 4, "clustered"),
                                        (123_457, 40, 6, "one-interval"), (200_000, 300, 1, "knots")])
def test_accum_binned_matches_oracle(cuda, n, m, k, dist):
    """Shuffled inputs through the partition path: several intervals per bucket (m > 256), empty buckets, every point
    in one interval, points exactly on knots (quirk Q2), unit and tile boundaries."""
    from asvgp_b200 import ops

    rng = np.random.default_rng(n + m)
    basis = _basis(k, -1, m + 1, m)
    lo, hi = 0.0, float(m)
    if dist == "uniform":
        x = rng.uniform(lo, hi, n)
    elif dist == "clustered":
        x = np.concatenate([rng.normal(0.31 * m, 0.002 * m, n // 2), rng.uniform(0.8 * m, 0.85 * m, n - n // 2)])
    elif dist == "one-interval":
        x = rng.uniform(7.01, 7.49, n)
    else:
        x = rng.choice(np.asarray(basis.mesh)[1:-1], n)
    y = np.cos(x / 7.0) + 0.1 * rng.standard_normal(n)
    acc = ops.accum_1d(x, y, basis, binned=True)
    G, b, scal = [t.cpu().numpy() for t in ops.split_accum_1d(acc, basis)]
    G0, b0, yy0 = O.precompute_1d_chunked(basis.mesh, basis.delta, k, m, x, y)
    _close(G, G0)
    _close(b, b0.ravel())
    assert abs(scal[0] - yy0) <= RTOL * yy0 and scal[1] == n


def test_accum_order_probe_and_auto(cuda):
    from asvgp_b200 import ops

    rng = np.random.default_rng(5)
    m, n = 2000, 1 << 19
    basis = _basis(3, -1, m + 1, m)
    xs = np.sort(rng.uniform(0, m, n))
    xr = rng.permutation(xs)
    assert ops.order_probe_1d(xs, basis) == 0.0
    assert ops.order_probe_1d(xr, basis) > 0.9
    y = np.sin(xr / 11)
    a_auto = ops.accum_1d(xr, y, basis, binned="auto").cpu().numpy()
    a_plain = ops.accum_1d(xr, y, basis).cpu().numpy()
    _close(a_auto, a_plain)


def test_accum_snelson_golden(cuda, golden):
    g = golden("snelson")
    basis = _basis(3, -3.5, 10.5, 100)
    np.testing.assert_array_equal(basis.mesh, g["mesh"])
    G, b, scal = _accum(g["X"], g["y"], basis)
    _close(G, g["G"])
    _close(b, g["Kuf_y"].ravel())
    assert abs(scal[0] - float(g["tr_yTy"])) <= RTOL * float(g["tr_yTy"])


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("order", ["sorted", "shuffled"])
def test_accum_matches_oracle(cuda, k, order):
    rng = np.random.default_rng(100 + k)
    m, n = 300, 200_003
    basis = _basis(k, -1, m + 1, m)
    x = rng.uniform(0.0, m, n)
    if order == "sorted":
        x.sort()
    y = np.cos(x / 7.0) + 0.2 * rng.standard_normal(n)
    G0, b0, yy0 = O.precompute_1d(basis.mesh, basis.delta, k, m, x, y)
    G, b, scal = _accum(x, y, basis)
    _close(G, G0)
    _close(b, b0.ravel())
    assert abs(scal[0] - yy0) <= RTOL * yy0 and scal[1] == n


@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 32, 33, 255, 257, 1025])
def test_accum_ragged_sizes_and_misaligned_views(cuda, n):
    import torch

    from asvgp_b200 import ops

    rng = np.random.default_rng(n)
    m, k = 40, 3
    basis = _basis(k, -1, m + 1, m)
    x = np.sort(rng.uniform(0.0, m, n + 1))
    y = rng.standard_normal(n + 1)
    xd, yd = ops.to_device(x), ops.to_device(y)
    for off in (0, 1):                      # off=1: 8-byte aligned only -> scalar-load variant
        xs, ys = xd[off:off + n], yd[off:off + n]
        acc = ops.accum_1d(xs, ys, basis)
        G, b, scal = [t.cpu().numpy() for t in ops.split_accum_1d(acc, basis)]
        G0, b0, yy0 = O.precompute_1d(basis.mesh, basis.delta, k, m, x[off:off + n], y[off:off + n]) if n else (
            np.zeros((k + 1, m)), np.zeros((m, 1)), 0.0)
        np.testing.assert_allclose(G, G0, rtol=RTOL, atol=1e-13)
        np.testing.assert_allclose(b, b0.ravel(), rtol=RTOL, atol=1e-13)
        assert abs(scal[0] - yy0) <= RTOL * max(yy0, 1e-300) and scal[1] == n
    torch.cuda.synchronize()


@pytest.mark.parametrize("n", [65536, 65537, 65538, 2 * 148 * 1024 + 1, 2 * 148 * 1024 * 3, 1_000_001])
def test_accum_chunk_boundary_sizes(cuda, n):
    """Sizes around the per-CTA / per-warp tile boundaries of the accumulate kernel, odd and even."""
    rng = np.random.default_rng(n % 1000)
    m, k = 500, 3
    basis = _basis(k, -1, m + 1, m)
    x = np.sort(rng.uniform(0.0, m, n))
    y = np.cos(x / 5.0) + 0.2 * rng.standard_normal(n)
    G0, b0, yy0 = O.precompute_1d_chunked(basis.mesh, basis.delta, k, m, x, y)
    G, b, scal = _accum(x, y, basis)
    _close(G, G0)
    _close(b, b0.ravel())
    assert abs(scal[0] - yy0) <= RTOL * yy0 and scal[1] == n


def test_accum_points_on_knots_and_edges(cuda):
    """Points exactly on knots go to the LEFT interval (searchsorted side='left'); x=a+eps, x=b-eps."""
    m, k = 30, 3
    basis = _basis(k, 0, 27, m)           # integer endpoints, delta = 1 exactly
    x = np.concatenate([basis.mesh[1:-1], basis.mesh[1:-1], [1e-9, 27 - 1e-9, 13.5]])
    y = np.arange(x.shape[0], dtype=np.float64) / 10.0
    G0, b0, yy0 = O.precompute_1d(basis.mesh, basis.delta, k, m, x, y)
    G, b, scal = _accum(x, y, basis)
    _close(G, G0, 1e-13)
    _close(b, b0.ravel(), 1e-13)


def test_accum_is_additive_over_shards(cuda):
    """Accumulating two halves into the same packed buffer equals one pass (what the multi-GPU path relies on)."""
    from asvgp_b200 import ops

    rng = np.random.default_rng(5)
    m, k, n = 200, 3, 100_000
    basis = _basis(k, -1, m + 1, m)
    x = np.sort(rng.uniform(0.0, m, n))
    y = rng.standard_normal(n)
    whole = ops.accum_1d(x, y, basis)
    acc = ops.accum_1d(x[: n // 2], y[: n // 2], basis)
    acc = ops.accum_1d(x[n // 2:], y[n // 2:], basis, acc=acc)
    np.testing.assert_allclose(acc.cpu().numpy(), whole.cpu().numpy(), rtol=1e-12, atol=1e-12)


def test_accum_large_sorted_properties(cuda):
    """N = 2^25 sorted points: partition of unity gives sum(G_full) = N and sum(b) = sum(y) (size-independent)."""
    import torch

    from asvgp_b200 import ops

    m, k, n = 10_000, 3, 1 << 25
    basis = _basis(k, -1, m + 1, m)
    gen = torch.Generator(device="cuda").manual_seed(1997)
    x = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) * m
    x = torch.sort(x).values
    y = torch.sin(x / 37.0)
    acc = ops.accum_1d(x, y, basis)
    G, b, scal = ops.split_accum_1d(acc, basis)
    full = G[0].sum() + 2.0 * G[1:].sum()
    assert abs(full.item() - n) <= 1e-10 * n
    assert abs(b.sum().item() - y.sum().item()) <= 1e-9 * n
    assert abs(scal[0].item() - (y * y).sum().item()) <= 1e-10 * n and scal[1].item() == n


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_predict_matches_reference(cuda, golden, k):
    import scipy.linalg as sla

    from asvgp_b200 import ops

    g = golden("synth_1d")
    key = "k%d" % k
    m = int(g[key + "_m"])
    basis = _basis(k, -1, m + 1, m)
    var, ell, s2 = g[key + "_pred_hypers"]
    kind = str(g[key + "_pred_kind"])
    tables = O.static_bands(k, m, basis.delta)
    Kuu = O.make_Kuu(kind, ell, var, tables)
    G, b = g[key + "_G"], g[key + "_Kuf_y"]
    P = G / s2 + Kuu
    LP = sla.cholesky_banded(P, lower=True)
    LK = sla.cholesky_banded(Kuu, lower=True)
    alpha = sla.cho_solve_banded((LP, True), b) / s2
    S = O.takahashi_band(LP) - O.takahashi_band(LK)
    mean, v = ops.predict_1d(g[key + "_xs"], basis, alpha, S, var)
    np.testing.assert_allclose(mean.cpu().numpy(), g[key + "_mean"].ravel(), atol=1e-9, rtol=0)
    np.testing.assert_allclose(v.cpu().numpy(), g[key + "_var"].ravel(), atol=1e-9, rtol=0)


def test_accum_binned_large_matches_streaming(cuda):
    """2e7 shuffled points on the C3 mesh (M = 1e4: 40 intervals per bucket, ~4900 units): the partition path against the
    streaming kernel on the same device data, and the size-independent checks of the sorted test (count, partition of unity)."""
    import torch
    from asvgp_b200 import ops

    m, n = 10_000, 20_000_000
    basis = _basis(3, -1, m + 1, m)
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    x = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * m
    y = torch.sin(x / 37.0) + 0.25
    a_bin = ops.accum_1d(x, y, basis, binned=True)
    a_str = ops.accum_1d(x, y, basis)
    scale = float(a_str[: 4 * m].abs().max())
    assert float((a_bin - a_str).abs().max()) <= 1e-11 * max(scale, float(a_str.abs().max()))
    G, b, scal = ops.split_accum_1d(a_bin, basis)
    assert scal[1].item() == n
    # partition of unity: sum over the band of G (off-diagonals twice) = number of points; sum of b = sum of y
    total = float(G[0].sum() + 2 * G[1:].sum())
    assert abs(total - n) <= 1e-10 * n
    assert abs(float(b.sum()) - float(y.sum())) <= 1e-10 * float(y.abs().sum())

"""GPU parity of the 2-D Kronecker model through the reference-shaped API (GPR_kron): Gram stencil / projection vs
the golden vectors produced by the unmodified reference under the shim (rel 1e-10), ELBO vs golden (rel 1e-10),
gradients vs the torch-autograd dense oracle (1e-8), predictions vs golden (1e-9 abs), plus a mid-size case against
the LAPACK-band oracle and size-independent properties (order invariance, additivity over shards)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import asvgp_oracle as O

pytestmark = pytest.mark.gpu
DOMAINS = ((0, 1), (0, 2))


def _model(X, y, k, ms, kinds=("Matern32", "Matern32"), hyp=None, sigma2=None, domains=DOMAINS, **kw):
    from asvgp_b200 import basis as B, kernels as Kn
    from asvgp_b200.gpr import GPR_kron

    cls = getattr(B, "B%dSpline" % k)
    bases = [cls(domains[0][0], domains[0][1], ms[0]), cls(domains[1][0], domains[1][1], ms[1])]
    kerns = [getattr(Kn, kind)() for kind in kinds]
    model = GPR_kron((X, y.reshape(-1, 1)), kerns, bases, **kw)
    if hyp is not None:
        for kern, (v, l) in zip(kerns, hyp):
            kern.variance.assign(v); kern.lengthscales.assign(l)
    if sigma2 is not None:
        model.likelihood.variance.assign(sigma2)
    return model


def _golden_G(g, key, m):
    return sp.coo_matrix((g[key + "_G_val"], (g[key + "_G_row"], g[key + "_G_col"])), shape=(m * m, m * m)).toarray()


@pytest.mark.parametrize("k", [2, 3, 4])
def test_accumulate_matches_reference(cuda, golden, k):
    g = golden("kron_2d")
    key = "k%d" % k
    m = int(g[key + "_m"])
    model = _model(g["X"], g["y"], k, (m, m))
    Gref = _golden_G(g, key, m)
    scale = np.abs(Gref).max()
    np.testing.assert_allclose(model.KufKfu_sparse.toarray(), Gref, rtol=1e-10, atol=1e-10 * scale)
    np.testing.assert_allclose(model.Kuf_y, g[key + "_Kuf_y"], rtol=1e-10, atol=1e-10 * np.abs(g[key + "_Kuf_y"]).max())
    assert abs(model.tr_yTy - float(g[key + "_tr_yTy"])) <= 1e-12 * float(g[key + "_tr_yTy"])
    assert model.num_data == g["X"].shape[0]
    assert model.bandwidth == int(g[key + "_bandwidth"])


@pytest.mark.parametrize("k", [2, 3, 4])
def test_elbo_golden(cuda, golden, k):
    g = golden("kron_2d")
    key = "k%d" % k
    m = int(g[key + "_m"])
    model = _model(g["X"], g["y"], k, (m, m), hyp=[(.7, .3), (1.3, .5)], sigma2=.05)
    want = float(g[key + "_elbo"])
    assert abs(model.elbo() - want) <= 1e-10 * abs(want)
    e, _ = model.elbo_and_grad()
    assert abs(e - want) <= 1e-10 * abs(want)
    if key + "_elbo_m52_m12" in g.files:
        model = _model(g["X"], g["y"], k, (m, m), kinds=("Matern52", "Matern12"), hyp=[(.9, .4), (1.1, .6)], sigma2=.2)
        want = float(g[key + "_elbo_m52_m12"])
        assert abs(model.elbo() - want) <= 1e-10 * abs(want)


@pytest.mark.parametrize("k", [2, 3, 4])
def test_predict_golden(cuda, golden, k):
    g = golden("kron_2d")
    key = "k%d" % k
    m = int(g[key + "_m"])
    model = _model(g["X"], g["y"], k, (m, m), hyp=[(.7, .3), (1.3, .5)], sigma2=.05)
    mean, var = model.predict_f(g[key + "_Xs"])
    np.testing.assert_allclose(mean, g[key + "_mean"], atol=1e-9, rtol=0)
    np.testing.assert_allclose(var, g[key + "_var"], atol=1e-9, rtol=0)
    mean2, var2 = model.predict_f_sparse(g[key + "_Xs"])
    np.testing.assert_array_equal(mean, mean2)


@pytest.mark.parametrize("k,kinds", [(3, ("Matern32", "Matern32")), (3, ("Matern52", "Matern12")),
                                     (4, ("Matern32", "Matern52")), (2, ("Matern12", "Matern32"))])
def test_gradients_match_autograd_oracle(cuda, golden, k, kinds):
    g = golden("kron_2d")
    key = "k%d" % k
    m = int(g[key + "_m"])
    hyp, s2 = [(.7, .3), (1.3, .5)], .05
    model = _model(g["X"], g["y"], k, (m, m), kinds=kinds, hyp=hyp, sigma2=s2)
    elbo, grads = model.elbo_and_grad()
    deltas = [b.delta for b in model.bases]
    T = [O.static_bands(k, m, d) for d in deltas]
    G = sp.coo_matrix((g[key + "_G_val"], (g[key + "_G_row"], g[key + "_G_col"])), shape=(m * m, m * m)).tocsr()
    e0, g0 = O.elbo_grad_kron_dense(kinds, T, G, g[key + "_Kuf_y"], float(g[key + "_tr_yTy"]), g["X"].shape[0], hyp, s2)
    assert abs(elbo - e0) <= 1e-10 * abs(e0)
    got = np.array([grads[id(p)] for p in model.trainable_variables])
    np.testing.assert_allclose(got, g0, rtol=1e-8, atol=1e-8 * np.abs(g0).max())


def _raster(n1, n2, rng, domains=((-80, -25), (15, 55)), inner=((-75, -30), (20, 50))):
    x1 = np.linspace(inner[0][0], inner[0][1], n1)
    x2 = np.linspace(inner[1][0], inner[1][1], n2)
    X = np.stack(np.meshgrid(x1, x2, indexing="ij"), -1).reshape(-1, 2)          # x1 slow (raster order)
    y = np.sin(X[:, 0] / 4.0) * np.cos(X[:, 1] / 3.0) + 0.05 * rng.standard_normal(X.shape[0])
    return X, y


def test_midsize_rectangular_vs_banded_oracle(cuda):
    """eNATL60-shaped raster (x1 slow), m1 != m2, several block columns and row tiles in the band factorisation."""
    rng = np.random.default_rng(5)
    k, ms = 3, (40, 36)
    domains = ((-80, -25), (15, 55))
    X, y = _raster(500, 400, rng)
    hyp, s2 = [(1.0, 6.0), (0.8, 5.0)], 0.01
    model = _model(X, y, k, ms, hyp=hyp, sigma2=s2, domains=domains)
    meshes = [b.mesh for b in model.bases]
    deltas = [b.delta for b in model.bases]
    G0, b0, yy0 = O.precompute_kron(meshes, deltas, k, list(ms), X, y)
    scale = abs(G0).max()
    assert abs(model.KufKfu_sparse - G0).max() <= 1e-10 * scale
    np.testing.assert_allclose(model.Kuf_y, b0, rtol=1e-10, atol=1e-10 * np.abs(b0).max())
    T = [O.static_bands(k, m, d) for m, d in zip(ms, deltas)]
    Ks = [O.make_Kuu("Matern32", l, v, t) for (v, l), t in zip(hyp, T)]
    want = O.elbo_kron_banded(Ks, G0, b0, yy0, X.shape[0], [h[0] for h in hyp], s2, k, list(ms))
    assert abs(model.elbo() - want) <= 1e-10 * abs(want)
    e, grads = model.elbo_and_grad()
    assert abs(e - want) <= 1e-10 * abs(want)
    # gradient vs central differences of the banded oracle (the dense autograd oracle is too slow at M = 1440)
    params = model.trainable_variables
    got = np.array([grads[id(p)] for p in params])
    th0 = np.array([hyp[0][0], hyp[0][1], hyp[1][0], hyp[1][1], s2])

    def f(th):
        Ks = [O.make_Kuu("Matern32", th[1], th[0], T[0]), O.make_Kuu("Matern32", th[3], th[2], T[1])]
        return O.elbo_kron_banded(Ks, G0, b0, yy0, X.shape[0], [th[0], th[2]], th[4], k, list(ms))

    for i in range(5):
        h = 1e-4 * th0[i]
        e_ = np.zeros(5); e_[i] = h
        fd = (-f(th0 + 2 * e_) + 8 * f(th0 + e_) - 8 * f(th0 - e_) + f(th0 - 2 * e_)) / (12 * h)
        assert abs(fd - got[i]) <= 1e-6 * max(1.0, abs(got[i])), (i, fd, got[i])
    Xs = np.stack([rng.uniform(-74, -31, 300), rng.uniform(21, 49, 300)], 1)
    mean, var = model.predict_f(Xs)
    mean0, var0 = O.predict_kron_banded(meshes, deltas, k, list(ms), Ks, G0, b0, [h[0] for h in hyp], s2, Xs)
    np.testing.assert_allclose(mean, mean0, atol=1e-9, rtol=0)
    np.testing.assert_allclose(var, var0, atol=1e-9, rtol=0)
    # gridded test points (x1 constant along a row: the predictor's run path), ragged row length
    g1, g2 = np.linspace(-74.5, -30.5, 37), np.linspace(20.5, 49.5, 53)
    Xg = np.stack(np.meshgrid(g1, g2, indexing="ij"), -1).reshape(-1, 2)
    mean, var = model.predict_f(Xg)
    mean0, var0 = O.predict_kron_banded(meshes, deltas, k, list(ms), Ks, G0, b0, [h[0] for h in hyp], s2, Xg)
    np.testing.assert_allclose(mean, mean0, atol=1e-9, rtol=0)
    np.testing.assert_allclose(var, var0, atol=1e-9, rtol=0)


def test_accumulate_order_invariance_and_shard_additivity(cuda):
    import torch

    from asvgp_b200 import basis as B, ops

    rng = np.random.default_rng(11)
    X, y = _raster(700, 300, rng)
    bases = [B.B3Spline(-80, -25, 30), B.B3Spline(15, 55, 24)]

    def run(parts):
        acc = torch.zeros(ops.accum_size_2d(bases), dtype=torch.float64, device="cuda")
        cm = ops.moment_table_2d(bases)
        for Xp, yp in parts:
            ops.accum_2d(Xp, yp, bases, cm, ops.split_accum_2d(acc, bases)[2])
        ops.expand_moments_2d(cm, bases, acc)
        return acc.cpu().numpy()

    full = run([(X, y)])
    perm = rng.permutation(X.shape[0])
    shuffled = run([(X[perm], y[perm])])
    cut = 123457
    sharded = run([(X[:cut], y[:cut]), (X[cut:], y[cut:])])
    scale = np.abs(full).max()
    np.testing.assert_allclose(shuffled, full, rtol=0, atol=1e-11 * scale)
    np.testing.assert_allclose(sharded, full, rtol=0, atol=1e-11 * scale)
    assert full[-1] == X.shape[0]
    # partition of unity: sum of all entries of Kuf Kuf^T = N, sum of Kuf y = sum y
    Gs, b, _ = (t for t in np.split(full, [ops.stencil_rows(3) * 30 * 24, (ops.stencil_rows(3) + 1) * 30 * 24]))
    Gs = Gs.reshape(ops.stencil_rows(3), -1)
    diag = Gs[3].sum()
    assert abs(2 * Gs.sum() - diag - X.shape[0]) <= 1e-9 * X.shape[0]
    assert abs(b.sum() - y.sum()) <= 1e-9 * np.abs(y).sum()


@pytest.mark.parametrize("variant,n1,n2", [("separable", 130, 300), ("separable", 67, 301), ("curvilinear", 130, 300),
                                            ("curvilinear", 67, 301), ("x1-runs", 97, 211), ("separable-dirty", 130, 300), ("separable-dirty-x1", 130, 300)])
def test_accumulate_raster_variants(cuda, variant, n1, n2):
    """Every input class of asvgp_accum_2d's probe (separable raster -> column sweep, raster with row-dependent x2 ->
    row streaming with 256-bit or per-point loads, ragged x1 runs) against the SciPy oracle."""
    import torch

    from asvgp_b200 import basis as B, ops, utils

    rng = np.random.default_rng(n1 * 1000 + n2)
    x1 = np.sort(rng.uniform(-74.9, -30.1, n1))
    x2 = np.sort(rng.uniform(20.1, 49.9, n2))
    X = np.stack(np.meshgrid(x1, x2, indexing="ij"), -1)
    if variant == "curvilinear":
        X[:, :, 1] += 0.01 * np.sin(np.arange(n1))[:, None]            # x2 depends on the row
    if variant == "separable-dirty-x1":
        for r, c in ((3, 17), (50, 150), (51, 151), (129, 298)):        # interior points only: the probe cannot see them
            X[r, c, 0] += 0.013
    if variant == "separable-dirty":
        # a few interior points off their row's x1 / their column's x2: the probe still sees a separable raster, the
        # kernel has to notice point by point
        for r, c in ((3, 17), (50, 150), (51, 151), (129, 298)):
            X[r, c, 0] += 0.013
        for r, c in ((7, 5), (64, 200), (100, 31)):
            X[r, c, 1] -= 0.021
    X = X.reshape(-1, 2)
    if variant == "x1-runs":
        keep = rng.uniform(size=X.shape[0]) > 0.1                      # rows of different lengths
        X = X[keep]
    y = np.sin(X[:, 0] / 4.0) * np.cos(X[:, 1] / 3.0) + 0.05 * rng.standard_normal(X.shape[0])
    bases = [B.B3Spline(-80, -25, 19), B.B3Spline(15, 55, 23)]
    acc = torch.zeros(ops.accum_size_2d(bases), dtype=torch.float64, device="cuda")
    cm = ops.moment_table_2d(bases)
    ops.accum_2d(X, y, bases, cm, ops.split_accum_2d(acc, bases)[2])
    select = int(cm[-1:].view(torch.int32)[0].item())
    want = {"separable": (4,), "separable-dirty-x1": (4,), "separable-dirty": (3, 4), "curvilinear": (3,) if n2 % 4 == 0 else (2,), "x1-runs": (1,)}[variant]
    assert select in want, "probe chose path %d for a %s input" % (select, variant)
    ops.expand_moments_2d(cm, bases, acc)
    Gs, b, scal = [t.cpu().numpy() for t in ops.split_accum_2d(acc, bases)]
    G0, b0, yy0 = O.precompute_kron([bb.mesh for bb in bases], [bb.delta for bb in bases], 3, [19, 23], X, y)
    scale = abs(G0).max()
    assert abs(utils.stencil_to_sparse(Gs, 19, 23, 3) - G0).max() <= 1e-11 * scale
    np.testing.assert_allclose(b, b0.ravel(), rtol=0, atol=1e-11 * np.abs(b0).max())
    assert abs(scal[0] - yy0) <= 1e-12 * yy0 and scal[1] == X.shape[0]


@pytest.mark.parametrize("n", [0, 1, 2, 3, 255, 1025])
def test_accumulate_ragged_sizes(cuda, n):
    import torch

    from asvgp_b200 import basis as B, ops

    rng = np.random.default_rng(n)
    bases = [B.B3Spline(0, 1, 12), B.B3Spline(0, 2, 13)]
    X = np.stack([rng.uniform(.01, .99, n), rng.uniform(.01, 1.99, n)], 1).reshape(n, 2)
    y = rng.standard_normal(n)
    acc = torch.zeros(ops.accum_size_2d(bases), dtype=torch.float64, device="cuda")
    cm = ops.moment_table_2d(bases)
    ops.accum_2d(X, y, bases, cm, ops.split_accum_2d(acc, bases)[2])
    ops.expand_moments_2d(cm, bases, acc)
    Gs, b, scal = [t.cpu().numpy() for t in ops.split_accum_2d(acc, bases)]
    assert scal[1] == n
    if n == 0:
        assert not Gs.any() and not b.any()
        return
    from asvgp_b200 import utils

    meshes = [bb.mesh for bb in bases]
    deltas = [bb.delta for bb in bases]
    G0, b0, yy0 = O.precompute_kron(meshes, deltas, 3, [12, 13], X, y)
    assert abs(utils.stencil_to_sparse(Gs, 12, 13, 3) - G0).max() <= 1e-12
    np.testing.assert_allclose(b, b0.ravel(), atol=1e-12)
    assert abs(scal[0] - yy0) <= 1e-12 * max(yy0, 1.0)


@pytest.mark.parametrize("k,m1,m2,n,dist", [(3, 19, 23, 70_001, "uniform"), (2, 30, 11, 4097, "uniform"), (4, 14, 14, 50_000, "uniform"),
                                            (3, 400, 12, 150_000, "uniform"), (3, 19, 23, 60_000, "clustered"),
                                            (1, 25, 40, 30_000, "knots"), (5, 16, 16, 20_000, "uniform"), (6, 15, 15, 9_000, "uniform")])
def test_accumulate_binned_matches_oracle(cuda, k, m1, m2, n, dist):
    """Shuffled points through the partition path (asvgp_accum_2d_binned): several dim-1 intervals per bucket (m1 > 256),
    clustered points, points exactly on knots in both dimensions, unit and tile boundaries, every spline order."""
    import torch

    from asvgp_b200 import basis as B, ops, utils

    rng = np.random.default_rng(n + 7 * k)
    cls = getattr(B, "B%dSpline" % k)
    bases = [cls(-80, -25, m1), cls(15, 55, m2)]
    if dist == "uniform":
        X = np.stack([rng.uniform(-79.5, -25.5, n), rng.uniform(15.5, 54.5, n)], 1)
    elif dist == "clustered":
        X = np.stack([rng.normal(-50.0, 0.4, n), rng.normal(30.0, 6.0, n)], 1).clip((-79, 16), (-26, 54))
    else:
        X = np.stack([rng.choice(np.asarray(bases[0].mesh)[1:-1], n), rng.choice(np.asarray(bases[1].mesh)[1:-1], n)], 1)
    y = np.sin(X[:, 0] / 4.0) * np.cos(X[:, 1] / 3.0) + 0.05 * rng.standard_normal(n)
    acc = torch.zeros(ops.accum_size_2d(bases), dtype=torch.float64, device="cuda")
    cm = ops.moment_table_2d(bases)
    ops.accum_2d(X, y, bases, cm, ops.split_accum_2d(acc, bases)[2], binned=True)
    assert bool(torch.isfinite(cm).all()), "moment table has %d non-finite entries" % int((~torch.isfinite(cm)).sum())
    ops.expand_moments_2d(cm, bases, acc)
    assert bool(torch.isfinite(acc).all()), "expand_moments_2d produced %d non-finite entries" % int((~torch.isfinite(acc)).sum())
    Gs, b, scal = [t.cpu().numpy() for t in ops.split_accum_2d(acc, bases)]
    G0, b0, yy0 = O.precompute_kron([bb.mesh for bb in bases], [bb.delta for bb in bases], k, [m1, m2], X, y)
    scale = abs(G0).max()
    assert abs(utils.stencil_to_sparse(Gs, m1, m2, k) - G0).max() <= 1e-11 * scale
    np.testing.assert_allclose(b, b0.ravel(), rtol=0, atol=1e-11 * np.abs(b0).max())
    assert abs(scal[0] - yy0) <= 1e-12 * yy0 and scal[1] == n


def test_accumulate_order_probe_2d(cuda):
    import torch

    from asvgp_b200 import basis as B, ops

    rng = np.random.default_rng(3)
    bases = [B.B3Spline(-80, -25, 40), B.B3Spline(15, 55, 40)]
    X = np.stack(np.meshgrid(np.linspace(-75, -30, 600), np.linspace(20, 50, 700), indexing="ij"), -1).reshape(-1, 2)
    assert ops.order_probe_2d(X, bases) < 0.01
    Xs = rng.permutation(X)
    assert ops.order_probe_2d(Xs, bases) > 0.8
    y = np.cos(Xs[:, 0]) + Xs[:, 1]
    res = []
    for mode in ("auto", False):
        acc = torch.zeros(ops.accum_size_2d(bases), dtype=torch.float64, device="cuda")
        cm = ops.moment_table_2d(bases)
        ops.accum_2d(Xs, y, bases, cm, ops.split_accum_2d(acc, bases)[2], binned=mode)
        ops.expand_moments_2d(cm, bases, acc)
        res.append(acc.cpu().numpy())
    np.testing.assert_allclose(res[0], res[1], rtol=0, atol=1e-11 * np.abs(res[1]).max())


def test_accumulate_binned_large_matches_streaming(cuda):
    """4e6 shuffled raster points on the C4 meshes (200 x 200, k = 3: one dim-1 interval per bucket, 197 cells per unit row):
    the partition path against the streaming kernels on the same device data, plus partition of unity."""
    import torch
    from asvgp_b200 import basis as B, ops

    n1 = 2000
    bases = [B.B3Spline(-80, -25, 200), B.B3Spline(15, 55, 200)]
    x1 = torch.linspace(-75, -30, n1, dtype=torch.float64, device="cuda")
    x2 = torch.linspace(20, 50, n1, dtype=torch.float64, device="cuda")
    X = torch.stack([x1[:, None].expand(n1, n1), x2[None, :].expand(n1, n1)], -1).reshape(-1, 2)
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    X = X[torch.randperm(n1 * n1, device="cuda", generator=g)].contiguous()
    y = torch.sin(X[:, 0] / 4) * torch.cos(X[:, 1] / 3) + 0.5
    res = []
    for mode in (True, False):
        acc = torch.zeros(ops.accum_size_2d(bases), dtype=torch.float64, device="cuda")
        cm = ops.moment_table_2d(bases)
        ops.accum_2d(X, y, bases, cm, ops.split_accum_2d(acc, bases)[2], binned=mode)
        ops.expand_moments_2d(cm, bases, acc)
        res.append(acc)
    assert float((res[0] - res[1]).abs().max()) <= 1e-11 * float(res[1].abs().max())
    Gs, b, scal = ops.split_accum_2d(res[0], bases)
    assert scal[1].item() == n1 * n1
    assert abs(float(b.sum()) - float(y.sum())) <= 1e-10 * float(y.abs().sum())


def test_raster_statement_paths_match_the_classified_ones(cuda):
    """asvgp_accum_2d_raster / asvgp_predict_2d_apply(row_len): the caller states the layout instead of the on-device probe.
    Same sums and predictions as the classified path for a true raster — and still the right numbers when the statement is
    WRONG (shuffled points passed as a 'raster'): every lane re-checks each point and falls back."""
    import torch

    from asvgp_b200 import basis as B, kernels as Kn, ops
    from asvgp_b200.gpr import GPR_kron

    rng = np.random.default_rng(21)
    n1, n2 = 300, 260
    X, y = _raster(n1, n2, rng)
    bases = [B.B3Spline(-80, -25, 30), B.B3Spline(15, 55, 24)]

    def run(Xa, ya, **kw):
        acc = torch.zeros(ops.accum_size_2d(bases), dtype=torch.float64, device="cuda")
        cm = ops.moment_table_2d(bases)
        ops.accum_2d(Xa, ya, bases, cm, ops.split_accum_2d(acc, bases)[2], **kw)
        ops.expand_moments_2d(cm, bases, acc)
        return acc.cpu().numpy()

    auto = run(X, y)
    stated = run(X, y, raster_row_len=n2)
    scale = np.abs(auto).max()
    np.testing.assert_allclose(stated, auto, rtol=0, atol=1e-12 * scale)
    perm = rng.permutation(X.shape[0])
    lied = run(X[perm], y[perm], raster_row_len=n2)                 # not a raster at all
    np.testing.assert_allclose(lied, auto, rtol=0, atol=1e-11 * scale)
    wrong_len = run(X, y, raster_row_len=n2 // 2)                   # a raster, but not with that row length
    np.testing.assert_allclose(wrong_len, auto, rtol=0, atol=1e-11 * scale)

    kerns = [Kn.Matern32(variance=1.0, lengthscales=6.0), Kn.Matern32(variance=0.8, lengthscales=5.0)]
    model = GPR_kron((torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda().view(-1, 1)), kerns, bases, raster_shape=(n1, n2))
    ref = GPR_kron((X, y.reshape(-1, 1)), kerns, bases)
    model.likelihood.variance.assign(0.01); ref.likelihood.variance.assign(0.01)
    assert abs(model.elbo() - ref.elbo()) <= 1e-11 * abs(ref.elbo())
    g1, g2 = np.linspace(-74.5, -30.5, 41), np.linspace(20.5, 49.5, 67)
    Xg = np.stack(np.meshgrid(g1, g2, indexing="ij"), -1).reshape(-1, 2)
    m0, v0 = ref.predict_f(Xg)
    m1, v1 = model.predict_f(Xg, raster_shape=(41, 67))
    np.testing.assert_allclose(m1, m0, atol=1e-12, rtol=0)
    np.testing.assert_allclose(v1, v0, atol=1e-12, rtol=0)
    Xw = Xg[rng.permutation(Xg.shape[0])][: 41 * 60]
    m2, v2 = model.predict_f(Xw, raster_shape=(41, 60))             # wrong statement: shuffled points are no raster
    m3, v3 = ref.predict_f(Xw)
    np.testing.assert_allclose(m2, m3, atol=1e-12, rtol=0)
    np.testing.assert_allclose(v2, v3, atol=1e-12, rtol=0)
    # the cached table follows the hyper-parameters
    model.kernels[0].lengthscales.assign(4.0); ref.kernels[0].lengthscales.assign(4.0)
    m4, _ = model.predict_f(Xg[:50]); m5, _ = ref.predict_f(Xg[:50])
    np.testing.assert_allclose(m4, m5, atol=1e-12, rtol=0)
    assert np.abs(m4 - m0[:50]).max() > 1e-6

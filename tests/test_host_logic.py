"""CPU tests of the host-side logic: exact spline tables vs the golden reference tables, mesh precision rule,
parameter transforms, Kuu coefficient derivatives, shard bounds, and the 2-rank gloo all-reduce of the packed
accumulator (the only collective of the path)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tables_match_reference_golden(golden):
    from asvgp_b200 import basis as B

    g = golden("basis_eval")
    for k in range(1, 7):
        for tag, a, b in (("f32", -3.5, 10.5), ("f64", -1, 41)):
            key = "k%d_%s" % (k, tag)
            basis = getattr(B, "B%dSpline" % k)(a, b, 40)
            np.testing.assert_array_equal(basis.mesh, g[key + "_mesh"])
            for name in ("A", "B", "C", "D", "BC", "BC_grad", "BC_ggrad"):
                if key + "_" + name in g.files:
                    want = g[key + "_" + name]
                    np.testing.assert_allclose(getattr(basis, name), want, rtol=1e-13, atol=1e-13 * np.abs(want).max())
    s = golden("snelson")
    basis = B.B3Spline(-3.5, 10.5, 100)
    assert basis.delta == float(s["delta"])
    for name in ("A", "B", "C", "D", "BC", "BC_grad", "BC_ggrad", "BC_ggrad_none", "BC_none_ggrad"):
        np.testing.assert_allclose(getattr(basis, name), s["tab_" + name], rtol=1e-13, atol=1e-13 * max(np.abs(s["tab_" + name]).max(), 1e-300))


def test_cubic_interior_gram_is_the_textbook_rational():
    from fractions import Fraction

    from asvgp_b200 import _spline_tables as T

    W = T.interval_gram(3, 0)
    full = [sum(W[r + d][r] for r in range(4 - d)) for d in range(4)]
    assert full == [Fraction(151, 315), Fraction(397, 1680), Fraction(1, 42), Fraction(1, 5040)]   # basis.py:290-293
    # partition of unity on every interval, every order
    for k in range(1, 7):
        P = T.piece_coeffs(k, 0)
        col = [sum(P[r][c] for r in range(k + 1)) for c in range(k + 1)]
        assert col == [1] + [0] * k


def test_mesh_precision_rule():
    from asvgp_b200.basis import B3Spline, tf_style_linspace

    f32 = B3Spline(-3.5, 10.5, 100)
    f64 = B3Spline(-3.5, 10.5, 100, mesh_dtype="float64")
    assert f32.delta == 0.14432978630065918 and abs(f64.delta - 0.14432989690721643) < 1e-16
    ints = B3Spline(-1, 101, 1000)
    np.testing.assert_array_equal(ints.mesh, np.linspace(-1.0, 101.0, 998))
    assert tf_style_linspace(0.0, 1.0, 5).dtype == np.float64
    with pytest.raises(NameError):
        __import__("asvgp_b200.basis", fromlist=["B4Spline"]).B4Spline(0, 1, 10)


def test_parameter_transform_roundtrip_and_bounds():
    from asvgp_b200.kernels import Gaussian, Matern52, Parameter, kernel_kind

    p = Parameter(0.37)
    assert abs(p.value - 0.37) < 1e-15
    h = 1e-6
    u = p.unconstrained
    p.unconstrained = u + h; hi = p.value
    p.unconstrained = u - h; lo = p.value
    p.unconstrained = u
    assert abs((hi - lo) / (2 * h) - p.dvalue_dunconstrained()) < 1e-8
    lik = Gaussian()
    assert lik.variance.value == 1.0 and lik.variance.lower == 1e-6
    lik.variance.unconstrained = -800.0
    assert lik.variance.value >= 1e-6
    assert kernel_kind(Matern52()) == "Matern52"
    with pytest.raises(AssertionError):
        kernel_kind(object())


@pytest.mark.parametrize("kind", ["Matern12", "Matern32", "Matern52"])
def test_kuu_coefficients_match_oracle_and_derivatives(kind):
    from asvgp_b200.inducing_features import kuu_terms
    from oracle import asvgp_oracle as O

    l, v = 1.7, 0.6
    terms = kuu_terms(kind, l, v)
    want = O.kuu_coefficients(kind, l, v)
    assert {n for n, _, _ in terms} == set(want)
    h = 1e-6
    for name, c, dc in terms:
        assert abs(c - want[name]) <= 1e-15 * abs(c)
        fd = (O.kuu_coefficients(kind, l + h, v)[name] - O.kuu_coefficients(kind, l - h, v)[name]) / (2 * h)
        assert abs(dc - fd) <= 1e-8 * max(1.0, abs(fd))


def test_shard_bounds_cover_and_align():
    from asvgp_b200.dist import shard_bounds

    for n in (0, 1, 7, 100, 10**8, 10**8 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            assert all(lo % 2 == 0 for lo, hi in spans if hi > lo)


def _gloo_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from asvgp_b200.dist import allreduce_packed, is_distributed, rank_world, shard_bounds
    from oracle import asvgp_oracle as O          # stands in for the CUDA accumulate on this CPU-only box

    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert is_distributed("auto") and rank_world() == (rank, world)
    rng = np.random.default_rng(11)
    n, m, k = 20001, 50, 3
    x = np.sort(rng.uniform(0, m, n)); y = rng.standard_normal(n)
    mesh, delta = O.make_mesh(-1, m + 1, m, k)
    lo, hi = shard_bounds(n, rank, world)
    G, b, yy = O.precompute_1d(mesh, delta, k, m, x[lo:hi], y[lo:hi])
    acc = torch.from_numpy(np.concatenate([G.ravel(), b.ravel(), [yy, hi - lo]]))
    allreduce_packed(acc)
    G0, b0, yy0 = O.precompute_1d(mesh, delta, k, m, x, y)
    want = np.concatenate([G0.ravel(), b0.ravel(), [yy0, n]])
    np.testing.assert_allclose(acc.numpy(), want, rtol=1e-12, atol=1e-12)
    open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_of_packed_accumulator(tmp_path):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def _gloo_worker_2d(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from asvgp_b200 import utils
    from asvgp_b200.dist import allreduce_packed, shard_bounds
    from oracle import asvgp_oracle as O          # stands in for the CUDA accumulate on this CPU-only box

    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(12)
    n, k, ms = 6001, 3, [11, 13]
    X = np.stack([rng.uniform(.01, .99, n), rng.uniform(.01, 1.99, n)], 1)
    y = rng.standard_normal(n)
    meshes, deltas = zip(*[O.make_mesh(0, 1, ms[0], k), O.make_mesh(0, 2, ms[1], k)])

    def packed(Xs, ys):          # [G stencil | Kuf_y | sum y^2 | count]: the layout of ops.accum_size_2d
        G, b, yy = O.precompute_kron(meshes, deltas, k, ms, Xs, ys)
        return np.concatenate([utils.sparse_to_stencil(G, ms[0], ms[1], k).ravel(), b.ravel(), [yy, Xs.shape[0]]])

    lo, hi = shard_bounds(n, rank, world)
    acc = torch.from_numpy(packed(X[lo:hi], y[lo:hi]))
    allreduce_packed(acc)
    want = packed(X, y)
    np.testing.assert_allclose(acc.numpy(), want, rtol=1e-12, atol=1e-12)
    # and the stencil layout round-trips through the sparse form the reference exposes (KufKfu_sparse)
    ne = (k + 1) * (2 * k + 1)
    G = utils.stencil_to_sparse(acc.numpy()[: ne * ms[0] * ms[1]].reshape(ne, -1), ms[0], ms[1], k)
    G0, _, _ = O.precompute_kron(meshes, deltas, k, ms, X, y)
    assert abs(G - G0).max() < 1e-11
    open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_of_packed_2d_accumulator(tmp_path):
    import torch.multiprocessing as mp

    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker_2d, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()

"""GPU parity of GPR_additive (reference gpr.py:139-236) through the reference-shaped API, and of the dense SPD solver it
rides on (asvgp_dense_factor / asvgp_dense_selinv): Gram, projection, bound and predictions against the golden vectors of the
unmodified reference under the shim (rel 1e-10 / abs 1e-9), gradients against torch autograd through the dense algebra (1e-8)."""
import numpy as np
import pytest

from oracle import asvgp_oracle as O

pytestmark = pytest.mark.gpu
CASES = {"a": (("Matern32",) * 3, [(1.0, 1.0)] * 3, 1.0),
         "b": (("Matern52", "Matern12", "Matern32"), [(.7, .3), (1.3, .5), (.9, .8)], .05)}


def _model(g, tag):
    from asvgp_b200 import basis as B, kernels as Kn
    from asvgp_b200.gpr import GPR_additive

    kinds, hypers, s2 = CASES[tag]
    m = int(g["m"])
    bases = [B.B3Spline(int(a), int(b), m) for a, b in g["doms"]]
    kerns = [getattr(Kn, kind)(variance=v, lengthscales=l) for kind, (v, l) in zip(kinds, hypers)]
    model = GPR_additive((g["X"], g["y"]), kerns, bases)
    model.likelihood.variance.assign(s2)
    return model


@pytest.mark.parametrize("n", [1, 63, 64, 65, 300, 1000])
def test_dense_spd_solver_matches_numpy(cuda, n):
    import torch

    from asvgp_b200 import ops

    rng = np.random.default_rng(n)
    B_ = rng.standard_normal((n, n))
    A = B_ @ B_.T + n * np.eye(n)
    b = rng.standard_normal(n)
    ws = ops.dense_workspace(n)
    Ad, bd = torch.from_numpy(np.tril(A)).cuda(), torch.from_numpy(b).cuda()       # only the lower triangle is read
    ops.dense_factor(Ad, bd, ws)
    x, inv = ops.dense_selinv(ws)
    scal = ws.scal.cpu().numpy()
    assert scal[2] == 0
    x0 = np.linalg.solve(A, b)
    assert abs(scal[0] - np.linalg.slogdet(A)[1]) <= 1e-12 * abs(scal[0]) + 1e-12
    assert abs(scal[1] - b @ x0) <= 1e-12 * abs(b @ x0)
    np.testing.assert_allclose(x.cpu().numpy(), x0, rtol=0, atol=1e-12 * np.abs(x0).max())
    inv0 = np.linalg.inv(A)
    np.testing.assert_allclose(inv.cpu().numpy(), inv0, rtol=0, atol=1e-12 * np.abs(inv0).max())
    # not positive definite: the failing row is reported
    A2 = A.copy(); A2[n // 2, n // 2] = -1.0
    ops.dense_factor(torch.from_numpy(np.tril(A2)).cuda(), bd, ws)
    assert ws.scal.cpu().numpy()[2] > 0


def test_precompute_matches_reference(cuda, golden):
    g = golden("additive_3d")
    model = _model(g, "a")
    scale = np.abs(g["KufKfu"]).max()
    np.testing.assert_allclose(model.KufKfu, g["KufKfu"], rtol=1e-10, atol=1e-10 * scale)
    np.testing.assert_allclose(model.Kuf_y, g["Kuf_y"], rtol=1e-10, atol=1e-10 * np.abs(g["Kuf_y"]).max())
    assert abs(model.tr_yTy - float(g["tr_yTy"])) <= 1e-12 * float(g["tr_yTy"])
    assert model.num_data == g["X"].shape[0] and model.bandwidth == 3
    assert len(model.trainable_variables) == 7


@pytest.mark.parametrize("tag", ["a", "b"])
def test_elbo_and_predictions_golden(cuda, golden, tag):
    g = golden("additive_3d")
    model = _model(g, tag)
    want = float(g["elbo_" + tag])
    assert abs(model.elbo() - want) <= 1e-10 * abs(want)
    e, _ = model.elbo_and_grad()
    assert abs(e - want) <= 1e-10 * abs(want)
    mean, var = model.predict_f(g["Xs"])
    assert mean.shape == var.shape == (g["Xs"].shape[0], 1)
    np.testing.assert_allclose(mean, g["mean_" + tag], atol=1e-9, rtol=0)
    np.testing.assert_allclose(var, g["var_" + tag], atol=1e-9, rtol=0)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_gradients_match_autograd_oracle(cuda, golden, tag):
    g = golden("additive_3d")
    kinds, hypers, s2 = CASES[tag]
    model = _model(g, tag)
    elbo, grads = model.elbo_and_grad()
    m = int(g["m"])
    T = [O.static_bands(3, m, b.delta) for b in model.bases]
    e0, g0 = O.elbo_grad_additive_dense(kinds, T, g["KufKfu"], g["Kuf_y"], float(g["tr_yTy"]), g["X"].shape[0], hypers, s2)
    assert abs(elbo - e0) <= 1e-10 * abs(e0)
    got = np.array([grads[id(p)] for p in model.trainable_variables])
    np.testing.assert_allclose(got, g0, rtol=1e-8, atol=1e-8 * np.abs(g0).max())


def test_two_dimensions_ordered_input_and_optimiser(cuda):
    """D = 2, sorted first coordinate (long runs for the cross accumulation), different m per dimension; a few L-BFGS steps
    must increase the bound."""
    from asvgp_b200 import basis as B, kernels as Kn
    from asvgp_b200.gpr import GPR_additive
    from asvgp_b200.optimizers import Scipy

    rng = np.random.default_rng(3)
    n = 20000
    X = np.stack([np.sort(rng.uniform(0.01, 9.99, n)), rng.uniform(-1.99, 1.99, n)], 1)
    y = (np.sin(X[:, 0]) + 0.5 * X[:, 1] ** 2 + 0.1 * rng.standard_normal(n)).reshape(-1, 1)
    bases = [B.B3Spline(0, 10, 40), B.B3Spline(-2, 2, 17)]
    kerns = [Kn.Matern32(), Kn.Matern52()]
    model = GPR_additive((X, y), kerns, bases)
    meshes, deltas = [b.mesh for b in bases], [b.delta for b in bases]
    G0, b0, yy0 = O.precompute_additive(meshes, deltas, 3, [40, 17], X, y)
    np.testing.assert_allclose(model.KufKfu, G0, rtol=1e-10, atol=1e-10 * np.abs(G0).max())
    T = [O.static_bands(3, b.m, b.delta) for b in bases]
    Ks = [O.make_Kuu("Matern32", 1.0, 1.0, T[0]), O.make_Kuu("Matern52", 1.0, 1.0, T[1])]
    want = O.elbo_additive_dense(Ks, G0, b0, yy0, n, [1.0, 1.0], 1.0)
    e0 = model.elbo()
    assert abs(e0 - want) <= 1e-10 * abs(want)
    Scipy().minimize(model.training_loss, model.trainable_variables, options=dict(maxiter=15))
    assert model.elbo() > e0 + 100

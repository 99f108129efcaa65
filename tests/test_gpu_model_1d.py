"""GPU parity of the 1-D model through the reference-shaped API: ELBO vs golden (reference-under-shim) at rel 1e-10,
gradients vs the torch-autograd oracle at 1e-8, predictions at 1e-9 abs, the notebook known answer after L-BFGS."""
import numpy as np
import pytest

from oracle import asvgp_oracle as O

pytestmark = pytest.mark.gpu
KINDS = ("Matern12", "Matern32", "Matern52")


def _model(X, y, kind, k, a, b, m, hyp=None, **kw):
    from asvgp_b200 import basis as B, kernels as Kn
    from asvgp_b200.gpr import GPR_1d

    basis = getattr(B, "B%dSpline" % k)(a, b, m)
    kern = getattr(Kn, kind)()
    model = GPR_1d((X.reshape(-1, 1), y.reshape(-1, 1)), kern, basis, **kw)
    if hyp is not None:
        kern.variance.assign(hyp[0]); kern.lengthscales.assign(hyp[1]); model.likelihood.variance.assign(hyp[2])
    return model


@pytest.mark.parametrize("kind", KINDS)
def test_snelson_elbo_golden(cuda, golden, kind):
    g = golden("snelson")
    model = _model(g["X"], g["y"], kind, 3, -3.5, 10.5, 100)
    np.testing.assert_allclose(model.inducing_features.make_Kuu(model.kernel), g["Kuu111_" + kind], rtol=1e-12, atol=1e-13)
    assert abs(model.elbo() - float(g["elbo111_" + kind])) <= 1e-10 * abs(float(g["elbo111_" + kind]))
    model = _model(g["X"], g["y"], kind, 3, -3.5, 10.5, 100, hyp=(0.8, 1.03, 0.08))
    assert abs(model.elbo() - float(g["elbo_b_" + kind])) <= 1e-10 * abs(float(g["elbo_b_" + kind]))


@pytest.mark.parametrize("chunks", [1, 2, 3, 7])
def test_snelson_chunking_invariance(cuda, golden, chunks):
    g = golden("snelson")
    model = _model(g["X"], g["y"], "Matern32", 3, -3.5, 10.5, 100, hyp=(0.8, 1.03, 0.08), chunks=chunks)
    assert abs(model.elbo() - float(g["elbo_b_Matern32"])) <= 1e-10 * abs(float(g["elbo_b_Matern32"]))


def test_snelson_predict_golden(cuda, golden):
    g = golden("snelson")
    model = _model(g["X"], g["y"], "Matern32", 3, -3.5, 10.5, 100, hyp=tuple(g["opt_hypers"]))
    assert abs(model.elbo() - float(g["elbo_opt"])) <= 1e-10 * abs(float(g["elbo_opt"]))
    mean, var = model.predict_f(g["Xtest"])
    np.testing.assert_allclose(mean, g["pred_mean"], atol=1e-9, rtol=0)
    np.testing.assert_allclose(var, g["pred_var"], atol=1e-9, rtol=0)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_synth_elbo_and_predict_golden(cuda, golden, k):
    g = golden("synth_1d")
    key = "k%d" % k
    m = int(g[key + "_m"])
    for kind in KINDS:
        for tag, hyp in (("a", (1.0, 1.0, 0.1)), ("b", (1.3, 2.5, 0.7))):
            name = "%s_%s_elbo_%s" % (key, kind, tag)
            if name not in g.files:
                continue
            model = _model(g[key + "_x"], g[key + "_y"], kind, k, -1, m + 1, m, hyp=hyp)
            want = float(g[name])
            assert abs(model.elbo() - want) <= 1e-10 * abs(want), (kind, tag)
    kind = str(g[key + "_pred_kind"])
    model = _model(g[key + "_x"], g[key + "_y"], kind, k, -1, m + 1, m, hyp=tuple(g[key + "_pred_hypers"]))
    mean, var = model.predict_f(g[key + "_xs"])
    np.testing.assert_allclose(mean, g[key + "_mean"], atol=1e-9, rtol=0)
    np.testing.assert_allclose(var, g[key + "_var"], atol=1e-9, rtol=0)


@pytest.mark.parametrize("kind,k", [("Matern12", 1), ("Matern12", 3), ("Matern32", 2), ("Matern32", 3),
                                    ("Matern52", 3), ("Matern32", 4), ("Matern52", 5)])
@pytest.mark.parametrize("chunks", [0, 5])
def test_gradients_match_autograd_oracle(cuda, golden, kind, k, chunks):
    g = golden("synth_1d")
    key = "k%d" % k
    m = int(g[key + "_m"])
    hyp = (1.3, 2.5, 0.7)
    model = _model(g[key + "_x"], g[key + "_y"], kind, k, -1, m + 1, m, hyp=hyp, chunks=chunks)
    elbo, grads = model.elbo_and_grad()
    tables = O.static_bands(k, m, model.basis.delta)
    e0, g0 = O.elbo_grad_1d_dense(kind, tables, g[key + "_G"], g[key + "_Kuf_y"], float(g[key + "_tr_yTy"]),
                                  g[key + "_x"].shape[0], hyp[0], hyp[1], hyp[2])
    assert abs(elbo - e0) <= 1e-10 * abs(e0)
    got = np.array([grads[id(model.kernel.variance)], grads[id(model.kernel.lengthscales)],
                    grads[id(model.likelihood.variance)]])
    np.testing.assert_allclose(got, g0, rtol=1e-8, atol=1e-8 * np.abs(g0).max())


def test_multi_output_golden_and_gradients(cuda, golden):
    """y with D = 3 columns (reference gpr.py:39-44,78-87): bound vs reference-under-shim at rel 1e-10, gradients vs the
    autograd oracle at 1e-8, mean (n*, D) / var (n*, 1) at 1e-9; numpy and device inputs agree."""
    import torch
    from asvgp_b200 import basis as B, kernels as Kn
    from asvgp_b200.gpr import GPR_1d

    g = golden("multi_output_1d")
    m, k = int(g["m"]), int(g["order"])
    X, y = g["x"].reshape(-1, 1), g["y"]
    for kind in KINDS:
        for tag, hyp in (("a", (1.0, 1.0, 0.1)), ("b", (0.6, 3.5, 0.3))):
            kern = getattr(Kn, kind)()
            model = GPR_1d((X, y), kern, B.B3Spline(-1, m + 1, m))
            kern.variance.assign(hyp[0]); kern.lengthscales.assign(hyp[1]); model.likelihood.variance.assign(hyp[2])
            want = float(g["elbo_%s_%s" % (kind, tag)])
            elbo, grads = model.elbo_and_grad()
            assert abs(elbo - want) <= 1e-10 * abs(want), (kind, tag)
            np.testing.assert_allclose(model.Kuf_y, g["Kuf_y"], rtol=1e-11, atol=1e-11)
            assert abs(model.tr_yTy - float(g["tr_yTy"])) <= 1e-12 * model.tr_yTy
            tables = O.static_bands(k, m, model.basis.delta)
            e0, g0 = O.elbo_grad_1d_dense(kind, tables, model.KufKfu, g["Kuf_y"], float(g["tr_yTy"]), X.shape[0], *hyp)
            assert abs(elbo - e0) <= 1e-10 * abs(e0)
            got = np.array([grads[id(kern.variance)], grads[id(kern.lengthscales)], grads[id(model.likelihood.variance)]])
            np.testing.assert_allclose(got, g0, rtol=1e-8, atol=1e-8 * np.abs(g0).max())
    mean, var = model.predict_f(g["xs"])                      # last model: Matern52 at pred_hypers
    assert mean.shape == (50, 3) and var.shape == (50, 1)
    np.testing.assert_allclose(mean, g["mean"], atol=1e-9, rtol=0)
    np.testing.assert_allclose(var, g["var"], atol=1e-9, rtol=0)
    kern = Kn.Matern52()
    dev = GPR_1d((torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()), kern, B.B3Spline(-1, m + 1, m))
    kern.variance.assign(0.6); kern.lengthscales.assign(3.5); dev.likelihood.variance.assign(0.3)
    assert abs(dev.elbo() - elbo) <= 1e-12 * abs(elbo)


def test_medium_m_partitioned_vs_oracle(cuda):
    """C2 shape (N=1e6, M=1000): default chunking (P>1) against the SciPy banded oracle."""
    rng = np.random.default_rng(1997)
    n, m, k = 1_000_000, 1000, 3
    x = np.sort(rng.uniform(0.0, m, n))
    y = np.sin(2 * np.pi * x / 37) + 0.5 * np.sin(2 * np.pi * x / 3.1) + 0.3 * rng.standard_normal(n)
    y = (y - y.mean()) / y.std()
    for kind, hyp in (("Matern32", (1.0, 1.0, 0.1)), ("Matern52", (1.0, 1.0, 1.0)), ("Matern32", (1.0, 10.0, 0.1))):
        model = _model(x, y, kind, k, -1, m + 1, m, hyp=hyp)
        tables = O.static_bands(k, m, model.basis.delta)
        G0, b0, yy0 = O.precompute_1d_chunked(model.basis.mesh, model.basis.delta, k, m, x, y)
        Kuu = O.make_Kuu(kind, hyp[1], hyp[0], tables)
        want = O.elbo_1d(Kuu, G0, b0, yy0, n, hyp[0], hyp[2])
        assert abs(model.elbo() - want) <= 1e-10 * abs(want), (kind, hyp)
        xs = rng.uniform(1.0, m - 1.0, 500)
        mean, var = model.predict_f(xs.reshape(-1, 1))
        mean0, var0 = O.predict_1d(model.basis.mesh, model.basis.delta, k, m, Kuu, G0, b0, hyp[0], hyp[2], xs)
        np.testing.assert_allclose(mean, mean0, atol=1e-9, rtol=0)
        np.testing.assert_allclose(var, var0, atol=1e-9, rtol=0)


def test_snelson_notebook_known_answer(cuda, golden):
    """L-BFGS-B from GPflow's defaults must land on the reference notebook's stored ELBO (example.ipynb cell 3)."""
    from asvgp_b200.optimizers import Scipy

    g = golden("snelson")
    model = _model(g["X"], g["y"], "Matern32", 3, -3.5, 10.5, 100)
    Scipy().minimize(model.training_loss, model.trainable_variables)
    elbo = model.elbo()
    assert abs(elbo - float(g["notebook_elbo"])) < 1e-6
    assert elbo < float(g["notebook_exact_gp"])


@pytest.mark.parametrize("m,chunks", [(100, 0), (3000, 0), (10000, 0), (10000, 16), (10000, 512), (10000, 256), (700, 512)])
def test_split_bound_equals_single_call(cuda, m, chunks):
    """asvgp_kuu_chain_1d (side stream) + asvgp_elbo_grad_1d_prepared give the bits of the one-call asvgp_elbo_grad_1d, also
    when the Kuu chain is launched before the accumulate it overlaps with (single CTA, 2/4/8-CTA cluster layouts)."""
    import torch

    from asvgp_b200 import basis as B, kernels as Kn, ops
    from asvgp_b200.inducing_features import SplineFeatures1D

    rng = np.random.default_rng(5)
    n = 400000
    x = np.sort(rng.uniform(0.0, m, n))
    y = np.sin(x / 7.0) + 0.1 * rng.standard_normal(n)
    basis = B.B3Spline(-1, m + 1, m)
    kern = Kn.Matern52(variance=1.2, lengthscales=1.7)
    feats = SplineFeatures1D(kern, basis)
    xd, yd = ops.to_device(x), ops.to_device(y)
    Kuu, dKuu = feats.make_Kuu_device(kern, want_grad=True)
    acc = ops.accum_1d(xd, yd, basis)
    one = ops.elbo_grad_1d_single_stream(Kuu, dKuu, acc, basis, 1.2, 0.3, chunks=chunks).cpu().numpy()
    forked = ops.elbo_grad_1d(Kuu, dKuu, acc, basis, 1.2, 0.3, chunks=chunks).cpu().numpy()
    # the order the bench uses: Kuu chain first, accumulate next to it, then the P chains
    kuu = ops.kuu_chain_1d(Kuu, dKuu, basis, chunks=chunks, gate=True)
    acc2 = ops.accum_1d(xd, yd, basis)
    early = ops.elbo_grad_1d(Kuu, dKuu, acc2, basis, 1.2, 0.3, chunks=chunks, kuu=kuu).cpu().numpy()
    one2 = ops.elbo_grad_1d_single_stream(Kuu, dKuu, acc2, basis, 1.2, 0.3, chunks=chunks).cpu().numpy()   # acc2 != acc bitwise (fp64 REDs)
    torch.cuda.synchronize()
    assert one[8] == 0
    np.testing.assert_array_equal(forked[:9], one[:9])
    assert forked[15] == one[15]
    # beside a streaming kernel the Kuu chain takes a smaller cluster (another chunking of the same sweeps: last-bit differences)
    np.testing.assert_allclose(early[:8], one2[:8], rtol=1e-11)
    np.testing.assert_allclose(early[15], one2[15], rtol=1e-9)
    assert early[8] == 0
    # and the trace term itself against the dense algebra
    G, b, scal = [t.cpu().numpy() for t in ops.split_accum_1d(acc, basis)]
    if m <= 3000:
        from scipy.linalg import solveh_banded

        Kb = Kuu.cpu().numpy()
        dense = lambda band: sum(np.diag(band[d, : m - d], -d) + (np.diag(band[d, : m - d], d) if d else 0) for d in range(4))
        tr = np.trace(solveh_banded(Kb, dense(G), lower=True))
        assert abs(one[7] - tr) <= 1e-9 * abs(tr)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("m", [300, 3000, 30000])
def test_bound_layouts_agree(cuda, k, m):
    """The default layout of the chains (single CTA, or 2/4/8-CTA clusters where 128 separator nodes per CTA fit in shared
    memory) against a 7-chunk single-CTA evaluation of the same bound, for every spline order."""
    from asvgp_b200 import basis as B, kernels as Kn, ops
    from asvgp_b200.inducing_features import SplineFeatures1D

    rng = np.random.default_rng(11 * k + m)
    n = 200000
    x = np.sort(rng.uniform(0.0, m, n))
    y = np.cos(x / 9.0) + 0.2 * rng.standard_normal(n)
    basis = getattr(B, "B%dSpline" % k)(-1, m + 1, m)
    kind = {1: "Matern12", 2: "Matern32", 6: "Matern32"}.get(k, "Matern52")      # B6 has no BC_ggrad table (as in the reference)
    kern = getattr(Kn, kind)(variance=0.9, lengthscales=2.2)
    feats = SplineFeatures1D(kern, basis)
    Kuu, dKuu = feats.make_Kuu_device(kern, want_grad=True)
    acc = ops.accum_1d(ops.to_device(x), ops.to_device(y), basis)
    a = ops.elbo_grad_1d(Kuu, dKuu, acc, basis, 0.9, 0.4).cpu().numpy()
    b = ops.elbo_grad_1d(Kuu, dKuu, acc, basis, 0.9, 0.4, chunks=7).cpu().numpy()
    assert a[8] == 0 and b[8] == 0
    np.testing.assert_allclose(a[:8], b[:8], rtol=1e-9)
    np.testing.assert_allclose(a[15], b[15], rtol=1e-7, atol=1e-7 * abs(a[7]))

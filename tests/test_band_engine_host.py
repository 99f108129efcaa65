"""CPU check of the partitioned banded engine (asvgp_b200/csrc/band_engine.cuh) compiled for the host by g++.
The same templates run inside the CUDA kernels of banded_1d.cu; this test pins their algebra — log-det, quadratic
form, solve, Takahashi band and the forward-mode tangents — against dense numpy for every order and for P = 1 .. many
chunks.  Test infrastructure only: the product never runs this code on the CPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_harness.cpp")
OUT = os.path.join(ROOT, "tests", "_build", "libhostcheck.so")


@pytest.fixture(scope="module")
def hh():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC] + [os.path.join(ROOT, "asvgp_b200", "csrc", f) for f in ("band_engine.cuh", "dual.cuh", "common.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", SRC, "-o", OUT])
    lib = ctypes.CDLL(OUT)
    lib.hh_chain.restype = ctypes.c_int
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def random_spd_band(rng, M, K, cond_boost=0.5):
    """Strictly diagonally dominant => SPD."""
    band = rng.standard_normal((K + 1, M))
    for d in range(1, K + 1):
        band[d, M - d:] = 0.0
    rowsum = np.zeros(M)
    for d in range(1, K + 1):
        rowsum[: M - d] += np.abs(band[d, : M - d])
        rowsum[d:] += np.abs(band[d, : M - d])
    band[0] = rowsum + cond_boost + np.abs(band[0])
    return band


def dense(band):
    K, M = band.shape[0] - 1, band.shape[1]
    A = np.zeros((M, M))
    for d in range(K + 1):
        i = np.arange(M - d)
        A[i + d, i] = band[d, : M - d]
        A[i, i + d] = band[d, : M - d]
    return A


def to_band(A, K):
    M = A.shape[0]
    out = np.zeros((K + 1, M))
    for d in range(K + 1):
        out[d, : M - d] = np.diagonal(A, -d)
    return out


def run(lib, band, dband, rhs, P, dual):
    K, M = band.shape[0] - 1, band.shape[1]
    nt = 2 if dual else 1
    scal = np.zeros(8)
    x = np.zeros(nt * M)
    sig = np.zeros(nt * (K + 1) * M)
    band, dband, rhs = [np.ascontiguousarray(a, dtype=np.float64) for a in (band, dband, rhs)]
    info = lib.hh_chain(M, K, P, int(dual), _ptr(band), _ptr(dband), _ptr(rhs), _ptr(scal), _ptr(x), _ptr(sig))
    return info, scal, x.reshape(nt, M), sig.reshape(nt, K + 1, M)


@pytest.mark.parametrize("K", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("M,P", [(40, 1), (57, 2), (200, 5), (203, 7), (1000, 31), (64, 64)])
def test_chain_matches_dense(hh, K, M, P):
    rng = np.random.default_rng(1000 * K + M + P)
    band = random_spd_band(rng, M, K)
    dband = rng.standard_normal((K + 1, M))
    for d in range(1, K + 1):
        dband[d, M - d:] = 0.0
    rhs = rng.standard_normal(M)
    A, dA = dense(band), dense(dband)
    Ainv = np.linalg.inv(A)
    x0 = Ainv @ rhs
    for dual in (False, True):
        info, scal, x, sig = run(hh, band, dband, rhs, P, dual)
        assert info == 0
        np.testing.assert_allclose(scal[0], np.linalg.slogdet(A)[1], rtol=1e-12)
        np.testing.assert_allclose(scal[1], rhs @ x0, rtol=1e-11)
        np.testing.assert_allclose(x[0], x0, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(sig[0], to_band(Ainv, K), rtol=1e-10, atol=1e-13)
        if dual:
            np.testing.assert_allclose(scal[2], np.trace(Ainv @ dA), rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(scal[3], -x0 @ dA @ x0, rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(x[1], -Ainv @ dA @ x0, rtol=1e-9, atol=1e-11)
            np.testing.assert_allclose(sig[1], to_band(-Ainv @ dA @ Ainv, K), rtol=1e-9, atol=1e-12)


def test_chain_reports_non_positive_pivot(hh):
    rng = np.random.default_rng(3)
    band = random_spd_band(rng, 120, 3)
    band[0, 77] = -5.0
    info, *_ = run(hh, band, band * 0, rng.standard_normal(120), 4, False)
    assert info != 0


@pytest.mark.parametrize("K", [1, 2, 3, 4, 5, 6])
def test_pieces_and_locate_match_oracle(hh, K):
    from oracle import asvgp_oracle as O

    rng = np.random.default_rng(K)
    t = np.concatenate([rng.uniform(0, 1, 200), [0.0, 1.0, 0.5]])
    out = np.zeros((K + 1, t.shape[0]))
    hh.hh_pieces(K, t.shape[0], _ptr(t), _ptr(out))
    np.testing.assert_allclose(out, O.pieces(K, t * 0.37, 0.37), rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(out.sum(0), 1.0, rtol=1e-14)
    for mesh_dtype in ("float32", "float64"):
        mesh, delta = O.make_mesh(-3.5, 10.5, 100, K, mesh_dtype)
        x = np.concatenate([rng.uniform(-3.5, 10.5, 5000), mesh, mesh + 1e-12, mesh - 1e-12, [-4.0, 11.0]])
        got = np.zeros(x.shape[0], dtype=np.int32)
        hh.hh_locate(_ptr(mesh), mesh.shape[0], x.shape[0], _ptr(x), _ptr(got))
        want = np.minimum(O.locate(mesh, x)[0], mesh.shape[0] - 2)
        np.testing.assert_array_equal(got, want)

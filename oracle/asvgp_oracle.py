"""CPU oracle: a numpy/SciPy restatement of the reference's algorithm for the ASVGP hot path.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference
arm may import this module; nothing under `asvgp_b200/` does (the product path has no CPU fallback).

Parity pinning: the reference ships no tests, one known answer (the Snelson notebook ELBO,
experiments/snelson/example.ipynb cell 3 = -60.8356263428725).  This restatement is pinned (a) against that
value and (b) against outputs of the *unmodified reference files run under numpy stand-ins* in the build
container (`oracle/make_golden.py` -> tests/golden/*.npz; checked by tests/test_oracle_golden.py).

Every function cites the reference file:line it follows (paths relative to /root/reference).  It deliberately
shares no code with asvgp_b200/: basis pieces come from the Cox-de Boor recursion in x-units, Gram tables from
Gauss-Legendre quadrature, the O(N) precompute uses the same SciPy sparse calls the reference makes.
"""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

SQRT3, SQRT5 = np.sqrt(3.0), np.sqrt(5.0)


# ---------------------------------------------------------------------------------------------------------------------
# a1: mesh  (asvgp/basis.py:13-18)
# ---------------------------------------------------------------------------------------------------------------------
def make_mesh(a, b, m, k, mesh_dtype="tf"):
    """mesh = cast(tf.linspace(a, b, m-(k-1)), f64); delta = mesh[1]-mesh[0]  (basis.py:17-18).
    'tf' emulates TF dtype inference: Python floats -> float32 linspace, ints -> float64."""
    n = m - (k - 1)
    if mesh_dtype == "tf":
        mesh_dtype = "float64" if all(isinstance(v, (int, np.integer)) for v in (a, b)) else "float32"
    if mesh_dtype == "float64":
        mesh = np.linspace(float(a), float(b), n)
    else:
        f = np.float32
        step = f(f(f(b) - f(a)) / f(n - 1))
        mesh = np.empty(n, dtype=f)
        mesh[0], mesh[-1] = f(a), f(b)
        mesh[1:-1] = f(a) + step * np.arange(1, n - 1, dtype=f)
        mesh = mesh.astype(np.float64)
    return mesh, float(mesh[1] - mesh[0])


# ---------------------------------------------------------------------------------------------------------------------
# a2: basis evaluation  (asvgp/basis.py:51-80 and the per-order `_evaluate*`)
# ---------------------------------------------------------------------------------------------------------------------
def locate(mesh, x):
    """idx = relu(searchsorted_left(mesh, x) - 1), u = mesh[idx]  (basis.py:58-59)."""
    idx = np.maximum(np.searchsorted(mesh, x, side="left") - 1, 0).astype(np.int64)
    return idx, mesh[idx]


def pieces(k, s, delta, dx=0):
    """Values of the dx-th x-derivative of the k+1 non-zero degree-k B-splines at offset s = x - u from the
    left knot; row r of the result belongs to basis row idx + r.  Cox-de Boor on the uniform knots u + i*delta
    (the reference's `_evaluate`, `_evaluate_grad`, ... are expanded forms of the same polynomials: its b_i is
    row k+1-i here, basis.py:72)."""
    s = np.asarray(s, dtype=np.float64)
    # polynomial coefficient arrays in s for each piece, built by the recursion on polynomials
    P = [np.poly1d([1.0])]                      # degree 0: one piece (the interval itself)
    for d in range(1, k + 1):
        new = []
        for r in range(d + 1):
            # function with support starting (d - r) intervals to the left of u: knots t_0 = -(d-r)*delta
            t0 = -(d - r) * delta
            acc = np.poly1d([0.0])
            if r - 1 >= 0:                      # left parent (same support start), degree d-1
                acc = acc + np.poly1d([1.0, -t0]) / (d * delta) * P[r - 1]
            if r <= d - 1:                      # right parent (support start one interval later)
                acc = acc + np.poly1d([-1.0, t0 + (d + 1) * delta]) / (d * delta) * P[r]
            new.append(acc)
        P = new
    out = np.empty((k + 1,) + s.shape)
    for r in range(k + 1):
        p = P[r]
        for _ in range(dx):
            p = p.deriv()
        out[r] = p(s)
    return out


def make_Kuf(mesh, delta, k, m, X, dx=0):
    """Sparse (m, n) Kuf exactly as the reference assembles it (basis.py:72-76, inducing_features.py:47-48)."""
    x = np.asarray(X, dtype=np.float64).reshape(-1)
    n = x.shape[0]
    idx, u = locate(mesh, x)
    vals = pieces(k, x - u, delta, dx)
    rows = (idx[None, :] + np.arange(k + 1)[:, None]).reshape(-1)
    cols = np.tile(np.arange(n), k + 1)
    return sp.csr_matrix((vals.reshape(-1), (rows, cols)), shape=(m, n))


# ---------------------------------------------------------------------------------------------------------------------
# a3/a4: static Gram and boundary bands  (asvgp/basis.py:31-45, 82-114 and the l2_*_inner_product tables)
# ---------------------------------------------------------------------------------------------------------------------
def _unit_pieces_exact(k):
    """Coefficient lists (ascending powers of t, fractions.Fraction) of the k+1 non-zero degree-k B-spline pieces on
    one knot interval of a UNIT mesh; entry r belongs to basis row idx + r.  Same Cox-de Boor recursion as `pieces`,
    in exact rational arithmetic."""
    from fractions import Fraction as F

    def mul(p, q):
        out = [F(0)] * (len(p) + len(q) - 1)
        for i, a in enumerate(p):
            for j, b in enumerate(q):
                out[i + j] += a * b
        return out

    def add(p, q):
        n = max(len(p), len(q))
        return [(p[i] if i < len(p) else F(0)) + (q[i] if i < len(q) else F(0)) for i in range(n)]

    P = [[F(1)]]
    for d in range(1, k + 1):
        new = []
        for r in range(d + 1):
            t0 = F(-(d - r))
            acc = [F(0)]
            if r - 1 >= 0:
                acc = add(acc, mul([-t0 / d, F(1, d)], P[r - 1]))
            if r <= d - 1:
                acc = add(acc, mul([(t0 + d + 1) / d, F(-1, d)], P[r]))
            new.append(acc)
        P = new
    return P


def _poly_deriv(p, times):
    for _ in range(times):
        p = [i * c for i, c in enumerate(p)][1:] or [0 * p[0]]
    return p


def gram_band(k, m, delta, q):
    """Lower band (k+1, m) of int_a^b phi_i^(q) phi_j^(q) dx, band[d, j] = S[j+d, j]; truncated edge functions
    (basis.py:31-45).  The per-interval Gram of the unit-mesh pieces is an exact rational (the reference tabulates the
    same rationals in closed form, basis.py:143-798); the contributions of the intervals an entry touches are summed as
    rationals and scaled by delta^(1-2q) with ONE rounding, so every entry is the correctly rounded value.  (A
    Gauss-Legendre version of this function was good to 4e-15 only, and at l / delta ~ 18 the bound moves by 1e-10
    relative per ulp of Kuu.)"""
    from fractions import Fraction as F

    P = [_poly_deriv(p, q) for p in _unit_pieces_exact(k)]
    W = [[F(0)] * (k + 1) for _ in range(k + 1)]
    for r in range(k + 1):
        for t in range(r + 1):
            acc = F(0)
            for i, a in enumerate(P[r]):
                for j, b in enumerate(P[t]):
                    acc += a * b / (i + j + 1)                  # int_0^1 t^(i+j) dt
            W[r][t] = acc
    scale = F(float(delta)) ** (1 - 2 * q)
    exact = {}
    band = np.zeros((k + 1, m))
    for c in range(m - k):                                      # interval c touches rows c..c+k
        for r in range(k + 1):
            for t in range(r + 1):
                key = (r - t, c + t)
                exact[key] = exact.get(key, F(0)) + W[r][t]
    cache = {}
    for (d, j), v in exact.items():
        if v not in cache:
            cache[v] = float(v * scale)
        band[d, j] = cache[v]
    return band


def boundary_band(k, m, delta, dx):
    """make_boundary_conditions(dx) for dx=0,1,2 (basis.py:82-114): outer product of the first k boundary values
    at x=a, d-th diagonal written at both ends of row d, last row zero.  dx=3,4 vanish for m>2k.
    Exact rationals scaled by delta^(-2 dx) with one rounding, like gram_band."""
    from fractions import Fraction as F

    band = np.zeros((k + 1, m))
    if dx in (3, 4):
        return band
    v = [_poly_deriv(p, dx)[0] for p in _unit_pieces_exact(k)][:k]      # values of the dx-th derivative at t = 0
    scale = F(float(delta)) ** (-2 * dx)
    for d in range(k):
        l = np.array([float(v[d + i] * v[i] * scale) for i in range(k - d)])
        band[d, : k - d] = l
        band[d, m - d - (k - d): m - d] = l
    return band


def static_bands(k, m, delta):
    t = {n: gram_band(k, m, delta, q) for q, n in enumerate("ABCD") if q <= k}
    t["BC"] = boundary_band(k, m, delta, 0)
    t["BC_grad"] = boundary_band(k, m, delta, 1)
    t["BC_ggrad"] = boundary_band(k, m, delta, 2)
    return t


# ---------------------------------------------------------------------------------------------------------------------
# a5: Kuu  (asvgp/inducing_features.py:12-44)
# ---------------------------------------------------------------------------------------------------------------------
def kuu_coefficients(kind, ell, var):
    """{table name: coefficient} with Kuu = sum coeff * table  (inducing_features.py:16-44)."""
    if kind == "Matern12":
        return {"A": 1 / (2 * ell * var), "B": ell / (2 * var), "BC": 1 / (2 * var)}
    if kind == "Matern32":
        return {"A": SQRT3 / (4 * ell * var), "B": ell / (2 * SQRT3 * var), "C": ell**3 / (12 * SQRT3 * var),
                "BC": 1 / (2 * var), "BC_grad": ell**2 / (2 * var)}
    if kind == "Matern52":
        return {"A": 3 * SQRT5 / (16 * ell * var), "B": 9 * ell / (16 * SQRT5 * var),
                "C": 9 * ell**3 / (80 * SQRT5 * var), "D": 3 * ell**5 / (400 * SQRT5 * var),
                "BC": 9 / (16 * var), "BC_grad": 3 * ell**2 / (10 * var), "BC_ggrad": 9 * ell**4 / (400 * var)}
    raise ValueError(kind)


def make_Kuu(kind, ell, var, tables):
    return sum(c * tables[n] for n, c in kuu_coefficients(kind, ell, var).items())


# ---------------------------------------------------------------------------------------------------------------------
# a7: O(N) precompute  (asvgp/gpr.py:39-44, asvgp/utils.py:24-30)
# ---------------------------------------------------------------------------------------------------------------------
def sparse_to_band(K_sparse, bw):
    m = K_sparse.shape[0]
    band = np.zeros((bw + 1, m))
    for d in range(bw + 1):
        band[d, : m - d] = K_sparse.diagonal(k=-d)
    return band


def precompute_1d(mesh, delta, k, m, X, y):
    """Kuf_y = Kuf@y, KufKfu = band(Kuf@Kuf.T), tr_yTy = sum(y^2) with the reference's own SciPy sparse calls
    (gpr.py:40-44).  This function is also the 1-thread CPU baseline of bench.py."""
    Kuf = make_Kuf(mesh, delta, k, m, X)
    y = np.asarray(y, dtype=np.float64).reshape(Kuf.shape[1], -1)
    Kuf_y = Kuf @ y
    G = sparse_to_band(Kuf @ Kuf.T, k)
    return G, Kuf_y, float(np.sum(np.square(y)))


def precompute_1d_chunked(mesh, delta, k, m, X, y, chunk=1 << 20):
    """Same sums accumulated chunk by chunk (bounded memory) — used for large-N parity checks."""
    x = np.asarray(X, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(x.shape[0], -1)
    G = np.zeros((k + 1, m))
    b = np.zeros((m, y.shape[1]))
    yy = 0.0
    for s in range(0, x.shape[0], chunk):
        g, bb, t = precompute_1d(mesh, delta, k, m, x[s:s + chunk], y[s:s + chunk])
        G += g
        b += bb
        yy += t
    return G, b, yy


# ---------------------------------------------------------------------------------------------------------------------
# banded helpers (stand for banded_matrices ops at gpr.py:56-75)
# ---------------------------------------------------------------------------------------------------------------------
def band_to_dense_sym(band):
    k, m = band.shape[0] - 1, band.shape[1]
    A = np.zeros((m, m))
    for d in range(k + 1):
        i = np.arange(m - d)
        A[i + d, i] = band[d, : m - d]
        A[i, i + d] = band[d, : m - d]
    return A


def _c_oracle():
    """oracle/_build/liboracle.so (plain C, built by oracle/Makefile) if present, else None."""
    global _C_LIB
    if _C_LIB is False:
        import ctypes
        import os

        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liboracle.so")
        _C_LIB = ctypes.CDLL(path) if os.path.exists(path) else None
    return _C_LIB


_C_LIB = False


def takahashi_band(L):
    """Lower band of (L L^T)^-1 from the lower band of L (banded.inverse_from_cholesky_band, gpr.py:59).
    Uses the plain-C twin (oracle/band_ref.c) when built, else the same recursion in Python."""
    k, m = L.shape[0] - 1, L.shape[1]
    S = np.zeros((k + 1, m))
    lib = _c_oracle()
    if lib is not None:
        import ctypes

        Lc = np.ascontiguousarray(L, dtype=np.float64)
        lib.takahashi_band(Lc.ctypes.data_as(ctypes.c_void_p), k, m, S.ctypes.data_as(ctypes.c_void_p))
        return S
    for j in range(m - 1, -1, -1):
        ljj = L[0, j]
        hi = min(m - 1, j + k)
        for i in range(hi, j - 1, -1):
            acc = 0.0
            for r in range(j + 1, hi + 1):
                a, c = (r, i) if r >= i else (i, r)
                acc += L[r - j, j] * S[a - c, c]
            S[i - j, j] = (1.0 / ljj**2 if i == j else 0.0) - acc / ljj
    return S


# ---------------------------------------------------------------------------------------------------------------------
# a8: collapsed ELBO  (asvgp/gpr.py:49-89)
# ---------------------------------------------------------------------------------------------------------------------
def elbo_1d(Kuu, G, Kuf_y, tr_yTy, n, var, sigma2, return_terms=False):
    """The 7-term bound of gpr.py:81-87 from lower bands (SciPy LAPACK band routines)."""
    k = Kuu.shape[0] - 1
    D = Kuf_y.shape[1]
    L_Kuu = sla.cholesky_banded(Kuu, lower=True)                     # gpr.py:56
    log_det_Kuu = np.sum(np.log(np.square(L_Kuu[0])))                # gpr.py:57
    Kuu_inv = takahashi_band(L_Kuu)                                  # gpr.py:59
    w = np.full((k + 1, 1), 2.0)
    w[0] = 1.0
    trace_term = np.sum(w * Kuu_inv * G)                             # gpr.py:60-70 (diag of band x band product)
    P = G / sigma2 + Kuu                                             # gpr.py:72
    L_P = sla.cholesky_banded(P, lower=True)                         # gpr.py:73
    log_det_P = np.sum(np.log(np.square(L_P[0])))                    # gpr.py:74
    c = sla.solve_banded((k, 0), L_P, Kuf_y) / sigma2                # gpr.py:75
    ND = float(n * D)
    elbo = -0.5 * ND * np.log(2 * np.pi * sigma2)                    # gpr.py:81
    elbo -= 0.5 * D * log_det_P
    elbo += 0.5 * D * log_det_Kuu
    elbo -= 0.5 * tr_yTy / sigma2
    elbo += 0.5 * np.sum(np.square(c))
    elbo -= 0.5 * n * var / sigma2                                   # gpr.py:86 (K_diag = var per point)
    elbo += 0.5 * trace_term / sigma2
    if return_terms:
        return elbo, dict(log_det_Kuu=log_det_Kuu, log_det_P=log_det_P, trace=trace_term, quad=np.sum(np.square(c)))
    return elbo


def elbo_grad_1d_dense(kind, tables, G, Kuf_y, tr_yTy, n, var, ell, sigma2):
    """ELBO and d/d(var, ell, sigma2) by torch-fp64 autograd through a dense restatement of gpr.py:49-89 — the
    gradient oracle (the reference gets these from TF reverse mode through the banded ops)."""
    import torch

    th = torch.tensor([var, ell, sigma2], dtype=torch.float64, requires_grad=True)
    v, l, s2 = th[0], th[1], th[2]
    dense = {nme: torch.from_numpy(band_to_dense_sym(t)) for nme, t in tables.items()}
    r3, r5 = float(SQRT3), float(SQRT5)
    if kind == "Matern12":
        co = {"A": 1 / (2 * l * v), "B": l / (2 * v), "BC": 1 / (2 * v)}
    elif kind == "Matern32":
        co = {"A": r3 / (4 * l * v), "B": l / (2 * r3 * v), "C": l**3 / (12 * r3 * v), "BC": 1 / (2 * v),
              "BC_grad": l**2 / (2 * v)}
    else:
        co = {"A": 3 * r5 / (16 * l * v), "B": 9 * l / (16 * r5 * v), "C": 9 * l**3 / (80 * r5 * v),
              "D": 3 * l**5 / (400 * r5 * v), "BC": 9 / (16 * v), "BC_grad": 3 * l**2 / (10 * v),
              "BC_ggrad": 9 * l**4 / (400 * v)}
    Kuu = sum(c * dense[nme] for nme, c in co.items())
    Gd = torch.from_numpy(band_to_dense_sym(G))
    b = torch.from_numpy(np.asarray(Kuf_y, dtype=np.float64))
    D = b.shape[1]
    LK = torch.linalg.cholesky(Kuu)
    P = Kuu + Gd / s2
    LP = torch.linalg.cholesky(P)
    c = torch.linalg.solve_triangular(LP, b, upper=False) / s2
    elbo = (-0.5 * n * D * torch.log(2 * np.pi * s2) - D * torch.log(torch.diagonal(LP)).sum()
            + D * torch.log(torch.diagonal(LK)).sum() - 0.5 * tr_yTy / s2 + 0.5 * (c**2).sum()
            - 0.5 * n * v / s2 + 0.5 * torch.trace(torch.cholesky_solve(Gd, LK)) / s2)
    elbo.backward()
    return float(elbo.detach()), th.grad.numpy().copy()


# ---------------------------------------------------------------------------------------------------------------------
# a10: 1-D predictor  (asvgp/gpr.py:91-136)
# ---------------------------------------------------------------------------------------------------------------------
def predict_1d(mesh, delta, k, m, Kuu, G, Kuf_y, var, sigma2, Xnew):
    """mean = Kus^T P^-1 Kuf_y / sigma2, var = v + diag(Kus^T P^-1 Kus) - diag(Kus^T Kuu^-1 Kus)
    (gpr.py:96-118; CHOLMOD calls replaced by LAPACK band solves)."""
    P = G / sigma2 + Kuu
    cP = sla.cholesky_banded(P, lower=True)
    cK = sla.cholesky_banded(Kuu, lower=True)
    alpha = sla.cho_solve_banded((cP, True), Kuf_y) / sigma2
    Kus = make_Kuf(mesh, delta, k, m, Xnew)
    mean = Kus.T @ alpha
    Kd = Kus.toarray()
    v = var + np.sum(Kd * sla.cho_solve_banded((cP, True), Kd), axis=0) \
        - np.sum(Kd * sla.cho_solve_banded((cK, True), Kd), axis=0)
    return mean, v.reshape(-1, 1)


# ---------------------------------------------------------------------------------------------------------------------
# a11-a15: Kronecker model  (asvgp/kronecker.py:7-33, asvgp/gpr.py:240-359, asvgp/utils.py:45-57)
# ---------------------------------------------------------------------------------------------------------------------
def khatri_rao_rows(A, B):
    """Kuf[i1*m2+i2, n] = A[i1,n]*B[i2,n] (kronecker.py:24-33: sparse_repeats x sparse_tile, elementwise)."""
    A, B = sp.csc_matrix(A), sp.csc_matrix(B)
    m1, m2, n = A.shape[0], B.shape[0], A.shape[1]
    rows, cols, data = [], [], []
    for j in range(n):
        a_r, a_v = A.indices[A.indptr[j]:A.indptr[j + 1]], A.data[A.indptr[j]:A.indptr[j + 1]]
        b_r, b_v = B.indices[B.indptr[j]:B.indptr[j + 1]], B.data[B.indptr[j]:B.indptr[j + 1]]
        rows.append((a_r[:, None] * m2 + b_r[None, :]).reshape(-1))
        data.append((a_v[:, None] * b_v[None, :]).reshape(-1))
        cols.append(np.full(rows[-1].shape, j))
    return sp.csr_matrix((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))), shape=(m1 * m2, n))


def precompute_kron(meshes, deltas, k, ms, X, y, chunk=200000):
    """Kuf_y and sparse KufKfu of GPR_kron.__init__ (gpr.py:268-274), accumulated in chunks."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    M = ms[0] * ms[1]
    G = sp.csr_matrix((M, M))
    b = np.zeros((M, 1))
    for s in range(0, X.shape[0], chunk):
        K1 = make_Kuf(meshes[0], deltas[0], k, ms[0], X[s:s + chunk, 0])
        K2 = make_Kuf(meshes[1], deltas[1], k, ms[1], X[s:s + chunk, 1])
        # vectorised Khatri-Rao (every column has exactly (k+1) non-zeros per factor)
        n = K1.shape[1]
        K1c, K2c = K1.tocsc(), K2.tocsc()
        K1c.sort_indices(); K2c.sort_indices()
        r1 = K1c.indices.reshape(n, k + 1); v1 = K1c.data.reshape(n, k + 1)
        r2 = K2c.indices.reshape(n, k + 1); v2 = K2c.data.reshape(n, k + 1)
        rows = (r1[:, :, None].astype(np.int64) * ms[1] + r2[:, None, :]).reshape(-1)
        data = (v1[:, :, None] * v2[:, None, :]).reshape(-1)
        cols = np.repeat(np.arange(n), (k + 1) ** 2)
        Kuf = sp.csr_matrix((data, (rows, cols)), shape=(M, n))
        G = G + Kuf @ Kuf.T
        b += Kuf @ y[s:s + chunk]
    return G.tocsr(), b, float(np.sum(np.square(y)))


def elbo_kron_dense(Kuu_bands, G_sparse, Kuf_y, tr_yTy, n, variances, sigma2):
    """GPR_kron.elbo (gpr.py:282-308) with dense M^2 x M^2 algebra — small sizes only."""
    Ks = [band_to_dense_sym(b) for b in Kuu_bands]
    Kuu = np.kron(Ks[0], Ks[1])                                   # utils.py:45-51
    L_Kuu = np.kron(np.linalg.cholesky(Ks[0]), np.linalg.cholesky(Ks[1]))
    Gd = G_sparse.toarray()
    P = Kuu + Gd / sigma2
    L_P = np.linalg.cholesky(P)
    c = sla.solve_triangular(L_P, Kuf_y, lower=True) / sigma2
    elbo = -0.5 * n * np.log(2 * np.pi * sigma2)
    elbo -= np.sum(np.log(np.diag(L_P)))
    elbo += np.sum(np.log(np.diag(L_Kuu)))
    elbo -= 0.5 * tr_yTy / sigma2
    elbo += 0.5 * np.sum(np.square(c))
    elbo -= 0.5 * n * np.prod(variances) / sigma2
    elbo += 0.5 * np.trace(sla.cho_solve((L_Kuu, True), Gd)) / sigma2
    return elbo


def predict_kron_dense(meshes, deltas, k, ms, Kuu_bands, G_sparse, Kuf_y, variances, sigma2, Xnew):
    """GPR_kron.predict_f (gpr.py:310-334), dense."""
    Ks = [band_to_dense_sym(b) for b in Kuu_bands]
    Kuu = np.kron(Ks[0], Ks[1])
    P = Kuu + G_sparse.toarray() / sigma2
    K1 = make_Kuf(meshes[0], deltas[0], k, ms[0], Xnew[:, 0]).toarray()
    K2 = make_Kuf(meshes[1], deltas[1], k, ms[1], Xnew[:, 1]).toarray()
    Kus = (K1[:, None, :] * K2[None, :, :]).reshape(ms[0] * ms[1], -1)
    cP = sla.cho_factor(P, lower=True)
    alpha = sla.cho_solve(cP, Kuf_y) / sigma2
    mean = Kus.T @ alpha
    var = np.prod(variances) + np.sum(Kus * sla.cho_solve(cP, Kus), axis=0) \
        - np.sum(Kus * np.linalg.solve(Kuu, Kus), axis=0)
    return mean, var.reshape(-1, 1)


def _kuu_coefficients_torch(kind, l, v):
    r3, r5 = float(SQRT3), float(SQRT5)
    if kind == "Matern12":
        return {"A": 1 / (2 * l * v), "B": l / (2 * v), "BC": 1 / (2 * v)}
    if kind == "Matern32":
        return {"A": r3 / (4 * l * v), "B": l / (2 * r3 * v), "C": l**3 / (12 * r3 * v), "BC": 1 / (2 * v),
                "BC_grad": l**2 / (2 * v)}
    return {"A": 3 * r5 / (16 * l * v), "B": 9 * l / (16 * r5 * v), "C": 9 * l**3 / (80 * r5 * v),
            "D": 3 * l**5 / (400 * r5 * v), "BC": 9 / (16 * v), "BC_grad": 3 * l**2 / (10 * v),
            "BC_ggrad": 9 * l**4 / (400 * v)}


def elbo_grad_kron_dense(kinds, tables, G_sparse, Kuf_y, tr_yTy, n, hypers, sigma2):
    """GPR_kron.elbo (gpr.py:282-308) and d/d(v1, l1, v2, l2, sigma2) by torch-fp64 autograd through the dense
    algebra the reference itself uses (the reference gets them from TF reverse mode).  hypers = [(v1, l1), (v2, l2)].
    Small sizes only."""
    import torch

    th = torch.tensor([hypers[0][0], hypers[0][1], hypers[1][0], hypers[1][1], sigma2], dtype=torch.float64,
                      requires_grad=True)
    Ks = []
    for i in range(2):
        v, l = th[2 * i], th[2 * i + 1]
        dense = {nme: torch.from_numpy(band_to_dense_sym(t)) for nme, t in tables[i].items()}
        Ks.append(sum(c * dense[nme] for nme, c in _kuu_coefficients_torch(kinds[i], l, v).items()))
    s2 = th[4]
    Kuu = torch.kron(Ks[0], Ks[1])                                           # utils.py:45-51
    L_Kuu = torch.kron(torch.linalg.cholesky(Ks[0]), torch.linalg.cholesky(Ks[1]))
    Gd = torch.from_numpy(np.asarray(G_sparse.toarray(), dtype=np.float64))
    b = torch.from_numpy(np.asarray(Kuf_y, dtype=np.float64).reshape(-1, 1))
    L_P = torch.linalg.cholesky(Kuu + Gd / s2)
    c = torch.linalg.solve_triangular(L_P, b, upper=False) / s2
    elbo = (-0.5 * n * torch.log(2 * np.pi * s2) - torch.log(torch.diagonal(L_P)).sum()
            + torch.log(torch.diagonal(L_Kuu)).sum() - 0.5 * tr_yTy / s2 + 0.5 * (c**2).sum()
            - 0.5 * n * th[0] * th[2] / s2 + 0.5 * torch.trace(torch.cholesky_solve(Gd, L_Kuu)) / s2)
    elbo.backward()
    return float(elbo.detach()), th.grad.numpy().copy()


def kron_band(Kuu_bands, G_sparse, sigma2, k, ms):
    """Lower band (bw+1, M) of P = K1 (x) K2 + G / sigma2 with bw = k (m2 + 1) (gpr.py:262, 292)."""
    Ks = [sp.csr_matrix(band_to_dense_sym(b)) for b in Kuu_bands]
    P = (sp.kron(Ks[0], Ks[1]) + G_sparse / sigma2).tocsr()
    return sparse_to_band(P, k * (ms[1] + 1))


def elbo_kron_banded(Kuu_bands, G_sparse, Kuf_y, tr_yTy, n, variances, sigma2, k, ms):
    """Same bound as elbo_kron_dense with LAPACK band routines on the scalar band of P and the Kronecker identities
    log|K1 (x) K2| = m2 log|K1| + m1 log|K2|, trace((K1 (x) K2)^-1 G) = sum G .* (K1^-1 (x) K2^-1) —
    for sizes where the dense M x M algebra of the reference is too slow.  Validated against elbo_kron_dense."""
    m1, m2 = ms
    Pb = kron_band(Kuu_bands, G_sparse, sigma2, k, ms)
    L_P = sla.cholesky_banded(Pb, lower=True)
    log_det_P = 2.0 * np.sum(np.log(L_P[0]))
    c = sla.solve_banded((Pb.shape[0] - 1, 0), L_P, np.asarray(Kuf_y).reshape(-1, 1)) / sigma2
    cK = [sla.cholesky_banded(b, lower=True) for b in Kuu_bands]
    log_det_Kuu = m2 * 2.0 * np.sum(np.log(cK[0][0])) + m1 * 2.0 * np.sum(np.log(cK[1][0]))
    Kinv = [sla.cho_solve_banded((c_, True), np.eye(c_.shape[1])) for c_ in cK]
    G = sp.coo_matrix(G_sparse)
    tr = np.sum(G.data * Kinv[0][G.row // m2, G.col // m2] * Kinv[1][G.row % m2, G.col % m2])
    elbo = -0.5 * n * np.log(2 * np.pi * sigma2) - 0.5 * log_det_P + 0.5 * log_det_Kuu - 0.5 * tr_yTy / sigma2
    elbo += 0.5 * np.sum(np.square(c)) - 0.5 * n * np.prod(variances) / sigma2 + 0.5 * tr / sigma2
    return elbo


def predict_kron_banded(meshes, deltas, k, ms, Kuu_bands, G_sparse, Kuf_y, variances, sigma2, Xnew):
    """GPR_kron.predict_f_sparse (gpr.py:336-359) with LAPACK band solves instead of CHOLMOD."""
    m1, m2 = ms
    Pb = kron_band(Kuu_bands, G_sparse, sigma2, k, ms)
    cP = sla.cholesky_banded(Pb, lower=True)
    alpha = sla.cho_solve_banded((cP, True), np.asarray(Kuf_y).reshape(-1, 1)) / sigma2
    K1 = make_Kuf(meshes[0], deltas[0], k, m1, Xnew[:, 0]).toarray()
    K2 = make_Kuf(meshes[1], deltas[1], k, m2, Xnew[:, 1]).toarray()
    Kus = (K1[:, None, :] * K2[None, :, :]).reshape(m1 * m2, -1)
    mean = Kus.T @ alpha
    cK = [sla.cholesky_banded(b, lower=True) for b in Kuu_bands]
    q1 = np.sum(K1 * sla.cho_solve_banded((cK[0], True), K1), axis=0)
    q2 = np.sum(K2 * sla.cho_solve_banded((cK[1], True), K2), axis=0)
    var = np.prod(variances) + np.sum(Kus * sla.cho_solve_banded((cP, True), Kus), axis=0) - q1 * q2
    return mean, var.reshape(-1, 1)


def predict_kron_banded_cells(meshes, deltas, k, ms, Kuu_bands, G_sparse, Kuf_y, variances, sigma2, Xnew):
    """predict_kron_banded for MANY test points that share few knot cells, at sizes where a dense M x n* right-hand side
    is too large (M = 200 x 200): Kus[:, n] has its (k+1)^2 non-zeros at the rows (idx1 + r, idx2 + s) of the point's
    cell (kronecker.py:24-33), so Kus^T P^-1 Kus (gpr.py:353-356) only reads the entries of P^-1 between those rows —
    (k+1)^2 LAPACK band solves with unit vectors per distinct cell.  Same numbers as predict_kron_banded (checked in
    tests/test_oracle_golden.py)."""
    m1, m2 = ms
    Xnew = np.asarray(Xnew, dtype=np.float64)
    Pb = kron_band(Kuu_bands, G_sparse, sigma2, k, ms)
    cP = sla.cholesky_banded(Pb, lower=True)
    alpha = sla.cho_solve_banded((cP, True), np.asarray(Kuf_y).reshape(-1, 1)) / sigma2
    cK = [sla.cholesky_banded(b, lower=True) for b in Kuu_bands]
    i1, u1 = locate(meshes[0], Xnew[:, 0])
    i2, u2 = locate(meshes[1], Xnew[:, 1])
    w1 = pieces(k, Xnew[:, 0] - u1, deltas[0])               # (k+1, n*)
    w2 = pieces(k, Xnew[:, 1] - u2, deltas[1])
    mean = np.zeros((Xnew.shape[0], 1))
    var = np.zeros((Xnew.shape[0], 1))
    cells = np.unique(np.stack([i1, i2], 1), axis=0)
    for c1, c2 in cells:
        rows = ((c1 + np.arange(k + 1))[:, None] * m2 + (c2 + np.arange(k + 1))[None, :]).reshape(-1)
        E = np.zeros((m1 * m2, rows.size))
        E[rows, np.arange(rows.size)] = 1.0
        Pinv = sla.cho_solve_banded((cP, True), E)[rows]     # (k+1)^2 x (k+1)^2 block of P^-1
        E1 = np.zeros((m1, k + 1)); E1[c1 + np.arange(k + 1), np.arange(k + 1)] = 1.0
        E2 = np.zeros((m2, k + 1)); E2[c2 + np.arange(k + 1), np.arange(k + 1)] = 1.0
        K1inv = sla.cho_solve_banded((cK[0], True), E1)[c1:c1 + k + 1]
        K2inv = sla.cho_solve_banded((cK[1], True), E2)[c2:c2 + k + 1]
        sel = np.nonzero((i1 == c1) & (i2 == c2))[0]
        a, b = w1[:, sel], w2[:, sel]
        w = (a[:, None, :] * b[None, :, :]).reshape(-1, sel.size)
        mean[sel, 0] = w.T @ alpha[rows, 0]
        q1 = np.sum(a * (K1inv @ a), axis=0)
        q2 = np.sum(b * (K2inv @ b), axis=0)
        var[sel, 0] = np.prod(variances) + np.sum(w * (Pinv @ w), axis=0) - q1 * q2
    return mean, var


def stencil_columns_of_inverse(Kuu_bands, G_sparse, sigma2, k, ms, cols):
    """Entries of P^-1 = (K1 (x) K2 + G / sigma2)^-1 in the stencil layout of include/asvgp_b200.h for the given columns
    j (what the reference's dense cholesky_solve would hold at those positions, gpr.py:293-307): out[e, c] with
    e = d1 (2k+1) + (d2 + k) is P^-1[(j1 + d1, j2 + d2), (j1, j2)], zero outside the matrix or for d1 = 0, d2 < 0.
    One LAPACK band solve per column."""
    m1, m2 = ms
    Pb = kron_band(Kuu_bands, G_sparse, sigma2, k, ms)
    cP = sla.cholesky_banded(Pb, lower=True)
    cols = np.asarray(cols, dtype=np.int64)
    E = np.zeros((m1 * m2, cols.size))
    E[cols, np.arange(cols.size)] = 1.0
    Pinv = sla.cho_solve_banded((cP, True), E)
    out = np.zeros(((k + 1) * (2 * k + 1), cols.size))
    for c, j in enumerate(cols):
        j1, j2 = divmod(int(j), m2)
        for d1 in range(k + 1):
            for d2 in range(-k, k + 1):
                if (d1 == 0 and d2 < 0) or j1 + d1 >= m1 or not (0 <= j2 + d2 < m2):
                    continue
                out[d1 * (2 * k + 1) + d2 + k, c] = Pinv[(j1 + d1) * m2 + j2 + d2, c]
    return out


# ---------------------------------------------------------------------------------------------------------------------
# a9 at large M: closed-form gradients on LAPACK band routines
# ---------------------------------------------------------------------------------------------------------------------
KUU_LENGTHSCALE_POWERS = {          # coefficient of each table is c * ell^p / var (inducing_features.py:17-44)
    "Matern12": {"A": -1, "B": 1, "BC": 0},
    "Matern32": {"A": -1, "B": 1, "C": 3, "BC": 0, "BC_grad": 2},
    "Matern52": {"A": -1, "B": 1, "C": 3, "D": 5, "BC": 0, "BC_grad": 2, "BC_ggrad": 4},
}


def _band_matmul(band, V):
    """(symmetric matrix given by its lower band) @ V, V dense M x r."""
    k, m = band.shape[0] - 1, band.shape[1]
    out = band[0][:, None] * V
    for d in range(1, k + 1):
        out[d:] += band[d, : m - d, None] * V[: m - d]
        out[: m - d] += band[d, : m - d, None] * V[d:]
    return out


def _band_dot(A, B):
    """trace(A B) of two symmetric matrices given by their lower bands (same bandwidth)."""
    w = np.full((A.shape[0], 1), 2.0)
    w[0] = 1.0
    return float(np.sum(w * A * B))


def elbo_grad_1d_banded(kind, tables, G, Kuf_y, tr_yTy, n, var, ell, sigma2, block=1000):
    """ELBO (elbo_1d) and d/d(var, ell, sigma2) in closed form (what TF reverse mode returns for
    gpr.py:49-89) on LAPACK band routines — the gradient oracle at M where the dense autograd restatement
    (elbo_grad_1d_dense) is too slow; validated against it in tests/test_oracle_golden.py.  One output column."""
    k, m = G.shape[0] - 1, G.shape[1]
    b = np.asarray(Kuf_y, dtype=np.float64).reshape(m, 1)
    co = kuu_coefficients(kind, ell, var)
    Kuu = sum(c * tables[nme] for nme, c in co.items())
    dK_l = sum(KUU_LENGTHSCALE_POWERS[kind][nme] * c / ell * tables[nme] for nme, c in co.items())
    dK_v = -Kuu / var
    cK = sla.cholesky_banded(Kuu, lower=True)
    P = G / sigma2 + Kuu
    cP = sla.cholesky_banded(P, lower=True)
    Kinv, Pinv = takahashi_band(cK), takahashi_band(cP)
    alpha = sla.cho_solve_banded((cP, True), b) / sigma2
    # band of W = Kuu^-1 G Kuu^-1, block of columns by block of columns (three band operations per block)
    W = np.zeros((k + 1, m))
    for s in range(0, m, block):
        cols = np.arange(s, min(s + block, m))
        E = np.zeros((m, cols.size))
        E[cols, np.arange(cols.size)] = 1.0
        Wc = sla.cho_solve_banded((cK, True), _band_matmul(G, sla.cho_solve_banded((cK, True), E)))
        for d in range(k + 1):
            ok = cols + d < m
            W[d, cols[ok]] = Wc[cols[ok] + d, np.arange(cols.size)[ok]]
    elbo = elbo_1d(Kuu, G, b, tr_yTy, n, var, sigma2)
    bPb = (b.T @ alpha).item() * sigma2
    grads = []
    for dK, dvar in ((dK_v, 1.0), (dK_l, 0.0)):
        g = (-0.5 * _band_dot(Pinv, dK) + 0.5 * _band_dot(Kinv, dK) - 0.5 * (alpha.T @ _band_matmul(dK, alpha)).item()
             - 0.5 * n * dvar / sigma2 - 0.5 * _band_dot(W, dK) / sigma2)
        grads.append(g)
    tr = _band_dot(Kinv, G)
    g_s2 = (-0.5 * n / sigma2 + 0.5 * _band_dot(Pinv, G) / sigma2**2 + 0.5 * tr_yTy / sigma2**2 - bPb / sigma2**3
            + 0.5 * (alpha.T @ _band_matmul(G, alpha)).item() / sigma2**2 + 0.5 * n * var / sigma2**2 - 0.5 * tr / sigma2**2)
    return elbo, np.array(grads + [g_s2])


# ---------------------------------------------------------------------------------------------------------------------
# GPR_additive  (asvgp/gpr.py:139-236): sum of 1-D models, Kuu block diagonal, Kuf stacked
# ---------------------------------------------------------------------------------------------------------------------
def precompute_additive(meshes, deltas, k, ms, X, y):
    """Kuf = vstack of the per-dimension feature matrices; dense KufKfu, Kuf_y, tr_yTy (gpr.py:171-176)."""
    Kuf = sp.vstack([make_Kuf(meshes[i], deltas[i], k, ms[i], X[:, i]) for i in range(X.shape[1])]).tocsr()
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    return np.asarray((Kuf @ Kuf.T).todense()), np.asarray(Kuf @ y), float(np.sum(np.square(y)))


def elbo_additive_dense(Kuu_bands, G, Kuf_y, tr_yTy, n, variances, sigma2):
    """GPR_additive.elbo (gpr.py:181-210) with dense algebra, as the reference does it."""
    Kuu = sla.block_diag(*[band_to_dense_sym(b) for b in Kuu_bands])
    P = Kuu + G / sigma2
    L = np.linalg.cholesky(P)
    c = sla.solve_triangular(L, Kuf_y, lower=True) / sigma2
    elbo = -0.5 * n * np.log(2 * np.pi * sigma2) - np.sum(np.log(np.diag(L))) + 0.5 * np.linalg.slogdet(Kuu)[1]
    elbo += -0.5 * tr_yTy / sigma2 + 0.5 * np.sum(np.square(c)) - 0.5 * n * np.sum(variances) / sigma2
    elbo += 0.5 * np.trace(np.linalg.solve(Kuu, G)) / sigma2
    return elbo


def elbo_grad_additive_dense(kinds, tables, G, Kuf_y, tr_yTy, n, hypers, sigma2):
    """The same bound and d/d(v_1, l_1, ..., v_D, l_D, sigma2) by torch-fp64 autograd (the reference gets them from TF
    reverse mode).  hypers = [(v_i, l_i)]."""
    import torch

    flat = [h for vl in hypers for h in vl] + [sigma2]
    th = torch.tensor(flat, dtype=torch.float64, requires_grad=True)
    Ks = []
    for i, kind in enumerate(kinds):
        v, l = th[2 * i], th[2 * i + 1]
        dense = {nme: torch.from_numpy(band_to_dense_sym(t)) for nme, t in tables[i].items()}
        Ks.append(sum(c * dense[nme] for nme, c in _kuu_coefficients_torch(kind, l, v).items()))
    s2 = th[-1]
    Kuu = torch.block_diag(*Ks)
    Gd = torch.from_numpy(np.asarray(G, dtype=np.float64))
    b = torch.from_numpy(np.asarray(Kuf_y, dtype=np.float64).reshape(-1, 1))
    L = torch.linalg.cholesky(Kuu + Gd / s2)
    c = torch.linalg.solve_triangular(L, b, upper=False) / s2
    total_var = sum(th[2 * i] for i in range(len(kinds)))
    elbo = (-0.5 * n * torch.log(2 * np.pi * s2) - torch.log(torch.diagonal(L)).sum() + 0.5 * torch.logdet(Kuu)
            - 0.5 * tr_yTy / s2 + 0.5 * (c**2).sum() - 0.5 * n * total_var / s2
            + 0.5 * torch.trace(torch.linalg.solve(Kuu, Gd)) / s2)
    elbo.backward()
    return float(elbo.detach()), th.grad.numpy().copy()


def predict_additive_dense(meshes, deltas, k, ms, Kuu_bands, G, Kuf_y, variances, sigma2, Xnew):
    """GPR_additive.predict_f (gpr.py:212-236), dense."""
    Kuu = sla.block_diag(*[band_to_dense_sym(b) for b in Kuu_bands])
    P = G / sigma2 + Kuu
    Kus = np.asarray(sp.vstack([make_Kuf(meshes[i], deltas[i], k, ms[i], Xnew[:, i]) for i in range(Xnew.shape[1])]).todense())
    cP = sla.cho_factor(P, lower=True)
    mean = Kus.T @ sla.cho_solve(cP, Kuf_y) / sigma2
    var = np.sum(variances) + np.sum(Kus * sla.cho_solve(cP, Kus), axis=0) - np.sum(Kus * np.linalg.solve(Kuu, Kus), axis=0)
    return mean, var.reshape(-1, 1)

"""Exact (rational-arithmetic) tables for uniform B-splines of degree k = 1..6.

The reference hard-codes per-order expanded formulas (reference asvgp/basis.py:117-798).  Here everything is
*derived* once from the Cox-de Boor recursion with `fractions.Fraction`, so any order works and the tables are
exact rationals (e.g. the cubic interior Gram 151/315, 397/1680, 1/42, 1/5040 of basis.py:290-293 falls out).

Conventions.  On knot interval c (left knot u = mesh[c], t = (x-u)/delta in [0,1]) the k+1
non-zero basis functions are rows c .. c+k.  `piece r` (r = 0..k) is the polynomial of row c+r on that
interval: piece_r(t) = N_k(t + k - r) with N_k the cardinal B-spline on [0, k+1].  The reference's b_i
(i = 1..k+1, basis.py:72) is piece k+1-i.
"""
from fractions import Fraction
from functools import lru_cache

import numpy as np


def _poly_mul(p, q):
    out = [Fraction(0)] * (len(p) + len(q) - 1)
    for i, a in enumerate(p):
        for j, b in enumerate(q):
            out[i + j] += a * b
    return out


def _poly_add(p, q):
    n = max(len(p), len(q))
    return [(p[i] if i < len(p) else 0) + (q[i] if i < len(q) else 0) for i in range(n)]


def _poly_der(p):
    return [p[i] * i for i in range(1, len(p))] or [Fraction(0)]


def _poly_int01(p):
    return sum(c / (i + 1) for i, c in enumerate(p))


@lru_cache(maxsize=None)
def cardinal_pieces(k):
    """Q[s] = coefficients (ascending powers of t) of N_k(s + t), t in [0,1], s = 0..k."""
    if k == 0:
        return ((Fraction(1),),)
    prev = cardinal_pieces(k - 1)
    out = []
    for s in range(k + 1):
        # N_k(x) = x/k N_{k-1}(x) + (k+1-x)/k N_{k-1}(x-1), with x = s + t
        acc = [Fraction(0)]
        if s <= k - 1:
            acc = _poly_add(acc, _poly_mul([Fraction(s, k), Fraction(1, k)], list(prev[s])))
        if s - 1 >= 0:
            acc = _poly_add(acc, _poly_mul([Fraction(k + 1 - s, k), Fraction(-1, k)], list(prev[s - 1])))
        acc = acc + [Fraction(0)] * (k + 1 - len(acc))
        out.append(tuple(acc[: k + 1]))
    return tuple(out)


@lru_cache(maxsize=None)
def piece_coeffs(k, dx=0):
    """(k+1) x (k+1) Fractions: row r = coefficients in t of d^dx/dt^dx piece_r(t) (row c+r on interval c)."""
    Q = cardinal_pieces(k)
    rows = []
    for r in range(k + 1):
        p = list(Q[k - r])
        for _ in range(dx):
            p = _poly_der(p)
        p = p + [Fraction(0)] * (k + 1 - len(p))
        rows.append(tuple(p))
    return tuple(rows)


def piece_coeffs_float(k, dx=0):
    return np.array([[float(c) for c in row] for row in piece_coeffs(k, dx)], dtype=np.float64)


@lru_cache(maxsize=None)
def interval_gram(k, q):
    """W[r][s] = int_0^1 piece_r^(q)(t) piece_s^(q)(t) dt (exact).  Multiply by delta**(1-2q) for x-units."""
    P = piece_coeffs(k, q)
    return tuple(tuple(_poly_int01(_poly_mul(list(P[r]), list(P[s]))) for s in range(k + 1)) for r in range(k + 1))


def gram_band(k, m, q, delta):
    """Lower band (k+1) x m of S^(q)[i,j] = int_a^b phi_i^(q) phi_j^(q) dx, layout band[d, j] = S[j+d, j]
    with d trailing zeros in row d.  Edge functions are truncated to [a,b] (reference basis.py:31-45:
    per-interval contributions cumsum'd in and out).  Row j lives on intervals j-k .. j, clipped to
    0 .. m-k-1."""
    W = interval_gram(k, q)
    n_int = m - k
    band = np.zeros((k + 1, m), dtype=np.float64)
    scale = float(delta) ** (1 - 2 * q)
    # number of in-domain intervals shared by rows j and j+d only takes O(k) distinct configurations
    cache = {}
    for d in range(k + 1):
        for j in range(m - d):
            lo = max(j + d - k, 0)
            hi = min(j, n_int - 1)
            key = (d, lo - j, hi - j)
            if key not in cache:
                tot = Fraction(0)
                for c in range(lo, hi + 1):
                    tot += W[j + d - c][j - c]
                cache[key] = float(tot)
            band[d, j] = cache[key] * scale
    return band


def boundary_values(k, dx):
    """Values at the left edge x = a of d^dx phi_r / dt^dx for rows r = 0..k (t = 0 on interval 0)."""
    P = piece_coeffs(k, dx)
    return [P[r][0] for r in range(k + 1)]


def boundary_band(k, m, dx, delta):
    """Reference `make_boundary_conditions(dx)` for dx in {0,1,2} (basis.py:82-114): outer product of the first
    k boundary values, its d-th diagonal written at BOTH ends of band row d (left corner and columns
    m-d-len .. m-d-1, exactly where the reference's concat([l, zero_fill, l, zero_pad]) puts it), last band
    row zero."""
    vals = boundary_values(k, dx)[:k]
    band = np.zeros((k + 1, m), dtype=np.float64)
    scale = float(delta) ** (-2 * dx)
    for d in range(k):
        diag = [float(vals[i + d] * vals[i]) * scale for i in range(k - d)]
        n = len(diag)
        band[d, :n] = diag
        band[d, m - d - n: m - d] = diag
    return band


# ---- Bernstein-type expansions used by the 2-D accumulate kernel -------------------------------------------------------
def _comb(n, k):
    from math import comb

    return comb(n, k)


def _to_bernstein(poly, n):
    """Coefficients c_p with poly(t) = sum_p c_p t^p (1-t)^(n-p), from ascending monomial coefficients
    (t^k = t^k (t + (1-t))^(n-k))."""
    out = [Fraction(0)] * (n + 1)
    for k, a in enumerate(poly):
        if a == 0:
            continue
        for j in range(n - k + 1):
            out[k + j] += a * _comb(n - k, j)
    return out


@lru_cache(maxsize=None)
def product_bernstein(k):
    """C[r][s][p] (exact): piece_r(t) piece_s(t) = sum_p C[r][s][p] t^p (1-t)^(2k-p).  All entries are >= 0 (uniform
    B-spline pieces have non-negative Bezier coefficients), so expanding per-cell moments of t^p (1-t)^(2k-p) back
    into the Gram stencil involves no cancellation."""
    P = piece_coeffs(k, 0)
    out = tuple(tuple(tuple(_to_bernstein(_poly_mul(list(P[r]), list(P[s])), 2 * k)) for s in range(k + 1))
                for r in range(k + 1))
    assert all(c >= 0 for a in out for b in a for c in b)
    return out


@lru_cache(maxsize=None)
def piece_bernstein(k):
    """D[r][p] (exact): piece_r(t) = sum_p D[r][p] t^p (1-t)^(k-p), all >= 0."""
    P = piece_coeffs(k, 0)
    out = tuple(tuple(_to_bernstein(list(P[r]), k)) for r in range(k + 1))
    assert all(c >= 0 for a in out for c in a)
    return out


def product_bernstein_float(k):
    return np.array([[[float(c) for c in s] for s in r] for r in product_bernstein(k)], dtype=np.float64)


def piece_bernstein_float(k):
    return np.array([[float(c) for c in r] for r in piece_bernstein(k)], dtype=np.float64)

"""numpy prototype of the nested-dissection multifrontal factorisation + selected inverse of
P = K1 (x) K2 + G / sigma2 that csrc/ndfront_2d.cu implements on the GPU.  Dense numpy per front; used
to validate the algebra and the index plan against the LAPACK-band oracle on the CPU (tests/test_nd_plan.py) before the
CUDA kernels existed.  TEST / DESIGN INFRASTRUCTURE — the product never imports this.

Graph: node (i1, i2) of the m1 x m2 basis grid, coupled to (j1, j2) iff |i1 - j1| <= k and |i2 - j2| <= k, plus one
"rhs node" coupled to every grid node (the right-hand side Kuf_y rides along as an extra row of every front, so that
y = L^-1 b, ||y||^2 and x = P^-1 b fall out of the same factorisation and selected inverse: A' = [[P, b], [b^T, gamma]]).
"""
import numpy as np


class Front:
    __slots__ = ("level", "sep", "bnd", "children", "parent", "region")


def build_tree(m1, m2, k, leaf):
    """Recursive bisection of the grid by k-wide separators; returns the list of fronts in post-order (children first).
    Each front: sep = array of grid node ids (i1 * m2 + i2) eliminated here, region = (r0, r1, c0, c1) bounding box of
    everything eliminated in its subtree."""
    fronts = []

    def rec(r0, r1, c0, c1, level, parent_holder):
        nr, nc = r1 - r0, c1 - c0
        f = Front()
        f.level, f.region, f.children, f.parent = level, (r0, r1, c0, c1), [], None
        can_r, can_c = nr >= 3 * k + 2 and nr > leaf, nc >= 3 * k + 2 and nc > leaf
        if not (can_r or can_c):
            rr, cc = np.meshgrid(np.arange(r0, r1), np.arange(c0, c1), indexing="ij")
            f.sep = (rr * m2 + cc).reshape(-1)
        else:
            split_rows = can_r and (nr >= nc or not can_c)
            if split_rows:
                mid = r0 + (nr - k) // 2
                rr, cc = np.meshgrid(np.arange(mid, mid + k), np.arange(c0, c1), indexing="ij")
                kids = [(r0, mid, c0, c1), (mid + k, r1, c0, c1)]
            else:
                mid = c0 + (nc - k) // 2
                rr, cc = np.meshgrid(np.arange(r0, r1), np.arange(mid, mid + k), indexing="ij")
                kids = [(r0, r1, c0, mid), (r0, r1, mid + k, c1)]
            f.sep = (rr * m2 + cc).reshape(-1)
            for kid in kids:
                f.children.append(rec(*kid, level + 1, f))
        fronts.append(f)
        return f

    root = rec(0, m1, 0, m2, 0, None)
    for f in fronts:
        for c in f.children:
            c.parent = f
    # boundaries: ancestors' separator nodes within distance k (both dimensions) of the subtree's region
    for f in fronts:
        r0, r1, c0, c1 = f.region
        bnd = []
        a = f.parent
        while a is not None:
            s1, s2 = a.sep // m2, a.sep % m2
            near = (s1 >= r0 - k) & (s1 <= r1 - 1 + k) & (s2 >= c0 - k) & (s2 <= c1 - 1 + k)
            bnd.append(a.sep[near])
            a = a.parent
        f.bnd = np.concatenate(bnd) if bnd else np.zeros(0, dtype=np.int64)
    return fronts, root


def dense_lookup(m1, m2, k, K1, K2, Gs, sigma2):
    """A(i, j) for arrays of node ids, from the per-dimension lower bands and the stencil layout of G."""
    def band_sym(B, i, j):
        d = np.abs(i - j)
        return np.where(d <= k, B[np.minimum(d, k), np.minimum(i, j)], 0.0)

    NS = 2 * k + 1

    def A(i, j):
        i1, i2, j1, j2 = i // m2, i % m2, j // m2, j % m2
        inside = (np.abs(i1 - j1) <= k) & (np.abs(i2 - j2) <= k)
        kv = band_sym(K1, i1, j1) * band_sym(K2, i2, j2)
        # stencil entry: lower part stored at column = the "smaller" node (d1 > 0 or d1 == 0 and d2 >= 0)
        lo_is_j = (i1 > j1) | ((i1 == j1) & (i2 >= j2))
        ci = np.where(lo_is_j, i, j); cj = np.where(lo_is_j, j, i)
        d1 = ci // m2 - cj // m2; d2 = ci % m2 - cj % m2
        e = np.clip(d1, 0, k) * NS + np.clip(d2, -k, k) + k
        g = Gs[e, cj]
        return np.where(inside, kv + g / sigma2, 0.0)

    return A


def factor_and_selinv(m1, m2, k, K1, K2, Gs, b, sigma2, leaf=12, gamma=None):
    """Returns dict(logdet, quad, x, sig) with sig = P^-1 entries in stencil layout."""
    M = m1 * m2
    fronts, root = build_tree(m1, m2, k, leaf)
    A = dense_lookup(m1, m2, k, K1, K2, Gs, sigma2)
    RHS = M                                              # id of the rhs node
    if gamma is None:
        gamma = 1.0 + 2.0 * float(b @ b)                 # anything that keeps A' positive definite; cancels out below
    order = {id(f): n for n, f in enumerate(fronts)}
    idx, ns = {}, {}
    for f in fronts:
        sep = f.sep if f is not root else np.concatenate([f.sep, [RHS]])
        bnd = f.bnd if f is root else np.concatenate([f.bnd, [RHS]])
        idx[id(f)] = np.concatenate([sep, bnd]).astype(np.int64)
        ns[id(f)] = sep.size

    def entry(i, j):                                     # A' on id arrays (broadcast)
        ii, jj = np.broadcast_arrays(i, j)
        out = np.zeros(ii.shape)
        gi, gj = ii < M, jj < M
        both = gi & gj
        out[both] = A(ii[both], jj[both])
        out[gi & ~gj] = b[ii[gi & ~gj]]
        out[~gi & gj] = b[jj[~gi & gj]]
        out[~gi & ~gj] = gamma
        return out

    L, U = {}, {}
    logdet = 0.0
    for f in fronts:                                     # post-order: children before parents
        I = idx[id(f)]
        n_s = ns[id(f)]
        F = np.zeros((I.size, I.size))
        F[:, :n_s] = entry(I[:, None], I[None, :n_s])
        F[:n_s, n_s:] = F[n_s:, :n_s].T
        pos = {g: p for p, g in enumerate(I)}
        for c in f.children:                             # extend-add
            Ic = idx[id(c)][ns[id(c)]:]
            mp = np.array([pos[g] for g in Ic])
            F[np.ix_(mp, mp)] += U[id(c)]
        Lss = np.linalg.cholesky(F[:n_s, :n_s])
        Lbs = np.linalg.solve(Lss, F[:n_s, n_s:]).T      # L_bs = F_bs L_ss^-T
        U[id(f)] = F[n_s:, n_s:] - Lbs @ Lbs.T
        L[id(f)] = (Lss, Lbs)
        d = np.diag(Lss)
        logdet += 2 * np.sum(np.log(d[:-1] if f is root else d))
    Lroot = L[id(root)][0]
    y_last = Lroot[-1, -1]                               # sqrt(gamma - ||y||^2)
    quad = gamma - y_last**2
    # selected inverse, top down
    Sig = {}
    for f in reversed(fronts):
        I = idx[id(f)]
        n_s = ns[id(f)]
        Lss, Lbs = L[id(f)]
        S = np.zeros((I.size, I.size))
        if f.parent is not None:
            Ip = idx[id(f.parent)]
            pos = {g: p for p, g in enumerate(Ip)}
            mp = np.array([pos[g] for g in I[n_s:]])
            S[n_s:, n_s:] = Sig[id(f.parent)][np.ix_(mp, mp)]
        Li = np.linalg.inv(Lss)
        Y = Lbs @ Li
        S[n_s:, :n_s] = -S[n_s:, n_s:] @ Y
        S[:n_s, n_s:] = S[n_s:, :n_s].T
        S[:n_s, :n_s] = Li.T @ Li - Y.T @ S[n_s:, :n_s]
        Sig[id(f)] = S
    tau = Sig[id(root)][ns[id(root)] - 1, ns[id(root)] - 1]          # Sigma'(rhs, rhs) = 1 / (gamma - b^T P^-1 b)
    x = np.zeros(M)
    for f in fronts:
        I, n_s = idx[id(f)], ns[id(f)]
        rhs_pos = I.size - 1 if f is not root else n_s - 1
        sel = I[:n_s] < M
        x[I[:n_s][sel]] = -Sig[id(f)][rhs_pos, :n_s][sel] / tau
    # stencil entries of P^-1 = Sigma' - tau x x^T
    NS = 2 * k + 1
    sig = np.zeros(((k + 1) * NS, M))
    for f in fronts:
        I, n_s = idx[id(f)], ns[id(f)]
        S = Sig[id(f)]
        grid = I < M
        pos_all = np.nonzero(grid)[0]
        for pj in np.nonzero(grid[:n_s])[0]:
            j = I[pj]
            j1, j2 = j // m2, j % m2
            i = I[pos_all]
            d1, d2 = i // m2 - j1, i % m2 - j2
            ok = (np.abs(d1) <= k) & (np.abs(d2) <= k)
            for pi, ii, a1, a2 in zip(pos_all[ok], i[ok], d1[ok], d2[ok]):
                v = S[pi, pj] - tau * x[ii] * x[j]
                if a1 > 0 or (a1 == 0 and a2 >= 0):
                    sig[a1 * NS + a2 + k, j] = v
                else:
                    sig[-a1 * NS + (-a2) + k, ii] = v
    return dict(logdet=logdet, quad=quad, x=x, sig=sig, n_fronts=len(fronts),
                front_sizes=[(f.level, ns[id(f)], idx[id(f)].size) for f in fronts])


if __name__ == "__main__":
    import os
    import sys
    import time

    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    import scipy.linalg as sla

    from asvgp_b200 import utils as U
    from oracle import asvgp_oracle as O

    m1, m2, k = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (40, 36, 3)
    leaf = int(sys.argv[4]) if len(sys.argv) > 4 else 12
    rng = np.random.default_rng(0)
    n = 40 * m1 * m2
    meshes, deltas = zip(*[O.make_mesh(0, m, m, k) for m in (m1, m2)])
    X = np.stack([rng.uniform(0.01, m1 - 0.01, n), rng.uniform(0.01, m2 - 0.01, n)], 1)
    y = np.sin(X[:, 0] / 3) * np.cos(X[:, 1] / 4) + 0.1 * rng.standard_normal(n)
    G, b, yy = O.precompute_kron(meshes, deltas, k, [m1, m2], X, y)
    T = [O.static_bands(k, m, d) for m, d in zip((m1, m2), deltas)]
    Ks = [O.make_Kuu("Matern32", 6.0, 1.0, T[0]), O.make_Kuu("Matern32", 5.0, 0.9, T[1])]
    s2 = 0.05
    Gs = U.sparse_to_stencil(G, m1, m2, k)
    t0 = time.time()
    out = factor_and_selinv(m1, m2, k, Ks[0], Ks[1], Gs, b[:, 0], s2, leaf=leaf)
    print("fronts", out["n_fronts"], "time %.1f s" % (time.time() - t0))
    Pb = O.kron_band(Ks, G, s2, k, [m1, m2])
    cP = sla.cholesky_banded(Pb, lower=True)
    ld = 2 * np.sum(np.log(cP[0]))
    xx = sla.cho_solve_banded((cP, True), b)[:, 0]
    print("logdet rel err", abs(out["logdet"] - ld) / abs(ld))
    print("quad rel err", abs(out["quad"] - b[:, 0] @ xx) / abs(b[:, 0] @ xx))
    print("x rel err", np.abs(out["x"] - xx).max() / np.abs(xx).max())
    cols = rng.choice(m1 * m2, 12, replace=False)
    want = O.stencil_columns_of_inverse(Ks, G, s2, k, [m1, m2], cols)
    print("sigma stencil rel err", np.abs(out["sig"][:, cols] - want).max() / np.abs(want).max())
    lv = {}
    for level, n_s, n_f in out["front_sizes"]:
        lv.setdefault(level, []).append((n_s, n_f))
    for level in sorted(lv):
        a = np.array(lv[level])
        print("level", level, "fronts", len(a), "sep max", a[:, 0].max(), "front max", a[:, 1].max(),
              "flops %.2e" % np.sum(a[:, 0] * a[:, 1].astype(float) ** 2))

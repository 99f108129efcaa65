"""B-spline bases `B1Spline` .. `B6Spline` — same names, constructor and attributes as the reference
(`asvgp/basis.py:8-114` base class, `:117-798` per-order classes):

    B3Spline(a, b, m)  ->  .a .b .m .order .mesh .delta
                           .A .B .C .D                      (static Gram bands, basis.py:31-45)
                           .BC .BC_grad .BC_ggrad .BC_ggrad_none .BC_none_ggrad   (basis.py:82-114)
                           .evaluate_basis(X, dx=0, sparse=True)                  (basis.py:51-80)

Differences in *how* (not what): the per-order closed forms of the reference are replaced by exact rational
tables derived from the Cox-de Boor recursion (`_spline_tables.py`), and `evaluate_basis` runs on the GPU
(`asvgp_basis_eval_1d`, csrc/accum_1d.cu) — there is no CPU fallback.

Mesh precision: the reference builds `tf.cast(tf.linspace(a, b, n), tf.float64)`; with
Python-float a, b TF computes the linspace in *float32*.  `mesh_dtype="tf"` (default) reproduces that rule
(floats -> float32 arithmetic, ints -> float64) so results match the reference on the same inputs;
`"float64"` / `"float32"` force one or the other.
"""
import numpy as np

from . import _spline_tables as _tab


def tf_style_linspace(a, b, n, mesh_dtype="tf"):
    """Knot mesh as the reference's `tf.cast(tf.linspace(a, b, n), tf.float64)` would produce it (basis.py:17)."""
    if mesh_dtype == "tf":
        ints = isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer))
        mesh_dtype = "float64" if ints else "float32"
    if mesh_dtype == "float64":
        return np.linspace(float(a), float(b), int(n), dtype=np.float64)
    if mesh_dtype != "float32":
        raise ValueError("mesh_dtype must be 'tf', 'float64' or 'float32'")
    f = np.float32
    a32, b32 = f(a), f(b)
    step = f(f(b32 - a32) / f(n - 1))
    out = np.empty(int(n), dtype=f)
    out[0], out[-1] = a32, b32
    out[1:-1] = (a32 + step * np.arange(1, n - 1, dtype=f)).astype(f)
    return out.astype(np.float64)


class SplineBasis:
    """Parent class of the B-spline bases (reference basis.py:8-114)."""

    order = None
    _n_gram = 0          # how many of A, B, C, D the reference defines for this order
    _bc_names = ()       # which boundary bands the reference defines for this order

    def __init__(self, a, b, m, mesh_dtype="tf"):
        k = self.order
        if m < 2 * (k + 1):
            raise ValueError("m >= 2*(order+1) basis functions are required (reference basis.py:36)")
        self.a = a
        self.b = b
        self.m = int(m)
        self.mesh = tf_style_linspace(a, b, self.m - (k - 1), mesh_dtype)      # basis.py:17  (m-k+1 knots)
        self.delta = float(self.mesh[1] - self.mesh[0])                        # basis.py:18  (first gap, Q3)
        for q, name in enumerate("ABCD"[: self._n_gram]):
            setattr(self, name, _tab.gram_band(k, self.m, q, self.delta))
        for name in self._bc_names:
            setattr(self, name, self.make_boundary_conditions(_BC_DX[name]))

    # -- static tables -------------------------------------------------------------------------------------
    def make_boundary_conditions(self, dx=0, pad="right"):
        """Boundary-condition bands (reference basis.py:82-114).  dx=3,4 (`BC_ggrad_none`, `BC_none_ggrad`)
        multiply values at `a` with values at `b` restricted to the first k rows, which is identically zero
        for m > 2k — reproduced as zeros."""
        if pad != "right":
            raise NotImplementedError("only pad='right' is used by the reference")
        if dx in (3, 4):
            return np.zeros((self.order + 1, self.m), dtype=np.float64)
        if dx not in (0, 1, 2):
            raise NotImplementedError
        return _tab.boundary_band(self.order, self.m, dx, self.delta)

    def piece_coefficients(self, dx=0):
        """(k+1)x(k+1) float64: row r = coefficients in t=(x-u)/delta of the dx-th t-derivative of basis row
        idx+r on interval idx.  This is what the CUDA kernels evaluate (Horner) for dx > 0."""
        return _tab.piece_coeffs_float(self.order, dx)

    # -- Kuf ---------------------------------------------------------------------------------------------------
    def evaluate_basis(self, X, dx=0, sparse=True):
        """Evaluations of the basis functions (or their dx-th derivative) as an (m, n) matrix
        (reference basis.py:51-80).  Runs `asvgp_basis_eval_1d` on the GPU; the CSR assembly of the
        (k+1) n non-zeros happens on the host only because the reference's return type is a SciPy matrix."""
        if dx not in (0, 1, 2, 3) or int(dx) != dx:
            raise NotImplementedError
        from . import ops

        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64).reshape(-1))
        n = X.shape[0]
        idx, vals = ops.basis_eval_1d(X, self, int(dx))          # idx[n] int64, vals[(k+1), n] rows idx+r
        k = self.order
        rows = (idx[None, :] + np.arange(k + 1, dtype=np.int64)[:, None]).reshape(-1)
        cols = np.tile(np.arange(n, dtype=np.int64), k + 1)
        data = vals.reshape(-1)
        if sparse:
            from scipy.sparse import csr_matrix

            return csr_matrix((data, (rows, cols)), shape=(self.m, n))
        out = np.zeros((self.m, n), dtype=np.float64)
        np.add.at(out, (rows, cols), data)
        return out


_BC_DX = {"BC": 0, "BC_grad": 1, "BC_ggrad": 2, "BC_ggrad_none": 3, "BC_none_ggrad": 4}
_ALL_BC = ("BC", "BC_grad", "BC_ggrad", "BC_ggrad_none", "BC_none_ggrad")


class B1Spline(SplineBasis):
    """Degree-1 (hat) basis; reference basis.py:117-167 defines A, B, BC."""
    order, _n_gram, _bc_names = 1, 2, ("BC",)


class B2Spline(SplineBasis):
    """Degree-2 basis; reference basis.py:170-249 defines A, B, C, BC, BC_grad."""
    order, _n_gram, _bc_names = 2, 3, ("BC", "BC_grad")


class B3Spline(SplineBasis):
    """Degree-3 basis; reference basis.py:252-369."""
    order, _n_gram, _bc_names = 3, 4, _ALL_BC


class B4Spline(SplineBasis):
    """Degree-4 basis; reference basis.py:372-503 (raises NameError for m < 12, :379-380)."""
    order, _n_gram, _bc_names = 4, 4, _ALL_BC

    def __init__(self, a, b, m, mesh_dtype="tf"):
        if m < 12:
            raise NameError("Not enough basis functions m >= 12")
        super().__init__(a, b, m, mesh_dtype)


class B5Spline(SplineBasis):
    """Degree-5 basis; reference basis.py:506-646."""
    order, _n_gram, _bc_names = 5, 4, _ALL_BC


class B6Spline(SplineBasis):
    """Degree-6 basis; reference basis.py:649-798 defines only BC, BC_grad (so no Matern52)."""
    order, _n_gram, _bc_names = 6, 4, ("BC", "BC_grad")

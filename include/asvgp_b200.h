/* libasvgp_sm100a — C ABI of the B200-native ASVGP hot path.
 *
 * This is the drop-in boundary.  Each entry point replaces the arithmetic that the
 * reference delegates to SciPy sparsetools / banded_matrices / CHOLMOD / dense TF linalg at the cited lines
 * (paths relative to the reference repository root).  The reference-side binding a maintainer would add is a
 * ctypes stub: see INTEGRATION.md.
 *
 * Conventions
 *  - plain C: pointers and sizes only, no torch/CUDA types in signatures; `stream` is a cudaStream_t passed as void*
 *    (NULL = the legacy default stream).
 *  - every data pointer is a DEVICE pointer to caller-owned, contiguous memory unless it is named `h_*`.
 *  - all floating point is IEEE fp64; indices are int64 (reference basis.py:58,72-73).
 *  - symmetric banded matrices travel as LOWER bands, row-major (k+1) x M with band[d*M + j] = A[j+d, j] and the last
 *    d entries of row d zero — exactly the layout of banded_matrices / reference utils.py:24-30.
 *  - return value: 0 = ok, <0 = error (asvgp_last_error() gives the message).  No entry point synchronises the host
 *    with the device; numerical failure (non-positive pivot) is reported through the `info` slot of the output
 *    buffer (0 = ok, j+1 = first failing pivot), LAPACK style.
 *  - nothing here falls back to the CPU: without a CUDA device every compute entry point returns an error.
 */
#ifndef ASVGP_B200_H
#define ASVGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASVGP_ABI_VERSION 1

#if defined(__GNUC__)
#define ASVGP_API __attribute__((visibility("default")))
#else
#define ASVGP_API
#endif

/* Matern kernel kinds (reference inducing_features.py:16,22,32). */
#define ASVGP_MATERN12 0
#define ASVGP_MATERN32 1
#define ASVGP_MATERN52 2

ASVGP_API int asvgp_abi_version(void);
/* Message of the last failing call on this thread (empty string if none). */
ASVGP_API const char* asvgp_last_error(void);

/* ---- a2/a6: basis evaluation = the non-zeros of Kuf -------------------------------------------------------------
 * Replaces SplineBasis.evaluate_basis (basis.py:51-80) / SplineFeatures1D.make_Kuf (inducing_features.py:47-48).
 * For point i: idx[i] = max(searchsorted_left(mesh, x[i]) - 1, 0) and vals[r*n + i] = d^dx/dx^dx of basis row
 * idx[i]+r at x[i], r = 0..order (the reference's b_{order+1-r}).  `coef` = (order+1)^2 exact piece coefficients in
 * t (ascending powers), required for dx > 0, ignored (may be NULL) for dx = 0. */
ASVGP_API int asvgp_basis_eval_1d(const double* x, int64_t n, const double* mesh, int n_knots, int order, int dx,
                        const double* coef, int64_t* idx, double* vals, void* stream);

/* ---- a7: O(N) accumulation of the banded Gram, the projection and sum(y^2) ------------------------------------------
 * Replaces GPR_1d.__init__'s `Kuf @ y`, `Kuf @ Kuf.T`, utils.sparse_to_band and sum(y^2) (gpr.py:39-44,
 * utils.py:24-30) without materialising Kuf.  Adds into (does NOT zero) the packed accumulator
 *   acc = [ G_band (order+1)*M | b M | sum(y^2) | count ]            ((order+2)*M + 2 doubles)
 * so that shards / chunks can be accumulated by repeated calls and all-reduced as one buffer.  M = n_knots+order-1. */
ASVGP_API int asvgp_accum_1d(const double* x, const double* y, int64_t n, const double* mesh, int n_knots, int order,
                   double* acc, void* stream);

/* ---- a10: 1-D posterior mean / variance over test points ----------------------------------------------------------------
 * Replaces GPR_1d.predict_f (gpr.py:91-136): mean[i] = sum_r w_r alpha[idx+r],
 * var[i] = variance + sum_{r,s} w_r w_s S[idx+r, idx+s] with S = band(P^-1) - band(Kuu^-1) (lower band, (order+1) x M). */
ASVGP_API int asvgp_predict_1d(const double* xnew, int64_t n, const double* mesh, int n_knots, int order, const double* alpha,
                     const double* S_band, double variance, double* mean, double* var, void* stream);

/* ---- a5: Kuu assembly ----------------------------------------------------------------------------------------------------
 * Replaces SplineFeatures1D.make_Kuu (inducing_features.py:12-44): Kuu = sum_t coef[t] * tables[t] over the static
 * bands A, B, C, D, BC, BC_grad, BC_ggrad of the basis (basis.py:31-45,82-114), each (order+1) x M, stacked
 * contiguously in `tables`.  `h_coef` / `h_dcoef` are HOST arrays of the Matern coefficients and of their
 * derivatives w.r.t. the lengthscale; dKuu (may be NULL) receives sum_t dcoef[t] * tables[t]. */
ASVGP_API int asvgp_kuu_assemble(const double* tables, int n_terms, const double* h_coef, const double* h_dcoef, int M,
                                 int order, double* Kuu, double* dKuu, void* stream);

/* Scratch bytes needed by asvgp_elbo_grad_1d / asvgp_posterior_1d (chunks = 0: library default). */
ASVGP_API int64_t asvgp_workspace_bytes_1d(int M, int order, int chunks);

/* ---- a8 + a9: collapsed ELBO and its hyper-parameter gradients ------------------------------------------------------------
 * Replaces GPR_1d.elbo (gpr.py:49-89: cholesky_band x2, inverse_from_cholesky_band, product_band_band,
 * solve_triang_mat) and the TF reverse-mode gradient the optimiser asks for (example.py:31-32).
 * `acc` is the packed accumulator of asvgp_accum_1d (after any all-reduce).  `chunks`: number of partitions of the
 * banded sweeps (0 = default).  out[16] (device):
 *   [0] ELBO  [1] dELBO/dvariance  [2] dELBO/dlengthscale  [3] dELBO/dsigma2
 *   [4] log|Kuu|  [5] log|P|  [6] b^T P^-1 b  [7] trace(Kuu^-1 G)  [8] info (0 ok, j+1 first non-positive pivot)
 *   [9..15] diagnostics (individual derivatives). */
ASVGP_API int asvgp_elbo_grad_1d(const double* Kuu, const double* dKuu, const double* acc, int M, int order,
                                 double variance, double sigma2, int chunks, double* out, void* work,
                                 int64_t work_bytes, void* stream);

/* ---- a10 (factorisation half): posterior weights -------------------------------------------------------------------------
 * Replaces the CHOLMOD factorisations and solves of GPR_1d.predict_f (gpr.py:96-108):
 * alpha = P^-1 Kuf_y / sigma2 and S = band(P^-1) - band(Kuu^-1), P = Kuu + G / sigma2.  info[2] (device): failing
 * pivot of the Kuu / P factorisation or 0. */
ASVGP_API int asvgp_posterior_1d(const double* Kuu, const double* acc, int M, int order, double sigma2, int chunks,
                                 double* alpha, double* S_band, double* info, void* work, int64_t work_bytes,
                                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ASVGP_B200_H */

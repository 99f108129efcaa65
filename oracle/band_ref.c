/* Plain-C restatement of banded_matrices' inverse_from_cholesky_band (the Takahashi / sparse-inverse-subset
 * recursion the reference calls at asvgp/gpr.py:59) — TEST INFRASTRUCTURE / CPU baseline only.
 * Layout: lower bands, row-major (k+1) x m, band[d*m + j] = A[j+d, j].
 *   S[i,j] = delta_ij / L_jj^2 - (1/L_jj) * sum_{r=j+1}^{min(m-1,j+k)} L[r,j] * S[max(r,i), min(r,i)]   
 */
void takahashi_band(const double* L, int k, int m, double* S) {
    for (int j = m - 1; j >= 0; --j) {
        const double ljj = L[j];
        const int hi = (j + k < m - 1) ? j + k : m - 1;
        for (int i = hi; i >= j; --i) {
            double acc = 0.0;
            for (int r = j + 1; r <= hi; ++r) {
                const int a = r >= i ? r : i, c = r >= i ? i : r;
                acc += L[(r - j) * m + j] * S[(a - c) * m + c];
            }
            S[(i - j) * m + j] = (i == j ? 1.0 / (ljj * ljj) : 0.0) - acc / ljj;
        }
    }
}

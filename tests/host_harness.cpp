// CPU harness around the ASVGP_HD templates of asvgp_b200/csrc/band_engine.cuh — TEST INFRASTRUCTURE ONLY.
// Compiled with g++ by tests/test_band_engine_host.py so that the partitioned banded algebra (the part of the CUDA
// path that is hardest to debug remotely) can be checked against dense numpy on a machine without a GPU.  The
// product never links this file.
#include <cstdlib>
#include <vector>
#include "../asvgp_b200/csrc/band_engine.cuh"

namespace asvgp { void set_last_error(const char*, ...) {} }
using namespace asvgp;

template <class T> struct HostMat;
template <> struct HostMat<double> {
    const double* band; const double* dband; int M;
    double operator()(int d, int j) const { return (j >= 0 && j < M && j + d < M) ? band[(size_t)d * M + j] : 0.0; }
};
template <> struct HostMat<Dual<1>> {
    const double* band; const double* dband; int M;
    Dual<1> operator()(int d, int j) const {
        Dual<1> r; r.v = 0; r.d[0] = 0;
        if (j >= 0 && j < M && j + d < M) { r.v = band[(size_t)d * M + j]; r.d[0] = dband[(size_t)d * M + j]; }
        return r;
    }
};
template <class T> struct HostRhs {
    const double* b; int M;
    T operator()(int j) const { return make_scalar<T>((j >= 0 && j < M) ? b[j] : 0.0, 0.0); }
};

template <class T, int K>
static int run_chain(int M, int P, const double* band, const double* dband, const double* rhs, double* scal,
                     double* x, double* sig) {
    ChunkLayout lay = make_layout(M, K, P);
    const int ns = lay.max_size();
    std::vector<T> cols(ColumnStore<T, K, true>::count(ns, lay.P));
    std::vector<char> small(ChainSmall<T, K>::bytes(lay.P) + 64);
    ChainWork<T, K> w;
    w.cols = ColumnStore<T, K, true>{cols.data(), ns, lay.P};
    ChainSmall<T, K>::carve(lay.P, small.data(), w);
    HostMat<T> A{band, dband, M};
    HostRhs<T> b{rhs, M};
    for (int p = 0; p < lay.P; ++p) chain_phase1<T, K, true>(lay, p, A, b, w);
    // the separator system by block cyclic reduction; every inner loop over nodes is one barrier-separated phase of the kernel
    const int n = lay.P - 1;
    for (int q = 0; q < n; ++q) cr_assemble<T, K>(lay, q, w);
    for (int s = 1; s < n; s *= 2) {
        for (int i = s; i < n; i += 2 * s) cr_eliminate<T, K>(n, s, i, lay.M, w);
        for (int c = 0; c < n; c += 2 * s) cr_update<T, K>(n, s, c, w);
    }
    if (n > 0) cr_eliminate<T, K>(n, 0, 0, lay.M, w);
    ChainTotals<T, K> tot = chain_totals<T, K>(lay, w);
    if (n > 0) {
        cr_back<T, K, true, true>(n, 0, 0, w);
        for (int s = cr_top_stride(n); s >= 1; s /= 2)
            for (int i = s; i < n; i += 2 * s) cr_back<T, K, true, true>(n, s, i, w);
        for (int q = 0; q < n; ++q) cr_export<T, K, true, true>(lay, q, w);
    }
    std::vector<T> xo(M), so((size_t)(K + 1) * M, zero_of<T>());
    BandSink<T> sink{so.data(), M};
    for (int p = 0; p < lay.P; ++p) chain_phase3<T, K, true, true>(lay, p, w, xo.data(), sink);
    const int nt = sizeof(T) / sizeof(double);
    scal[0] = value_of(tot.logdet); scal[1] = value_of(tot.quad);
    scal[2] = tangent_of(tot.logdet, 0); scal[3] = tangent_of(tot.quad, 0);
    scal[4] = lay.P;
    for (int i = 0; i < M; ++i) { x[i] = value_of(xo[i]); if (nt > 1) x[M + i] = tangent_of(xo[i], 0); }
    for (size_t i = 0; i < (size_t)(K + 1) * M; ++i) {
        sig[i] = value_of(so[i]);
        if (nt > 1) sig[(size_t)(K + 1) * M + i] = tangent_of(so[i], 0);
    }
    return tot.info;
}

#define DISPATCH(K_) case K_: return dual ? run_chain<Dual<1>, K_>(M, P, band, dband, rhs, scal, x, sig) \
                                          : run_chain<double, K_>(M, P, band, dband, rhs, scal, x, sig);
extern "C" int hh_chain(int M, int K, int P, int dual, const double* band, const double* dband, const double* rhs,
                        double* scal, double* x, double* sig) {
    switch (K) { DISPATCH(1) DISPATCH(2) DISPATCH(3) DISPATCH(4) DISPATCH(5) DISPATCH(6) }
    return -1;
}

extern "C" void hh_pieces(int K, int n, const double* t, double* out) {
    for (int i = 0; i < n; ++i) {
        switch (K) {
#define PC(K_) case K_: { double w[K_ + 1]; bspline_pieces<K_>(t[i], w); for (int r = 0; r <= K_; ++r) out[(size_t)r * n + i] = w[r]; } break;
            PC(1) PC(2) PC(3) PC(4) PC(5) PC(6)
        }
    }
}

extern "C" void hh_locate(const double* mesh, int n_knots, int n, const double* x, int* out) {
    Mesh m; m.knots = mesh; m.n_knots = n_knots; m.x0 = mesh[0]; m.inv_delta = 1.0 / (mesh[1] - mesh[0]);
    for (int i = 0; i < n; ++i) out[i] = locate_interval(m, x[i], [](const double* p) { return *p; });
}

// Forward-mode dual numbers (value + NT tangents) for the banded factorisation kernels.
//
// The reference obtains d ELBO / d(variance, lengthscale, sigma^2) from TensorFlow reverse mode through the
// registered gradients of banded_matrices' cholesky_band / inverse_from_cholesky_band / solve_triang_mat
// (reference gpr.py:56-75 via example.py:31-32).  Here the same derivatives come from pushing one tangent
// through the very same elimination sweeps: exact, band-closed and with no extra pass.
#pragma once
#include "common.cuh"

namespace asvgp {

template <int NT>
struct Dual {
    double v;
    double d[NT];
};

// ---- plain doubles share the generic code through these overloads ------------------------------------------------
ASVGP_HD double value_of(double a) { return a; }
ASVGP_HD double sqrt_of(double a) { return sqrt(a); }
ASVGP_HD double log_of(double a) { return log(a); }
ASVGP_HD double recip_of(double a) { return 1.0 / a; }
ASVGP_HD double rsqrt_of(double a) {
#if defined(__CUDA_ARCH__)
    return rsqrt(a);
#else
    return 1.0 / sqrt(a);
#endif
}
template <class T> ASVGP_HD T zero_of();
template <> ASVGP_HD double zero_of<double>() { return 0.0; }
template <class T> ASVGP_HD T make_scalar(double v, double tangent);
template <> ASVGP_HD double make_scalar<double>(double v, double) { return v; }
ASVGP_HD double tangent_of(double, int) { return 0.0; }

// ---- Dual<NT> ----------------------------------------------------------------------------------------------------
template <int NT> ASVGP_HD double value_of(const Dual<NT>& a) { return a.v; }
template <int NT> ASVGP_HD double tangent_of(const Dual<NT>& a, int i) { return a.d[i]; }

template <int NT> ASVGP_HD Dual<NT> operator+(const Dual<NT>& a, const Dual<NT>& b) {
    Dual<NT> r; r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <int NT> ASVGP_HD Dual<NT> operator-(const Dual<NT>& a, const Dual<NT>& b) {
    Dual<NT> r; r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <int NT> ASVGP_HD Dual<NT> operator-(const Dual<NT>& a) {
    Dual<NT> r; r.v = -a.v;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = -a.d[i];
    return r;
}
template <int NT> ASVGP_HD Dual<NT> operator*(const Dual<NT>& a, const Dual<NT>& b) {
    Dual<NT> r; r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = fma(a.v, b.d[i], a.d[i] * b.v);
    return r;
}
template <int NT> ASVGP_HD Dual<NT>& operator+=(Dual<NT>& a, const Dual<NT>& b) { a = a + b; return a; }
template <int NT> ASVGP_HD Dual<NT>& operator-=(Dual<NT>& a, const Dual<NT>& b) { a = a - b; return a; }

template <int NT> ASVGP_HD Dual<NT> recip_of(const Dual<NT>& a) {
    Dual<NT> r; r.v = 1.0 / a.v;
    const double m = -r.v * r.v;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = m * a.d[i];
    return r;
}
template <int NT> ASVGP_HD Dual<NT> sqrt_of(const Dual<NT>& a) {
    Dual<NT> r; r.v = sqrt(a.v);
    const double m = 0.5 / r.v;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = m * a.d[i];
    return r;
}
// 1/sqrt(a): one MUFU+Newton instead of a sqrt followed by two divisions on the elimination's critical chain
template <int NT> ASVGP_HD Dual<NT> rsqrt_of(const Dual<NT>& a) {
    Dual<NT> r; r.v = rsqrt_of(a.v);
    const double m = -0.5 * r.v * r.v * r.v;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = m * a.d[i];
    return r;
}
template <int NT> ASVGP_HD Dual<NT> log_of(const Dual<NT>& a) {
    Dual<NT> r; r.v = log(a.v);
    const double m = 1.0 / a.v;
#pragma unroll
    for (int i = 0; i < NT; ++i) r.d[i] = m * a.d[i];
    return r;
}

template <> ASVGP_HD Dual<1> zero_of<Dual<1>>() { Dual<1> r; r.v = 0.0; r.d[0] = 0.0; return r; }
template <> ASVGP_HD Dual<1> make_scalar<Dual<1>>(double v, double tangent) {
    Dual<1> r; r.v = v; r.d[0] = tangent; return r;
}

// ---- running log-determinant without a log() per pivot ---------------------------------------------------------------
// sum_j log(w_j) is kept as log(m) + e*ln2 with the product m renormalised into [1,2) by integer exponent surgery
// (one DMUL + a few ALU ops per pivot instead of a ~45-instruction fp64 log on the single-lane critical path); the
// tangent of log(w) is w'/w = w' * ip^2 with ip = 1/sqrt(w) already at hand.
struct LogAccumCore {
    double m;
    long long e;
    ASVGP_HD void init() { m = 1.0; e = 0; }
    ASVGP_HD void mul(double w) {
        m *= w;
#if defined(__CUDA_ARCH__)
        const long long bits = __double_as_longlong(m);
        const long long ex = ((bits >> 52) & 0x7ff) - 1023;
        m = __longlong_as_double((bits & ~(0x7ffLL << 52)) | (1023LL << 52));
        e += ex;
#else
        int ex;
        m = frexp(m, &ex);     // m in [0.5, 1)
        e += ex;
#endif
    }
    ASVGP_HD double result() const { return log(m) + (double)e * 0.693147180559945309417232121458; }
};

template <class T> struct LogAccum;
template <> struct LogAccum<double> {
    LogAccumCore c;
    ASVGP_HD void init() { c.init(); }
    ASVGP_HD void add(double w, double) { c.mul(w); }
    ASVGP_HD double result() const { return c.result(); }
};
template <int NT> struct LogAccum<Dual<NT>> {
    LogAccumCore c;
    double d[NT];
    ASVGP_HD void init() {
        c.init();
#pragma unroll
        for (int i = 0; i < NT; ++i) d[i] = 0.0;
    }
    ASVGP_HD void add(const Dual<NT>& w, const Dual<NT>& ip) {
        c.mul(w.v);
        const double r = ip.v * ip.v;
#pragma unroll
        for (int i = 0; i < NT; ++i) d[i] = fma(w.d[i], r, d[i]);
    }
    ASVGP_HD Dual<NT> result() const {
        Dual<NT> r; r.v = c.result();
#pragma unroll
        for (int i = 0; i < NT; ++i) r.d[i] = d[i];
        return r;
    }
};

}  // namespace asvgp

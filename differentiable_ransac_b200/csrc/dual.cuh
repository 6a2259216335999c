// Forward-mode dual numbers with a fixed number of partials.  The device math headers are templates on the
// scalar type, so instantiating them with Dual<double, 9> yields d(output)/d(E[0..8]) of a pose error in the
// same pass that computes it -- the backward of PoseLoss (loss.py:11-68) without hand-derived adjoints.
// Comparisons look at the value only, so clamps and branch selections behave like torch.min / torch.max /
// Python `if` under autograd: the untaken side contributes no derivative.
#pragma once

#include "drb_common.cuh"

namespace drb {

template <class T, int P>
struct Dual {
    T v;
    T d[P];
    DRB_HD Dual() {}
    DRB_HD Dual(T x) : v(x) {
        for (int i = 0; i < P; ++i) d[i] = T(0);
    }
    DRB_HD Dual(int x) : v(T(x)) {
        for (int i = 0; i < P; ++i) d[i] = T(0);
    }
    DRB_HD static Dual variable(T x, int which) {
        Dual r(x);
        r.d[which] = T(1);
        return r;
    }
    // a constant with the same value: what torch.tensor([...]) of tensor elements does (cv_utils.py:146-150)
    DRB_HD Dual detached() const { return Dual(v); }
};

#define DRB_DUAL template <class T, int P> DRB_HD
DRB_DUAL Dual<T, P> operator+(const Dual<T, P>& a, const Dual<T, P>& b) {
    Dual<T, P> r;
    r.v = a.v + b.v;
    for (int i = 0; i < P; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
DRB_DUAL Dual<T, P> operator-(const Dual<T, P>& a, const Dual<T, P>& b) {
    Dual<T, P> r;
    r.v = a.v - b.v;
    for (int i = 0; i < P; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
DRB_DUAL Dual<T, P> operator-(const Dual<T, P>& a) {
    Dual<T, P> r;
    r.v = -a.v;
    for (int i = 0; i < P; ++i) r.d[i] = -a.d[i];
    return r;
}
DRB_DUAL Dual<T, P> operator*(const Dual<T, P>& a, const Dual<T, P>& b) {
    Dual<T, P> r;
    r.v = a.v * b.v;
    for (int i = 0; i < P; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
DRB_DUAL Dual<T, P> operator/(const Dual<T, P>& a, const Dual<T, P>& b) {
    Dual<T, P> r;
    const T inv = T(1) / b.v;
    r.v = a.v * inv;
    for (int i = 0; i < P; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
DRB_DUAL Dual<T, P>& operator+=(Dual<T, P>& a, const Dual<T, P>& b) { a = a + b; return a; }
DRB_DUAL Dual<T, P>& operator-=(Dual<T, P>& a, const Dual<T, P>& b) { a = a - b; return a; }
DRB_DUAL Dual<T, P>& operator/=(Dual<T, P>& a, const Dual<T, P>& b) { a = a / b; return a; }
DRB_DUAL bool operator<(const Dual<T, P>& a, const Dual<T, P>& b) { return a.v < b.v; }
DRB_DUAL bool operator>(const Dual<T, P>& a, const Dual<T, P>& b) { return a.v > b.v; }
DRB_DUAL bool operator>=(const Dual<T, P>& a, const Dual<T, P>& b) { return a.v >= b.v; }
DRB_DUAL bool operator==(const Dual<T, P>& a, const Dual<T, P>& b) { return a.v == b.v; }

DRB_DUAL Dual<T, P> t_sqrt(const Dual<T, P>& a) {
    Dual<T, P> r;
    r.v = t_sqrt(a.v);
    const T g = T(0.5) / r.v;
    for (int i = 0; i < P; ++i) r.d[i] = a.d[i] * g;
    return r;
}
// `x` as a constant (no derivative); the identity for plain scalars.
template <class T>
DRB_HD T detach_value(const T& x) { return x; }
DRB_DUAL Dual<T, P> detach_value(const Dual<T, P>& x) { return x.detached(); }
#undef DRB_DUAL

}  // namespace drb

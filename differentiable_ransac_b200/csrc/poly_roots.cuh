// Real roots of a degree-10 polynomial by Sturm-sequence isolation + safeguarded
// Newton.  Everything is held in registers (fully unrolled, compile-time degrees).
//
// The reference finds the roots with a companion-matrix `eigvals` per sample
// (nister.py:361-370) and documents the Sturm scheme in math_utils.py:111-291
// (sign-change counting :168, bracket refinement :191); this is an independent
// implementation of that classical scheme.  Roots with |z| <= 1 are isolated on
// p(z); roots with |z| > 1 on the reversed polynomial w^10 p(1/w), w in (-1, 1),
// so that every evaluation happens on the unit interval (no overflow in fp32 and
// bounded condition numbers).
#pragma once

#include "drb_common.cuh"

namespace drb {

// p(z) and p'(z) by Horner; p has 11 ascending coefficients.
template <class T>
DRB_HD void poly10_eval(const T* p, T z, T& f, T& df) {
    f = p[10];
    df = T(0);
    DRB_UNROLL
    for (int i = 9; i >= 0; --i) {
        df = df * z + f;
        f = f * z + p[i];
    }
}

// Safeguarded Newton on p inside a bracket (lo, hi] that holds exactly one simple root.
template <class T>
DRB_HD T refine_bracket(const T* p, T lo, T hi) {
    T flo, fhi, d;
    poly10_eval(p, lo, flo, d);
    poly10_eval(p, hi, fhi, d);
    T z = T(0.5) * (lo + hi);
    if ((flo < T(0)) == (fhi < T(0))) return z;  // no sign change (cluster / rounding): keep the midpoint
    // fp32: the caller polishes every root by Gauss-Newton on the original constraints, so a root
    // good to ~1e-4 is enough here and the loop stays short; fp64 goes to rounding.
    const T tol = sizeof(T) == 4 ? T(1e-4) : T(4.8e-16);
    const int max_it = sizeof(T) == 4 ? 8 : 24;
    for (int it = 0; it < max_it; ++it) {
        T f, df;
        poly10_eval(p, z, f, df);
        if ((f < T(0)) == (flo < T(0))) {
            lo = z;
        } else {
            hi = z;
        }
        T zn = z - f * t_rcp(df);
        if (!(zn > lo && zn < hi)) zn = T(0.5) * (lo + hi);
        const T dz = t_abs(zn - z);
        z = zn;
        if (dz <= tol * t_max(t_abs(z), T(1e-3))) break;
    }
    return z;
}

template <class T>
struct SturmChain10 {
    // f_{i-1} = (a[i] z + b[i]) f_i - m[i] f_{i+1},  i = 1..9  (every f_i positively rescaled)
    T a[10], b[10], m[10];
    T l1, l0;  // f_9 = l1 z + l0
    T c;       // f_10
    T p[11];   // f_0 (ascending coefficients)

    DRB_HD void build(const T* coef) {
        T u[11], v[11];
        T sc = T(0);
        DRB_UNROLL
        for (int i = 0; i <= 10; ++i) sc = t_max(sc, t_abs(coef[i]));
        const T inv = sc > T(0) ? t_rcp(sc) : T(1);
        DRB_UNROLL
        for (int i = 0; i <= 10; ++i) {
            p[i] = coef[i] * inv;
            u[i] = p[i];
        }
        DRB_UNROLL
        for (int i = 0; i < 10; ++i) v[i] = T(i + 1) * p[i + 1];
        v[10] = T(0);
        {   // rescale f_1 (positive factor keeps signs)
            T s1 = T(0);
            DRB_UNROLL
            for (int i = 0; i < 10; ++i) s1 = t_max(s1, t_abs(v[i]));
            const T i1 = s1 > T(0) ? t_rcp(s1) : T(1);
            DRB_UNROLL
            for (int i = 0; i < 10; ++i) v[i] *= i1;
        }
        const T tiny = T(1e-30);
        DRB_UNROLL
        for (int i = 1; i <= 9; ++i) {
            const int d = 10 - i;  // deg v = d, deg u = d + 1
            T lead = v[d];
            if (t_abs(lead) < tiny) lead = lead < T(0) ? -tiny : tiny;
            const T il = t_rcp(lead);
            const T ai = u[d + 1] * il;
            const T bi = (u[d] - ai * v[d - 1]) * il;
            T r[11];
            T mx = T(0);
            DRB_UNROLL
            for (int j = 0; j < d; ++j) {
                T t = u[j] - bi * v[j];
                if (j > 0) t -= ai * v[j - 1];
                r[j] = -t;  // f_{i+1} = -remainder
                mx = t_max(mx, t_abs(r[j]));
            }
            if (!(mx > tiny)) mx = T(1);
            const T im = t_rcp(mx);
            a[i] = ai;
            b[i] = bi;
            m[i] = mx;
            DRB_UNROLL
            for (int j = 0; j <= d; ++j) u[j] = v[j];
            DRB_UNROLL
            for (int j = 0; j < d; ++j) v[j] = r[j] * im;
            v[d] = T(0);
        }
        // after the loop u = f_9 (degree 1), v = f_10 (degree 0)
        l0 = u[0];
        l1 = u[1];
        c = v[0];
        a[0] = b[0] = m[0] = T(0);
    }

    // shift the sign bit of v into the low end of `mask` (one funnel shift on the device)
    DRB_HD static uint32_t push_sign(uint32_t mask, float v) {
#if defined(__CUDA_ARCH__)
        return __funnelshift_l(__float_as_uint(v), mask, 1);
#else
        return (mask << 1) | (uint32_t)(v < 0.f || (v == 0.f && 1.f / v < 0.f));
#endif
    }
    DRB_HD static uint32_t push_sign(uint32_t mask, double v) {
        return (mask << 1) | (uint32_t)(v < 0.0 || (v == 0.0 && 1.0 / v < 0.0));
    }

    // Number of sign changes of the chain f_10, f_9, ..., f_0 at z.  The eleven sign bits are collected
    // in one word and adjacent bits compared at the end (an exact zero counts as positive: measure-zero
    // event, and the grid / bisection points never coincide with a root of a chain member in practice).
    DRB_HD int count(T z) const {
        T s_next = c;
        T s_cur = l1 * z + l0;
        uint32_t mask = push_sign(push_sign(0u, s_next), s_cur);
        DRB_UNROLL
        for (int i = 9; i >= 1; --i) {
            const T s_prev = (a[i] * z + b[i]) * s_cur - m[i] * s_next;
            s_next = s_cur;
            s_cur = s_prev;
            mask = push_sign(mask, s_cur);
        }
        const uint32_t flips = (mask ^ (mask >> 1)) & 0x3ffu;  // 11 values -> 10 adjacent pairs
#if defined(__CUDA_ARCH__)
        return __popc(flips);
#else
        int n = 0;
        for (uint32_t f = flips; f; f &= f - 1) ++n;
        return n;
#endif
    }

    DRB_HD void eval(T z, T& f, T& df) const { poly10_eval<T>(p, z, f, df); }

    // Safeguarded Newton on p inside a bracket (lo, hi] that holds exactly one simple root.
    DRB_HD T refine(T lo, T hi) const { return refine_bracket<T>(p, lo, hi); }

    // Real roots in (-1, 1]; returns how many were written to out[0..max_out).
    // Phase 1 -- the same work for every thread (no divergence inside a warp): Sturm counts on a
    // fixed grid of kGrid cells.  Cells holding one root become brackets directly; a cell holding
    // several is split by Sturm bisection (uncommon).  Phase 2 refines each bracket by Newton.
    static constexpr int kGrid = 16;
    DRB_HD int roots_unit(T* out, int max_out) const {
        T blo[10], bhi[10];
        const int nb = brackets_unit(blo, bhi, max_out);
        for (int r = 0; r < nb; ++r) out[r] = refine(blo[r], bhi[r]);
        return nb;
    }

    // Grid pass over cells [cell0, cell0 + ncells) of the kGrid-cell grid on (-1, 1]: roots per cell, 4 bits each
    // (a cell holds at most 10); `c_left` receives the Sturm count at the left edge of cell0.  Returns the packed
    // word.  Rolled loops keep the code small -- this routine is instruction-cache bound when unrolled.
    DRB_HD unsigned long long grid_cells(int cell0, int ncells, int& c_left) const {
        unsigned long long cells = 0ull;
        c_left = count(T(-1) + T(2 * cell0) / T(kGrid));
        int prev = c_left;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int i = 1; i <= ncells; ++i) {
            const int cur = count(T(-1) + T(2 * (cell0 + i)) / T(kGrid));
            int n = prev - cur;
            n = n < 0 ? 0 : (n > 15 ? 15 : n);
            cells |= (unsigned long long)n << (4 * (i - 1));
            prev = cur;
        }
        return cells;
    }

    // Roots in the cells of a packed word.
    DRB_HD static int cells_total(unsigned long long cells, int ncells) {
        int n = 0;
        for (int i = 0; i < ncells; ++i) n += (int)((cells >> (4 * i)) & 15ull);
        return n;
    }

    // Emit pass: `out(lo, hi)` for every root of the cells of `cells` (at most `max_out` calls): a cell holding
    // one root is its own bracket, a cell holding several is split by Sturm bisection (uncommon).
    template <class Out>
    DRB_HD int emit_brackets(int cell0, int ncells, unsigned long long cells, int c_left, int max_out, Out& out) const {
        int nb = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int i = 0; i < ncells; ++i) {
            const int n = (int)((cells >> (4 * i)) & 15ull);
            if (n == 0) continue;
            const T clo_x = T(-1) + T(2 * (cell0 + i)) / T(kGrid), chi_x = T(-1) + T(2 * (cell0 + i) + 2) / T(kGrid);
            if (n == 1) {
                if (nb < max_out) { out(clo_x, chi_x); ++nb; }
            } else {
                // several roots in this cell: peel them off from the left by bisection on the count
                T start = clo_x;
                for (int r = 0; r < n && nb < max_out; ++r) {
                    T lo = start, hi = chi_x;
                    int clo = c_left - r, chi = c_left - n;
                    for (int it = 0; it < 40; ++it) {
                        if (clo - chi <= 1) break;
                        const T mid = T(0.5) * (lo + hi);
                        if (!(mid > lo) || !(mid < hi)) break;
                        const int cm = count(mid);
                        if (c_left - cm >= r + 1) {
                            hi = mid;
                            chi = cm;
                        } else {
                            lo = mid;
                            clo = cm;
                        }
                    }
                    out(lo, hi);
                    ++nb;
                    start = hi;
                }
            }
            c_left -= n;
        }
        return nb;
    }

    struct ArrayOut {
        T* lo;
        T* hi;
        int n;
        DRB_HD void operator()(T a, T b) { lo[n] = a; hi[n] = b; ++n; }
    };

    // Phase 1 only: brackets (blo[r], bhi[r]], each holding one real root (or a cluster).
    DRB_HD int brackets_unit(T* blo, T* bhi, int max_out) const {
        int c_left;
        const unsigned long long cells = grid_cells(0, kGrid, c_left);
        ArrayOut out{blo, bhi, 0};
        return emit_brackets(0, kGrid, cells, c_left, max_out, out);
    }
};

// Root isolation only, both domains: brackets 0..n0-1 are intervals of z on p (|z| <= 1), brackets n0..nb-1
// are intervals of w = 1/z on the reversed polynomial.  Returns nb (<= 10) and n0.
template <class T>
DRB_HD int isolate_deg10(const T* coef, T* blo, T* bhi, int& n0) {
    int nb = 0;
    n0 = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int dom = 0; dom < 2; ++dom) {
        T c[11];
        DRB_UNROLL
        for (int i = 0; i <= 10; ++i) c[i] = dom ? coef[10 - i] : coef[i];
        SturmChain10<T> s;
        s.build(c);
        nb += s.brackets_unit(blo + nb, bhi + nb, 10 - nb);
        if (dom == 0) n0 = nb;
    }
    return nb;
}

// Refine bracket r of isolate_deg10 into a root z of the ORIGINAL polynomial; false if it must be dropped.
template <class T>
DRB_HD bool root_from_bracket(const T* coef, bool reversed, T lo, T hi, T& z) {
    T c[11];
    T sc = T(0);
    DRB_UNROLL
    for (int i = 0; i <= 10; ++i) sc = t_max(sc, t_abs(coef[i]));
    const T inv = sc > T(0) ? t_rcp(sc) : T(1);
    DRB_UNROLL
    for (int i = 0; i <= 10; ++i) c[i] = (reversed ? coef[10 - i] : coef[i]) * inv;
    const T w = refine_bracket<T>(c, lo, hi);
    if (!reversed) {
        z = w;
        return true;
    }
    const T aw = t_abs(w);
    if (!(aw < T(1)) || !(aw > T(1e-7))) return false;
    z = t_rcp(w);
    return true;
}

// All real roots of sum_i coef[i] z^i (degree <= 10).  Returns the count (<= 10).
template <class T>
DRB_HD int real_roots_deg10(const T* coef, T* roots) {
    int n = 0;
    // domain 0: z in (-1, 1] on p;  domain 1: w = 1/z in (-1, 1) on the reversed polynomial.
    // One rolled loop so that the chain builder / root isolation exist once in the binary.
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int dom = 0; dom < 2; ++dom) {
        T c[11];
        DRB_UNROLL
        for (int i = 0; i <= 10; ++i) c[i] = dom ? coef[10 - i] : coef[i];
        SturmChain10<T> s;
        s.build(c);
        T w[10];
        const int nw = s.roots_unit(w, 10 - n);
        for (int i = 0; i < nw; ++i) {
            if (dom == 0) {
                if (n < 10) roots[n++] = w[i];
            } else {
                const T aw = t_abs(w[i]);
                if (aw < T(1) && aw > T(1e-7) && n < 10) roots[n++] = T(1) / w[i];
            }
        }
    }
    return n;
}

}  // namespace drb

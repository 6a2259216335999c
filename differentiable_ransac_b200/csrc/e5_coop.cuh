// Five-point solver, per-SAMPLE stages run by a group of FOUR lanes (a "quad") instead of one thread.
//
// Why: with one hypothesis per thread, 32 000 hypotheses are 1 000 warps -- 1.7 per sub-partition of a B200 -- and
// the per-sample chain (null space -> 10 x 20 constraints -> Gauss-Jordan -> z-polynomials -> Sturm isolation) is
// one long dependent instruction stream per warp: latency-bound by construction (round 1: sm__warps_active 10 %,
// issue slots 36 % busy).  Four lanes per sample make 4 000 warps of the same work and take the 10 x 20 matrix out of
// shared memory into registers (5 columns x 10 rows per lane).  Replaces the same reference lines as e5_math.cuh
// (nister.py:69-176, 355-370); the per-ROOT stage (one bracket per lane, pooled) is unchanged.
//
// Everything here is written against a small group interface G so that tests/hostcheck can run the identical
// arithmetic on the CPU (four threads and a barrier stand in for the quad):
//     g.q                 lane within the quad, 0..3
//     g.shfl(v, src)      value of `v` held by lane `src` of the quad (every lane of the quad must call it)
//     g.sync()            memory ordering + convergence of the quad
// and a per-sample scratch area S (shared memory on the device), laid out by the kCo* offsets below.
//
// Work split:
//   null space     Householder factorisation redundantly in the four lanes, null vector q in lane q
//   constraints    the nine quadratic forms (six entries of E E^T - tr/2 I, three cofactors) and then the ten
//                  cubic rows are dealt to the lanes round-robin; operands come from S by lane-dependent
//                  ADDRESSES, so the instruction stream is the same in every lane (no divergence, no dynamic
//                  register indexing)
//   elimination    column-cyclic: lane q owns columns q, q+4, ..., q+16 of all ten rows (50 registers); per
//                  step the pivot column is broadcast, every lane finds the same pivot row, swaps it in (the
//                  row index stays compile-time) and updates its own columns
//   z-polynomials  lane i < 3 forms equation i; lane t < 3 expands one cofactor term of the 3 x 3 determinant;
//                  a two-step butterfly leaves the degree-10 polynomial in all four lanes
//   isolation      lanes (dom, half): two lanes per domain (|z| <= 1 on P, |z| > 1 on the reversed P) build the
//                  same Sturm chain and count sign changes on one half of the grid each
#pragma once

#include "drb_common.cuh"
#include "e5_math.cuh"

namespace drb {

constexpr int kQuad = 4;
// per-sample scratch map, in scalars
constexpr int kCoN = 0;        // N[4][9]: the basis (E5Sample::N); alive until the sample's last root is done
constexpr int kCoQ = 36;       // Q[9][10]: quadratic forms; dead after the rows -> the sample's solutions [10][9]
constexpr int kCoRows = 126;   // rows[10][20]; dead once every lane holds its columns -> the tail below
constexpr int kCoRB = 126;     //   right block of rows 4..9 after the elimination, RB[6][10]
constexpr int kCoCx = 186;     //   cx[3][4] | cy[3][4] | cq[3][5] | P[11]  (E5Sample order after N)
constexpr int kCoLo = 236;     //   bracket lower ends [10]
constexpr int kCoHi = 246;     //   bracket upper ends [10]
constexpr int kCoRev = 256;    //   bracket domain [10]: 0 = z on P, 1 = w = 1/z on the reversed polynomial
constexpr int kCoNb = 266;     //   number of brackets (as a scalar)
constexpr int kCoMask = 267;   //   bit j: solution slot j is valid            (device: reinterpreted as int)
constexpr int kCoCount = 268;  //   number of solutions after compaction      (device: reinterpreted as int)
constexpr int kCoPos = 269;    //   position of the sample in the compact list (device: reinterpreted as int)
constexpr int kCoStride = 327; // odd: "same element, consecutive samples" hits 32 different banks

// (a, b, sign) of term t of quadratic form r:  r < 6: entry (i, j), i <= j, of E E^T = sum_t e[3i+t] e[3j+t];
// r = 6 + k: cofactor of e[6+k] in det E expanded along the third row, e[a0] e[b0] - e[a1] e[b1].
DRB_HD void co_quad_term(int r, int t, int& a, int& b, float& sign) {
    if (r < 6) {
        const int i = r < 3 ? 0 : (r < 5 ? 1 : 2);
        const int j = r < 3 ? r : (r < 5 ? r - 2 : 2);
        a = 3 * i + t;
        b = 3 * j + t;
        sign = 1.f;
    } else {
        const int k = r - 6;
        const int u = (k + 1) % 3, v = (k + 2) % 3;
        a = t == 0 ? u : v;
        b = 3 + (t == 0 ? v : u);
        sign = t == 0 ? 1.f : (t == 1 ? -1.f : 0.f);
    }
}

// (quadratic form, linear form) of term k of cubic row r: rows 0..8 = entry (i, j) of Lambda E, row 9 = det E.
DRB_HD void co_row_term(int r, int k, int& qi, int& ei) {
    if (r < 9) {
        const int i = r / 3, j = r - 3 * i;
        const int lo = i < k ? i : k, hi = i < k ? k : i;
        qi = lo == 0 ? hi : (lo == 1 ? 2 + hi : 5);
        ei = 3 * k + j;
    } else {
        qi = 6 + k;
        ei = 6 + k;
    }
}

// Column-cyclic elimination of the 10 x 20 system held as M[row][local column], column = 4 * local + q.
// Forward elimination with partial pivoting (rows physically swapped, so every register index is a compile-time
// constant), then the back-substitution restricted to rows 4..8 of the right block -- all the z-polynomials need
// (the same reductions as e5_eliminate).  Returns false when a pivot is negligible; identical in the four lanes.
template <class T, class G>
DRB_HD bool co_eliminate(G& g, T (*M)[5]) {
    T scale = T(0);
    DRB_UNROLL
    for (int r = 0; r < 10; ++r) {
        DRB_UNROLL
        for (int cl = 0; cl < 5; ++cl) scale = t_max(scale, t_abs(M[r][cl]));
    }
    scale = t_max(scale, g.shfl(scale, g.q ^ 1));
    scale = t_max(scale, g.shfl(scale, g.q ^ 2));
    const T tol = scale * (sizeof(T) == 4 ? T(1e-6) : T(1e-13));
    bool ok = true;
    DRB_UNROLL
    for (int p = 0; p < 10; ++p) {
        const int owner = p & 3, pl = p >> 2;
        T col[10];
        DRB_UNROLL
        for (int r = p; r < 10; ++r) col[r] = g.shfl(M[r][pl], owner);
        int piv = p;
        T best = t_abs(col[p]), pv = col[p];
        DRB_UNROLL
        for (int r = p + 1; r < 10; ++r) {
            const T v = t_abs(col[r]);
            if (v > best) { best = v; piv = r; pv = col[r]; }
        }
        if (!(best > tol)) ok = false;
        const T ip = t_rcp(pv);
        T f[10];                                   // column p of the row that sits at r after the swap
        DRB_UNROLL
        for (int r = p + 1; r < 10; ++r) f[r] = (r == piv) ? col[p] : col[r];
        DRB_UNROLL
        for (int cl = 0; cl < 5; ++cl) {
            if (4 * cl + 3 < p) continue;          // columns left of the pivot are already zero in rows >= p
            const T a = M[p][cl];
            T prow = a;
            DRB_UNROLL
            for (int r = p + 1; r < 10; ++r) {
                if (r == piv) { prow = M[r][cl]; M[r][cl] = a; }        // swap rows p <-> piv
            }
            prow *= ip;                            // normalised pivot row
            M[p][cl] = prow;
            DRB_UNROLL
            for (int r = p + 1; r < 10; ++r) M[r][cl] -= f[r] * prow;
        }
    }
    // back-substitution restricted to rows 4..8, right block only: local columns 2..4 hold columns >= 10 (for
    // q < 2 local column 2 is a left-block column: updated too, never read again)
    DRB_UNROLL
    for (int p = 9; p >= 5; --p) {
        const int owner = p & 3, pl = p >> 2;
        DRB_UNROLL
        for (int r = 4; r < p; ++r) {
            const T f = g.shfl(M[r][pl], owner);
            DRB_UNROLL
            for (int cl = 2; cl < 5; ++cl) M[r][cl] -= f * M[p][cl];
        }
    }
    return ok;
}

// Stages 1 + 1b for one sample by its quad: points -> S (N, cx / cy / cq, P).  P[11] is also returned in every
// lane.  false: degenerate sample (identical in the four lanes).
template <class T, class G>
DRB_HD bool e5_coop_prepare(G& g, const T (*pts)[4], T* S, T* P) {
    const int q = g.q;
    {   // ---- null space: vector q in lane q ----
        T rows[5][9];
        DRB_UNROLL
        for (int j = 0; j < 5; ++j) epipolar_row(pts[j][0], pts[j][1], pts[j][2], pts[j][3], rows[j]);
        T W[9][5], beta[5], nv[9];
        householder_factor<T, 5>(rows, W, beta);
        householder_null_vector<T, 5>(W, beta, q, nv);
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) S[kCoN + q * 9 + i] = nv[i];
    }
    g.sync();
    // ---- quadratic forms: r = q, q + 4, q + 8 ----
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int r = q; r < 9; r += 4) {
        T acc[10];
        DRB_UNROLL
        for (int c = 0; c < 10; ++c) acc[c] = T(0);
        DRB_UNROLL
        for (int t = 0; t < 3; ++t) {
            int a, b;
            float sign;
            co_quad_term(r, t, a, b, sign);
            T ea[4], eb[4];
            DRB_UNROLL
            for (int c = 0; c < 4; ++c) {
                ea[c] = S[kCoN + c * 9 + a] * T(sign);     // e[i] = sum_c N[c][i] (x, y, z, 1)_c
                eb[c] = S[kCoN + c * 9 + b];
            }
            poly_mul11_acc(ea, eb, acc);
        }
        DRB_UNROLL
        for (int c = 0; c < 10; ++c) S[kCoQ + r * 10 + c] = acc[c];
    }
    g.sync();
    // Lambda = E E^T - 1/2 tr(E E^T) I: the diagonal forms are 0, 3, 5
    for (int c = q; c < 10; c += 4) {
        const T ht = T(0.5) * (S[kCoQ + c] + S[kCoQ + 30 + c] + S[kCoQ + 50 + c]);
        S[kCoQ + c] -= ht;
        S[kCoQ + 30 + c] -= ht;
        S[kCoQ + 50 + c] -= ht;
    }
    g.sync();
    // ---- cubic rows: r = q, q + 4, q + 8 ----
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int r = q; r < 10; r += 4) {
        T row[20];
        DRB_UNROLL
        for (int c = 0; c < 20; ++c) row[c] = T(0);
        DRB_UNROLL
        for (int k = 0; k < 3; ++k) {
            int qi, ei;
            co_row_term(r, k, qi, ei);
            T a[10], b[4];
            DRB_UNROLL
            for (int c = 0; c < 10; ++c) a[c] = S[kCoQ + qi * 10 + c];
            DRB_UNROLL
            for (int c = 0; c < 4; ++c) b[c] = S[kCoN + c * 9 + ei];
            poly_mul21_acc(a, b, row);
        }
        DRB_UNROLL
        for (int c = 0; c < 20; ++c) S[kCoRows + r * 20 + c] = row[c];
    }
    g.sync();
    // ---- every lane takes its five columns, then the rows' storage is free ----
    T M[10][5];
    DRB_UNROLL
    for (int r = 0; r < 10; ++r) {
        DRB_UNROLL
        for (int cl = 0; cl < 5; ++cl) M[r][cl] = S[kCoRows + r * 20 + 4 * cl + q];
    }
    g.sync();
    bool ok = co_eliminate<T, G>(g, M);
    // ---- right block of rows 4..9 -> S ----
    DRB_UNROLL
    for (int r = 4; r < 10; ++r) {
        DRB_UNROLL
        for (int cl = 2; cl < 5; ++cl) {
            const int c = 4 * cl + q;
            if (c >= 10) S[kCoRB + (r - 4) * 10 + (c - 10)] = M[r][cl];
        }
    }
    g.sync();
    if (q < 3) {   // equation q:  cx x + cy y + cq = 0 with polynomial coefficients in z (ascending powers)
        T ra[10], rb[10];
        DRB_UNROLL
        for (int c = 0; c < 10; ++c) {
            ra[c] = S[kCoRB + (2 * q) * 10 + c];
            rb[c] = S[kCoRB + (2 * q + 1) * 10 + c];
        }
        T* cx = S + kCoCx + 4 * q;
        T* cy = S + kCoCx + 12 + 4 * q;
        T* cq = S + kCoCx + 24 + 5 * q;
        cx[0] = ra[2]; cx[1] = ra[1] - rb[2]; cx[2] = ra[0] - rb[1]; cx[3] = -rb[0];
        cy[0] = ra[5]; cy[1] = ra[4] - rb[5]; cy[2] = ra[3] - rb[4]; cy[3] = -rb[3];
        cq[0] = ra[9]; cq[1] = ra[8] - rb[9]; cq[2] = ra[7] - rb[8]; cq[3] = ra[6] - rb[7];
        cq[4] = -rb[6];
    }
    g.sync();
    // ---- determinant polynomial: lane t < 3 expands the cofactor term of cq[t]; butterfly sum ----
    DRB_UNROLL
    for (int i = 0; i <= 10; ++i) P[i] = T(0);
    {
        const int t = q < 3 ? q : 0;
        const int r = (t == 0) ? 1 : 0;
        const int s = (t == 2) ? 1 : 2;
        const T sign = (q >= 3) ? T(0) : ((t == 1) ? T(-1) : T(1));
        T xr[4], yr[4], xs[4], ys[4], qt[5];
        DRB_UNROLL
        for (int i = 0; i < 4; ++i) {
            xr[i] = S[kCoCx + 4 * r + i];
            yr[i] = S[kCoCx + 12 + 4 * r + i];
            xs[i] = S[kCoCx + 4 * s + i];
            ys[i] = S[kCoCx + 12 + 4 * s + i];
        }
        DRB_UNROLL
        for (int i = 0; i < 5; ++i) qt[i] = S[kCoCx + 24 + 5 * t + i] * sign;
        T mn[7];
        DRB_UNROLL
        for (int i = 0; i < 7; ++i) mn[i] = T(0);
        DRB_UNROLL
        for (int i = 0; i < 4; ++i) {
            DRB_UNROLL
            for (int j = 0; j < 4; ++j) mn[i + j] += xr[i] * ys[j] - xs[i] * yr[j];
        }
        DRB_UNROLL
        for (int i = 0; i < 7; ++i) {
            DRB_UNROLL
            for (int j = 0; j < 5; ++j) P[i + j] += mn[i] * qt[j];
        }
    }
    DRB_UNROLL
    for (int i = 0; i <= 10; ++i) {
        P[i] += g.shfl(P[i], q ^ 1);
        P[i] += g.shfl(P[i], q ^ 2);
    }
    DRB_UNROLL
    for (int i = 0; i <= 10; ++i) ok = ok && (P[i] == P[i]) && (t_abs(P[i]) < T(1e30));
    if (q == 0) {
        DRB_UNROLL
        for (int i = 0; i <= 10; ++i) S[kCoCx + 39 + i] = P[i];
    }
    return ok;
}

// Appends (lo, hi, domain) records to the sample's bracket list in S, at most 10 in all.
template <class T>
struct CoBracketOut {
    T* S;
    int pos;
    T dom;
    DRB_HD void operator()(T lo, T hi) {
        if (pos < 10) {
            S[kCoLo + pos] = lo;
            S[kCoHi + pos] = hi;
            S[kCoRev + pos] = dom;
        }
        ++pos;
    }
};

// Stage 2 for one sample by its quad: P -> brackets in S.  Lane (dom, half) = (q >> 1, q & 1).  Returns the number
// of brackets (<= 10), identical in the four lanes; S[kCoNb] is written by the caller.
template <class T, class G>
DRB_HD int e5_coop_isolate(G& g, const T* P, T* S) {
    const int q = g.q, dom = q >> 1, half = q & 1;
    constexpr int kHalf = SturmChain10<T>::kGrid / 2;
    T c[11];
    DRB_UNROLL
    for (int i = 0; i <= 10; ++i) c[i] = dom ? P[10 - i] : P[i];
    SturmChain10<T> s;
    s.build(c);
    int c_left;
    const unsigned long long cells = s.grid_cells(half * kHalf, kHalf, c_left);
    const int mine = SturmChain10<T>::cells_total(cells, kHalf);
    const int n0 = g.shfl(mine, 0), n1 = g.shfl(mine, 1), n2 = g.shfl(mine, 2), n3 = g.shfl(mine, 3);
    const int before = (q > 0 ? n0 : 0) + (q > 1 ? n1 : 0) + (q > 2 ? n2 : 0);
    CoBracketOut<T> out{S, before, T(dom)};
    const int room = before < 10 ? 10 - before : 0;
    s.emit_brackets(half * kHalf, kHalf, cells, c_left, room, out);
    const int total = n0 + n1 + n2 + n3;
    return total < 10 ? total : 10;
}

// The per-sample data of a parked sample -> E5Sample (what e5_model_from_root takes).
template <class T>
DRB_HD void co_fetch_sample(const T* S, E5Sample<T>& smp) {
    DRB_UNROLL
    for (int a = 0; a < 4; ++a) {
        DRB_UNROLL
        for (int i = 0; i < 9; ++i) smp.N[a][i] = S[kCoN + a * 9 + i];
    }
    DRB_UNROLL
    for (int a = 0; a < 3; ++a) {
        DRB_UNROLL
        for (int i = 0; i < 4; ++i) {
            smp.cx[a][i] = S[kCoCx + 4 * a + i];
            smp.cy[a][i] = S[kCoCx + 12 + 4 * a + i];
        }
        DRB_UNROLL
        for (int i = 0; i < 5; ++i) smp.cq[a][i] = S[kCoCx + 24 + 5 * a + i];
    }
    DRB_UNROLL
    for (int i = 0; i <= 10; ++i) smp.P[i] = S[kCoCx + 39 + i];
}

}  // namespace drb

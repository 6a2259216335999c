// Operation profile of the five-point solver's math (csrc/e5_math.cuh, poly_roots.cuh) per stage, on the host: the
// templates are instantiated with a scalar type that counts its arithmetic, and with a matrix type that counts its
// element accesses (each one is a shared-memory load or store in solve_e5_kernel).  Tells where the ~29 000
// warp-instructions per hypothesis of the kernel go.   g++ -O1 -std=c++17 -I differentiable_ransac_b200/csrc
//   profiles/microbench/e5_opcount.cpp -o /tmp/e5_opcount && /tmp/e5_opcount
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "drb_common.cuh"

namespace drb {
struct Ops {
    long long add = 0, mul = 0, div = 0, cmp = 0, sqrt = 0, mat = 0;
    long long total() const { return add + mul + div + cmp + sqrt; }
};
static Ops g_ops;

struct Counted {
    float v;
    Counted() : v(0.f) {}
    Counted(float x) : v(x) {}
    Counted(double x) : v((float)x) {}
    Counted(int x) : v((float)x) {}
    operator float() const { return v; }
};
inline Counted operator+(Counted a, Counted b) { ++g_ops.add; return Counted(a.v + b.v); }
inline Counted operator-(Counted a, Counted b) { ++g_ops.add; return Counted(a.v - b.v); }
inline Counted operator*(Counted a, Counted b) { ++g_ops.mul; return Counted(a.v * b.v); }
inline Counted operator/(Counted a, Counted b) { ++g_ops.div; return Counted(a.v / b.v); }
inline Counted operator-(Counted a) { return Counted(-a.v); }
inline Counted& operator+=(Counted& a, Counted b) { ++g_ops.add; a.v += b.v; return a; }
inline Counted& operator-=(Counted& a, Counted b) { ++g_ops.add; a.v -= b.v; return a; }
inline Counted& operator*=(Counted& a, Counted b) { ++g_ops.mul; a.v *= b.v; return a; }
inline Counted& operator/=(Counted& a, Counted b) { ++g_ops.div; a.v /= b.v; return a; }
#define DRB_CMP(op) inline bool operator op(Counted a, Counted b) { ++g_ops.cmp; return a.v op b.v; }
DRB_CMP(<) DRB_CMP(>) DRB_CMP(<=) DRB_CMP(>=) DRB_CMP(==) DRB_CMP(!=)
inline Counted t_sqrt(Counted x) { ++g_ops.sqrt; return Counted(sqrtf(x.v)); }
inline Counted t_rsqrt(Counted x) { ++g_ops.sqrt; return Counted(1.f / sqrtf(x.v)); }
}  // namespace drb

#include "e5_math.cuh"

namespace {
using drb::Counted;
struct CountedMat {
    Counted v[10][20];
    Counted& operator()(int r, int c) { ++drb::g_ops.mat; return v[r][c]; }
};
struct Stage {
    const char* name;
    drb::Ops ops;
};
drb::Ops diff(const drb::Ops& a, const drb::Ops& b) {
    drb::Ops d;
    d.add = a.add - b.add; d.mul = a.mul - b.mul; d.div = a.div - b.div; d.cmp = a.cmp - b.cmp; d.sqrt = a.sqrt - b.sqrt;
    d.mat = a.mat - b.mat;
    return d;
}
void acc(drb::Ops& t, const drb::Ops& d) { t.add += d.add; t.mul += d.mul; t.div += d.div; t.cmp += d.cmp; t.sqrt += d.sqrt; t.mat += d.mat; }
}  // namespace

int main() {
    using namespace drb;
    static_assert(sizeof(Counted) == 4, "the math headers switch tolerances on sizeof(T)");
    const int K = 2000;
    srand(7);
    auto rnd = []() { return (float)rand() / RAND_MAX - 0.5f; };
    Stage st[7] = {{"null space (Householder 5x9)", {}}, {"constraints (10x20 coefficients)", {}}, {"elimination (Gauss-Jordan 10x20)", {}},
                   {"z-polynomials + determinant", {}}, {"Sturm chain + root isolation", {}}, {"roots: Newton in the bracket", {}},
                   {"roots: back-substitution + polish + normalise", {}}};
    long long roots = 0, models = 0;
    for (int k = 0; k < K; ++k) {
        // a random relative pose and five exact correspondences
        float R[9], t[3], q[4] = {rnd(), rnd(), rnd(), rnd() + 1.5f};
        const float qn = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        for (float& x : q) x /= qn;
        R[0] = 1 - 2 * (q[2] * q[2] + q[3] * q[3]); R[1] = 2 * (q[1] * q[2] - q[0] * q[3]); R[2] = 2 * (q[1] * q[3] + q[0] * q[2]);
        R[3] = 2 * (q[1] * q[2] + q[0] * q[3]); R[4] = 1 - 2 * (q[1] * q[1] + q[3] * q[3]); R[5] = 2 * (q[2] * q[3] - q[0] * q[1]);
        R[6] = 2 * (q[1] * q[3] - q[0] * q[2]); R[7] = 2 * (q[2] * q[3] + q[0] * q[1]); R[8] = 1 - 2 * (q[1] * q[1] + q[2] * q[2]);
        t[0] = rnd(); t[1] = rnd(); t[2] = rnd();
        Counted pts[5][4];
        for (int j = 0; j < 5; ++j) {
            const float X[3] = {2 * rnd(), 2 * rnd(), 5.f + rnd()};
            float Y[3];
            for (int i = 0; i < 3; ++i) Y[i] = R[3 * i] * X[0] + R[3 * i + 1] * X[1] + R[3 * i + 2] * X[2] + t[i];
            pts[j][0] = X[0] / X[2]; pts[j][1] = X[1] / X[2]; pts[j][2] = Y[0] / Y[2]; pts[j][3] = Y[1] / Y[2];
        }
        CountedMat M;
        E5Sample<Counted> S;
        Ops o0 = g_ops;
        {
            Counted rows[5][9];
            for (int j = 0; j < 5; ++j) epipolar_row(pts[j][0], pts[j][1], pts[j][2], pts[j][3], rows[j]);
            null_space_rows<Counted, 5>(rows, S.N);
        }
        Ops o1 = g_ops; acc(st[0].ops, diff(o1, o0));
        e5_constraints<Counted, CountedMat>(S.N, M);
        Ops o2 = g_ops; acc(st[1].ops, diff(o2, o1));
        bool ok = e5_eliminate<Counted, CountedMat>(M);
        Ops o3 = g_ops; acc(st[2].ops, diff(o3, o2));
        {   // the tail of e5_prepare_from_null, re-done on a copy so the stages can be separated
            CountedMat M2;
            e5_constraints<Counted, CountedMat>(S.N, M2);
            Ops before = g_ops;
            (void)before;
        }
        g_ops = o3;   // discard the re-done constraints
        {
            CountedMat M3;
            std::memcpy(&M3, &M, sizeof(M));
            // e5_prepare_from_null = constraints + eliminate + this tail; run it whole and subtract the first two
            Ops b0 = g_ops;
            ok = e5_prepare_from_null<Counted, CountedMat>(M3, S) && ok;
            Ops b1 = g_ops;
            Ops whole = diff(b1, b0), first = diff(o3, o1);
            Ops tail;
            tail.add = whole.add - first.add; tail.mul = whole.mul - first.mul; tail.div = whole.div - first.div;
            tail.cmp = whole.cmp - first.cmp; tail.sqrt = whole.sqrt - first.sqrt; tail.mat = whole.mat - first.mat;
            acc(st[3].ops, tail);
        }
        if (!ok) continue;
        Counted Pr[11], blo[10], bhi[10];
        for (int i = 0; i <= 10; ++i) Pr[i] = S.P[i];
        int n0 = 0;
        Ops o4 = g_ops;
        const int nb = isolate_deg10<Counted>(Pr, blo, bhi, n0);
        Ops o5 = g_ops; acc(st[4].ops, diff(o5, o4));
        for (int r = 0; r < nb; ++r) {
            Counted zr;
            Ops a0 = g_ops;
            const bool got = root_from_bracket<Counted>(Pr, r >= n0, blo[r], bhi[r], zr);
            Ops a1 = g_ops; acc(st[5].ops, diff(a1, a0));
            ++roots;
            if (!got) continue;
            Counted E[9];
            const bool good = e5_model_from_root<Counted>(S, zr, 2, E);
            Ops a2 = g_ops; acc(st[6].ops, diff(a2, a1));
            models += good;
        }
    }
    printf("five-point solver, %d random minimal samples: %.2f brackets, %.2f models per sample\n", K, (double)roots / K, (double)models / K);
    printf("%-48s %9s %9s %7s %7s %6s %9s %9s\n", "stage (per sample)", "add/sub", "mul", "div", "cmp", "sqrt", "arith", "mat acc");
    Ops tot;
    for (const Stage& s : st) {
        printf("%-48s %9.1f %9.1f %7.1f %7.1f %6.1f %9.1f %9.1f\n", s.name, (double)s.ops.add / K, (double)s.ops.mul / K, (double)s.ops.div / K,
               (double)s.ops.cmp / K, (double)s.ops.sqrt / K, (double)s.ops.total() / K, (double)s.ops.mat / K);
        acc(tot, s.ops);
    }
    printf("%-48s %9.1f %9.1f %7.1f %7.1f %6.1f %9.1f %9.1f\n", "total", (double)tot.add / K, (double)tot.mul / K, (double)tot.div / K,
           (double)tot.cmp / K, (double)tot.sqrt / K, (double)tot.total() / K, (double)tot.mat / K);
    return 0;
}

// TEST INFRASTRUCTURE ONLY.  Compiles the device math headers
// (differentiable_ransac_b200/csrc/*_math.cuh) for the HOST so the CPU test
// suite (no GPU in the build container) can run the exact arithmetic of the
// CUDA kernels against the oracle.  The package never loads this library; the
// product path is the CUDA library and fails loudly without it.
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "e5_math.cuh"
#include "e5_coop.cuh"
#include "e5_backward.cuh"
#include "f8_math.cuh"
#include "rigid_math.cuh"
#include "refit_math.cuh"
#include "pose_math.cuh"
#include "msac_tc_layout.cuh"

namespace {
template <class T>
struct HostMat {
    T v[10][20];
    T& operator()(int r, int c) { return v[r][c]; }
};

template <class T, class RT = T>
void run_e5(const T* pts, int K, T* models, int* nsol, int polish) {
    for (int k = 0; k < K; ++k) {
        T p[5][4];
        for (int j = 0; j < 5; ++j)
            for (int c = 0; c < 4; ++c) p[j][c] = pts[(k * 5 + j) * 4 + c];
        HostMat<T> M;
        T out[10][9];
        for (int s = 0; s < 10; ++s)
            for (int i = 0; i < 9; ++i) out[s][i] = (i % 4 == 0) ? T(1) : T(0);   // identity padding
        drb::ArrayModelSink<T> sink{out};
        nsol[k] = drb::e5_solve<T, HostMat<T>, RT, drb::ArrayModelSink<T>>(p, M, sink, polish);
        std::memcpy(models + (size_t)k * 90, out, sizeof(out));
    }
}
}  // namespace

// ---- the cooperative five-point stages (e5_coop.cuh) on the host: four threads are the quad's lanes, a barrier is
// the warp's lock step.  A shuffle is publish / barrier / read / barrier; g.sync() is one barrier. ----------------
struct QuadShared {
    std::atomic<int> arrived{0};
    std::atomic<int> phase{0};
    float fslot[4];
    int islot[4];
    void barrier(int& local_phase) {
        const int next = local_phase ^ 1;
        if (arrived.fetch_add(1, std::memory_order_acq_rel) == 3) {
            arrived.store(0, std::memory_order_relaxed);
            phase.store(next, std::memory_order_release);
        } else {
            while (phase.load(std::memory_order_acquire) != next) std::this_thread::yield();
        }
        local_phase = next;
    }
};
struct QuadHost {
    int q;
    QuadShared* sh;
    int local_phase = 0;
    float shfl(float v, int src) {
        sh->fslot[q] = v;
        sh->barrier(local_phase);
        const float r = sh->fslot[src];
        sh->barrier(local_phase);
        return r;
    }
    int shfl(int v, int src) {
        sh->islot[q] = v;
        sh->barrier(local_phase);
        const int r = sh->islot[src];
        sh->barrier(local_phase);
        return r;
    }
    void sync() { sh->barrier(local_phase); }
};

static void run_e5_coop(const float* pts, int K, float* models, int* nsol, int polish, float* aux) {
    QuadShared sh;
    std::vector<float> S(drb::kCoStride);
    int nb_shared = 0;
    auto lane = [&](int q) {
        QuadHost g{q, &sh};
        for (int k = 0; k < K; ++k) {
            float p[5][4], P[11];
            for (int j = 0; j < 5; ++j)
                for (int c = 0; c < 4; ++c) p[j][c] = pts[(k * 5 + j) * 4 + c];
            const bool ok = drb::e5_coop_prepare<float, QuadHost>(g, p, S.data(), P);
            int nb = 0;
            if (ok) nb = drb::e5_coop_isolate<float, QuadHost>(g, P, S.data());
            if (q == 0) nb_shared = nb;
            g.sync();
            if (q == 0) {   // the per-root stage, serially (on the device: one bracket per lane, pooled)
                float out[10][9];
                for (int s = 0; s < 10; ++s)
                    for (int i = 0; i < 9; ++i) out[s][i] = (i % 4 == 0) ? 1.f : 0.f;
                int n = 0;
                drb::E5Sample<float> smp;
                drb::co_fetch_sample<float>(S.data(), smp);
                for (int j = 0; j < nb_shared; ++j) {
                    float z, E[9];
                    if (!drb::root_from_bracket<float>(smp.P, S[drb::kCoRev + j] != 0.f, S[drb::kCoLo + j],
                                                       S[drb::kCoHi + j], z))
                        continue;
                    if (!drb::e5_model_from_root<float>(smp, z, polish, E)) continue;
                    std::memcpy(out[n++], E, sizeof(E));
                }
                nsol[k] = n;
                std::memcpy(models + (size_t)k * 90, out, sizeof(out));
                if (aux) {   // P (11), N (36) for stage-level comparisons
                    std::memcpy(aux + (size_t)k * 47, smp.P, 11 * sizeof(float));
                    std::memcpy(aux + (size_t)k * 47 + 11, smp.N, 36 * sizeof(float));
                }
            }
            g.sync();
        }
    };
    std::thread t1(lane, 1), t2(lane, 2), t3(lane, 3);
    lane(0);
    t1.join();
    t2.join();
    t3.join();
}

extern "C" {
void hc_e5_coop_f32(const float* pts, int K, float* models, int* nsol, int polish, float* aux) {
    run_e5_coop(pts, K, models, nsol, polish, aux);
}
void hc_e5_solve_f32(const float* pts, int K, float* models, int* nsol, int polish) {
    run_e5<float>(pts, K, models, nsol, polish);
}
void hc_e5_solve_f32_r64(const float* pts, int K, float* models, int* nsol, int polish) {
    run_e5<float, double>(pts, K, models, nsol, polish);
}
void hc_e5_solve_f64(const double* pts, int K, double* models, int* nsol, int polish) {
    run_e5<double>(pts, K, models, nsol, polish);
}
int hc_e5_backward_f64(const double* pts, const double* E, const double* g, double* gp) {
    double p[5][4], o[5][4];
    std::memcpy(p, pts, sizeof(p));
    bool ok = drb::e5_backward<double, double>(p, E, g, o);
    std::memcpy(gp, o, sizeof(o));
    return ok ? 1 : 0;
}
int hc_e5_backward_f32(const float* pts, const float* E, const float* g, float* gp) {
    float p[5][4], o[5][4];
    std::memcpy(p, pts, sizeof(p));
    bool ok = drb::e5_backward<float, double>(p, E, g, o);
    std::memcpy(gp, o, sizeof(o));
    return ok ? 1 : 0;
}
// ---- 8-point / 7-point / rigid ------------------------------------------------------------
#define HC_F8(NAME, T)                                                                         \
    void NAME(const T* pts, int K, T* F, int* ok) {                                            \
        for (int k = 0; k < K; ++k) {                                                          \
            T p[8][4];                                                                         \
            std::memcpy(p, pts + (size_t)k * 32, sizeof(p));                                   \
            ok[k] = drb::f8_solve<T>(p, F + (size_t)k * 9) ? 1 : 0;                            \
        }                                                                                      \
    }
HC_F8(hc_f8_solve_f32, float)
HC_F8(hc_f8_solve_f64, double)
#define HC_F8B(NAME, T)                                                                        \
    void NAME(const T* pts, const T* g, int K, T* gp, int* ok) {                               \
        for (int k = 0; k < K; ++k) {                                                          \
            T p[8][4], o[8][4];                                                                \
            std::memcpy(p, pts + (size_t)k * 32, sizeof(p));                                   \
            ok[k] = drb::f8_backward<T, double>(p, g + (size_t)k * 9, o) ? 1 : 0;              \
            std::memcpy(gp + (size_t)k * 32, o, sizeof(o));                                    \
        }                                                                                      \
    }
HC_F8B(hc_f8_backward_f32, float)
HC_F8B(hc_f8_backward_f64, double)
#define HC_F7(NAME, T)                                                                         \
    void NAME(const T* pts, int K, T* F, int* n) {                                             \
        for (int k = 0; k < K; ++k) {                                                          \
            T p[7][4], o[3][9];                                                                \
            std::memcpy(p, pts + (size_t)k * 28, sizeof(p));                                   \
            n[k] = drb::f7_solve<T>(p, o);                                                     \
            std::memcpy(F + (size_t)k * 27, o, sizeof(o));                                     \
        }                                                                                      \
    }
HC_F7(hc_f7_solve_f32, float)
HC_F7(hc_f7_solve_f64, double)
#define HC_RIGID(NAME, T)                                                                      \
    void NAME(const T* pts, int K, int flag, T* model, int* ok) {                              \
        for (int k = 0; k < K; ++k) {                                                          \
            T p[3][6];                                                                         \
            std::memcpy(p, pts + (size_t)k * 18, sizeof(p));                                   \
            ok[k] = drb::rigid3_solve<T>(p, flag, model + (size_t)k * 16) ? 1 : 0;             \
        }                                                                                      \
    }
HC_RIGID(hc_rigid3_solve_f32, float)
HC_RIGID(hc_rigid3_solve_f64, double)
#define HC_RIGIDB(NAME, T)                                                                     \
    void NAME(const T* pts, const T* g, int K, int flag, T* gp, int* ok) {                     \
        for (int k = 0; k < K; ++k) {                                                          \
            T p[3][6], o[3][6];                                                                \
            std::memcpy(p, pts + (size_t)k * 18, sizeof(p));                                   \
            ok[k] = drb::rigid3_backward<T, double>(p, flag, g + (size_t)k * 16, o) ? 1 : 0;   \
            std::memcpy(gp + (size_t)k * 18, o, sizeof(o));                                    \
        }                                                                                      \
    }
HC_RIGIDB(hc_rigid3_backward_f32, float)
HC_RIGIDB(hc_rigid3_backward_f64, double)

// rigid residual from second moments (score.cu: rigid_residual_moments_kernel): points [N,6] float, models [K,16]
// float -> res [K], g [K,12] (= d res / d [R | t], i.e. twice the header's g12), accumulated in double
void hc_rigid_residual_moments(const float* points, int N, const float* models, int K, double* res, double* g) {
    double mom[drb::kRigidMoments] = {0};
    for (int n = 0; n < N; ++n) {
        const double p[3] = {points[6 * n], points[6 * n + 1], points[6 * n + 2]};
        const double q[3] = {points[6 * n + 3], points[6 * n + 4], points[6 * n + 5]};
        drb::rigid_moments_add<double>(p, q, mom);
    }
    for (int k = 0; k < K; ++k) {
        double g12[12];
        drb::rigid_residual_from_moments<float, double>(mom, (double)N, models + (size_t)k * 16, res[k], g12);
        for (int i = 0; i < 12; ++i) g[(size_t)k * 12 + i] = 2.0 * g12[i];
    }
}

// Serial restatement of refit.cu's moment accumulation + the shared serial tail (refit_math.cuh).
int hc_refit(int fmat, const float* matches, const unsigned char* mask, const float* weights, int N, float* models) {
    drb::HartleyNorm<double> h;
    h.m[0] = h.m[1] = h.m[2] = h.m[3] = 0.0;
    h.r1 = h.r2 = 1.0;
    double count = 0.0, s[4] = {0, 0, 0, 0};
    for (int n = 0; n < N; ++n) {
        if (mask && !mask[n]) continue;
        count += 1.0;
        for (int c = 0; c < 4; ++c) s[c] += matches[n * 4 + c];
    }
    if (count < (fmat ? 8 : 5)) return 0;
    if (fmat) {
        for (int c = 0; c < 4; ++c) h.m[c] = s[c] / count;
        double d1 = 0.0, d2 = 0.0;
        for (int n = 0; n < N; ++n) {
            if (mask && !mask[n]) continue;
            const double a = matches[n * 4] - h.m[0], b = matches[n * 4 + 1] - h.m[1];
            const double c = matches[n * 4 + 2] - h.m[2], e = matches[n * 4 + 3] - h.m[3];
            d1 += std::sqrt(a * a + b * b);
            d2 += std::sqrt(c * c + e * e);
        }
        h.r1 = 1.4142135623730951 / (d1 / count);
        h.r2 = 1.4142135623730951 / (d2 / count);
    }
    double acc[45];
    for (int i = 0; i < 45; ++i) acc[i] = 0.0;
    for (int n = 0; n < N; ++n) {
        if (mask && !mask[n]) continue;
        double row[9];
        drb::epipolar_row<double>((matches[n * 4] - h.m[0]) * h.r1, (matches[n * 4 + 1] - h.m[1]) * h.r1,
                                  (matches[n * 4 + 2] - h.m[2]) * h.r2, (matches[n * 4 + 3] - h.m[3]) * h.r2, row);
        const double ww = weights ? (double)weights[n] * (double)weights[n] : 1.0;
        int e = 0;
        for (int i = 0; i < 9; ++i)
            for (int j = i; j < 9; ++j) acc[e++] += ww * row[i] * row[j];
    }
    if (fmat) {
        double F[9];
        if (!drb::f8_refit_from_moments<double>(acc, h, F)) return 0;
        for (int i = 0; i < 9; ++i) models[i] = (float)F[i];
        return 1;
    }
    double E[10][9];
    const int n_out = drb::e5_refit_from_moments<double>(acc, E);
    for (int k = 0; k < n_out; ++k)
        for (int i = 0; i < 9; ++i) models[k * 9 + i] = (float)E[k][i];
    return n_out;
}

// Serial restatement of pose.cu: decompose, vote over the correspondences, pick the pose, mask, errors.
int hc_recover_pose(const double* E, const float* matches, int N, double dist, const double* R_gt,
                    const double* t_gt, double* R, double* t, unsigned char* mask, int* counts, double* err) {
    drb::PoseCandidates<double> pc;
    if (!drb::decompose_essential<double>(E, pc)) return -1;
    std::vector<int> bits(N);
    for (int c = 0; c < 4; ++c) counts[c] = 0;
    for (int n = 0; n < N; ++n) {
        bits[n] = drb::cheirality_bits<double>(pc, matches[n * 4], matches[n * 4 + 1], matches[n * 4 + 2],
                                               matches[n * 4 + 3], dist);
        for (int c = 0; c < 4; ++c) counts[c] += (bits[n] >> c) & 1;
    }
    int best = 0;
    for (int c = 1; c < 4; ++c)
        if (counts[c] > counts[best]) best = c;
    for (int i = 0; i < 9; ++i) R[i] = (best & 1) ? pc.R2[i] : pc.R1[i];
    for (int i = 0; i < 3; ++i) t[i] = (best & 2) ? -pc.t[i] : pc.t[i];
    for (int n = 0; n < N; ++n) mask[n] = (bits[n] >> best) & 1;
    if (R_gt && t_gt) drb::pose_errors_deg<double>(R, t, R_gt, t_gt, err[0], err[1]);
    return best;
}

// PoseLoss term of one model (loss.py:57-63 with svd=False): Horn decomposition, cheirality vote, then
// (err_R + err_t) / 2 in degrees and its gradient with respect to E by forward-mode duals.
int hc_pose_loss(const double* E, const float* matches, int N, double dist, const double* R_gt, const double* t_gt,
                 double* err, double* grad) {
    typedef drb::Dual<double, 9> D;
    drb::PoseCandidates<double> pc;
    if (!drb::decompose_essential_horn<double>(E, pc)) return -1;
    int counts[4] = {0, 0, 0, 0};
    for (int n = 0; n < N; ++n) {
        const int bits = drb::cheirality_bits<double>(pc, matches[n * 4], matches[n * 4 + 1], matches[n * 4 + 2],
                                                      matches[n * 4 + 3], dist);
        for (int c = 0; c < 4; ++c) counts[c] += (bits >> c) & 1;
    }
    int best = 0;
    for (int c = 1; c < 4; ++c)
        if (counts[c] > counts[best]) best = c;
    D Ed[9], Rg[9], tg[3], R[9], t[3], er, et;
    for (int i = 0; i < 9; ++i) { Ed[i] = D::variable(E[i], i); Rg[i] = D(R_gt[i]); }
    for (int i = 0; i < 3; ++i) tg[i] = D(t_gt[i]);
    drb::PoseCandidates<D> pd;
    drb::decompose_essential_horn<D>(Ed, pd);
    for (int i = 0; i < 9; ++i) R[i] = (best & 1) ? pd.R2[i] : pd.R1[i];
    for (int i = 0; i < 3; ++i) t[i] = (best & 2) ? -pd.t[i] : pd.t[i];
    drb::pose_errors_deg<D>(R, t, Rg, tg, er, et);
    err[0] = er.v;
    err[1] = et.v;
    for (int i = 0; i < 9; ++i) grad[i] = 0.5 * (er.d[i] + et.d[i]);
    return best;
}

// ---- tensor-core scorer (score_tc.cu): the operand images, descriptors and column mapping on the host ----
// A software model of what the kernel asks the hardware to do: the operands are read back from the images
// THROUGH THE DESCRIPTOR FIELDS (start address, leading / stride byte offsets of the canonical K-major
// no-swizzle layout, M and N of the instruction descriptor), eight K values per MMA step, TF32 inputs (low 13
// bits ignored), fp32 accumulation; the epilogue follows the kernel's thread mapping (lane quarter, column
// half, four 32-column loads of eight model pairs).  Exact division stands in for rcp.approx.
uint64_t hc_tc_smem_desc(uint32_t addr) { return drb::tc::smem_desc(addr); }
uint32_t hc_tc_instr_desc(void) { return drb::tc::instr_desc(); }
int hc_tc_abytes(void) { return drb::tc::kABytes; }
int hc_tc_bbytes(void) { return drb::tc::kBBytes; }

static float tf32_trunc(float x) {
    uint32_t u;
    std::memcpy(&u, &x, 4);
    u &= 0xffffe000u;
    std::memcpy(&x, &u, 4);
    return x;
}

// element (row, kk) of one MMA K step of the operand described by `desc`: 32-bit containers (tf32, 8 per step)
// or 16-bit ones (bf16, 16 per step); either way a 16-byte chunk per row and two chunks per step
static float desc_fetch(const std::vector<uint32_t>& smem, uint64_t desc, int row, int kk, bool bf16) {
    const uint32_t start = (uint32_t)(desc & 0x3fff) << 4;
    const uint32_t lbo = (uint32_t)((desc >> 16) & 0x3fff) << 4;
    const uint32_t sbo = (uint32_t)((desc >> 32) & 0x3fff) << 4;
    const int per_chunk = bf16 ? 8 : 4, bytes = bf16 ? 2 : 4;
    const uint32_t addr = start + (row >> 3) * sbo + (kk / per_chunk) * lbo + (row & 7) * 16 + (kk % per_chunk) * bytes;
    const uint32_t w = smem[addr / 4];
    if (bf16) return drb::tc::bf16_value((uint16_t)((addr & 2) ? (w >> 16) : (w & 0xffffu)));
    float f;
    std::memcpy(&f, &w, 4);
    return f;
}

int hc_msac_tc_scores(const float* matches, int N, const float* models, int M, float thr, int words, float* scores,
                      int* lossless) {
    using namespace drb::tc;
    const bool bf16 = (words & 15) == 3, pair = (words & 16) != 0, fold = (words & 128) != 0;
    const uint32_t idesc = bf16 ? instr_desc_bf16() : instr_desc();
    const int mmaN = (int)((idesc >> 17) & 0x3f) << 3, mmaM = (int)((idesc >> 24) & 0x1f) << 4;
    if (mmaN != kTileN || mmaM != kTileM) return -1;
    const uint32_t fmt = bf16 ? 1u : 2u;
    if (((idesc >> 4) & 3) != 1 || ((idesc >> 7) & 7) != fmt || ((idesc >> 10) & 7) != fmt) return -2;
    if (((idesc >> 15) & 3) != 0) return -3;                                                         // both K-major
    const int step_k = bf16 ? 16 : 8;
    const uint32_t a_addr = 0x400, b_addr = 0x400 + kABytes;      // a made-up shared-memory carve-up
    std::vector<uint32_t> smem((b_addr + kBBytes) / 4, 0u);
    const int tiles = (N + kTileM - 1) / kTileM;
    const float th = 1.5f * thr, nci = -1.f / (th * th);
    *lossless = 1;
    for (int m0 = 0; m0 < M; m0 += kTileModels) {
        // the builder warps
        for (int i = 0; i < kTileModels; ++i) {
            float m[9], cr[kFeat], cj[kFeat], cr15 = 0.f, cj15 = 0.f;
            uint32_t row48[kK];
            for (int q = 0; q < 9; ++q) m[q] = (m0 + i < M) ? models[(size_t)(m0 + i) * 9 + q] : 0.f;
            if (fold) model_rows_folded(m, m0 + i < M, -(th * th), cr, cj, cr15, cj15);
            else model_rows(m, m0 + i < M, pair, cr, cj);
            operand_row_words(cr, false, bf16, row48, cr15);
            for (int k = 0; k < kK; ++k) smem[b_addr / 4 + image_index(column_r(i), k)] = row48[k];
            operand_row_words(cj, false, bf16, row48, cj15);
            for (int k = 0; k < kK; ++k)
                smem[b_addr / 4 + image_index(pair ? column_j_swapped(i) : column_j(i), k)] = row48[k];
        }
        std::vector<float> lane_sum((size_t)kTileM * kTileModels, 0.f);   // per epilogue thread (lane, model)
        for (int t = 0; t < tiles; ++t) {
            // msac_tc_features_kernel
            for (int row = 0; row < kTileM; ++row) {
                const int n = t * kTileM + row;
                uint32_t row48[kK];
                bool real = n < N;
                if (real && fold)
                    for (int q = 0; q < 4; ++q) real = real && std::fabs(matches[n * 4 + q]) <= 1.0e18f;
                if (real) {
                    float f[kFeat];
                    features(matches[n * 4], matches[n * 4 + 1], matches[n * 4 + 2], matches[n * 4 + 3], f);
                    operand_row_words(f, true, bf16, row48);
                } else if (fold) {
                    float f[kFeat];
                    for (int k = 0; k < kFeat; ++k) f[k] = 0.f;
                    operand_row_words(f, true, bf16, row48, 1.f);
                } else {
                    for (int k = 0; k < kK; ++k) row48[k] = 0u;
                }
                for (int k = 0; k < kK; ++k) smem[a_addr / 4 + image_index(row, k)] = row48[k];
            }
            // the MMA warp: six K steps through the descriptors
            std::vector<float> D((size_t)kTileM * kTileN, 0.f);
            const uint64_t adesc = smem_desc(a_addr), bdesc = smem_desc(b_addr);
            for (int s = 0; s < kKSteps; ++s) {
                const uint64_t ad = smem_desc_kstep(adesc, s), bd = smem_desc_kstep(bdesc, s);
                for (int r = 0; r < kTileM; ++r)
                    for (int c = 0; c < kTileN; ++c) {
                        float acc = s ? D[(size_t)r * kTileN + c] : 0.f;
                        for (int kk = 0; kk < step_k; ++kk) {
                            const float a = desc_fetch(smem, ad, r, kk, bf16), b = desc_fetch(smem, bd, c, kk, bf16);
                            if (a != a || b != b) { acc = a * b; continue; }
                            if (!bf16 && (tf32_trunc(a) != a || tf32_trunc(b) != b)) *lossless = 0;
                            acc += a * b;
                        }
                        D[(size_t)r * kTileN + c] = acc;
                    }
            }
            // the epilogue warps
            for (int quarter = 0; quarter < 4; ++quarter)
                for (int half = 0; half < 2; ++half)
                    for (int lane = 0; lane < 32; ++lane) {
                        const int row = quarter * 32 + lane;
                        const float one = (t * kTileM + row < N) ? 1.f : 0.f;
                        for (int c = 0; c < 4; ++c)
                            for (int q = 0; q < 8; ++q) {
                                const int col = half * 128 + c * 32 + 4 * q;
                                const float ja = D[(size_t)row * kTileN + col + 2], jb = D[(size_t)row * kTileN + col + 3];
                                const float tn = (1.f / (ja * jb)) * nci;        // pair variant: one reciprocal
                                for (int h = 0; h < 2; ++h) {
                                    const float r = D[(size_t)row * kTileN + col + h];
                                    float v;
                                    if (fold) {
                                        // columns (r0, r1, j1', j0'): the term minus its "1 +", t max(r^2 j_other', -p)
                                        const float pp = ja * jb, tt = 1.f / pp, w = (r * r) * (h ? jb : ja);
                                        const float mx = (w != w) ? -pp : (w > -pp ? w : -pp);    // FMNMX: NaN loses
                                        lane_sum[(size_t)row * kTileModels + half * 64 + 2 * (c * 8 + q) + h] += tt * mx;
                                        continue;
                                    }
                                    if (pair) {
                                        v = ((r * r) * (h ? jb : ja)) * tn + one;     // columns (r0, r1, j1, j0)
                                    } else {
                                        const float j = h ? jb : ja;
                                        v = ((r * r) * (1.f / j)) * nci + one;
                                    }
                                    v = (v != v) ? 0.f : (v < 0.f ? 0.f : (v > 1.f ? 1.f : v));   // FFMA.SAT
                                    const int model = half * 64 + 2 * (c * 8 + q) + h;
                                    lane_sum[(size_t)row * kTileModels + model] += v;
                                }
                            }
                    }
        }
        for (int i = 0; i < kTileModels; ++i) {
            float q4[4] = {0.f, 0.f, 0.f, 0.f};
            for (int row = 0; row < kTileM; ++row)
                q4[row >> 5] += lane_sum[(size_t)row * kTileModels + i] + (fold ? (float)tiles : 0.f);
            float sc = ((q4[0] + q4[1]) + q4[2]) + q4[3];
            if (fold) sc = (sc != sc) ? 0.f : (sc < 0.f ? 0.f : sc);
            if (m0 + i < M) scores[m0 + i] = sc;
        }
    }
    return 0;
}
// ---- model-stationary arrangement (score_tc2.cu): A = the models' words in tensor memory (lane = model,
// column = K index, 16-bit elements two per column), B = a tile of 80 correspondences read through the descriptor,
// two MMAs per K step (r and j), epilogue lane = model summing over its columns.
int hc_msac_tc2_scores(const float* matches, int N, const float* models, int M, float thr, int words, float* scores) {
    using namespace drb::tc;
    const bool bf16 = (words & 15) == 3;
    const int kPts = 80, kPtBytes = (kPts / 8) * kSBO;
    const uint32_t idesc = instr_desc_mn(128, kPts, bf16);
    if (((int)((idesc >> 17) & 0x3f) << 3) != kPts || ((int)((idesc >> 24) & 0x1f) << 4) != 128) return -1;
    const uint32_t fmt = bf16 ? 1u : 2u;
    if (((idesc >> 4) & 3) != 1 || ((idesc >> 7) & 7) != fmt || ((idesc >> 10) & 7) != fmt || ((idesc >> 15) & 3) != 0) return -2;
    const int step_k = bf16 ? 16 : 8;
    const uint32_t p_addr = 0x800;
    std::vector<uint32_t> smem((p_addr + kPtBytes) / 4, 0u);
    const int tiles = (N + kPts - 1) / kPts;
    const float th = 1.5f * thr, nci = -1.f / (th * th);
    for (int m0 = 0; m0 < M; m0 += 128) {
        // builders: 48 columns of words per model and type
        std::vector<uint32_t> tm_r(128 * kK), tm_j(128 * kK);
        for (int i = 0; i < 128; ++i) {
            float m[9], cr[kFeat], cj[kFeat];
            for (int q = 0; q < 9; ++q) m[q] = (m0 + i < M) ? models[(size_t)(m0 + i) * 9 + q] : 0.f;
            model_rows(m, m0 + i < M, false, cr, cj);
            operand_row_words(cr, false, bf16, &tm_r[i * kK]);
            operand_row_words(cj, false, bf16, &tm_j[i * kK]);
        }
        auto a_elem = [&](const std::vector<uint32_t>& tm, int lane, int col0, int kk) -> float {
            // K step starting at column col0: element kk -> column col0 + kk (32-bit) or col0 + kk / 2 (16-bit)
            if (bf16) {
                const uint32_t w = tm[lane * kK + col0 + kk / 2];
                return bf16_value((uint16_t)((kk & 1) ? (w >> 16) : (w & 0xffffu)));
            }
            float f;
            std::memcpy(&f, &tm[lane * kK + col0 + kk], 4);
            return f;
        };
        std::vector<float> sum(128, 0.f);
        for (int t = 0; t < tiles; ++t) {
            for (int row = 0; row < kPts; ++row) {
                const int n = t * kPts + row;
                uint32_t row48[kK];
                if (n < N) {
                    float f[kFeat];
                    features(matches[n * 4], matches[n * 4 + 1], matches[n * 4 + 2], matches[n * 4 + 3], f);
                    operand_row_words(f, true, bf16, row48);
                } else {
                    for (int k = 0; k < kK; ++k) row48[k] = 0u;
                }
                for (int k = 0; k < kK; ++k) smem[p_addr / 4 + image_index(row, k)] = row48[k];
            }
            const uint64_t bdesc = smem_desc(p_addr);
            const bool paired = (words & 16) && !((N & 1) && t == tiles - 1);   // as the kernel decides
            for (int lane = 0; lane < 128; ++lane) {
                float rr[80], jj[80];
                for (int c = 0; c < kPts; ++c) {
                    float r = 0.f, j = 0.f;
                    for (int s = 0; s < kKSteps; ++s) {
                        const uint64_t bd = smem_desc_kstep(bdesc, s);
                        for (int kk = 0; kk < step_k; ++kk) {
                            const float p = desc_fetch(smem, bd, c, kk, bf16);
                            r += a_elem(tm_r, lane, 8 * s, kk) * p;
                            j += a_elem(tm_j, lane, 8 * s, kk) * p;
                        }
                    }
                    rr[c] = r;
                    jj[c] = j;
                }
                for (int c = 0; c < kPts; c += 2) {
                    float v0, v1;
                    if (paired) {
                        const float tn = (1.f / (jj[c] * jj[c + 1])) * nci;
                        v0 = ((rr[c] * rr[c]) * jj[c + 1]) * tn + 1.f;
                        v1 = ((rr[c + 1] * rr[c + 1]) * jj[c]) * tn + 1.f;
                    } else {
                        v0 = ((rr[c] * rr[c]) * (1.f / jj[c])) * nci + 1.f;
                        v1 = ((rr[c + 1] * rr[c + 1]) * (1.f / jj[c + 1])) * nci + 1.f;
                    }
                    v0 = (v0 != v0) ? 0.f : (v0 < 0.f ? 0.f : (v0 > 1.f ? 1.f : v0));   // FFMA.SAT (NaN -> 0: rows past N)
                    v1 = (v1 != v1) ? 0.f : (v1 < 0.f ? 0.f : (v1 > 1.f ? 1.f : v1));
                    sum[lane] += v0 + v1;
                }
            }
        }
        for (int i = 0; i < 128 && m0 + i < M; ++i) scores[m0 + i] = sum[i];
    }
    return 0;
}
uint32_t hc_tc_instr_desc_bf16(void) { return drb::tc::instr_desc_bf16(); }
// w0 + w1 + w2 of the BF16 split (exactness check)
double hc_tc_bf16_sum(float x) {
    uint16_t w[3];
    drb::tc::bf16_split3(x, w);
    return (double)drb::tc::bf16_value(w[0]) + (double)drb::tc::bf16_value(w[1]) + (double)drb::tc::bf16_value(w[2]);
}

int hc_roots_f32(const float* coef, float* roots) { return drb::real_roots_deg10<float>(coef, roots); }
int hc_roots_f64(const double* coef, double* roots) { return drb::real_roots_deg10<double>(coef, roots); }
}

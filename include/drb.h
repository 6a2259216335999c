/* drb.h -- C ABI of the B200 differentiable-RANSAC hot path (libdrb.so).
 *
 * The reference (weitong8591/differentiable_ransac) is pure Python and has no
 * native boundary; these entry points are what a ctypes / cffi binding inside
 * the reference's plugin classes calls instead of the torch tensor graph.  Each
 * entry cites the reference interface it replaces.  See INTEGRATION.md for the
 * reference-side stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates);
 *     the library never allocates, frees or keeps state between calls;
 *   - tensors are contiguous row-major fp32 unless stated; B pairs, K hypotheses
 *     (minimal samples) per pair, N correspondences per pair, s sample size;
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     synchronise, and are re-entrant;
 *   - return 0 on success or a negative drb_status; nothing is thrown across
 *     the boundary.  Launch errors are reported by the call that caused them;
 *     asynchronous execution errors by drb_last_error().
 *   - per-hypothesis numerical failure is NOT an error: the slot gets the
 *     identity model and valid = 0 (mirrors nister.py:400-405).
 *   - model layout: 3x3 row-major M with x2^T M x1 = 0, x1 = matches[:,0:2],
 *     x2 = matches[:,2:4] (the convention of every reference estimator).
 */
#ifndef DRB_H_
#define DRB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    DRB_OK = 0,
    DRB_ERR_NULL_POINTER = -1,
    DRB_ERR_BAD_SHAPE = -2,
    DRB_ERR_UNSUPPORTED = -3,
    DRB_ERR_CUDA = -4
} drb_status;

#define DRB_E5_SLOTS 10 /* models per 5-point sample, as the reference (nister.py:400) */
#define DRB_F7_SLOTS 3  /* real solutions of the 7-point cubic */

int drb_version(void);
const char* drb_status_string(int status);
/* cudaGetLastError()/cudaPeekAtLastError() of the calling thread's device, as an int. */
int drb_last_error(void);
/* Shared-memory / occupancy facts used by the host mirror and the benchmark. */
int drb_device_sm_count(void);

/* ---- a1 + a2: Gumbel top-s sampler and minimal-sample gather ------------------------------
 * Replaces GumbelSoftmaxSampler.sample (samplers/gumbel_sampler.py:25-42) and the gather
 * ransac.py:64-65.  For every (pair b, hypothesis k): keys[n] = (logits[b,n] + G[b,k,n]) / tau,
 * the s largest keys are selected and returned as point indices in ASCENDING order (the order
 * of the reference's boolean-mask gather).  G is read from `noise` ([B,K,N]) when non-null
 * (parity mode), otherwise generated in-kernel with Philox4x32-10 keyed by (seed, offset).
 * Optional outputs (nullable): lse[B,K] = logsumexp_n keys (needed by the backward),
 * sel_key[B,K,s] = keys of the selected points (same order as idx), noise_out[B,K,N] = the
 * Gumbel noise actually used (lets a test replay a Philox run through the oracle).
 * offset_dev (nullable): one uint64 in device memory ADDED to `offset` when the kernel runs, so a captured
 * CUDA graph of a training step draws fresh noise on every replay (the caller advances it on the stream;
 * drb_sample_backward must be given the same pointer before it is advanced).                */
int drb_sample(const float* logits, const float* noise, uint64_t seed, uint64_t offset,
               const uint64_t* offset_dev, float tau, int B, int K, int N, int s,
               int32_t* idx, float* lse, float* sel_key, float* noise_out, void* stream);

/* Test-mode set sampler: when only the K minimal sets are wanted (no log-sum-exp, no noise dump,
 * no injected noise) the Gumbel noise need not exist at all.  "Top-s of logits + Gumbel" is sampling
 * s items WITHOUT replacement from softmax(logits) (Plackett-Luce), drawn here directly by inverse
 * CDF on a per-pair prefix sum: one thread per hypothesis, O(s log N) instead of O(N).  Same
 * distribution as drb_sample, different realisation.  idx[B,K,s] ascending.  Requires N*4 bytes of
 * shared memory (N <= 51200), else DRB_ERR_UNSUPPORTED (use drb_sample).
 * offset_dev (nullable): one uint64 in device memory that is ADDED to `offset` when the kernel runs, so a
 * captured CUDA graph draws fresh sets on every replay (the caller advances it on the same stream).  */
int drb_sample_sets(const float* logits, uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int B,
                    int K, int N, int s, int32_t* idx, void* stream);

/* Straight-through backward of the sampler (autograd of gumbel_sampler.py:34-38 + ransac.py:64-65).
 * g_sel[B,K,s] = dL/d ret[b,k,idx[b,k,j]] = sum_c matches[n,c] * dL/d minimal[b,k,j,c].
 * Adds to grad_logits[B,N] (caller zero-initialises):
 *   (1/tau) * sum_k y[k,n] (g[k,n] - sum_m y[k,m] g[k,m]),  y = softmax_n(keys).
 * The noise is re-read (`noise` non-null) or regenerated from (seed, offset); the K x N
 * soft one-hot is never stored.  `scratch` is B*K floats of caller-owned workspace.          */
int drb_sample_backward(const float* logits, const float* noise, uint64_t seed, uint64_t offset,
                        const uint64_t* offset_dev /* nullable, as in drb_sample */, float tau,
                        int B, int K, int N, int s,
                        const int32_t* idx, const float* lse, const float* sel_key, const float* g_sel,
                        float* scratch /* [B,K] */, float* grad_logits, void* stream);

/* ---- a3 / a4: five-point essential matrix ---------------------------------------------------
 * Replaces EssentialMatrixEstimatorNister.estimate_minimal_model (nister.py:69-408) and
 * EssentialMatrixEstimator.estimate_minimal_model (stewenius.py:20-80).
 * Input is either gathered (idx == NULL: `matches` is [B*K,5,4] minimal samples, N ignored) or
 * indexed (idx[B,K,5] into matches[B,N,4]).  Outputs: models[B,K,10,9] (unit Frobenius norm),
 * nsol[B,K] real solutions in slots 0..nsol-1 (others = identity).  Optional compact list for
 * the scorer: cmodels[B,K*10,9], cids[B,K*10] (= k*10+slot), ccount[B] (caller zeroes).     */
int drb_solve_e5(const float* matches, const int32_t* idx, int B, int K, int N,
                 float* models, int32_t* nsol, float* cmodels, int32_t* cids, int32_t* ccount, void* stream);

/* Backward of the five-point solver by the implicit-function theorem at the solution
 * (replaces autograd through nister.py:117-399).  For every (b,k) and the slot `sel[b,k]`
 * (-1 = no gradient): g_model[B,K,9] = dL/dE  ->  g_pts[B,K,5,4] = dL/d minimal sample.     */
int drb_solve_e5_backward(const float* matches, const int32_t* idx, int B, int K, int N,
                          const float* models, const int32_t* sel, const float* g_model,
                          float* g_pts, void* stream);
/* Train mode in one launch (ransac.py:78-96): the five-point solve AND the choice of the slot closest to the ground
 * truth gt[B,9] (Frobenius; up to sign when sign_invariant), made while the solutions are still on chip.
 * sel[B,K] = chosen slot or -1, chosen[B,K,9] = that model (identity when none).  models (nullable) = the dense
 * [B,K,10,9] output of drb_solve_e5 when a caller wants all slots too.                        */
int drb_solve_e5_select(const float* matches, const int32_t* idx, const float* gt, int sign_invariant,
                        int B, int K, int N, float* models, int32_t* nsol, int32_t* sel, float* chosen, void* stream);
/* drb_solve_e5_backward given only the chosen models [B,K,9] (what drb_solve_e5_select leaves behind).  */
int drb_solve_e5_backward_chosen(const float* matches, const int32_t* idx, int B, int K, int N,
                                 const float* chosen, const int32_t* sel, const float* g_model, float* g_pts,
                                 void* stream);


/* Train-mode slot selection, ransac.py:87-96: per sample the slot closest to gt[B,9] in
 * Frobenius norm (sign_invariant != 0: min(||E-gt||, ||E+gt||)).  Writes
 * sel[B,K] (-1 when nsol == 0) and chosen[B,K,9].                                             */
int drb_select_closest(const float* models, const int32_t* nsol, const float* gt, int B, int K, int slots,
                       int sign_invariant, int32_t* sel, float* chosen, void* stream);

/* ---- a5 / a6: fundamental matrix ------------------------------------------------------------
 * Replaces FundamentalMatrixEstimatorNew.normalize + estimate_non_minimal_model
 * (fundamental_matrix_estimator.py:177-260): Hartley-normalised s-point (s = 8) null vector,
 * F = T2^T Fn T1, unit-norm Fn, no rank-2 projection.  models[B,K,9], valid[B,K].          */
int drb_solve_f8(const float* matches, const int32_t* idx, int B, int K, int N,
                 float* models, uint8_t* valid, void* stream);
int drb_solve_f8_backward(const float* matches, const int32_t* idx, int B, int K, int N,
                          const float* models, const float* g_model, float* g_pts, void* stream);
/* Correct 7-point (the reference's, fundamental_matrix_estimator.py:262-308, is broken: SURVEY D4).
 * models[B,K,3,9], nsol[B,K].                                                                 */
int drb_solve_f7(const float* matches, const int32_t* idx, int B, int K, int N,
                 float* models, int32_t* nsol, void* stream);

/* ---- a7: rigid 3-point ----------------------------------------------------------------------
 * Replaces RigidTransformationSVDBasedSolver.estimate_model
 * (rigid_transformation_SVD_based_solver.py:11-74), both `flag` branches.  points [B,N,6] or
 * gathered [B*K,3,6]; models[B,K,16] = 4x4 pose row-major; valid[B,K].                      */
int drb_solve_rigid3(const float* points, const int32_t* idx, int B, int K, int N, int flag,
                     float* models, uint8_t* valid, void* stream);
int drb_solve_rigid3_backward(const float* points, const int32_t* idx, int B, int K, int N, int flag,
                              const float* models, const float* g_model, float* g_pts, void* stream);

/* ---- a9: Sampson / soft-MSAC scoring with fused arg-max -------------------------------------
 * Replaces MSACScore.score (scorings/msac_score.py:12-55) and the arg-max of ransac.py:114.
 * models[B,M,9]; count[B] (nullable: all M) = number of leading models to score per pair;
 * ids[B,M] (nullable: identity) = caller's index of each model, used for the tie-break and
 * reported in best_id; thr[B] = the (already normalised) inlier threshold per pair;
 * scores[B,M] (nullable) in the order of `models`.  best_packed[B] (caller zeroes) receives
 * max over models of (score bits << 32 | ~id): decode with drb_best_finalize.  NaN scores
 * never win.                                                                     */
int drb_score_msac(const float* matches, const float* models, const int32_t* count, const int32_t* ids,
                   const float* thr, int B, int M, int N,
                   float* scores, unsigned long long* best_packed, void* stream);
/* The same contract on a work queue (score_stream.cu): persistent one-warp CTAs pull (32 models x 192
 * correspondences) units from an atomic counter, so no warp idles while another still has work; the pieces
 * of a model block are summed in piece order by the warp that delivers the last one (the scores do not
 * depend on the schedule).  B <= 1024.  Needs a 16-byte aligned workspace of
 * drb_score_msac_workspace_bytes(B, M, N) bytes whose first drb_score_msac_workspace_zeroed_bytes(B, M)
 * bytes are ZERO on entry; the kernel leaves them zero, so one memset serves every later call that uses
 * the workspace on the same stream.                                                              */
size_t drb_score_msac_workspace_bytes(int B, int M, int N);
size_t drb_score_msac_workspace_zeroed_bytes(int B, int M);
int drb_score_msac_stream(const float* matches, const float* models, const int32_t* count, const int32_t* ids,
                          const float* thr, int B, int M, int N,
                          float* scores, unsigned long long* best_packed, void* workspace,
                          size_t workspace_bytes, void* stream);
/* The same contract with the contraction on the tensor cores (score_tc.cu; 0.122 ms
 * against 0.215 ms at the headline shape).  r = x2' M x1 and the Sampson denominator are two polynomials
 * in the correspondence's coordinates, i.e. inner products of 15 monomials with per-model coefficients; a
 * (128 correspondences x 128 models) tile is one 128 x 256 x 48 tcgen05 MMA (split operands, fp32
 * accumulation in tensor memory) and the CUDA cores keep r^2 / j -> clamp -> sum.  `words` selects the variant:
 *    2   operands split into two TF32 words, three partial products (22-bit operands: scores within
 *        ~1.4e-4 relative of drb_score_msac) -- measured on B200
 *    3   three BF16 words, six partial products (exact operands: fp32-level scores), same MMA time
 *  + 16  one reciprocal per PAIR of neighbouring models, rcp(j0 j1) (j1, j0): half the work of the pipe that
 *        bounds the kernel; a model with a non-finite coefficient still scores 0 without touching its neighbour
 *  + 32  16 epilogue warps per CTA instead of 8
 *  + 64  the model-stationary arrangement of score_tc2.cu: a unit's 128 models live in tensor memory as the A
 *        operand, only tiles of 80 correspondences pass through shared memory (a third of the operand traffic,
 *        one accumulator register per thread); with + 16 it pairs neighbouring correspondences
 *  + 128 (with + 16, not with + 64) the threshold folded into the denominator rows, j' = -(1.5 thr)^2 j, and the
 *        term written 1 + t max(r^2 j_other', -p), p = j0' j1', t = 1 / p: the clamp runs on the ALU pipe and the
 *        multiply-add accumulates, 7 FMA-pipe cycles per two pairs instead of 10; rows past N and correspondences
 *        that are not finite are flagged in the sixteenth K slot and answer exactly -1 (msac_tc_layout.cuh)
 *  + 256 (with + 16 alone) the slim build: 128 registers per thread and three correspondence stages, which leaves 16 K
 *        registers and ~50 KB of shared memory of every SM to a five-point CTA of the next batch on another stream
 *  + 512 (with + 16 alone) two SMs per tile (score_tc_pair.cu): clusters of two CTAs, tcgen05 cta_group::2 MMAs of 256
 *        correspondences x 128 models, half of the operand stream and of the model operand per CTA; correct and slower
 *        (0.138 vs 0.108 ms at the headline shape: the cross-SM hand-over every six MMAs), opt-in
 * (every variant is pinned to the fp64 oracle on the B200, tests/test_gpu_score_tc.py.)  B <= 1024; matches 16-byte
 * aligned.  Needs a 128-byte aligned workspace of drb_score_msac_tc_workspace_bytes(B, N) bytes (contents
 * irrelevant on entry: the call writes the operand images of the correspondences there first).           */
size_t drb_score_msac_tc_workspace_bytes(int B, int N);
int drb_score_msac_tc(const float* matches, const float* models, const int32_t* count, const int32_t* ids,
                      const float* thr, int B, int M, int N, int words,
                      float* scores, unsigned long long* best_packed, void* workspace,
                      size_t workspace_bytes, void* stream);
/* Host -> device copy of a caller's buffer on `stream` (cudaMemcpyAsync; asynchronous when src is pinned): the
 * pipelined service (engine.E5TestService / RANSACLayer.submit) enqueues a batch's inputs through it.            */
int drb_copy_h2d_async(void* dst_device, const void* src_host, size_t bytes, void* stream);

/* ---- the same chain in double precision (`-pr 2`: utils.py:42, model_cl.py:164-170) -------------------------
 * In the reference the precision flag sets the dtype of the sampler's one-hot, and the minimal samples, the five-point
 * solver (nister.py:121-122) and MSAC (msac_score.py:12-55) follow by type promotion.  These entries run that chain in
 * float64 on the device with the same templated math as the fp32 kernels (fp64_path.cu): built for results, not for
 * the roofline.  matches[B,N,4] (or [B*K,5,4] when idx is null), thr[B] are DOUBLE.
 *   drb_solve_e5_f64      idx[B,K,5] -> models[B,K,10,9] (identity in the unused slots), nsol[B,K]
 *   drb_score_msac_f64    models[B,K*slots,9], nsol[B,K] (nullable: every slot holds a model) -> scores[B,K*slots]
 *                         (-1 for an empty slot)
 *   drb_best_finalize_f64 scores[B,M] -> best_id[B] (first maximum, like torch.argmax at ransac.py:114; -1 if none),
 *                         best_score[B], best_model[B,9], mask[B,N] (d2 < (1.5 thr)^2), ninl[B]                     */
int drb_solve_e5_f64(const double* matches, const int32_t* idx, int B, int K, int N, double* models, int32_t* nsol,
                     void* stream);
int drb_score_msac_f64(const double* matches, const double* models, const int32_t* nsol, const double* thr, int B,
                       int K, int slots, int N, double* scores, void* stream);
int drb_best_finalize_f64(const double* matches, const double* models, const double* scores, const double* thr,
                          int B, int M, int N, int32_t* best_id, double* best_score, double* best_model,
                          uint8_t* mask, int32_t* ninl, void* stream);
/* Decode best_packed and produce the winner's model, score, id and inlier mask
 * (d2 < (1.5 thr)^2, msac_score.py:44) -- the only mask ransac.py:116-118 ever uses.
 * models_dense[B,Md,9] is indexed by best id.  best_id[B], best_score[B], best_model[B,9],
 * mask[B,N] (uint8), ninl[B].                                                                */
int drb_best_finalize(const float* matches, const float* models_dense, const unsigned long long* best_packed,
                      const float* thr, int B, int Md, int N,
                      int32_t* best_id, float* best_score, float* best_model, uint8_t* mask, int32_t* ninl,
                      void* stream);

/* ---- a10: symmetric epipolar loss (MatchLoss core) ------------------------------------------
 * Replaces batch_episym (model_cl.py:13-26) + min(.,1) + mean of loss.py:138-151 over the
 * GT-inlier subset.  pts[B,P,4] (the inlier correspondences, P may be padded: npts[B] valid),
 * models[B,K,9], mvalid[B,K] (nullable).  row_sum[B,K] = sum_p min(episym, 1).            */
int drb_episym_forward(const float* pts, const int32_t* npts, const float* models, const uint8_t* mvalid,
                       int B, int K, int P, float* row_sum, void* stream);
/* g_row[B,K] = dL/d row_sum  ->  g_models[B,K,9] (clamped entries contribute zero).          */
int drb_episym_backward(const float* pts, const int32_t* npts, const float* models, const uint8_t* mvalid,
                        const float* g_row, int B, int K, int P, float* g_models, void* stream);

/* Both in ONE pass over the points (the training step knows g_row before the loss value: it depends only on which
 * models are valid): row_sum[B,K] as drb_episym_forward, g_models[B,K,9] as drb_episym_backward.            */
int drb_episym_forward_backward(const float* pts, const int32_t* npts, const float* models, const uint8_t* mvalid,
                                const float* g_row, int B, int K, int P, float* row_sum, float* g_models,
                                void* stream);

/* ---- a8: rigid squared residual -------------------------------------------------------------
 * Replaces squared_residual (rigid_transformation_SVD_based_solver.py:76-89).
 * points[B,N,6], models[B,K,16] -> res_sum[B,K] = sum_n ||q - (R p + t)||^2,
 * ninl[B,K] (nullable) = #(d2 < threshold).                                                  */
int drb_rigid_residual_forward(const float* points, const float* models, int B, int K, int N, float threshold,
                               float* res_sum, int32_t* ninl, void* stream);
/* g_res[B,K] -> g_models[B,K,16] (only the 3x4 [R|t] block is written).                      */
int drb_rigid_residual_backward(const float* points, const float* models, const float* g_res,
                                int B, int K, int N, float* g_models, void* stream);

/* res_sum and g_models in one pass over the points (see drb_episym_forward_backward).          */
int drb_rigid_residual_forward_backward(const float* points, const float* models, const float* g_res,
                                        int B, int K, int N, float* res_sum, float* g_models, void* stream);

/* ---- a2 backward: scatter minimal-sample gradients ----------------------------------------------
 * g_pts[B,K,s,D] = dL/d minimal  ->  g_sel[B,K,s] = sum_c matches[n,c] g_pts[...,c] and (nullable)
 * grad_matches[B,N,D] += g_pts (atomic).                                                      */
int drb_gather_backward(const float* matches, const int32_t* idx, const float* g_pts,
                        int B, int K, int N, int s, int D, float* g_sel, float* grad_matches, void* stream);

/* ---- non-minimal refit / local optimisation on a set of correspondences --------
 * Replaces, for n > sample_size selected points, EssentialMatrixEstimatorNister.estimate_model without
 * pymagsac (essential_matrix_estimator_nister.py:51-65 -> :69-430: the four smallest right singular vectors
 * of A^T A, then the five-point polynomial system) and FundamentalMatrixEstimatorNew.normalize +
 * estimate_non_minimal_model (fundamental_matrix_estimator.py:177-260), as called by the final refit
 * (ransac.py:148-165) and by localOptimization (ransac.py:217-257, lo = 1, 2).
 * matches[B,N,4]; mask[B,N] (nullable = all points; nonzero = selected); weights[B,N] (nullable; rows are
 * scaled by the weight as the reference does).  Accumulation and solve in double.
 * drb_refit_e5: models[B,10,9] (slots >= nsol[b] = identity); drb_refit_f8: models[B,1,9], nsol[b] in {0,1}. */
int drb_refit_e5(const float* matches, const uint8_t* mask, const float* weights, int B, int N,
                 float* models, int32_t* nsol, void* stream);
int drb_refit_f8(const float* matches, const uint8_t* mask, const float* weights, int B, int N,
                 float* models, int32_t* nsol, void* stream);

/* ---- the chunked test-mode loop with adaptive termination, replayed on the device -
 * Replaces the bookkeeping of `while iterations < max_iters` (ransac.py:55-144): per-chunk arg-max (:114),
 * keep-if-strictly-better (:116), adaptive_iteration_number from the winner's inlier count (:134-142, :202-215).
 * Call after drb_score_msac with `scores` requested, for all C = ceil(max_iterations / rbs) chunks at once:
 * scores[B,M] with optional count[B] / ids[B,M] exactly as handed to drb_score_msac; models_dense[B,Md,9];
 * a chunk is `span` consecutive dense model ids (rbs * slots per sample).  chunk_best[B,C] and chunk_ninl[B,C]
 * are caller-provided scratch (left filled: per-chunk winner key and its inlier count).
 * Out: best_packed[B] (feed to drb_best_finalize) and iterations[B] = what the reference's loop would return. */
int drb_adaptive_select(const float* matches, const float* models_dense, const float* scores,
                        const int32_t* count, const int32_t* ids, const float* thr, int B, int M, int Md,
                        int N, int span, int rbs, int max_iterations, int sample_size, double confidence,
                        double eps, unsigned long long* chunk_best, int32_t* chunk_ninl,
                        unsigned long long* best_packed, int32_t* iterations, void* stream);

/* ---- pose from an essential matrix ------------------------------------------
 * Replaces cv_utils.recoverPose / decompose_E / cheirality_check (cv_utils.py:48-116, :179-189; host loop
 * around cv2.triangulatePoints), cv2.recoverPose as MatchLoss uses it for the ground-truth inlier mask
 * (loss.py:126-135) and evaluate_R_t_tensor behind eval_essential_matrix (cv_utils.py:361-378, :503-525).
 * E[B,M,9] (x2^T E x1 = 0, any scale), matches[B,N,4] in normalised camera coordinates, npts[B] nullable
 * (only the first npts[b] correspondences vote), dist = OpenCV's distanceThresh (the reference: 50).
 * Out: R[B,M,9], t[B,M,3] (unit) of the pose with the most correspondences in front of both cameras,
 * mask[B,M,N] (nullable) = those correspondences, ngood[B,M]; with R_gt[B,9], t_gt[B,3] also
 * err[B,M,2] = (rotation, translation) angular error in degrees.  counts[B,M,4] is caller-provided scratch
 * (left holding the votes of the four candidate poses).  Arithmetic in double.                      */
int drb_recover_pose(const float* E, const float* matches, const int32_t* npts, const float* R_gt,
                     const float* t_gt, int B, int M, int N, float dist, int32_t* counts, float* R, float* t,
                     uint8_t* mask, int32_t* ngood, float* err, void* stream);
/* One term of PoseLoss.forward_average per (pair, model) (loss.py:57-63 with its default svd=False): Horn's
 * closed-form decomposition (cv_utils.py:118-165), the same cheirality vote, err[B,M,2] in degrees and
 * (nullable) grad[B,M,9] = d((err_R + err_t)/2)/dE -- what autograd returns through the reference's chain,
 * including its constant [b]x (cv_utils.py:146-150).  Non-finite derivatives (arccos at +-1) are zeroed. */
int drb_pose_loss(const float* E, const float* matches, const int32_t* npts, const float* R_gt,
                  const float* t_gt, int B, int M, int N, float dist, int32_t* counts, float* err, float* grad,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRB_H_ */

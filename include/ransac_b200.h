/*
 * ransac_b200 — C ABI of the B200 (sm_100a) RANSAC camera-location hot path.
 *
 * Drop-in boundary.  The reference (Mendel0408/Code-Reproduction-RANSAC) has no FFI of its own: its hot
 * path is two calls into the cv2 Python binding of OpenCV calib3d,
 *
 *     M, mask = cv2.findHomography(pos2[good == 1], pixels[good == 1], cv2.RANSAC, ransacbound)
 *                                   /root/reference/main_v1.py:312   (process.py:200, testpro.py:350,
 *                                                                     test_pro.py:351, test02.py:263)
 *     ok, rvec, tvec, inliers = cv2.solvePnPRansac(pos3d, pixels, K, dist, iterationsCount=5000,
 *                                   reprojectionError=30.0, confidence=0.99)
 *                                   /root/reference/main_v1.py:497-502 (testpro.py:536, test_pro.py:515,
 *                                                                     testpro-K.py:72-75)
 *
 * so the entry points below are what a binding for those two calls (and for the loop around the first one,
 * find_homographies, main_v1.py:254-297) would bind.  Plain pointers and sizes only; no torch, numpy or
 * OpenCV types.  INTEGRATION.md shows the ctypes stub that replaces the cv2 attribute lookups.
 *
 * Conventions
 *   - Every function returns B2R_OK (0), B2R_NO_MODEL (1: the call succeeded and there is no model — cv2
 *     returns None / False), or a negative error (bad argument, CUDA failure); b2r_last_error() then
 *     returns a message.  There is NO CPU fallback: without a usable CUDA device every compute entry
 *     point fails with B2R_ERR_CUDA.
 *   - "host" pointers are ordinary host memory owned by the caller, never written unless documented as an
 *     output; "dev" pointers are device memory on the context's GPU.  Calls are synchronous: outputs are
 *     complete on return.  A context is not thread-safe; use one per thread / per GPU.
 *   - Point arrays are row-major (n, 2) / (n, 3) float64, exactly what the reference passes to cv2.
 */
#ifndef RANSAC_B200_H
#define RANSAC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2R_OK 0
#define B2R_NO_MODEL 1
#define B2R_ERR_ARG (-1)
#define B2R_ERR_CUDA (-2)
#define B2R_ERR_INTERNAL (-3)

/* sampler */
#define B2R_SAMPLER_CV_REPLAY 0 /* replay OpenCV's cv::RNG sample stream + adaptive termination (parity mode) */
#define B2R_SAMPLER_PHILOX 1    /* Philox4x32-10 counter-based sampler, fixed number of hypotheses           */
/* arithmetic of the hypotheses x points scoring kernel */
#define B2R_ARITH_EXACT 0 /* the reference's un-fused fp32 sequence: inlier sets bit-exact with cv2 */
#define B2R_ARITH_FAST 1  /* FMA-contracted, division-free form of the same inequality               */
#define B2R_ARITH_EXACT_UNFILTERED 2 /* the un-fused sequence for EVERY evaluation.  B2R_ARITH_EXACT returns the same counts:
                                        it takes the sign of the division-free margin where a proved error bound allows and
                                        runs this sequence elsewhere; this value exists so that tests can hold the two equal */
/* minimal solver */
#define B2R_SOLVER_EXACT 0 /* OpenCV's normalised DLT + Jacobi eigen-solver, fp64, bit-identical models */
#define B2R_SOLVER_FAST 1  /* closed-form 4-point solve in registers (fp64), ~1e-12 relative agreement   */
#define B2R_SOLVER_EXACT_WARP 2 /* b2r_solve_h4 only: the exact solver's one-warp-per-solve kernel (same bits; the
                                   pipeline picks it by itself for small batches, where latency matters)  */
/* which mask cv2.findHomography returns */
#define B2R_REFINE_NONE 0
#define B2R_REFINE_CV 1        /* as cv2; with the exact solver and n <= 128 the sums run in OpenCV's order and every LM step is
                                  solved by eigen-decomposition: the refined H is bit-identical to the CPU restatement   */
#define B2R_REFINE_PARALLEL 2  /* same algorithm, parallel reductions and Cholesky for damped steps at every size (last-bit
                                  differences; 0.2 ms less latency on a 12-point problem)                                */
#define B2R_MASK_CV413 0  /* OpenCV 4.13: mask re-derived from the refined H   */
#define B2R_MASK_LEGACY 1 /* older OpenCV (the reference's debug.log): RANSAC-stage mask */

typedef struct b2r_ctx b2r_ctx;

typedef struct {
    double thr;          /* ransacReprojThreshold, pixels (main_v1.py:862 passes 75.0, process.py:374 120.0) */
    int32_t max_iters;   /* cv2 default 2000; with B2R_SAMPLER_PHILOX: number of hypotheses scored           */
    double confidence;   /* cv2 default 0.995                                                                */
    int32_t sampler;     /* B2R_SAMPLER_*                                                                    */
    uint64_t seed;       /* Philox key (ignored for CV_REPLAY, whose seed OpenCV fixes at 2^64-1)            */
    int32_t arith;       /* B2R_ARITH_*                                                                      */
    int32_t mask_semantics; /* B2R_MASK_*                                                                    */
    int32_t refine;      /* B2R_REFINE_* : 1 refit on inliers + 10 Levenberg-Marquardt iterations, as cv2 does; 0 skip */
    /* hypothesis-id shard of this rank (PHILOX only): ids [hyp_begin, hyp_begin + max_iters) are scored.  Ids must lie
     * in [0, 2^32) (the arg-max key carries the id in 32 bits): hyp_begin < 0 or hyp_begin + max_iters > 2^32 -> B2R_ERR_ARG */
    int64_t hyp_begin;
    int32_t solver;      /* B2R_SOLVER_*                                                                     */
    int32_t reserved;
} b2r_h_params;

typedef struct {
    int32_t status;       /* B2R_OK / B2R_NO_MODEL                                          */
    int32_t iters_run;    /* RANSAC iterations executed (CV_REPLAY) / hypotheses scored      */
    int32_t best_iter;    /* 0-based iteration (CV_REPLAY) or hypothesis id - hyp_begin; -1 after a sharded
                             finish whose winner lies in another rank's shard                 */
    int32_t best_count;   /* RANSAC-stage inlier count of the winning hypothesis             */
    int32_t sample[4];    /* its minimal sample (point indices)                              */
    int32_t n_inliers;    /* number of ones in the returned mask                             */
    int32_t lm_iters;     /* LM iterations executed by the refinement                        */
    int32_t reserved[2];  /* [0]: PHILOX sampler: global hypothesis id of the winner, low 31 bits (the same on
                             every rank of a sharded run)                                     */
} b2r_h_info;

/* ---- life cycle -------------------------------------------------------------------------------------- */
int b2r_version(void);
const char* b2r_last_error(void);
int b2r_device_count(void);
/* Creates a context on CUDA device `device` (own stream, growable workspaces).  NULL + last_error on failure. */
b2r_ctx* b2r_ctx_create(int device);
void b2r_ctx_destroy(b2r_ctx* ctx);
/* The context's cudaStream_t, as an opaque pointer (for callers that time with their own CUDA events). */
void* b2r_ctx_stream(b2r_ctx* ctx);
int b2r_ctx_synchronize(b2r_ctx* ctx);
void b2r_default_h_params(b2r_h_params* p);

/* ---- cv2.findHomography(src, dst, cv2.RANSAC, thr)  —  main_v1.py:312 ------------------------------- */
/* src_host, dst_host: (n,2) float64.  H_out: 9 doubles (row-major; all zero when there is no model).
 * mask_out: n bytes.  info_out may be NULL.  n < 4 -> B2R_ERR_ARG (cv2 raises).  End to end: uploads the
 * points, runs sampler -> solver -> scoring -> selection -> refinement on the GPU, downloads H and the mask. */
int b2r_find_homography(b2r_ctx* ctx, const double* src_host, const double* dst_host, int32_t n,
                        const b2r_h_params* params, double* H_out, uint8_t* mask_out, b2r_h_info* info_out);

/* ---- the loop of find_homographies (main_v1.py:254-297): Q independent problems of n points each ---- */
/* src_host: (Q,n,2) float64.  dst_host: (n,2) if dst_shared != 0 (the sweep: same pixels for every candidate
 * camera), else (Q,n,2).  H_out: (Q,9).  mask_out: (Q,n).  info_out: Q entries or NULL.  Returns B2R_OK even
 * when some problems have no model (see info_out[q].status; their H is zero and mask zero). */
int b2r_find_homography_batch(b2r_ctx* ctx, const double* src_host, const double* dst_host, int32_t dst_shared,
                              int32_t Q, int32_t n, const b2r_h_params* params, double* H_out, uint8_t* mask_out,
                              b2r_h_info* info_out);

/* ---- find_homographies + arg-min (main_v1.py:254-297, :863-866), fused on the device ---------------------------------- */
/* pos3d_host (n,3), pixels_host (n,2): the landmarks with a non-zero pixel (main_v1.py:308 drops the others), cams_host (Q,3):
 * candidate camera positions; all float64.  Per candidate the device computes the reference's projection pos2
 * (main_v1.py:304-311), runs cv2.findHomography(pos2, pixels, RANSAC, params->thr), and scores it (main_v1.py:314, 327-348,
 * 419).  scores_out (Q,2) = err1, err2 (0,0 for a candidate without model — the reference fails there, see info_out).
 * M_out (Q,9) = inv(H), H_out (Q,9), mask_out (Q,n), info_out (Q) may be NULL.  *best_out = argmin(err2 | 0 -> 1e6). */
int b2r_camera_sweep(b2r_ctx* ctx, const double* pos3d_host, const double* pixels_host, int32_t n, const double* cams_host,
                     int32_t Q, const b2r_h_params* params, double* scores_out, double* M_out, double* H_out,
                     uint8_t* mask_out, b2r_h_info* info_out, int32_t* best_out);

/* ---- device-resident form: the points already live in HBM -------------------------------------------- */
/* Uploads and packs Q problems once (fp32 quantisation as cv2 does); returns a handle. */
typedef struct b2r_h_problem b2r_h_problem;
b2r_h_problem* b2r_h_problem_upload(b2r_ctx* ctx, const double* src_host, const double* dst_host, int32_t dst_shared,
                                    int32_t Q, int32_t n);
/* Replaces the points of an existing handle (same or different Q, n), reusing its device buffers. */
int b2r_h_problem_reupload(b2r_ctx* ctx, b2r_h_problem* prob, const double* src_host, const double* dst_host,
                           int32_t dst_shared, int32_t Q, int32_t n);
void b2r_h_problem_free(b2r_ctx* ctx, b2r_h_problem* prob);
/* Runs the whole RANSAC on a resident problem set; results stay on the device until fetched. */
int b2r_h_problem_run(b2r_ctx* ctx, b2r_h_problem* prob, const b2r_h_params* params);
int b2r_h_problem_fetch(b2r_ctx* ctx, b2r_h_problem* prob, double* H_out, uint8_t* mask_out, b2r_h_info* info_out);
/* Multi-GPU (PHILOX): stage 1 scores this rank's hypothesis shard and returns, per problem, the packed key
 * (count << 32) | (0xFFFFFFFF - global hypothesis id) — the value ranks reduce with MAX (one 8-byte NCCL
 * all-reduce per problem).  Stage 2 re-derives the winning hypothesis from its global id and finishes
 * (mask, refit, LM) identically on every rank. */
int b2r_h_problem_score_shard(b2r_ctx* ctx, b2r_h_problem* prob, const b2r_h_params* params, uint64_t* keys_out);
int b2r_h_problem_finish(b2r_ctx* ctx, b2r_h_problem* prob, const b2r_h_params* params, const uint64_t* keys);
/* The same two stages with the keys staying in DEVICE memory (Q uint64 on the context's GPU), asynchronous on the context's
 * stream: a collective library can reduce them in place (NCCL all-reduce enqueued on that stream) without a host round trip. */
int b2r_h_problem_score_shard_dev(b2r_ctx* ctx, b2r_h_problem* prob, const b2r_h_params* params, uint64_t* keys_dev_out);
int b2r_h_problem_finish_dev(b2r_ctx* ctx, b2r_h_problem* prob, const b2r_h_params* params, const uint64_t* keys_dev);
/* Diagnostics / parity tests: copies hypothesis slots [first, first + count) of problem q as the LAST run left them on the
 * device — minimal samples (count,4) int32, fp32 models (count,8) as K3 scored them (NaN: rejected), inlier counts (count)
 * int32.  Any output may be NULL.  Slot i holds hypothesis id hyp_begin + i (PHILOX) / iteration i (CV_REPLAY). */
int b2r_h_problem_peek_hyps(b2r_ctx* ctx, b2r_h_problem* prob, int32_t q, int32_t first, int32_t count, int32_t* samples_out,
                            float* models8_out, int32_t* counts_out);
/* Device time (ms, CUDA events on the context stream) of the stages of the last run:
 * [0] sample+solve, [1] scoring kernel, [2] select, [3] finalize (mask/refit/LM), [4] total. */
int b2r_h_problem_stage_ms(b2r_ctx* ctx, b2r_h_problem* prob, float ms_out[5]);
/* Number of kernels the last run launched (for bench.py's gpu_launches). */
int b2r_ctx_launch_count(b2r_ctx* ctx);

/* ---- building blocks (used by the parity tests; each is one kernel of the pipeline) ------------------ */
/* K3: inlier counts of n_models fp32 models (n_models,8) over n points.  src/dst (n,2) float32 (already
 * quantised).  counts_out: n_models int32. */
int b2r_score_h(b2r_ctx* ctx, const float* models_host, int32_t n_models, const float* src_host,
                const float* dst_host, int32_t n, float thr_sq, int32_t arith, int32_t* counts_out);
/* K2: 4-point solves.  idx_host (n_samples,4) int32 into the n points.  H_out (n_samples,9) fp64,
 * ok_out (n_samples) 1/0 (0: degenerate normalisation, or checkSubset rejected the sample). */
int b2r_solve_h4(b2r_ctx* ctx, const float* src_host, const float* dst_host, int32_t n, const int32_t* idx_host,
                 int32_t n_samples, int32_t solver, double* H_out, uint8_t* ok_out, uint8_t* subset_ok_out);
/* K1: the first n_iters minimal samples OpenCV's RANSAC would draw on these points (replay sampler).
 * idx_out (n_iters,4).  *n_generated_out < n_iters when getSubset gave up. */
int b2r_sample_cv(b2r_ctx* ctx, const float* src_host, const float* dst_host, int32_t n, int32_t n_iters,
                  int32_t* idx_out, int32_t* n_generated_out);
/* K1: Philox samples for hypothesis ids [hyp_begin, hyp_begin + n_hyp) of problem q.  idx_out (n_hyp,4);
 * an all -1 row marks a hypothesis whose 16 attempts were all rejected by checkSubset. */
int b2r_sample_philox(b2r_ctx* ctx, const float* src_host, const float* dst_host, int32_t n, uint64_t seed,
                      int32_t q, int64_t hyp_begin, int32_t n_hyp, int32_t* idx_out);
/* K4 refinement only: refit on the masked points + LM(10).  H_io: in = RANSAC model, out = refined. */
int b2r_refine_h(b2r_ctx* ctx, const float* src_host, const float* dst_host, int32_t n, const uint8_t* mask_host,
                 double* H_io, int32_t* lm_iters_out);
/* The symmetric eigen-solver every exact-mode stage rests on (cv::eigen's Jacobi as calib3d calls it: runKernel's
 * L^T L, cv::solve(DECOMP_EIG) of the LM steps), n x n with 2 <= n <= 9.  A_host (n_mat, n*n) row-major symmetric;
 * W_out (n_mat, n) eigenvalues descending, V_out (n_mat, n*n) eigenvectors in rows.
 * form 0: one thread per matrix in registers/local memory (reference form); 1: one warp per matrix (finalize kernels);
 * 2: one thread per matrix on packed strided shared memory (K2 exact solver; n = 9 only).  All bit-identical. */
int b2r_jacobi_eig(b2r_ctx* ctx, const double* A_host, int32_t n_mat, int32_t n, int32_t form, double* W_out,
                   double* V_out);
/* Sweeps all 2^32 bit patterns of x and compares the scoring kernel's reciprocal with IEEE 1.0f/x inside the
 * kernel's fast range; *mismatches_out must be 0. */
int b2r_selftest_rcp(b2r_ctx* ctx, uint64_t* mismatches_out, uint64_t* tested_out);
/* Register-resident FFMA peak of this GPU, in fp32 FMA lane-operations per second (x2 = FLOP/s). */
int b2r_probe_fp32_peak(b2r_ctx* ctx, double* fma_per_s_out, double* ffma2_per_s_out);

/* ==== cv2.solvePnPRansac / cv2.solvePnPRefineLM  —  main_v1.py:497-502, :508-509; testpro-K.py:72-75, :122-125 ==== */
typedef struct {
    double thr;          /* reprojectionError, pixels (the reference passes 30.0; cv2 default 8.0)                    */
    int32_t max_iters;   /* iterationsCount (reference 5000; cv2 default 100); PHILOX: hypotheses scored              */
    double confidence;   /* reference 0.99 (= cv2 default)                                                            */
    int32_t sampler;     /* B2R_SAMPLER_*                                                                             */
    uint64_t seed;       /* Philox key (ignored for CV_REPLAY)                                                        */
    int32_t arith;       /* B2R_ARITH_EXACT: cv::projectPoints' fp64 projection rounded to fp32 + un-fused fp32 error,
                            bit-exact inlier sets; B2R_ARITH_FAST: all-fp32 FMA on re-centred points                 */
    int32_t refine;      /* 1: pose = LM (CvLevMarq, as solvePnP(ITERATIVE, useExtrinsicGuess)) on the inliers seeded
                            with the best RANSAC model, as cv2 does; 0: return the best RANSAC model                 */
    int64_t hyp_begin;   /* hypothesis-id shard of this rank (PHILOX only)                                            */
    int32_t solver;      /* B2R_SOLVER_EXACT: OpenCV's EPnP on the five points, restated operation for operation (the
                            hypotheses cv2 scores); B2R_SOLVER_FAST: depth-parametrised 5-point solver + 3 Gauss-Newton
                            steps, in registers (throughput path; its hypotheses are not OpenCV's)                      */
    int32_t reserved;
} b2r_p_params;

typedef struct {
    int32_t status;       /* B2R_OK / B2R_NO_MODEL (cv2: retval False)                         */
    int32_t iters_run;    /* RANSAC iterations executed (CV_REPLAY) / hypotheses scored         */
    int32_t best_iter;    /* 0-based iteration, or hypothesis id - hyp_begin                    */
    int32_t best_count;   /* RANSAC-stage inlier count of the winning hypothesis                */
    int32_t sample[5];    /* its minimal sample                                                 */
    int32_t n_inliers;    /* length of the returned inlier list                                 */
    int32_t lm_iters;     /* LM iterations of the pose refinement                               */
    int32_t reserved;
    double ransac_rvec[3], ransac_tvec[3]; /* the best minimal model (before refinement)        */
    double mean_inlier_err; /* mean reprojection error (px) of the inliers under the returned pose, evaluated on the
                               caller's un-quantised points: compute_reprojection_error, testpro-K.py:32-36, :80-82 */
    double sum_sq_err;      /* sum of squared inlier residuals under the returned pose          */
} b2r_p_info;

void b2r_default_p_params(b2r_p_params* p);

/* obj_host (n,3) float64, img_host (n,2) float64, K (3,3) row-major float64; distortion is zero on the reference's path
 * (it passes np.zeros((4,1))).  rvec_out/tvec_out: 3 doubles each.  inliers_out: capacity n int32 (ascending indices, as
 * cv2 returns them), *n_inliers_out their number.  n < 4 -> B2R_ERR_ARG (cv2 raises); n == 4 (OpenCV's P3P branch) is
 * not on the reference's path and returns B2R_ERR_ARG.  Returns B2R_NO_MODEL when cv2 would return retval False. */
int b2r_solve_pnp_ransac(b2r_ctx* ctx, const double* obj_host, const double* img_host, int32_t n, const double* K,
                         const b2r_p_params* params, double* rvec_out, double* tvec_out, int32_t* inliers_out,
                         int32_t* n_inliers_out, b2r_p_info* info_out);
/* The loop of estimate_camera_orientation (testpro-K.py:58-97): Q camera matrices K (Q,9).  pts_shared != 0: all Q problems
 * use the same points obj (n,3) / img (n,2) (the intrinsics grid); else obj (Q,n,3), img (Q,n,2).  Outputs are (Q,3),
 * (Q,3), (Q,n), (Q), Q entries.  Returns B2R_OK even when some problems have no model (info_out[q].status). */
int b2r_solve_pnp_ransac_batch(b2r_ctx* ctx, const double* obj_host, const double* img_host, int32_t pts_shared, int32_t Q,
                               int32_t n, const double* K, const b2r_p_params* params, double* rvec_out, double* tvec_out,
                               int32_t* inliers_out, int32_t* n_inliers_out, b2r_p_info* info_out);
/* cv2.solvePnPRefineLM(obj, img, K, 0, rvec, tvec): classic LMSolver, max_iters (<= 0: cv2's 20), eps FLT_EPSILON, on the
 * caller's fp64 points (k >= 3).  rvec_io/tvec_io: in = initial pose, out = refined. */
int b2r_solve_pnp_refine_lm(b2r_ctx* ctx, const double* obj_host, const double* img_host, int32_t k, const double* K,
                            double* rvec_io, double* tvec_io, int32_t max_iters, int32_t* iters_out);

/* device-resident form (same life cycle as b2r_h_problem) */
typedef struct b2r_p_problem b2r_p_problem;
b2r_p_problem* b2r_p_problem_upload(b2r_ctx* ctx, const double* obj_host, const double* img_host, int32_t pts_shared, int32_t Q,
                                    int32_t n, const double* K);
int b2r_p_problem_reupload(b2r_ctx* ctx, b2r_p_problem* prob, const double* obj_host, const double* img_host,
                           int32_t pts_shared, int32_t Q, int32_t n, const double* K);
void b2r_p_problem_free(b2r_ctx* ctx, b2r_p_problem* prob);
int b2r_p_problem_run(b2r_ctx* ctx, b2r_p_problem* prob, const b2r_p_params* params);
int b2r_p_problem_fetch(b2r_ctx* ctx, b2r_p_problem* prob, double* rvec_out, double* tvec_out, int32_t* inliers_out,
                        int32_t* n_inliers_out, b2r_p_info* info_out);
int b2r_p_problem_score_shard(b2r_ctx* ctx, b2r_p_problem* prob, const b2r_p_params* params, uint64_t* keys_out);
int b2r_p_problem_finish(b2r_ctx* ctx, b2r_p_problem* prob, const b2r_p_params* params, const uint64_t* keys);
/* device-resident keys, asynchronous on the context's stream (see b2r_h_problem_score_shard_dev) */
int b2r_p_problem_score_shard_dev(b2r_ctx* ctx, b2r_p_problem* prob, const b2r_p_params* params, uint64_t* keys_dev_out);
int b2r_p_problem_finish_dev(b2r_ctx* ctx, b2r_p_problem* prob, const b2r_p_params* params, const uint64_t* keys_dev);
int b2r_p_problem_stage_ms(b2r_ctx* ctx, b2r_p_problem* prob, float ms_out[5]);

/* building blocks of the PnP path (parity tests) */
/* K3: inlier counts of n_models poses, models_Rt (n_models,12) = R row-major | t, over n points (fp64 arrays, quantised to
 * fp32 inside as cv2 does).  thr_sq = (float)(thr*thr). */
int b2r_score_p(b2r_ctx* ctx, const double* models_Rt, int32_t n_models, const double* obj_host, const double* img_host,
                int32_t n, const double* K, float thr_sq, int32_t arith, int32_t* counts_out);
/* K2: EPnP minimal models of 5-point samples idx (n_samples,5).  Outputs (any may be NULL): rvec (n_samples,3), tvec
 * (n_samples,3), R = Rodrigues(rvec) (n_samples,9), ok (n_samples).  solver: B2R_SOLVER_EXACT / B2R_SOLVER_FAST. */
int b2r_pnp_minimal_models(b2r_ctx* ctx, const double* obj_host, const double* img_host, int32_t n, const double* K,
                           const int32_t* idx_host, int32_t n_samples, int32_t solver, double* rvec_out, double* tvec_out,
                           double* R_out, uint8_t* ok_out);
/* K1: the first n_iters 5-point subsets OpenCV's RANSAC draws for n points.  idx_out (n_iters,5). */
int b2r_sample_cv_p(b2r_ctx* ctx, int32_t n, int32_t n_iters, int32_t* idx_out);

/* ---- downstream georeferencing: the DEM ray-march -------------------------------------------
 * Replaces the per-vertex Python loops of the reference:
 *   ray_intersect_dem        /root/reference/main_v1.py:635-658  (10 000 x [pyproj UTM->WGS84, DEM bilinear lookup, test])
 *   pixel_to_geo             /root/reference/main_v1.py:661-684  (weights :577-596, weighted factors :627-632, pixel_to_ray :547-574)
 *   convert_boundary_to_geo  /root/reference/main_v1.py:765-785  (every polygon vertex of the annotation JSON)
 * A DEM is the grid of scipy's RegularGridInterpolator((dem_y, dem_x), dem_array) (main_v1.py:454): strictly ASCENDING
 * axes grid_y (latitude, ny) and grid_x (longitude, nx) in degrees — scipy flips descending axes, pass them flipped —
 * and values (ny, nx) float64, row-major.  utm_series16 = {k0 A, lon0 [rad], FE, FN, beta[6], delta[6]}: the constants of
 * the inverse transverse Mercator series (host layer: geo.utm_series_constants(), EPSG:32650).
 * Outputs per ray: status 0 = hit (geo_out = E, N, height of the first step s >= min_steps with height <= DEM;
 * the reference's rule with min_steps = 150), 1 = no intersection within int(max_search_dist / step) steps, 2 = a step
 * left the DEM (the interpolator raises; the reference returns None in both cases); hit_step_out = that step. */
typedef struct b2r_dem b2r_dem;
b2r_dem* b2r_dem_upload(b2r_ctx* ctx, const double* grid_y, int32_t ny, const double* grid_x, int32_t nx, const double* values);
void b2r_dem_free(b2r_ctx* ctx, b2r_dem* dem);
/* ray_intersect_dem for m rays: origins (m,3) or one shared origin (origin_shared != 0), unit directions (m,3), UTM metres. */
int b2r_ray_march_dem(b2r_ctx* ctx, const b2r_dem* dem, const double* origins, int32_t origin_shared, const double* dirs, int32_t m,
                      const double* utm_series16, double max_search_dist, double step, int32_t min_steps, double* geo_out,
                      int32_t* hit_step_out, int32_t* status_out);
/* pixel_to_geo for m pixels (m,2): Kinv = inverse camera matrix (3,3), R = world->camera rotation (3,3), ray_origin (3),
 * n_ctrl control points with their pixels (n_ctrl,2) and optimisation factors (n_ctrl,3).  dirs_out (m,3), optional: the
 * corrected unit ray directions. */
int b2r_pixels_to_geo(b2r_ctx* ctx, const b2r_dem* dem, const double* pixels, int32_t m, const double* Kinv, const double* R,
                      const double* ray_origin, const double* ctrl_pixels, const double* ctrl_factors, int32_t n_ctrl,
                      const double* utm_series16, double max_search_dist, double step, int32_t min_steps, double* geo_out,
                      int32_t* hit_step_out, int32_t* status_out, double* dirs_out);

#ifdef __cplusplus
}
#endif
#endif /* RANSAC_B200_H */

/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product path
 * (code-reproduction-ransac_b200/); only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may use it, and there only as the checker.
 *
 * CPU restatement, in plain C, of the arithmetic that the reference's hot path executes inside
 *     cv2.findHomography(pos2, pixels, cv2.RANSAC, thr)        /root/reference/main_v1.py:312
 *                                                              (process.py:200, testpro.py:350,
 *                                                               test_pro.py:351, test02.py:263)
 *     cv2.solvePnPRansac(pos3d, pixels, K, 0, 5000, 30.0, .99) /root/reference/main_v1.py:497-502
 *                                                              (testpro.py:536, test_pro.py:515,
 *                                                               testpro-K.py:72-75)
 * The arithmetic lives in a third-party dependency that is absent from /root/reference and that
 * the reference does not pin: OpenCV calib3d, here `opencv-python-headless 4.13.0.92`
 * (cv2.__version__ == "4.13.0").  Each function below restates the published OpenCV 4.x algorithm
 * (section numbers A.1..A.8 refer to SURVEY.md Appendix A, which lists every probe of the binary).
 *
 * PINNING.  The reference has no tests, golden vectors or fixtures for this path ("parity
 * unpinned" by the reference itself).  This oracle is pinned instead against (i) the cv2 4.13.0
 * binary run in the build container (tests/golden/make_golden.py writes the vectors,
 * tests/test_oracle_golden.py replays them without cv2), and (ii) the reference's own recorded run
 * /root/reference/debug.log (24 candidate-location blocks: refined matrix and RANSAC-stage mask).
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: the reference wheel is built without FMA, and
 * the fp32 scoring formula is only bit-reproducible un-fused).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------- */
/* A.2  cv::RNG — multiply-with-carry, re-seeded with 2^64-1 inside every RANSAC run            */
/* ------------------------------------------------------------------------------------------- */
typedef struct { uint64_t state; } orc_rng;

ORC_API void orc_rng_seed(orc_rng* r, uint64_t s) { r->state = s ? s : 0xffffffffull; }

ORC_API uint32_t orc_rng_next(orc_rng* r) {
    r->state = (uint64_t)(uint32_t)r->state * 4164903690u + (uint32_t)(r->state >> 32);
    return (uint32_t)r->state;
}

static int rng_uniform(orc_rng* r, int a, int b) { return a == b ? a : (int)(orc_rng_next(r) % (uint32_t)(b - a) + a); }

/* ------------------------------------------------------------------------------------------- */
/* Jacobi eigen-solver for a symmetric n x n fp64 matrix (cv::eigen), A.4 step 5                */
/* ------------------------------------------------------------------------------------------- */
static double cv_hypot(double a, double b) {
    a = fabs(a);
    b = fabs(b);
    if (a > b) { b /= a; return a * sqrt(1 + b * b); }
    if (b > 0) { a /= b; return b * sqrt(1 + a * a); }
    return 0;
}

#define MAXN 12
/* A: n*n row-major (destroyed), W: n eigenvalues (descending), V: n*n, rows are eigenvectors */
ORC_API void orc_jacobi(double* A, int n, double* W, double* V) {
    int indR[MAXN], indC[MAXN];
    int i, k, l, m, it;
    const double eps = DBL_EPSILON;
    for (i = 0; i < n * n; i++) V[i] = 0;
    for (i = 0; i < n; i++) V[i * n + i] = 1;
    for (k = 0; k < n; k++) {
        W[k] = A[k * n + k];
        if (k < n - 1) {
            double mv;
            for (m = k + 1, mv = fabs(A[k * n + m]), i = k + 2; i < n; i++) {
                double val = fabs(A[k * n + i]);
                if (mv < val) mv = val, m = i;
            }
            indR[k] = m;
        }
        if (k > 0) {
            double mv;
            for (m = 0, mv = fabs(A[k]), i = 1; i < k; i++) {
                double val = fabs(A[i * n + k]);
                if (mv < val) mv = val, m = i;
            }
            indC[k] = m;
        }
    }
    if (n > 1)
        for (it = 0; it < n * n * 30; it++) {
            double mv, p, y, t, s, c;
            for (k = 0, mv = fabs(A[indR[0]]), i = 1; i < n - 1; i++) {
                double val = fabs(A[i * n + indR[i]]);
                if (mv < val) mv = val, k = i;
            }
            l = indR[k];
            for (i = 1; i < n; i++) {
                double val = fabs(A[indC[i] * n + i]);
                if (mv < val) mv = val, k = indC[i], l = i;
            }
            p = A[k * n + l];
            if (fabs(p) <= eps) break;
            y = (W[l] - W[k]) * 0.5;
            t = fabs(y) + cv_hypot(p, y);
            s = cv_hypot(p, t);
            c = t / s;
            s = p / s;
            t = (p / t) * p;
            if (y < 0) s = -s, t = -t;
            A[k * n + l] = 0;
            W[k] -= t;
            W[l] += t;
#define ROT(v0, v1) { double a0 = (v0), b0 = (v1); (v0) = a0 * c - b0 * s; (v1) = a0 * s + b0 * c; }
            for (i = 0; i < k; i++) ROT(A[i * n + k], A[i * n + l]);
            for (i = k + 1; i < l; i++) ROT(A[k * n + i], A[i * n + l]);
            for (i = l + 1; i < n; i++) ROT(A[k * n + i], A[l * n + i]);
            for (i = 0; i < n; i++) ROT(V[k * n + i], V[l * n + i]);
#undef ROT
            for (int j = 0; j < 2; j++) {
                int idx = j == 0 ? k : l;
                if (idx < n - 1) {
                    for (m = idx + 1, mv = fabs(A[idx * n + m]), i = idx + 2; i < n; i++) {
                        double val = fabs(A[idx * n + i]);
                        if (mv < val) mv = val, m = i;
                    }
                    indR[idx] = m;
                }
                if (idx > 0) {
                    for (m = 0, mv = fabs(A[idx]), i = 1; i < idx; i++) {
                        double val = fabs(A[i * n + idx]);
                        if (mv < val) mv = val, m = i;
                    }
                    indC[idx] = m;
                }
            }
        }
    for (k = 0; k < n - 1; k++) {
        m = k;
        for (i = k + 1; i < n; i++)
            if (W[m] < W[i]) m = i;
        if (k != m) {
            double tmp = W[m]; W[m] = W[k]; W[k] = tmp;
            for (i = 0; i < n; i++) { tmp = V[m * n + i]; V[m * n + i] = V[k * n + i]; V[k * n + i] = tmp; }
        }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* A.4  HomographyEstimatorCallback::runKernel — anisotropic-L1-normalised DLT                  */
/*      src ("M"), dst ("m"): n fp32 (x,y) pairs.  Returns number of models (0 or 1).           */
/* ------------------------------------------------------------------------------------------- */
static void mat3_mul(const double* a, const double* b, double* o) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += a[i * 3 + k] * b[k * 3 + j];
            o[i * 3 + j] = s;
        }
}

ORC_API int orc_h_run_kernel(const float* M, const float* m, int count, double* H) {
    double LtL[81], W[9], V[81];
    double cMx = 0, cMy = 0, cmx = 0, cmy = 0, sMx = 0, sMy = 0, smx = 0, smy = 0;
    int i, j, k;
    for (i = 0; i < count; i++) {
        cmx += m[2 * i]; cmy += m[2 * i + 1];
        cMx += M[2 * i]; cMy += M[2 * i + 1];
    }
    cmx /= count; cmy /= count; cMx /= count; cMy /= count;
    for (i = 0; i < count; i++) {
        smx += fabs(m[2 * i] - cmx); smy += fabs(m[2 * i + 1] - cmy);
        sMx += fabs(M[2 * i] - cMx); sMy += fabs(M[2 * i + 1] - cMy);
    }
    if (fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON)
        return 0;
    smx = count / smx; smy = count / smy; sMx = count / sMx; sMy = count / sMy;
    {
        double invHnorm[9] = {1. / smx, 0, cmx, 0, 1. / smy, cmy, 0, 0, 1};
        double Hnorm2[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
        double Ht[9], H0[9];
        memset(LtL, 0, sizeof(LtL));
        for (i = 0; i < count; i++) {
            double x = (m[2 * i] - cmx) * smx, y = (m[2 * i + 1] - cmy) * smy;
            double X = (M[2 * i] - cMx) * sMx, Y = (M[2 * i + 1] - cMy) * sMy;
            double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
            double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
            for (j = 0; j < 9; j++)
                for (k = j; k < 9; k++) LtL[j * 9 + k] += Lx[j] * Lx[k] + Ly[j] * Ly[k];
        }
        for (j = 0; j < 9; j++)
            for (k = 0; k < j; k++) LtL[j * 9 + k] = LtL[k * 9 + j];
        orc_jacobi(LtL, 9, W, V);
        mat3_mul(invHnorm, V + 72, Ht);
        mat3_mul(Ht, Hnorm2, H0);
        {
            double sc = 1. / H0[8];
            for (i = 0; i < 9; i++) H[i] = H0[i] * sc;
        }
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------- */
/* A.3  checkSubset for the homography callback                                                 */
/* ------------------------------------------------------------------------------------------- */
static int have_collinear(const float* p, int count) {
    int j, k, i = count - 1;
    for (j = 0; j < i; j++) {
        double dx1 = (float)(p[2 * j] - p[2 * i]);
        double dy1 = (float)(p[2 * j + 1] - p[2 * i + 1]);
        for (k = 0; k < j; k++) {
            double dx2 = (float)(p[2 * k] - p[2 * i]);
            double dy2 = (float)(p[2 * k + 1] - p[2 * i + 1]);
            if (fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return 1;
        }
    }
    return 0;
}

static double det3_pts(const float* p, int a, int b, int c) {
    double a00 = p[2 * a], a01 = p[2 * a + 1], a10 = p[2 * b], a11 = p[2 * b + 1], a20 = p[2 * c], a21 = p[2 * c + 1];
    return a00 * (a11 * 1. - a21 * 1.) - a01 * (a10 * 1. - a20 * 1.) + 1. * (a10 * a21 - a20 * a11);
}

ORC_API int orc_h_check_subset(const float* ms1, const float* ms2, int count) {
    if (have_collinear(ms1, count) || have_collinear(ms2, count)) return 0;
    if (count == 4) {
        static const int tt[4][3] = {{0, 1, 2}, {1, 2, 3}, {0, 2, 3}, {0, 1, 3}};
        int negative = 0;
        for (int i = 0; i < 4; i++)
            negative += det3_pts(ms1, tt[i][0], tt[i][1], tt[i][2]) * det3_pts(ms2, tt[i][0], tt[i][1], tt[i][2]) < 0;
        if (negative != 0 && negative != 4) return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------- */
/* A.5  computeError (fp32, un-fused) and findInliers                                           */
/* ------------------------------------------------------------------------------------------- */
ORC_API void orc_h_compute_error(const double* H, const float* M, const float* m, int count, float* err) {
    const float Hf[8] = {(float)H[0], (float)H[1], (float)H[2], (float)H[3], (float)H[4], (float)H[5], (float)H[6], (float)H[7]};
    for (int i = 0; i < count; i++) {
        float ww = 1.f / (Hf[6] * M[2 * i] + Hf[7] * M[2 * i + 1] + 1.f);
        float dx = (Hf[0] * M[2 * i] + Hf[1] * M[2 * i + 1] + Hf[2]) * ww - m[2 * i];
        float dy = (Hf[3] * M[2 * i] + Hf[4] * M[2 * i + 1] + Hf[5]) * ww - m[2 * i + 1];
        err[i] = dx * dx + dy * dy;
    }
}

/* fp32 models in (the layout the CUDA scoring kernel consumes): counts only */
ORC_API void orc_h_count_inliers_f32(const float* models8, int n_models, const float* M, const float* m, int count,
                                     float thr_sq, int32_t* counts) {
    for (int k = 0; k < n_models; k++) {
        const float* Hf = models8 + 8 * (size_t)k;
        int c = 0;
        for (int i = 0; i < count; i++) {
            float ww = 1.f / (Hf[6] * M[2 * i] + Hf[7] * M[2 * i + 1] + 1.f);
            float dx = (Hf[0] * M[2 * i] + Hf[1] * M[2 * i + 1] + Hf[2]) * ww - m[2 * i];
            float dy = (Hf[3] * M[2 * i] + Hf[4] * M[2 * i + 1] + Hf[5]) * ww - m[2 * i + 1];
            float e = dx * dx + dy * dy;
            c += e <= thr_sq;
        }
        counts[k] = c;
    }
}

static int find_inliers(const double* H, const float* M, const float* m, int count, float* err, uint8_t* mask, double thresh) {
    float t = (float)(thresh * thresh);
    int good = 0;
    orc_h_compute_error(H, M, m, count, err);
    for (int i = 0; i < count; i++) good += (mask[i] = (uint8_t)(err[i] <= t));
    return good;
}

/* ------------------------------------------------------------------------------------------- */
/* A.6  RANSACUpdateNumIters                                                                     */
/* ------------------------------------------------------------------------------------------- */
ORC_API int orc_update_num_iters(double p, double ep, int modelPoints, int maxIters) {
    double num, denom;
    p = p < 0 ? 0 : p; p = p > 1 ? 1 : p;
    ep = ep < 0 ? 0 : ep; ep = ep > 1 ? 1 : ep;
    num = 1 - p; if (num < DBL_MIN) num = DBL_MIN;
    denom = 1 - pow(1 - ep, modelPoints);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return denom >= 0 || -num >= maxIters * (-denom) ? maxIters : (int)lrint(num / denom); /* cvRound: half-even */
}

/* ------------------------------------------------------------------------------------------- */
/* A.3 + A.6  the RANSAC stage of findHomography, with an optional trace                         */
/*   trace_idx   [max_iters*4]  sample indices of every executed iteration (or NULL)            */
/*   trace_count [max_iters]    inlier count of every executed iteration, -1 = kernel failed    */
/* returns 1 if a model was found; *iters_run = iterations executed                              */
/* ------------------------------------------------------------------------------------------- */
ORC_API int orc_h_ransac_stage(const float* M, const float* m, int count, double thresh, int maxIters, double confidence,
                               double* bestH, uint8_t* bestMask, int* iters_run, int32_t* trace_idx, int32_t* trace_count,
                               uint64_t* rng_draws) {
    const int modelPoints = 4;
    orc_rng rng;
    int iter, niters = maxIters > 1 ? maxIters : 1, maxGood = 0;
    float* err = (float*)malloc(sizeof(float) * (size_t)(count > 0 ? count : 1));
    uint8_t* mask = (uint8_t*)malloc((size_t)(count > 0 ? count : 1));
    uint64_t draws = 0;
    int result = 0;
    orc_rng_seed(&rng, 0xffffffffffffffffull);
    *iters_run = 0;
    if (count < modelPoints) goto done;
    if (count == modelPoints) {
        if (orc_h_run_kernel(M, m, count, bestH) <= 0) goto done;
        memset(bestMask, 1, (size_t)count);
        result = 1;
        goto done;
    }
    for (iter = 0; iter < niters; iter++) {
        int idx[4], i, attempts, found = 0, nmodels;
        float ms1[8], ms2[8];
        double H[9];
        for (attempts = 0; attempts < 10000; attempts++) {
            for (i = 0; i < modelPoints; i++) {
                int idx_i, dup;
                do {
                    idx_i = rng_uniform(&rng, 0, count);
                    draws++;
                    dup = 0;
                    for (int q = 0; q < i; q++) dup |= idx[q] == idx_i;
                } while (dup);
                idx[i] = idx_i;
                ms1[2 * i] = M[2 * idx_i]; ms1[2 * i + 1] = M[2 * idx_i + 1];
                ms2[2 * i] = m[2 * idx_i]; ms2[2 * i + 1] = m[2 * idx_i + 1];
            }
            if (!orc_h_check_subset(ms1, ms2, modelPoints)) continue;
            found = 1;
            break;
        }
        if (!found) {
            if (iter == 0) goto done;
            break;
        }
        if (trace_idx) for (i = 0; i < 4; i++) trace_idx[4 * iter + i] = idx[i];
        *iters_run = iter + 1;
        nmodels = orc_h_run_kernel(ms1, ms2, modelPoints, H);
        if (nmodels <= 0) {
            if (trace_count) trace_count[iter] = -1;
            continue;
        }
        {
            int good = find_inliers(H, M, m, count, err, mask, thresh);
            if (trace_count) trace_count[iter] = good;
            if (good > (maxGood > modelPoints - 1 ? maxGood : modelPoints - 1)) {
                memcpy(bestMask, mask, (size_t)count);
                memcpy(bestH, H, sizeof(double) * 9);
                maxGood = good;
                niters = orc_update_num_iters(confidence, (double)(count - good) / count, modelPoints, niters);
            }
        }
    }
    result = maxGood > 0;
done:
    if (rng_draws) *rng_draws = draws;
    free(err);
    free(mask);
    return result;
}

/* ------------------------------------------------------------------------------------------- */
/* A.7  refinement: cv::LMSolver (max 10 iterations, eps FLT_EPSILON) on the nine entries of H   */
/* ------------------------------------------------------------------------------------------- */
/* The building blocks below are pinned BIT FOR BIT against the cv2 4.13.0 binary through the functions it exposes
 * (cv2.mulTransposed, cv2.gemm, cv2.norm, cv2.solve / cv2.invert with DECOMP_EIG, cv2.eigen): an LM assembled from those
 * calls returns cv2.findHomography's H bit for bit, and each block here equals its cv2 call on every probe
 * (tests/golden/make_golden_lm.py -> tests/golden/cv2_lm_blocks.json, tests/test_oracle_golden.py).  The summation
 * orders are properties of that binary on an AVX2 host (cv::norm is dispatched; gemm / mulTransposed are not). */

/* cv::gemm's inner product (GEMMSingleMul, d_size.width <= 4): FOUR interleaved partial sums over k, the tail of
 * (n mod 4) terms into the first, combined as ((s0 + s1) + s2) + s3.  Products rounded (no FMA). */
static double acc4_dot(const double* a, int astep, const double* b, int bstep, int n) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int k = 0;
    for (; k <= n - 4; k += 4) {
        s0 += a[k * astep] * b[k * bstep];
        s1 += a[(k + 1) * astep] * b[(k + 1) * bstep];
        s2 += a[(k + 2) * astep] * b[(k + 2) * bstep];
        s3 += a[(k + 3) * astep] * b[(k + 3) * bstep];
    }
    for (; k < n; k++) s0 += a[k * astep] * b[k * bstep];
    return ((s0 + s1) + s2) + s3;
}

/* x = solve(A, b) for symmetric A through its eigen-decomposition (cv::solve DECOMP_EIG): Jacobi, then the SVD
 * back-substitution (SVBkSb): threshold 2 eps * sum(w) (signed sum), singular directions below it dropped,
 * s = (v_i . b) * (1 / w_i)  — a multiplication by the reciprocal, not a division —, x += s v_i. */
static void solve_sym_eig(const double* A, const double* b, int n, double* x) {
    double a[MAXN * MAXN], W[MAXN], V[MAXN * MAXN], thr = 0;
    int i, j;
    memcpy(a, A, sizeof(double) * n * n);
    orc_jacobi(a, n, W, V);
    for (i = 0; i < n; i++) thr += W[i];
    thr *= DBL_EPSILON * 2;
    for (i = 0; i < n; i++) x[i] = 0;
    for (i = 0; i < n; i++) {
        double wi = W[i], s = 0;
        if (fabs(wi) <= thr) continue;
        wi = 1 / wi;
        for (j = 0; j < n; j++) s += V[i * n + j] * b[j];
        s *= wi;
        for (j = 0; j < n; j++) x[j] = x[j] + s * V[i * n + j];
    }
}

/* diagonal of cv::invert(A, DECOMP_EIG): X[j][j] += (v_i[j] * (1 / w_i)) * v_i[j], same threshold. */
static void invert_sym_eig_diag(const double* A, int n, double* diag) {
    double a[MAXN * MAXN], W[MAXN], V[MAXN * MAXN], thr = 0;
    int i, j;
    memcpy(a, A, sizeof(double) * n * n);
    orc_jacobi(a, n, W, V);
    for (i = 0; i < n; i++) thr += W[i];
    thr *= DBL_EPSILON * 2;
    for (j = 0; j < n; j++) diag[j] = 0;
    for (i = 0; i < n; i++) {
        double wi = W[i];
        if (fabs(wi) <= thr) continue;
        wi = 1 / wi;
        for (j = 0; j < n; j++) diag[j] = diag[j] + (V[i * n + j] * wi) * V[i * n + j];
    }
}

/* residuals r[2k], and (optionally) the Jacobian J (2 count x 9, row-major), A = J^T J and v = J^T r, for the NINE
 * parameters h[0..8].  OpenCV 4.13's HomographyRefineCallback refines all nine entries (w = h6 X + h7 Y + h8).
 * A = cv::mulTransposed(J, aTa): every entry ONE running sum over the rows of J in order (x row, then y row, of each
 * point).  v = cv::gemm(J, r, GEMM_1_T): the four-partial-sum inner product above over the 2 count rows. */
static void h_refine_eval(const double* h, const float* M, const float* m, int count, double* r, double* A, double* v) {
    double* J = A ? (double*)malloc(sizeof(double) * 18 * (size_t)(count > 0 ? count : 1)) : NULL;
    if (A) memset(A, 0, sizeof(double) * 81);
    for (int i = 0; i < count; i++) {
        double Mx = M[2 * i], My = M[2 * i + 1];
        double ww = h[6] * Mx + h[7] * My + h[8];
        ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
        double xi = (h[0] * Mx + h[1] * My + h[2]) * ww;
        double yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
        r[2 * i] = xi - m[2 * i];
        r[2 * i + 1] = yi - m[2 * i + 1];
        if (A) {
            double* Jx = J + 18 * (size_t)i;
            double* Jy = Jx + 9;
            Jx[0] = Mx * ww; Jx[1] = My * ww; Jx[2] = ww; Jx[3] = Jx[4] = Jx[5] = 0;
            Jx[6] = -Mx * ww * xi; Jx[7] = -My * ww * xi; Jx[8] = -ww * xi;
            Jy[0] = Jy[1] = Jy[2] = 0; Jy[3] = Mx * ww; Jy[4] = My * ww; Jy[5] = ww;
            Jy[6] = -Mx * ww * yi; Jy[7] = -My * ww * yi; Jy[8] = -ww * yi;
            for (int j = 0; j < 9; j++)
                for (int k = 0; k < 9; k++) {
                    A[j * 9 + k] += Jx[j] * Jx[k];
                    A[j * 9 + k] += Jy[j] * Jy[k];
                }
        }
    }
    if (A) {
        for (int j = 0; j < 9; j++) v[j] = acc4_dot(J + j, 9, r, 1, 2 * count);
        free(J);
    }
}

/* cv::norm(r, NORM_L2SQR) for CV_64F as the binary's AVX2 code path computes it: 16 elements per step into four
 * 4-lane accumulators by FUSED multiply-add, combined ((r0 + r1) + r2) + r3 lane-wise and then (l0 + l1) + (l2 + l3);
 * the remaining elements in blocks of four ROUNDED squares added in order, the last (n mod 4) by fused multiply-add.
 * 20 000 / 20 000 random vectors of 1 ... 129 elements equal cv2.norm bit for bit. */
static double norm_l2sqr(const double* a, int n) {
    double r[4][4], t[4], s = 0;
    int j = 0, q, l;
    for (q = 0; q < 4; q++) for (l = 0; l < 4; l++) r[q][l] = 0;
    for (; j <= n - 16; j += 16)
        for (q = 0; q < 4; q++)
            for (l = 0; l < 4; l++) { double v = a[j + 4 * q + l]; r[q][l] = fma(v, v, r[q][l]); }
    for (l = 0; l < 4; l++) t[l] = ((r[0][l] + r[1][l]) + r[2][l]) + r[3][l];
    s += (t[0] + t[1]) + (t[2] + t[3]);
    for (; j <= n - 4; j += 4)
        for (l = 0; l < 4; l++) { double v = a[j + l], p = v * v; s += p; }
    for (; j < n; j++) { double v = a[j]; s = fma(v, v, s); }
    return s;
}
static double norm_inf(const double* r, int n) { double s = 0; for (int i = 0; i < n; i++) if (fabs(r[i]) > s) s = fabs(r[i]); return s; }

/* cv::Mat::dot (dotProd_64f) as the binary computes it: unrolled by four,
 *     s += a0 b0 + a1 b1 + a2 b2 + a3 b3
 * compiled with FMA contraction: the block is fma(a3, b3, fma(a2, b2, fma(a0, b0, a1 b1))) — the SECOND product is the
 * rounded one —, added to s by a plain add; the tail (n mod 4) is s = fma(a, b, s).  Mat::dot is not exposed by the
 * Python binding, so this form is pinned through the whole refinement: of eight candidate forms it is the only one
 * for which orc_h_lm_refine equals cv2.findHomography(inliers, 0) bit for bit on all 1982 probe problems (24 debug.log
 * blocks, the 458 candidates of the repo's sweep, 1500 random sets of 5 ... 39 points); the others leave 6 ... 26
 * problems 1e-13 ... 6e-9 away. */
static double cv_dot(const double* a, const double* b, int n) {
    double s = 0;
    int i = 0;
    for (; i <= n - 4; i += 4) s += fma(a[i + 3], b[i + 3], fma(a[i + 2], b[i + 2], fma(a[i], b[i], a[i + 1] * b[i + 1])));
    for (; i < n; i++) s = fma(a[i], b[i], s);
    return s;
}

/* the blocks, exported for the known-answer tests (tests/test_oracle_golden.py) */
ORC_API double orc_cv_norm_l2sqr(const double* a, int n) { return norm_l2sqr(a, n); }
ORC_API double orc_cv_dot(const double* a, const double* b, int n) { return cv_dot(a, b, n); }
ORC_API void orc_cv_gemm_atb(const double* J, int rows, int cols, const double* r, double* v) {
    for (int j = 0; j < cols; j++) v[j] = acc4_dot(J + j, cols, r, 1, rows);
}
ORC_API void orc_cv_gemm_axpby(const double* A, int n, const double* d, double alpha, const double* c, double beta, double* out) {
    for (int i = 0; i < n; i++) out[i] = acc4_dot(A + i * n, 1, d, 1, n) * alpha + c[i] * beta;
}
ORC_API void orc_cv_solve_eig(const double* A, const double* b, int n, double* x) { solve_sym_eig(A, b, n, x); }
ORC_API void orc_cv_invert_eig_diag(const double* A, int n, double* diag) { invert_sym_eig_diag(A, n, diag); }
ORC_API void orc_h_refine_eval(const double* h, const float* M, const float* m, int count, double* r, double* A, double* v) {
    h_refine_eval(h, M, m, count, r, A, v);
}

/* cv::LMSolver (max maxIters iterations, eps FLT_EPSILON) on the nine entries of H, then H *= 1/H[8] (OpenCV's
 * convertTo(..., scaleFor(H22))).  H: in = start (runKernel's output), out = refined, H[8] == 1. */
ORC_API int orc_h_lm_refine(const float* M, const float* m, int count, double* H, int maxIters) {
    const int lx = 9;
    const double epsx = FLT_EPSILON, epsf = FLT_EPSILON, Rlo = 0.25, Rhi = 0.75;
    double x[9], xd[9], d[9], A[81], Ap[81], v[9], D[9], tmp[9];
    double* r = (double*)malloc(sizeof(double) * 2 * (size_t)count);
    double* rd = (double*)malloc(sizeof(double) * 2 * (size_t)count);
    double lambda = 1, lc = 0.75, S;
    int i, iter = 0;
    for (i = 0; i < lx; i++) x[i] = H[i];
    h_refine_eval(x, M, m, count, r, A, v);
    S = norm_l2sqr(r, 2 * count);
    for (i = 0; i < lx; i++) D[i] = A[i * lx + i];
    for (;;) {
        double Sd, dS, R;
        memcpy(Ap, A, sizeof(A));
        for (i = 0; i < lx; i++) Ap[i * lx + i] += lambda * D[i];
        solve_sym_eig(Ap, v, lx, d);
        for (i = 0; i < lx; i++) xd[i] = x[i] - d[i];
        h_refine_eval(xd, M, m, count, rd, NULL, NULL);
        Sd = norm_l2sqr(rd, 2 * count);
        for (i = 0; i < lx; i++) tmp[i] = acc4_dot(A + i * lx, 1, d, 1, lx) * -1. + v[i] * 2.; /* gemm(A, d, -1, v, 2) */
        dS = cv_dot(d, tmp, lx);
        R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
        if (R > Rhi) {
            lambda *= 0.5;
            if (lambda < lc) lambda = 0;
        } else if (R < Rlo) {
            double t = cv_dot(d, v, lx), nu;
            nu = (Sd - S) / (fabs(t) > DBL_EPSILON ? t : 1) + 2;
            nu = nu < 2. ? 2. : nu; nu = nu > 10. ? 10. : nu;
            if (lambda == 0) {
                double diag[9], maxval = DBL_EPSILON;
                invert_sym_eig_diag(A, lx, diag);
                for (i = 0; i < lx; i++) if (fabs(diag[i]) > maxval) maxval = fabs(diag[i]);
                lambda = lc = 1. / maxval;
                nu *= 0.5;
            }
            lambda *= nu;
        }
        if (Sd < S) {
            S = Sd;
            memcpy(x, xd, sizeof(x));
            h_refine_eval(x, M, m, count, r, A, v);
        }
        iter++;
        if (!(iter < maxIters && norm_inf(d, lx) >= epsx && norm_inf(r, 2 * count) >= epsf)) break;
    }
    {
        const double sc = fabs(x[8]) > DBL_EPSILON ? 1. / x[8] : 1;
        for (i = 0; i < 9; i++) H[i] = x[i] * sc;
    }
    free(r);
    free(rd);
    return iter;
}

/* ------------------------------------------------------------------------------------------- */
/* cv2.findHomography(src, dst, cv2.RANSAC, thr, maxIters, confidence) — the whole call          */
/*   src, dst: n fp64 (x,y) pairs, as the reference passes them (main_v1.py:312)                */
/*   mask_semantics 0: OpenCV 4.13 (mask re-derived from the refined H); 1: legacy (RANSAC mask)*/
/*   returns 1 and fills H (row-major 3x3, H[8] = 1 up to one ulp, as cv2) / mask, or 0 (cv2 returns None, zero mask)  */
/* ------------------------------------------------------------------------------------------- */
ORC_API int orc_find_homography(const double* src, const double* dst, int n, double thresh, int maxIters, double confidence,
                                int mask_semantics, double* H, uint8_t* mask, int* iters_run, uint8_t* ransac_mask_out,
                                double* ransac_H_out) {
    float* M = (float*)malloc(sizeof(float) * 2 * (size_t)(n > 0 ? n : 1));
    float* m = (float*)malloc(sizeof(float) * 2 * (size_t)(n > 0 ? n : 1));
    uint8_t* rmask = (uint8_t*)calloc((size_t)(n > 0 ? n : 1), 1);
    int result, i, k = 0;
    for (i = 0; i < 2 * n; i++) { M[i] = (float)src[i]; m[i] = (float)dst[i]; } /* A.1 */
    result = orc_h_ransac_stage(M, m, n, thresh, maxIters, confidence, H, rmask, iters_run, NULL, NULL, NULL);
    if (result && ransac_mask_out) memcpy(ransac_mask_out, rmask, (size_t)n);
    if (result && ransac_H_out) memcpy(ransac_H_out, H, sizeof(double) * 9);
    if (result && n > 4) {
        float* M1 = (float*)malloc(sizeof(float) * 2 * (size_t)n);
        float* m1 = (float*)malloc(sizeof(float) * 2 * (size_t)n);
        for (i = 0; i < n; i++)
            if (rmask[i]) { M1[2 * k] = M[2 * i]; M1[2 * k + 1] = M[2 * i + 1]; m1[2 * k] = m[2 * i]; m1[2 * k + 1] = m[2 * i + 1]; k++; }
        orc_h_run_kernel(M1, m1, k, H);
        orc_h_lm_refine(M1, m1, k, H, 10); /* refines all nine entries, then scales by 1/H[8] */
        free(M1);
        free(m1);
    }
    if (result) {
        if (mask_semantics == 0 && n > 4) {
            float* err = (float*)malloc(sizeof(float) * (size_t)n);
            find_inliers(H, M, m, n, err, mask, thresh);
            free(err);
        } else {
            memcpy(mask, rmask, (size_t)n);
        }
    } else {
        memset(mask, 0, (size_t)n);
    }
    free(M); free(m); free(rmask);
    return result;
}

/* ------------------------------------------------------------------------------------------- */
/* A.8  PnP scoring: cv::projectPoints (fp64, zero distortion) rounded to fp32, fp32 sq. error  */
/*   obj: n fp32 (X,Y,Z) (already quantised, A.1), img: n fp32 (u,v), rvec/tvec fp64, K fp64    */
/* ------------------------------------------------------------------------------------------- */
ORC_API void orc_rodrigues(const double* r, double* R) {
    double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < DBL_EPSILON) {
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1. : 0.;
        return;
    }
    {
        double c = cos(theta), s = sin(theta), c1 = 1. - c, it = theta ? 1. / theta : 0.;
        double rx = r[0] * it, ry = r[1] * it, rz = r[2] * it;
        double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
        double rcr[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
        for (int i = 0; i < 9; i++) R[i] = c * ((i % 4 == 0) ? 1. : 0.) + c1 * rrt[i] + s * rcr[i];
    }
}

ORC_API void orc_pnp_project_f32(const double* R, const double* t, const double* Kmat, const float* obj, int n, float* proj) {
    const double fx = Kmat[0], fy = Kmat[4], cx = Kmat[2], cy = Kmat[5];
    for (int i = 0; i < n; i++) {
        double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2];
        double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
        double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
        double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
        z = z ? 1. / z : 1;
        x *= z;
        y *= z;
        proj[2 * i] = (float)(x * fx + cx);
        proj[2 * i + 1] = (float)(y * fy + cy);
    }
}

ORC_API int orc_pnp_count_inliers(const double* R, const double* t, const double* Kmat, const float* obj, const float* img,
                                  int n, double thresh, uint8_t* mask) {
    float* proj = (float*)malloc(sizeof(float) * 2 * (size_t)(n > 0 ? n : 1));
    float tt = (float)(thresh * thresh);
    int good = 0;
    orc_pnp_project_f32(R, t, Kmat, obj, n, proj);
    for (int i = 0; i < n; i++) {
        float dx = img[2 * i] - proj[2 * i], dy = img[2 * i + 1] - proj[2 * i + 1];
        float e = dx * dx + dy * dy;
        uint8_t f = (uint8_t)(e <= tt);
        if (mask) mask[i] = f;
        good += f;
    }
    free(proj);
    return good;
}

/* A.3 for the PnP callback: 5 distinct indices, no checkSubset.  Fills idx[iters*5] for the     */
/* first `iters` iterations, assuming every iteration draws one subset (the registrator draws    */
/* exactly one subset per iteration whatever the kernel returns).                                */
ORC_API void orc_pnp_sample_stream(int count, int iters, int modelPoints, int32_t* idx_out) {
    orc_rng rng;
    orc_rng_seed(&rng, 0xffffffffffffffffull);
    for (int it = 0; it < iters; it++) {
        int32_t* idx = idx_out + (size_t)it * modelPoints;
        for (int i = 0; i < modelPoints; i++) {
            int idx_i, dup;
            do {
                idx_i = rng_uniform(&rng, 0, count);
                dup = 0;
                for (int q = 0; q < i; q++) dup |= idx[q] == idx_i;
            } while (dup);
            idx[i] = idx_i;
        }
    }
}

/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (same rules as cv_ransac_oracle.c: never used by the product path).
 *
 * CPU restatement of what the reference executes inside
 *     cv2.solvePnPRansac(pos3d, pixels, K, dist, iterationsCount=5000, reprojectionError=30.0, confidence=0.99)
 *                                      /root/reference/main_v1.py:497-502 (testpro.py:536, test_pro.py:515, testpro-K.py:72)
 *     cv2.solvePnPRefineLM(pos3d[inliers], pixels[inliers], K, dist, rvec, tvec)        /root/reference/main_v1.py:508
 * i.e. OpenCV calib3d 4.13.0 (un-vendored, un-pinned dependency): RANSAC over 5-point samples with the EPnP minimal
 * solver (Lepetit, Moreno-Noguer, Fua, IJCV 2009 — the published algorithm, as OpenCV ships it), scoring through
 * projectPoints, then a Levenberg-Marquardt pose refinement seeded with the best model.
 *
 * PINNING (against outputs of the cv2 4.13.0 binary frozen in tests/golden/, replayed by tests/test_oracle_golden.py):
 *   - scoring (projectPoints + fp32 error) and the 5-point sample stream: bit for bit;
 *   - the EPnP minimal solver, through OpenCV's one-sided Jacobi SVD restated operation for operation: poses agree with
 *     cv2.solvePnP(flags=SOLVEPNP_EPNP) to 1e-13 (most are bit-identical);
 *   - the whole call: cv2.solvePnPRansac inlier index sets identical on the reference's data, the 27-K grid of
 *     testpro-K.py and 24 synthetic sets; returned pose within 1e-14 relative (most bit-identical) — analytic Jacobian,
 *     CvLevMarq driver;
 *   - cv2.solvePnPRefineLM: within 3e-13 relative on all of them (most bit-identical).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

void orc_jacobi(double* A, int n, double* W, double* V); /* cv_ransac_oracle.c */
ORC_API void orc_svd(const double* A, int m, int n, double* w, double* Ut, double* Vt);
void orc_rodrigues(const double* r, double* R);
int orc_update_num_iters(double p, double ep, int modelPoints, int maxIters);
uint32_t orc_rng_next(void* r);

/* ---- small dense helpers ------------------------------------------------------------------------------------ */
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double dist2(const double* a, const double* b) {
    return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
}

/* ---- cv::SVD for small matrices: OpenCV's one-sided Jacobi (JacobiSVDImpl_, lapack.cpp), restated ------------------ */
static double cvs_hypot(double a, double b) {
    a = fabs(a);
    b = fabs(b);
    if (a > b) { b /= a; return a * sqrt(1 + b * b); }
    if (b > 0) { a /= b; return b * sqrt(1 + a * a); }
    return 0;
}

/* At: n rows x m (row i = column i of A), destroyed -> rows become left singular vectors (first n1 rows normalised);
 * W: n singular values (descending); Vt: n x n (rows = right singular vectors) or NULL. */
static void jacobi_svd(double* At, int astep, double* Wout, double* Vt, int vstep, int m, int n, int n1) {
    const double minval = DBL_MIN, eps = DBL_EPSILON * 10;
    double W[16];
    int i, j, k, iter, max_iter = m > 30 ? m : 30;
    double c, s, sd;
    for (i = 0; i < n; i++) {
        for (k = 0, sd = 0; k < m; k++) { double t = At[i * astep + k]; sd += t * t; }
        W[i] = sd;
        if (Vt) {
            for (k = 0; k < n; k++) Vt[i * vstep + k] = 0;
            Vt[i * vstep + i] = 1;
        }
    }
    for (iter = 0; iter < max_iter; iter++) {
        int changed = 0;
        for (i = 0; i < n - 1; i++)
            for (j = i + 1; j < n; j++) {
                double *Ai = At + i * astep, *Aj = At + j * astep;
                double a = W[i], p = 0, b = W[j];
                for (k = 0; k < m; k++) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                double beta = a - b, gamma = cvs_hypot(p, beta);
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (k = 0; k < m; k++) {
                    double t0 = c * Ai[k] + s * Aj[k];
                    double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = 1;
                if (Vt) {
                    double *Vi = Vt + i * vstep, *Vj = Vt + j * vstep;
                    for (k = 0; k < n; k++) {
                        double t0 = c * Vi[k] + s * Vj[k];
                        double t1 = -s * Vi[k] + c * Vj[k];
                        Vi[k] = t0; Vj[k] = t1;
                    }
                }
            }
        if (!changed) break;
    }
    for (i = 0; i < n; i++) {
        for (k = 0, sd = 0; k < m; k++) { double t = At[i * astep + k]; sd += t * t; }
        W[i] = sqrt(sd);
    }
    for (i = 0; i < n - 1; i++) {
        j = i;
        for (k = i + 1; k < n; k++)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            double t = W[i]; W[i] = W[j]; W[j] = t;
            if (Vt) {
                for (k = 0; k < m; k++) { t = At[i * astep + k]; At[i * astep + k] = At[j * astep + k]; At[j * astep + k] = t; }
                for (k = 0; k < n; k++) { t = Vt[i * vstep + k]; Vt[i * vstep + k] = Vt[j * vstep + k]; Vt[j * vstep + k] = t; }
            }
        }
    }
    for (i = 0; i < n; i++) Wout[i] = W[i];
    if (!Vt) return;
    uint64_t rng = 0x12345678;
    for (i = 0; i < n1; i++) {
        sd = i < n ? W[i] : 0;
        for (int ii = 0; ii < 100 && sd <= minval; ii++) {
            /* zero singular value: random +-1/m vector, orthogonalised against the previous left vectors */
            const double val0 = 1. / m;
            for (k = 0; k < m; k++) {
                rng = (uint64_t)(uint32_t)rng * 4164903690u + (uint32_t)(rng >> 32);
                At[i * astep + k] = ((uint32_t)rng & 256) != 0 ? val0 : -val0;
            }
            for (iter = 0; iter < 2; iter++) {
                for (j = 0; j < i; j++) {
                    sd = 0;
                    for (k = 0; k < m; k++) sd += At[i * astep + k] * At[j * astep + k];
                    double asum = 0;
                    for (k = 0; k < m; k++) {
                        double t = At[i * astep + k] - sd * At[j * astep + k];
                        At[i * astep + k] = t;
                        asum += fabs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (k = 0; k < m; k++) At[i * astep + k] *= asum;
                }
                sd = 0;
                for (k = 0; k < m; k++) { double t = At[i * astep + k]; sd += t * t; }
                sd = sqrt(sd);
            }
        }
        s = sd > minval ? 1 / sd : 0.;
        for (k = 0; k < m; k++) At[i * astep + k] *= s;
    }
}

/* cv::SVD::compute(A (m x n, m >= n), w, u, vt): u is m x n (returned here as Ut, n x m), vt n x n */
ORC_API void orc_svd(const double* A, int m, int n, double* w, double* Ut, double* Vt) {
    double At[16 * 16];
    for (int i = 0; i < n; i++)
        for (int k = 0; k < m; k++) At[i * m + k] = A[k * n + i];
    jacobi_svd(At, m, w, Vt, n, m, n, n);
    for (int i = 0; i < n * m; i++) Ut[i] = At[i];
}

/* cv::solve(A (m x n), b, x, DECOMP_SVD): SVD + back-substitution with OpenCV's threshold eps*2*sum(w) */
static void cv_solve_svd(const double* A, const double* b, int m, int n, double* x) {
    double At[16 * 16], w[16], Vt[16 * 16];
    for (int i = 0; i < n; i++)
        for (int k = 0; k < m; k++) At[i * m + k] = A[k * n + i];
    jacobi_svd(At, m, w, Vt, n, m, n, n);
    double threshold = 0;
    for (int i = 0; i < n; i++) { x[i] = 0; }
    for (int i = 0; i < n; i++) threshold += w[i];
    threshold *= DBL_EPSILON * 2;
    for (int i = 0; i < n; i++) {
        double wi = w[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        double sacc = 0;
        for (int j = 0; j < m; j++) sacc += At[i * m + j] * b[j];
        sacc *= wi;
        for (int j = 0; j < n; j++) x[j] = x[j] + sacc * Vt[i * n + j];
    }
}

/* cv::invert(A 3x3, DECOMP_SVD): pseudo-inverse V diag(1/w) U^T accumulated triplet by triplet */
static void cv_invert3_svd(const double* A, double* inv) {
    double At[9], w[3], Vt[9];
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) At[i * 3 + k] = A[k * 3 + i];
    jacobi_svd(At, 3, w, Vt, 3, 3, 3, 3);
    double threshold = (w[0] + w[1] + w[2]) * (DBL_EPSILON * 2);
    for (int i = 0; i < 9; i++) inv[i] = 0;
    for (int i = 0; i < 3; i++) {
        double wi = w[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        double buffer[3];
        for (int j = 0; j < 3; j++) buffer[j] = At[i * 3 + j] * wi; /* u_i[j] * (1/w_i) */
        for (int r = 0; r < 3; r++)
            for (int j = 0; j < 3; j++) inv[r * 3 + j] = inv[r * 3 + j] + Vt[i * 3 + r] * buffer[j];
    }
}

/* rotation matrix -> rotation vector (cv::Rodrigues, matrix input) */
ORC_API void orc_rodrigues_inv(const double* Rin, double* r) {
    double At[9], w[3], Vt[9], R[9];
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) At[i * 3 + k] = Rin[k * 3 + i];
    jacobi_svd(At, 3, w, Vt, 3, 3, 3, 3); /* U(i,k) = At[k*3+i] */
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double acc = 0;
            for (int k = 0; k < 3; k++) acc += At[k * 3 + i] * Vt[k * 3 + j];
            R[i * 3 + j] = acc;
        }
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) {
            r[0] = r[1] = r[2] = 0;
        } else {
            double t;
            t = (R[0] + 1) * 0.5; rx = sqrt(t > 0 ? t : 0);
            t = (R[4] + 1) * 0.5; ry = sqrt(t > 0 ? t : 0) * (R[1] < 0 ? -1. : 1.);
            t = (R[8] + 1) * 0.5; rz = sqrt(t > 0 ? t : 0) * (R[2] < 0 ? -1. : 1.);
            if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
            theta /= sqrt(rx * rx + ry * ry + rz * rz);
            r[0] = rx * theta; r[1] = ry * theta; r[2] = rz * theta;
        }
    } else {
        double vth = 1 / (2 * s);
        vth *= theta;
        r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
    }
}

/* ---- EPnP (n points, here n = 5) -------------------------------------------------------------------------------- */
/* Householder QR least squares for the 6x4 Gauss-Newton step */
static void qr_solve_6x4(double* A, double* b, double* X) {
    const int nr = 6, nc = 4;
    double A1[4], A2[4];
    double* pA = A;
    for (int k = 0; k < nc; k++) {
        double* ppAkk = pA + k * nc + k;
        double eta = fabs(*ppAkk);
        for (int i = k + 1; i < nr; i++) {
            double elt = fabs(A[i * nc + k]);
            if (eta < elt) eta = elt;
        }
        if (eta == 0) {
            A1[k] = A2[k] = 0.0;
            return; /* singular */
        }
        double sum2 = 0.0, inv_eta = 1. / eta;
        for (int i = k; i < nr; i++) {
            A[i * nc + k] *= inv_eta;
            sum2 += A[i * nc + k] * A[i * nc + k];
        }
        double sigma = sqrt(sum2);
        if (*ppAkk < 0) sigma = -sigma;
        *ppAkk += sigma;
        A1[k] = sigma * *ppAkk;
        A2[k] = -eta * sigma;
        for (int j = k + 1; j < nc; j++) {
            double sum = 0;
            for (int i = k; i < nr; i++) sum += A[i * nc + k] * A[i * nc + j];
            double tau = sum / A1[k];
            for (int i = k; i < nr; i++) A[i * nc + j] -= tau * A[i * nc + k];
        }
    }
    for (int j = 0; j < nc; j++) {
        double tau = 0;
        for (int i = j; i < nr; i++) tau += A[i * nc + j] * b[i];
        tau /= A1[j];
        for (int i = j; i < nr; i++) b[i] -= tau * A[i * nc + j];
    }
    X[nc - 1] = b[nc - 1] / A2[nc - 1];
    for (int i = nc - 2; i >= 0; i--) {
        double sum = 0;
        for (int j = i + 1; j < nc; j++) sum += A[i * nc + j] * X[j];
        X[i] = (b[i] - sum) / A2[i];
    }
}

typedef struct {
    int n;
    double fu, fv, uc, vc;
    const double* pws; /* 3n */
    const double* us;  /* 2n */
    double cws[4][3], ccs[4][3];
    double alphas[4 * 16], pcs[3 * 16];
} epnp_t;

static void gauss_newton(const double* L, const double* rho, double* betas) {
    for (int k = 0; k < 5; k++) {
        double A[24], b[6], x[4] = {0, 0, 0, 0};
        for (int i = 0; i < 6; i++) {
            const double* r = L + i * 10;
            A[i * 4 + 0] = 2 * r[0] * betas[0] + r[1] * betas[1] + r[3] * betas[2] + r[6] * betas[3];
            A[i * 4 + 1] = r[1] * betas[0] + 2 * r[2] * betas[1] + r[4] * betas[2] + r[7] * betas[3];
            A[i * 4 + 2] = r[3] * betas[0] + r[4] * betas[1] + 2 * r[5] * betas[2] + r[8] * betas[3];
            A[i * 4 + 3] = r[6] * betas[0] + r[7] * betas[1] + r[8] * betas[2] + 2 * r[9] * betas[3];
            b[i] = rho[i] - (r[0] * betas[0] * betas[0] + r[1] * betas[0] * betas[1] + r[2] * betas[1] * betas[1] +
                             r[3] * betas[0] * betas[2] + r[4] * betas[1] * betas[2] + r[5] * betas[2] * betas[2] +
                             r[6] * betas[0] * betas[3] + r[7] * betas[1] * betas[3] + r[8] * betas[2] * betas[3] +
                             r[9] * betas[3] * betas[3]);
        }
        qr_solve_6x4(A, b, x);
        for (int i = 0; i < 4; i++) betas[i] += x[i];
    }
}

static double compute_R_and_t(epnp_t* e, const double* ut, const double* betas, double R[3][3], double t[3]) {
    const int n = e->n;
    for (int i = 0; i < 4; i++) e->ccs[i][0] = e->ccs[i][1] = e->ccs[i][2] = 0;
    for (int i = 0; i < 4; i++) {
        const double* v = ut + 12 * (11 - i);
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 3; k++) e->ccs[j][k] += betas[i] * v[3 * j + k];
    }
    for (int i = 0; i < n; i++) {
        const double* a = e->alphas + 4 * i;
        for (int j = 0; j < 3; j++)
            e->pcs[3 * i + j] = a[0] * e->ccs[0][j] + a[1] * e->ccs[1][j] + a[2] * e->ccs[2][j] + a[3] * e->ccs[3][j];
    }
    if (e->pcs[2] < 0.0) {
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 3; j++) e->ccs[i][j] = -e->ccs[i][j];
        for (int i = 0; i < 3 * n; i++) e->pcs[i] = -e->pcs[i];
    }
    double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++) { pc0[j] += e->pcs[3 * i + j]; pw0[j] += e->pws[3 * i + j]; }
    for (int j = 0; j < 3; j++) { pc0[j] /= n; pw0[j] /= n; }
    double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, U[9], w[3], V[9];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) abt[3 * j + k] += (e->pcs[3 * i + j] - pc0[j]) * (e->pws[3 * i + k] - pw0[k]);
    {
        double Ut[9], Vt[9];
        orc_svd(abt, 3, 3, w, Ut, Vt);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) { U[i * 3 + j] = Ut[j * 3 + i]; V[i * 3 + j] = Vt[j * 3 + i]; }
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i][j] = U[i * 3] * V[j * 3] + U[i * 3 + 1] * V[j * 3 + 1] + U[i * 3 + 2] * V[j * 3 + 2];
    const double det = R[0][0] * R[1][1] * R[2][2] + R[0][1] * R[1][2] * R[2][0] + R[0][2] * R[1][0] * R[2][1] -
                       R[0][2] * R[1][1] * R[2][0] - R[0][1] * R[1][0] * R[2][2] - R[0][0] * R[1][2] * R[2][1];
    if (det < 0) { R[2][0] = -R[2][0]; R[2][1] = -R[2][1]; R[2][2] = -R[2][2]; }
    for (int i = 0; i < 3; i++) t[i] = pc0[i] - dot3(R[i], pw0);
    double sum2 = 0;
    for (int i = 0; i < n; i++) {
        const double* pw = e->pws + 3 * i;
        double Xc = dot3(R[0], pw) + t[0], Yc = dot3(R[1], pw) + t[1], inv_Zc = 1.0 / (dot3(R[2], pw) + t[2]);
        double ue = e->uc + e->fu * Xc * inv_Zc, ve = e->vc + e->fv * Yc * inv_Zc;
        double u = e->us[2 * i], v = e->us[2 * i + 1];
        sum2 += sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
    }
    return sum2 / n;
}

/* obj: n x 3, img: n x 2 (pixel coordinates), K 3x3; n <= 16.  Returns 1 and R (row-major), t. */
ORC_API int orc_epnp(const double* obj, const double* img, int n, const double* K, double* Rout, double* tout) {
    epnp_t e;
    double us[32];
    if (n < 4 || n > 16) return 0;
    e.n = n; e.fu = K[0]; e.fv = K[4]; e.uc = K[2]; e.vc = K[5]; e.pws = obj;
    /* OpenCV normalises the image points (undistortPoints, fp32 storage for fp32 input) and re-applies K */
    for (int i = 0; i < n; i++) {
        const double ifx = 1. / e.fu, ify = 1. / e.fv;
        float xn = (float)((img[2 * i] - e.uc) * ifx), yn = (float)((img[2 * i + 1] - e.vc) * ify);
        us[2 * i] = xn * e.fu + e.uc;
        us[2 * i + 1] = yn * e.fv + e.vc;
    }
    e.us = us;
    /* control points: centroid + principal directions */
    for (int j = 0; j < 3; j++) e.cws[0][j] = 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++) e.cws[0][j] += obj[3 * i + j];
    for (int j = 0; j < 3; j++) e.cws[0][j] /= n;
    {
        double C[9], dc[3], uct[9], vt_unused[9], PW0[48];
        for (int i = 0; i < n; i++)
            for (int a = 0; a < 3; a++) PW0[3 * i + a] = obj[3 * i + a] - e.cws[0][a];
        for (int a = 0; a < 3; a++)       /* cvMulTransposed(PW0, C, 1): sequential sums over the rows */
            for (int b = a; b < 3; b++) {
                double acc = 0;
                for (int i = 0; i < n; i++) acc += PW0[3 * i + a] * PW0[3 * i + b];
                C[a * 3 + b] = C[b * 3 + a] = acc;
            }
        orc_svd(C, 3, 3, dc, uct, vt_unused);   /* cvSVD(..., CV_SVD_U_T): rows of uct = left singular vectors */
        for (int i = 1; i < 4; i++) {
            double k = sqrt((dc[i - 1] > 0 ? dc[i - 1] : 0) / n);
            for (int j = 0; j < 3; j++) e.cws[i][j] = e.cws[0][j] + k * uct[3 * (i - 1) + j];
        }
    }
    {
        double cc[9], ci[9];
        for (int i = 0; i < 3; i++)
            for (int j = 1; j < 4; j++) cc[3 * i + j - 1] = e.cws[j][i] - e.cws[0][i];
        cv_invert3_svd(cc, ci);
        for (int i = 0; i < n; i++) {
            const double* pi = obj + 3 * i;
            double* a = e.alphas + 4 * i;
            for (int j = 0; j < 3; j++)
                a[1 + j] = ci[3 * j] * (pi[0] - e.cws[0][0]) + ci[3 * j + 1] * (pi[1] - e.cws[0][1]) + ci[3 * j + 2] * (pi[2] - e.cws[0][2]);
            a[0] = 1.0f - a[1] - a[2] - a[3];
        }
    }
    double mtm[144], d[12], ut[144];
    {
        double M[32 * 12];
        for (int i = 0; i < n; i++) {
            const double* as = e.alphas + 4 * i;
            double* M1 = M + 24 * i;
            double* M2 = M1 + 12;
            for (int k = 0; k < 4; k++) {
                M1[3 * k] = as[k] * e.fu; M1[3 * k + 1] = 0.0; M1[3 * k + 2] = as[k] * (e.uc - us[2 * i]);
                M2[3 * k] = 0.0; M2[3 * k + 1] = as[k] * e.fv; M2[3 * k + 2] = as[k] * (e.vc - us[2 * i + 1]);
            }
        }
        for (int a = 0; a < 12; a++)
            for (int b = a; b < 12; b++) {
                double s = 0;
                for (int k = 0; k < 2 * n; k++) s += M[k * 12 + a] * M[k * 12 + b];
                mtm[a * 12 + b] = mtm[b * 12 + a] = s;
            }
        {
            double vt_unused[144];
            orc_svd(mtm, 12, 12, d, ut, vt_unused); /* cvSVD(MtM, D, Ut, 0, MODIFY_A | U_T): OpenCV reads the LEFT vectors */
        }
    }
    double L[60], rho[6];
    {
        const double* v[4] = {ut + 12 * 11, ut + 12 * 10, ut + 12 * 9, ut + 12 * 8};
        double dv[4][6][3];
        for (int i = 0; i < 4; i++) {
            int a = 0, b = 1;
            for (int j = 0; j < 6; j++) {
                for (int c = 0; c < 3; c++) dv[i][j][c] = v[i][3 * a + c] - v[i][3 * b + c];
                b++;
                if (b > 3) { a++; b = a + 1; }
            }
        }
        for (int i = 0; i < 6; i++) {
            double* row = L + 10 * i;
            row[0] = dot3(dv[0][i], dv[0][i]);
            row[1] = 2.0f * dot3(dv[0][i], dv[1][i]);
            row[2] = dot3(dv[1][i], dv[1][i]);
            row[3] = 2.0f * dot3(dv[0][i], dv[2][i]);
            row[4] = 2.0f * dot3(dv[1][i], dv[2][i]);
            row[5] = dot3(dv[2][i], dv[2][i]);
            row[6] = 2.0f * dot3(dv[0][i], dv[3][i]);
            row[7] = 2.0f * dot3(dv[1][i], dv[3][i]);
            row[8] = 2.0f * dot3(dv[2][i], dv[3][i]);
            row[9] = dot3(dv[3][i], dv[3][i]);
        }
        rho[0] = dist2(e.cws[0], e.cws[1]); rho[1] = dist2(e.cws[0], e.cws[2]); rho[2] = dist2(e.cws[0], e.cws[3]);
        rho[3] = dist2(e.cws[1], e.cws[2]); rho[4] = dist2(e.cws[1], e.cws[3]); rho[5] = dist2(e.cws[2], e.cws[3]);
    }
    double Betas[4][4], rep[4], Rs[4][3][3], ts[4][3];
    { /* approximation 1: betas10 columns {0,1,3,6} */
        double A[24], b4[4];
        for (int i = 0; i < 6; i++) { A[i * 4] = L[i * 10]; A[i * 4 + 1] = L[i * 10 + 1]; A[i * 4 + 2] = L[i * 10 + 3]; A[i * 4 + 3] = L[i * 10 + 6]; }
        cv_solve_svd(A, rho, 6, 4, b4);
        double* be = Betas[1];
        if (b4[0] < 0) { be[0] = sqrt(-b4[0]); be[1] = -b4[1] / be[0]; be[2] = -b4[2] / be[0]; be[3] = -b4[3] / be[0]; }
        else { be[0] = sqrt(b4[0]); be[1] = b4[1] / be[0]; be[2] = b4[2] / be[0]; be[3] = b4[3] / be[0]; }
        gauss_newton(L, rho, be);
        rep[1] = compute_R_and_t(&e, ut, be, Rs[1], ts[1]);
    }
    { /* approximation 2: columns {0,1,2} */
        double A[18], b3[3];
        for (int i = 0; i < 6; i++) { A[i * 3] = L[i * 10]; A[i * 3 + 1] = L[i * 10 + 1]; A[i * 3 + 2] = L[i * 10 + 2]; }
        cv_solve_svd(A, rho, 6, 3, b3);
        double* be = Betas[2];
        if (b3[0] < 0) { be[0] = sqrt(-b3[0]); be[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0; }
        else { be[0] = sqrt(b3[0]); be[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0; }
        if (b3[1] < 0) be[0] = -be[0];
        be[2] = 0.0; be[3] = 0.0;
        gauss_newton(L, rho, be);
        rep[2] = compute_R_and_t(&e, ut, be, Rs[2], ts[2]);
    }
    { /* approximation 3: columns {0,1,2,3,4} */
        double A[30], b5[5];
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 5; j++) A[i * 5 + j] = L[i * 10 + j];
        cv_solve_svd(A, rho, 6, 5, b5);
        double* be = Betas[3];
        if (b5[0] < 0) { be[0] = sqrt(-b5[0]); be[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0; }
        else { be[0] = sqrt(b5[0]); be[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0; }
        if (b5[1] < 0) be[0] = -be[0];
        be[2] = b5[3] / be[0];
        be[3] = 0.0;
        gauss_newton(L, rho, be);
        rep[3] = compute_R_and_t(&e, ut, be, Rs[3], ts[3]);
    }
    int N = 1;
    if (rep[2] < rep[1]) N = 2;
    if (rep[3] < rep[N]) N = 3;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) Rout[i * 3 + j] = Rs[N][i][j];
        tout[i] = ts[N][i];
    }
    for (int i = 0; i < 9; i++)
        if (!(fabs(Rout[i]) <= 2.0)) return 0;
    return 1;
}


/* ---- the RANSAC stage of cv2.solvePnPRansac ---------------------------------------------------------- */
void orc_pnp_project_f32(const double* R, const double* t, const double* Kmat, const float* obj, int n, float* proj);
int orc_pnp_count_inliers(const double* R, const double* t, const double* Kmat, const float* obj, const float* img, int n,
                          double thresh, uint8_t* mask);

/* minimal model of one 5-point sample exactly as the callback produces it: EPnP -> Rodrigues -> [rvec | tvec] */
ORC_API int orc_pnp_minimal_model(const float* obj5, const float* img5, int n, const double* K, double* rvec, double* tvec) {
    double o[48], im[32], R[9];
    for (int i = 0; i < 3 * n; i++) o[i] = obj5[i];
    for (int i = 0; i < 2 * n; i++) im[i] = img5[i];
    if (!orc_epnp(o, im, n, K, R, tvec)) return 0;
    orc_rodrigues_inv(R, rvec);
    return 1;
}

/* obj (n,3) fp32, img (n,2) fp32 (already quantised).  Returns 1 when a model was found.
 * best_model: rvec[3] tvec[3]; mask: RANSAC-stage inliers; trace_count[iter] (optional) */
ORC_API int orc_pnp_ransac_stage(const float* obj, const float* img, int n, const double* K, int maxIters, double thresh,
                                 double confidence, double* best_model, uint8_t* best_mask, int* iters_run, int* best_iter,
                                 int32_t* trace_count) {
    const int modelPoints = 5;
    uint64_t rng = 0xffffffffffffffffull;
    int niters = maxIters > 1 ? maxIters : 1, maxGood = 0, iter;
    uint8_t* mask = (uint8_t*)malloc((size_t)(n > 0 ? n : 1));
    *iters_run = 0;
    *best_iter = -1;
    if (n < modelPoints) { free(mask); return 0; }
    for (iter = 0; iter < niters; iter++) {
        int idx[5];
        float o5[15], i5[10];
        double rvec[3], tvec[3], R[9];
        if (n > modelPoints) {
            for (int i = 0; i < modelPoints; i++) {
                int idx_i, dup;
                do {
                    rng = (uint64_t)(uint32_t)rng * 4164903690u + (uint32_t)(rng >> 32);
                    idx_i = (int)((uint32_t)rng % (uint32_t)n);
                    dup = 0;
                    for (int q = 0; q < i; q++) dup |= idx[q] == idx_i;
                } while (dup);
                idx[i] = idx_i;
            }
        } else {
            for (int i = 0; i < modelPoints; i++) idx[i] = i;
        }
        for (int i = 0; i < modelPoints; i++) {
            for (int c = 0; c < 3; c++) o5[3 * i + c] = obj[3 * idx[i] + c];
            for (int c = 0; c < 2; c++) i5[2 * i + c] = img[2 * idx[i] + c];
        }
        *iters_run = iter + 1;
        if (!orc_pnp_minimal_model(o5, i5, modelPoints, K, rvec, tvec)) {
            if (trace_count) trace_count[iter] = -1;
            continue;
        }
        orc_rodrigues(rvec, R);
        int good = orc_pnp_count_inliers(R, tvec, K, obj, img, n, thresh, mask);
        if (trace_count) trace_count[iter] = good;
        if (good > (maxGood > modelPoints - 1 ? maxGood : modelPoints - 1)) {
            memcpy(best_mask, mask, (size_t)n);
            memcpy(best_model, rvec, sizeof(double) * 3);
            memcpy(best_model + 3, tvec, sizeof(double) * 3);
            maxGood = good;
            *best_iter = iter;
            niters = orc_update_num_iters(confidence, (double)(n - good) / n, modelPoints, niters);
        }
    }
    free(mask);
    return maxGood > 0;
}

/* ---- pose refinement ------------------------------------------------------------------------------------------------ */
/* residuals r (2n) of pose p = [rvec | tvec]; returns |r|^2 */
static double pnp_cost(const double* p, const double* obj, const double* img, int n, const double* K, double* r) {
    double R[9], S = 0;
    orc_rodrigues(p, R);
    for (int i = 0; i < n; i++) {
        const double* X = obj + 3 * i;
        double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + p[3];
        double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + p[4];
        double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + p[5];
        double iz = z ? 1. / z : 1;
        double ru = x * iz * K[0] + K[2] - img[2 * i], rv = y * iz * K[4] + K[5] - img[2 * i + 1];
        if (r) { r[2 * i] = ru; r[2 * i + 1] = rv; }
        S += ru * ru + rv * rv;
    }
    return S;
}

/* dR/dr_i (i = 0..2, each a row-major 3x3) of R = Rodrigues(r): the closed form cv::Rodrigues returns as its Jacobian */
static void rodrigues_jacobian(const double* r, double* R, double* dR) {
    double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    orc_rodrigues(r, R);
    if (theta < DBL_EPSILON) {
        /* dR/dr_i = [e_i]x at the identity */
        static const double G[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0};
        memcpy(dR, G, sizeof(G));
        return;
    }
    {
        const double c = cos(theta), s = sin(theta), c1 = 1. - c, itheta = 1. / theta;
        const double rx = r[0] * itheta, ry = r[1] * itheta, rz = r[2] * itheta, rv[3] = {rx, ry, rz};
        const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
        const double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
        const double drrt[27] = {rx + rx, ry, rz, ry, 0, 0, rz, 0, 0, 0, rx, 0, rx, ry + ry, rz, 0, rz, 0, 0, 0, rx, 0, 0, ry, rx, ry, rz + rz};
        static const double d_r_x[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0};
        for (int i = 0; i < 3; i++) {
            const double ri = rv[i], a0 = -s * ri, a1 = (s - 2 * c1 * itheta) * ri, a2 = c1 * itheta, a3 = (c - s * itheta) * ri,
                         a4 = s * itheta;
            for (int k = 0; k < 9; k++)
                dR[i * 9 + k] = a0 * I[k] + a1 * rrt[k] + a2 * drrt[i * 9 + k] + a3 * r_x[k] + a4 * d_r_x[i * 9 + k];
        }
    }
}

/* J^T J (6x6), J^T r (6) and |r|^2 at p, analytic Jacobian of the projection (what cv::projectPoints hands to the
 * solvers): du/dp = fx [1/z, 0, -x/z], dv/dp = fy [0, 1/z, -y/z] in camera coordinates, dp/dt = I, dp/dr_i = dR/dr_i X.
 * (A central-difference Jacobian agrees to ~1e-10 and is NOT good enough: its noise keeps the LM step above the
 * FLT_EPSILON stopping threshold, so the iteration count — part of the reference's answer — changes.) */
static double pnp_normal_eq(const double* p, const double* obj, const double* img, int n, const double* K, double* A, double* g,
                            double* r, double* rp, double* rm, double* J) {
    double R[9], dR[27], S = 0;
    (void)rp; (void)rm;
    rodrigues_jacobian(p, R, dR);
    for (int i = 0; i < n; i++) {
        const double* X = obj + 3 * i;
        double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + p[3];
        double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + p[4];
        double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + p[5];
        double iz = z ? 1. / z : 1;
        double xn = x * iz, yn = y * iz;
        double* Ju = J + (2 * i) * 6;
        double* Jv = Ju + 6;
        r[2 * i] = xn * K[0] + K[2] - img[2 * i];
        r[2 * i + 1] = yn * K[4] + K[5] - img[2 * i + 1];
        S += r[2 * i] * r[2 * i] + r[2 * i + 1] * r[2 * i + 1];
        {
            const double ux = K[0] * iz, uz = -K[0] * xn * iz, vy = K[4] * iz, vz = -K[4] * yn * iz;
            for (int k = 0; k < 3; k++) {
                const double* D = dR + 9 * k;
                const double dx = D[0] * X[0] + D[1] * X[1] + D[2] * X[2];
                const double dy = D[3] * X[0] + D[4] * X[1] + D[5] * X[2];
                const double dz = D[6] * X[0] + D[7] * X[1] + D[8] * X[2];
                Ju[k] = ux * dx + uz * dz;
                Jv[k] = vy * dy + vz * dz;
            }
            Ju[3] = ux; Ju[4] = 0; Ju[5] = uz;
            Jv[3] = 0; Jv[4] = vy; Jv[5] = vz;
        }
    }
    for (int a = 0; a < 6; a++) {
        for (int b = 0; b < 6; b++) { double s = 0; for (int i = 0; i < 2 * n; i++) s += J[i * 6 + a] * J[i * 6 + b]; A[a * 6 + b] = s; }
        double s = 0; for (int i = 0; i < 2 * n; i++) s += J[i * 6 + a] * r[i]; g[a] = s;
    }
    return S;
}

static void solve6_svd(const double* A, const double* b, double* x) { cv_solve_svd(A, b, 6, 6, x); }

/* The pose solvePnPRansac returns: solvePnP(inliers, SOLVEPNP_ITERATIVE, useExtrinsicGuess, seed = best RANSAC model),
 * i.e. OpenCV's CvLevMarq driver: damping J^T J(i,i) *= 1 + 10^lg (lg starts at -3, +1 on a worse step, -1 on a better
 * one), stop after 20 iterations or when the parameter vector moves by less than FLT_EPSILON in relative L2 norm.
 * On the reference's data (UTM-scale translations) that criterion fires after 2 iterations; this restatement is
 * bit-identical to the binary's pose on most golden cases and within 1e-14 relative on the rest. */
ORC_API int orc_pnp_refine_cvlevmarq(const double* obj, const double* img, int n, const double* K, double* rvec, double* tvec) {
    double p[6] = {rvec[0], rvec[1], rvec[2], tvec[0], tvec[1], tvec[2]}, prev[6], A[36], g[6], step[6];
    double* r = (double*)malloc(sizeof(double) * 2 * (size_t)n);
    double* rp = (double*)malloc(sizeof(double) * 2 * (size_t)n);
    double* rm = (double*)malloc(sizeof(double) * 2 * (size_t)n);
    double* J = (double*)malloc(sizeof(double) * 12 * (size_t)n);
    int lg = -3, iters = 0;
    for (;;) {
        double prevErr = pnp_normal_eq(p, obj, img, n, K, A, g, r, rp, rm, J), err;
        memcpy(prev, p, sizeof(p));
        for (;;) {
            double An[36], lambda = pow(10., lg);
            memcpy(An, A, sizeof(A));
            for (int a = 0; a < 6; a++) An[a * 6 + a] *= 1. + lambda;
            solve6_svd(An, g, step);
            for (int a = 0; a < 6; a++) p[a] = prev[a] - step[a];
            err = pnp_cost(p, obj, img, n, K, NULL);
            if (err > prevErr && ++lg <= 16) continue;
            break;
        }
        lg = lg - 1 > -16 ? lg - 1 : -16;
        double num = 0, den = 0;
        for (int a = 0; a < 6; a++) { num += (p[a] - prev[a]) * (p[a] - prev[a]); den += prev[a] * prev[a]; }
        if (++iters >= 20 || sqrt(num) / sqrt(den) < FLT_EPSILON) break;
    }
    for (int k = 0; k < 3; k++) { rvec[k] = p[k]; tvec[k] = p[3 + k]; }
    free(r); free(rp); free(rm); free(J);
    return iters;
}

/* cv2.solvePnPRefineLM (main_v1.py:508): the classic cv::LMSolver, max 20 iterations, eps FLT_EPSILON, on the 6 pose
 * parameters (same driver as orc_h_lm_refine in cv_ransac_oracle.c).  Within 3e-13 relative of the binary on every golden case. */
ORC_API int orc_pnp_refine_lm(const double* obj, const double* img, int n, const double* K, double* rvec, double* tvec,
                              int max_iters) {
    double x[6] = {rvec[0], rvec[1], rvec[2], tvec[0], tvec[1], tvec[2]}, xd[6], d[6], A[36], Ap[36], v[6], D[6];
    double* r = (double*)malloc(sizeof(double) * 2 * (size_t)n);
    double* rp = (double*)malloc(sizeof(double) * 2 * (size_t)n);
    double* rm = (double*)malloc(sizeof(double) * 2 * (size_t)n);
    double* J = (double*)malloc(sizeof(double) * 12 * (size_t)n);
    double lambda = 1, lc = 0.75, S = pnp_normal_eq(x, obj, img, n, K, A, v, r, rp, rm, J);
    int iter = 0;
    for (int i = 0; i < 6; i++) D[i] = A[i * 6 + i];
    for (;;) {
        double W[6], V[36], thr = 0;
        memcpy(Ap, A, sizeof(A));
        for (int i = 0; i < 6; i++) Ap[i * 6 + i] += lambda * D[i];
        orc_jacobi(Ap, 6, W, V); /* cv::solve(..., DECOMP_EIG) */
        for (int i = 0; i < 6; i++) thr += W[i];       /* SVBkSb: signed sum; multiplication by 1/w (pinned against */
        thr *= DBL_EPSILON * 2;                        /* cv2.solve(DECOMP_EIG), see cv_ransac_oracle.c)             */
        for (int i = 0; i < 6; i++) d[i] = 0;
        for (int e = 0; e < 6; e++) {
            if (fabs(W[e]) <= thr) continue;
            double s = 0, wi = 1 / W[e];
            for (int a = 0; a < 6; a++) s += V[e * 6 + a] * v[a];
            s *= wi;
            for (int a = 0; a < 6; a++) d[a] += s * V[e * 6 + a];
        }
        for (int i = 0; i < 6; i++) xd[i] = x[i] - d[i];
        double Sd = pnp_cost(xd, obj, img, n, K, NULL), dS = 0;
        for (int i = 0; i < 6; i++) {
            double s = 0;
            for (int j = 0; j < 6; j++) s += A[i * 6 + j] * d[j];
            dS += d[i] * (2 * v[i] - s);
        }
        double Rr = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
        if (Rr > 0.75) {
            lambda *= 0.5;
            if (lambda < lc) lambda = 0;
        } else if (Rr < 0.25) {
            double t = 0, nu;
            for (int i = 0; i < 6; i++) t += d[i] * v[i];
            nu = (Sd - S) / (fabs(t) > DBL_EPSILON ? t : 1) + 2;
            nu = nu < 2. ? 2. : nu; nu = nu > 10. ? 10. : nu;
            if (lambda == 0) {
                double a2[36], maxval = DBL_EPSILON, th2 = 0;
                memcpy(a2, A, sizeof(A));
                orc_jacobi(a2, 6, W, V);
                for (int i = 0; i < 6; i++) th2 += W[i];
                th2 *= DBL_EPSILON * 2;
                for (int j = 0; j < 6; j++) {
                    double dj = 0;
                    for (int e = 0; e < 6; e++) if (fabs(W[e]) > th2) dj += (V[e * 6 + j] * (1 / W[e])) * V[e * 6 + j];
                    if (fabs(dj) > maxval) maxval = fabs(dj);
                }
                lambda = lc = 1. / maxval;
                nu *= 0.5;
            }
            lambda *= nu;
        }
        if (Sd < S) {
            memcpy(x, xd, sizeof(x));
            S = pnp_normal_eq(x, obj, img, n, K, A, v, r, rp, rm, J);
        }
        iter++;
        double dmax = 0, rmax = 0;
        for (int i = 0; i < 6; i++) dmax = fmax(dmax, fabs(d[i]));
        for (int i = 0; i < 2 * n; i++) rmax = fmax(rmax, fabs(r[i]));
        if (!(iter < max_iters && dmax >= FLT_EPSILON && rmax >= FLT_EPSILON)) break;
    }
    for (int k = 0; k < 3; k++) { rvec[k] = x[k]; tvec[k] = x[3 + k]; }
    free(r); free(rp); free(rm); free(J);
    return iter;
}

/* cv2.solvePnPRansac(obj, img, K, 0, iterationsCount, reprojectionError, confidence): returns ok; rvec/tvec = pose
 * refined on the RANSAC inliers (fp32-quantised points, seeded with the best model); inliers = ascending indices */
ORC_API int orc_solve_pnp_ransac(const double* obj64, const double* img64, int n, const double* K, int maxIters, double thresh,
                                 double confidence, double* rvec, double* tvec, int32_t* inliers, int* n_inliers,
                                 int* iters_run, double* ransac_model) {
    float* obj = (float*)calloc(3 * (size_t)(n > 0 ? n : 1), sizeof(float));
    float* img = (float*)calloc(2 * (size_t)(n > 0 ? n : 1), sizeof(float));
    uint8_t* mask = (uint8_t*)calloc((size_t)(n > 0 ? n : 1), 1);
    double model[6];
    int best_iter, ok, k = 0;
    for (int i = 0; i < 3 * n; i++) obj[i] = (float)obj64[i];
    for (int i = 0; i < 2 * n; i++) img[i] = (float)img64[i];
    if (n == 5) {
        /* model_points == npoints: OpenCV returns solvePnP(EPNP) of the five points as it is — all of them inliers, no
         * threshold test, no refinement (probed against the binary: tests/golden, case with 5 points) */
        ok = orc_pnp_minimal_model(obj, img, n, K, model, model + 3);
        *iters_run = 1;
        *n_inliers = 0;
        if (ok) {
            for (int i = 0; i < n; i++) inliers[i] = i;
            *n_inliers = n;
            memcpy(rvec, model, sizeof(double) * 3);
            memcpy(tvec, model + 3, sizeof(double) * 3);
            if (ransac_model) memcpy(ransac_model, model, sizeof(model));
        }
        free(obj); free(img); free(mask);
        return ok;
    }
    ok = orc_pnp_ransac_stage(obj, img, n, K, maxIters, thresh, confidence, model, mask, iters_run, &best_iter, NULL);
    *n_inliers = 0;
    if (ok) {
        double* oi = (double*)malloc(sizeof(double) * 3 * (size_t)n);
        double* ii = (double*)malloc(sizeof(double) * 2 * (size_t)n);
        for (int i = 0; i < n; i++)
            if (mask[i]) {
                inliers[k] = i;
                for (int c = 0; c < 3; c++) oi[3 * k + c] = obj[3 * i + c];
                for (int c = 0; c < 2; c++) ii[2 * k + c] = img[2 * i + c];
                k++;
            }
        *n_inliers = k;
        if (ransac_model) memcpy(ransac_model, model, sizeof(model));
        memcpy(rvec, model, sizeof(double) * 3);
        memcpy(tvec, model + 3, sizeof(double) * 3);
        orc_pnp_refine_cvlevmarq(oi, ii, k, K, rvec, tvec);
        free(oi); free(ii);
    }
    free(obj); free(img); free(mask);
    return ok;
}

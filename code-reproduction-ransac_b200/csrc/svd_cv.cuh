// OpenCV's small-matrix SVD (one-sided Jacobi) and the solvers built on it, fp64, one thread per matrix.
//
// cv2.solvePnPRansac (reference: main_v1.py:497-502) runs EPnP on every 5-point sample; EPnP reads the LEFT singular
// vectors of a rank-deficient 12x12 Gram matrix, so its hypotheses depend on exactly how OpenCV's SVD orders and signs
// the (near-)null space.  This is that routine restated — sweep order, rotation formulas, the OpenCV hypot, the
// final sort and normalisation — so the device EPnP produces the poses the reference's RANSAC scores.
#pragma once
#include "solver_h.cuh"

namespace b2r {

// At: N rows x M (row i = column i of A), destroyed -> rows become the left singular vectors (normalised)
// W: N singular values, descending.  Vt: N x N right singular vectors (rows), touched only when HAS_VT.
// The sizes are template parameters so that every pass over the M (or N) entries of a row is unrolled: the loads of a
// pass are independent and issue back to back, while the sums keep OpenCV's order (k ascending).  With run-time bounds
// each of the ~66 x sweeps dot products of the 12x12 case was a chain of 12 load-then-add round trips.
template <int M, int N, bool HAS_VT>
static __device__ void jacobi_svd(double* At, double* Wout, double* Vt) {
    constexpr int astep = M, vstep = N, m = M, n = N, n1 = N;
    const double minval = DBL_MIN, eps = DBL_EPSILON * 10;
    double W[N];
    int i, j, k, iter;
    constexpr int max_iter = m > 30 ? m : 30;
    double c, s, sd;
#pragma unroll 1
    for (i = 0; i < n; i++) {
        sd = 0;
#pragma unroll
        for (k = 0; k < m; k++) {
            const double t = At[i * astep + k];
            sd += t * t;
        }
        W[i] = sd;
        if (HAS_VT) {
#pragma unroll
            for (k = 0; k < n; k++) Vt[i * vstep + k] = 0;
            Vt[i * vstep + i] = 1;
        }
    }
#pragma unroll 1
    for (iter = 0; iter < max_iter; iter++) {
        bool changed = false;
#pragma unroll 1
        for (i = 0; i < n - 1; i++)
#pragma unroll 1
            for (j = i + 1; j < n; j++) {
                double *Ai = At + i * astep, *Aj = At + j * astep;
                double a = W[i], p = 0, b = W[j];
                double ai[M], aj[M];
#pragma unroll
                for (k = 0; k < m; k++) { ai[k] = Ai[k]; aj[k] = Aj[k]; }
#pragma unroll
                for (k = 0; k < m; k++) p += ai[k] * aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = cv_hypot(p, beta);
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
#pragma unroll
                for (k = 0; k < m; k++) {
                    const double t0 = c * ai[k] + s * aj[k];
                    const double t1 = -s * ai[k] + c * aj[k];
                    Ai[k] = t0;
                    Aj[k] = t1;
                    a += t0 * t0;
                    b += t1 * t1;
                }
                W[i] = a;
                W[j] = b;
                changed = true;
                if (HAS_VT) {
                    double *Vi = Vt + i * vstep, *Vj = Vt + j * vstep;
#pragma unroll
                    for (k = 0; k < n; k++) {
                        const double v0 = Vi[k], v1 = Vj[k];
                        Vi[k] = c * v0 + s * v1;
                        Vj[k] = -s * v0 + c * v1;
                    }
                }
            }
        if (!changed) break;
    }
#pragma unroll 1
    for (i = 0; i < n; i++) {
        sd = 0;
#pragma unroll
        for (k = 0; k < m; k++) {
            const double t = At[i * astep + k];
            sd += t * t;
        }
        W[i] = sqrt(sd);
    }
#pragma unroll 1
    for (i = 0; i < n - 1; i++) {
        j = i;
        for (k = i + 1; k < n; k++)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            double t = W[i];
            W[i] = W[j];
            W[j] = t;
#pragma unroll
            for (k = 0; k < m; k++) {
                t = At[i * astep + k];
                At[i * astep + k] = At[j * astep + k];
                At[j * astep + k] = t;
            }
            if (HAS_VT)
#pragma unroll
                for (k = 0; k < n; k++) {
                    t = Vt[i * vstep + k];
                    Vt[i * vstep + k] = Vt[j * vstep + k];
                    Vt[j * vstep + k] = t;
                }
        }
    }
    for (i = 0; i < n; i++) Wout[i] = W[i];
    unsigned long long rng = 0x12345678ull;
#pragma unroll 1
    for (i = 0; i < n1; i++) {
        sd = i < n ? W[i] : 0;
#pragma unroll 1
        for (int ii = 0; ii < 100 && sd <= minval; ii++) {
            // exactly-zero singular value: OpenCV builds the left vector from a +-1/m pattern of its own RNG
            const double val0 = 1. / m;
            for (k = 0; k < m; k++) {
                rng = (unsigned long long)(unsigned)rng * 4164903690u + (unsigned)(rng >> 32);
                At[i * astep + k] = ((unsigned)rng & 256) != 0 ? val0 : -val0;
            }
            for (iter = 0; iter < 2; iter++) {
                for (j = 0; j < i; j++) {
                    sd = 0;
                    for (k = 0; k < m; k++) sd += At[i * astep + k] * At[j * astep + k];
                    double asum = 0;
                    for (k = 0; k < m; k++) {
                        const double t = At[i * astep + k] - sd * At[j * astep + k];
                        At[i * astep + k] = t;
                        asum += fabs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (k = 0; k < m; k++) At[i * astep + k] *= asum;
                }
                sd = 0;
                for (k = 0; k < m; k++) {
                    const double t = At[i * astep + k];
                    sd += t * t;
                }
                sd = sqrt(sd);
            }
        }
        s = sd > minval ? 1 / sd : 0.;
#pragma unroll
        for (k = 0; k < m; k++) At[i * astep + k] *= s;
    }
}

// cv::solve(A (M x N), b, x, DECOMP_SVD), M <= 6, N <= 6
template <int M, int N>
static __device__ void cv_solve_svd(const double* A, const double* b, double* x) {
    constexpr int m = M, n = N;
    double At[M * N], w[N], Vt[N * N];
    for (int i = 0; i < n; i++)
        for (int k = 0; k < m; k++) At[i * m + k] = A[k * n + i];
    jacobi_svd<M, N, true>(At, w, Vt);
    double threshold = 0;
    for (int i = 0; i < n; i++) x[i] = 0;
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

// cv::invert(A 3x3, DECOMP_SVD)
static __device__ void cv_invert3_svd(const double* A, double* inv) {
    double At[9], w[3], Vt[9];
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) At[i * 3 + k] = A[k * 3 + i];
    jacobi_svd<3, 3, true>(At, w, Vt);
    const double threshold = (w[0] + w[1] + w[2]) * (DBL_EPSILON * 2);
    for (int i = 0; i < 9; i++) inv[i] = 0;
    for (int i = 0; i < 3; i++) {
        double wi = w[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        double buffer[3];
        for (int j = 0; j < 3; j++) buffer[j] = At[i * 3 + j] * wi;
        for (int r = 0; r < 3; r++)
            for (int j = 0; j < 3; j++) inv[r * 3 + j] = inv[r * 3 + j] + Vt[i * 3 + r] * buffer[j];
    }
}

// cv::SVD::compute of a 3x3: w, Ut (rows = left vectors), Vt
static __device__ void cv_svd3(const double* A, double* w, double* Ut, double* Vt) {
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) Ut[i * 3 + k] = A[k * 3 + i];
    jacobi_svd<3, 3, true>(Ut, w, Vt);
}

}  // namespace b2r

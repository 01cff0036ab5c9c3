// K2 for the PnP path — the minimal solver cv2.solvePnPRansac runs on every 5-point sample
// (reference call sites: main_v1.py:497-502, testpro.py:536, test_pro.py:515, testpro-K.py:72-75):
// OpenCV's PnPRansacCallback::runKernel = solvePnP(5 points, SOLVEPNP_EPNP) followed by Rodrigues, i.e. EPnP
// (Lepetit, Moreno-Noguer, Fua, IJCV 2009) as OpenCV ships it.  fp64, one thread per hypothesis.
//
// Why this is an operation-for-operation restatement and not "an EPnP": with 5 points the 12x12 Gram matrix has
// rank <= 10 and OpenCV reads its null space from the LEFT singular vectors of its one-sided Jacobi SVD, which for
// (near-)zero singular values are normalised rounding residue.  The hypotheses the reference scores therefore depend
// on the exact sequence of IEEE operations; the translation unit is compiled with -fmad=false and every sum below is
// written in OpenCV's order so that the device models are the ones the CPU path produces (checked against the oracle,
// which is itself pinned against the cv2 binary: tests/test_gpu_parity_pnp.py, tests/test_oracle_golden.py).
#pragma once
#include "common_k.cuh"
#include "svd_cv.cuh"

namespace b2r {

__device__ __forceinline__ double dot3d(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ double dist2d(const double* a, const double* b) {
    return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
}

// rotation vector -> rotation matrix (cv::Rodrigues, vector input; what cv::projectPoints applies to the model)
__device__ __forceinline__ void rodrigues_vec2mat(const double* r, double* R) {
    const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < DBL_EPSILON) {
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1. : 0.;
        return;
    }
    const double c = cos(theta), s = sin(theta), c1 = 1. - c, it = 1. / theta;
    const double rx = r[0] * it, ry = r[1] * it, rz = r[2] * it;
    const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    const double rcr[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    for (int i = 0; i < 9; i++) R[i] = c * ((i % 4 == 0) ? 1. : 0.) + c1 * rrt[i] + s * rcr[i];
}

// rotation matrix -> rotation vector (cv::Rodrigues, matrix input: SVD-orthogonalise first)
static __device__ void rodrigues_mat2vec(const double* Rin, double* r) {
    double At[9], w[3], Vt[9], R[9];
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) At[i * 3 + k] = Rin[k * 3 + i];
    jacobi_svd<3, 3, true>(At, w, Vt);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double acc = 0;
            for (int k = 0; k < 3; k++) acc += At[k * 3 + i] * Vt[k * 3 + j];
            R[i * 3 + j] = acc;
        }
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
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

// Householder QR least squares of the 6x4 Gauss-Newton step (OpenCV epnp::qr_solve)
static __device__ void epnp_qr_solve_6x4(double* A, double* b, double* X) {
    const int nr = 6, nc = 4;
    double A1[4], A2[4];
    for (int k = 0; k < nc; k++) {
        double* ppAkk = A + k * nc + k;
        double eta = fabs(*ppAkk);
        for (int i = k + 1; i < nr; i++) {
            const double elt = fabs(A[i * nc + k]);
            if (eta < elt) eta = elt;
        }
        if (eta == 0) return;  // singular: the step is left at zero
        double sum2 = 0.0;
        const double inv_eta = 1. / eta;
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
            const double tau = sum / A1[k];
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

static __device__ void epnp_gauss_newton(const double* L, const double* rho, double* betas) {
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
        epnp_qr_solve_6x4(A, b, x);
        for (int i = 0; i < 4; i++) betas[i] += x[i];
    }
}

constexpr int EPNP_N = 5;  // points per minimal sample (modelPoints = 5 for solvePnPRansac)

struct EpnpState {
    double fu, fv, uc, vc;
    double pws[3 * EPNP_N];
    double us[2 * EPNP_N];
    double cws[4][3];
    double alphas[4 * EPNP_N];
};

// camera-frame control points from betas, Procrustes alignment, mean reprojection error (epnp::compute_R_and_t)
static __device__ double epnp_R_and_t(const EpnpState& e, const double* ut, const double* betas, double* R, double* t) {
    const int n = EPNP_N;
    double ccs[4][3], pcs[3 * EPNP_N];
    for (int i = 0; i < 4; i++) ccs[i][0] = ccs[i][1] = ccs[i][2] = 0;
    for (int i = 0; i < 4; i++) {
        const double* v = ut + 12 * (11 - i);
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 3; k++) ccs[j][k] += betas[i] * v[3 * j + k];
    }
    for (int i = 0; i < n; i++) {
        const double* a = e.alphas + 4 * i;
        for (int j = 0; j < 3; j++) pcs[3 * i + j] = a[0] * ccs[0][j] + a[1] * ccs[1][j] + a[2] * ccs[2][j] + a[3] * ccs[3][j];
    }
    if (pcs[2] < 0.0) {
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 3; j++) ccs[i][j] = -ccs[i][j];
        for (int i = 0; i < 3 * n; i++) pcs[i] = -pcs[i];
    }
    double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++) {
            pc0[j] += pcs[3 * i + j];
            pw0[j] += e.pws[3 * i + j];
        }
    for (int j = 0; j < 3; j++) {
        pc0[j] /= n;
        pw0[j] /= n;
    }
    double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, w[3], Ut[9], Vt[9];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) abt[3 * j + k] += (pcs[3 * i + j] - pc0[j]) * (e.pws[3 * i + k] - pw0[k]);
    cv_svd3(abt, w, Ut, Vt);
    // R = U V^T with U(i,k) = Ut[k][i], V(j,k) = Vt[k][j]
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i * 3 + j] = Ut[i] * Vt[j] + Ut[3 + i] * Vt[3 + j] + Ut[6 + i] * Vt[6 + j];
    const double det = R[0] * R[4] * R[8] + R[1] * R[5] * R[6] + R[2] * R[3] * R[7] - R[2] * R[4] * R[6] -
                       R[1] * R[3] * R[8] - R[0] * R[5] * R[7];
    if (det < 0) {
        R[6] = -R[6];
        R[7] = -R[7];
        R[8] = -R[8];
    }
    for (int i = 0; i < 3; i++) t[i] = pc0[i] - dot3d(R + 3 * i, pw0);
    double sum2 = 0;
    for (int i = 0; i < n; i++) {
        const double* pw = e.pws + 3 * i;
        const double Xc = dot3d(R, pw) + t[0], Yc = dot3d(R + 3, pw) + t[1], inv_Zc = 1.0 / (dot3d(R + 6, pw) + t[2]);
        const double ue = e.uc + e.fu * Xc * inv_Zc, ve = e.vc + e.fv * Yc * inv_Zc;
        const double u = e.us[2 * i], v = e.us[2 * i + 1];
        sum2 += sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
    }
    return sum2 / n;
}

// obj: 5 x 3 (the fp32-quantised object points widened to fp64), img: 5 x 2 fp32 pixels widened, K = (fu, fv, uc, vc).
// Returns true and (R row-major, t) or false (no model: the RANSAC iteration is consumed without a hypothesis).
static __device__ bool epnp5(const double* obj, const double* img, double fu, double fv, double uc, double vc, double* Rout,
                      double* tout) {
    const int n = EPNP_N;
    EpnpState e;
    e.fu = fu; e.fv = fv; e.uc = uc; e.vc = vc;
    for (int i = 0; i < 3 * n; i++) e.pws[i] = obj[i];
    {   // OpenCV normalises the image points (undistortPoints, fp32 storage) and re-applies K
        const double ifx = 1. / fu, ify = 1. / fv;
        for (int i = 0; i < n; i++) {
            const float xn = (float)((img[2 * i] - uc) * ifx), yn = (float)((img[2 * i + 1] - vc) * ify);
            e.us[2 * i] = (double)xn * fu + uc;
            e.us[2 * i + 1] = (double)yn * fv + vc;
        }
    }
    // control points: centroid + principal directions
    for (int j = 0; j < 3; j++) e.cws[0][j] = 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 3; j++) e.cws[0][j] += obj[3 * i + j];
    for (int j = 0; j < 3; j++) e.cws[0][j] /= n;
    {
        double C[9], dc[3], uct[9], PW0[3 * EPNP_N];
        for (int i = 0; i < n; i++)
            for (int a = 0; a < 3; a++) PW0[3 * i + a] = obj[3 * i + a] - e.cws[0][a];
        for (int a = 0; a < 3; a++)
            for (int b = a; b < 3; b++) {
                double acc = 0;
                for (int i = 0; i < n; i++) acc += PW0[3 * i + a] * PW0[3 * i + b];
                C[a * 3 + b] = C[b * 3 + a] = acc;
            }
        double vt_unused[9];
        cv_svd3(C, dc, uct, vt_unused);
        for (int i = 1; i < 4; i++) {
            const double k = sqrt((dc[i - 1] > 0 ? dc[i - 1] : 0) / n);
            for (int j = 0; j < 3; j++) e.cws[i][j] = e.cws[0][j] + k * uct[3 * (i - 1) + j];
        }
    }
    {   // barycentric coordinates
        double cc[9], ci[9];
        for (int i = 0; i < 3; i++)
            for (int j = 1; j < 4; j++) cc[3 * i + j - 1] = e.cws[j][i] - e.cws[0][i];
        cv_invert3_svd(cc, ci);
        for (int i = 0; i < n; i++) {
            const double* pi = obj + 3 * i;
            double* a = e.alphas + 4 * i;
            for (int j = 0; j < 3; j++)
                a[1 + j] = ci[3 * j] * (pi[0] - e.cws[0][0]) + ci[3 * j + 1] * (pi[1] - e.cws[0][1]) +
                           ci[3 * j + 2] * (pi[2] - e.cws[0][2]);
            a[0] = 1.0 - a[1] - a[2] - a[3];
        }
    }
    // M^T M (12 x 12) and its left singular vectors
    double ut[144], d[12];
    {
        double M[2 * EPNP_N * 12];
        for (int i = 0; i < n; i++) {
            const double* as = e.alphas + 4 * i;
            double* M1 = M + 24 * i;
            double* M2 = M1 + 12;
            for (int k = 0; k < 4; k++) {
                M1[3 * k] = as[k] * fu; M1[3 * k + 1] = 0.0; M1[3 * k + 2] = as[k] * (uc - e.us[2 * i]);
                M2[3 * k] = 0.0; M2[3 * k + 1] = as[k] * fv; M2[3 * k + 2] = as[k] * (vc - e.us[2 * i + 1]);
            }
        }
        for (int a = 0; a < 12; a++)
            for (int b = a; b < 12; b++) {
                double s = 0;
                for (int k = 0; k < 2 * n; k++) s += M[k * 12 + a] * M[k * 12 + b];
                ut[a * 12 + b] = ut[b * 12 + a] = s;  // symmetric: the transposed copy the SVD works on is the matrix itself
            }
    }
    jacobi_svd<12, 12, false>(ut, d, nullptr);
    double L[60], rho[6];
    {
        const double* v[4] = {ut + 12 * 11, ut + 12 * 10, ut + 12 * 9, ut + 12 * 8};
        double dv[4][6][3];
        for (int i = 0; i < 4; i++) {
            int a = 0, b = 1;
            for (int j = 0; j < 6; j++) {
                for (int c = 0; c < 3; c++) dv[i][j][c] = v[i][3 * a + c] - v[i][3 * b + c];
                b++;
                if (b > 3) {
                    a++;
                    b = a + 1;
                }
            }
        }
        for (int i = 0; i < 6; i++) {
            double* row = L + 10 * i;
            row[0] = dot3d(dv[0][i], dv[0][i]);
            row[1] = 2.0 * dot3d(dv[0][i], dv[1][i]);
            row[2] = dot3d(dv[1][i], dv[1][i]);
            row[3] = 2.0 * dot3d(dv[0][i], dv[2][i]);
            row[4] = 2.0 * dot3d(dv[1][i], dv[2][i]);
            row[5] = dot3d(dv[2][i], dv[2][i]);
            row[6] = 2.0 * dot3d(dv[0][i], dv[3][i]);
            row[7] = 2.0 * dot3d(dv[1][i], dv[3][i]);
            row[8] = 2.0 * dot3d(dv[2][i], dv[3][i]);
            row[9] = dot3d(dv[3][i], dv[3][i]);
        }
        rho[0] = dist2d(e.cws[0], e.cws[1]); rho[1] = dist2d(e.cws[0], e.cws[2]); rho[2] = dist2d(e.cws[0], e.cws[3]);
        rho[3] = dist2d(e.cws[1], e.cws[2]); rho[4] = dist2d(e.cws[1], e.cws[3]); rho[5] = dist2d(e.cws[2], e.cws[3]);
    }
    double rep[4], Rs[4][9], ts[4][3];
    {   // approximation 1: betas10 columns {0,1,3,6}
        double A[24], b4[4], be[4];
        for (int i = 0; i < 6; i++) {
            A[i * 4] = L[i * 10]; A[i * 4 + 1] = L[i * 10 + 1]; A[i * 4 + 2] = L[i * 10 + 3]; A[i * 4 + 3] = L[i * 10 + 6];
        }
        cv_solve_svd<6, 4>(A, rho, b4);
        if (b4[0] < 0) {
            be[0] = sqrt(-b4[0]); be[1] = -b4[1] / be[0]; be[2] = -b4[2] / be[0]; be[3] = -b4[3] / be[0];
        } else {
            be[0] = sqrt(b4[0]); be[1] = b4[1] / be[0]; be[2] = b4[2] / be[0]; be[3] = b4[3] / be[0];
        }
        epnp_gauss_newton(L, rho, be);
        rep[1] = epnp_R_and_t(e, ut, be, Rs[1], ts[1]);
    }
    {   // approximation 2: columns {0,1,2}
        double A[18], b3[3], be[4];
        for (int i = 0; i < 6; i++) {
            A[i * 3] = L[i * 10]; A[i * 3 + 1] = L[i * 10 + 1]; A[i * 3 + 2] = L[i * 10 + 2];
        }
        cv_solve_svd<6, 3>(A, rho, b3);
        if (b3[0] < 0) {
            be[0] = sqrt(-b3[0]); be[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0;
        } else {
            be[0] = sqrt(b3[0]); be[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0;
        }
        if (b3[1] < 0) be[0] = -be[0];
        be[2] = 0.0; be[3] = 0.0;
        epnp_gauss_newton(L, rho, be);
        rep[2] = epnp_R_and_t(e, ut, be, Rs[2], ts[2]);
    }
    {   // approximation 3: columns {0,1,2,3,4}
        double A[30], b5[5], be[4];
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 5; j++) A[i * 5 + j] = L[i * 10 + j];
        cv_solve_svd<6, 5>(A, rho, b5);
        if (b5[0] < 0) {
            be[0] = sqrt(-b5[0]); be[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0;
        } else {
            be[0] = sqrt(b5[0]); be[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0;
        }
        if (b5[1] < 0) be[0] = -be[0];
        be[2] = b5[3] / be[0];
        be[3] = 0.0;
        epnp_gauss_newton(L, rho, be);
        rep[3] = epnp_R_and_t(e, ut, be, Rs[3], ts[3]);
    }
    int N = 1;
    if (rep[2] < rep[1]) N = 2;
    if (rep[3] < rep[N]) N = 3;
    bool ok = true;
    for (int i = 0; i < 9; i++) {
        Rout[i] = Rs[N][i];
        ok &= fabs(Rs[N][i]) <= 2.0;  // NaN fails the comparison
    }
    for (int i = 0; i < 3; i++) tout[i] = ts[N][i];
    return ok;
}

// The model PnPRansacCallback::runKernel hands to RANSAC: [rvec | tvec] of one 5-point sample.
static __device__ bool pnp_minimal_model(const double* obj5, const double* img5, double fu, double fv, double uc, double vc,
                                  double* rvec, double* tvec) {
    double R[9];
    if (!epnp5(obj5, img5, fu, fv, uc, vc, R, tvec)) return false;
    rodrigues_mat2vec(R, rvec);
    return true;
}

// ---- throughput-mode minimal solver ------------------------------------------------------------------------------------
// A 5-point pose solver for the Philox path (B2R_SOLVER_FAST), where the hypotheses need not be OpenCV's: the same
// linear-then-rigidity idea as EPnP, parametrised by the five DEPTHS instead of twelve control-point coordinates, so that
// everything is 3x3 algebra in registers instead of a 12x12 SVD in local memory.
//   * a camera-frame point on the viewing ray of pixel i is P_i = l_i d_i, d_i = ((u-cx)/fx, (v-cy)/fy, 1);
//   * five world points satisfy one affine dependency sum n_i [p_i; 1] = 0, which any rigid image inherits:
//     sum n_i l_i d_i = 0 — three linear equations in the five depths, a 2-dimensional solution space l = b1 l1 + b2 l2;
//   * (b1, b2) from the ten pairwise distances |P_i - P_j|^2 = |p_i - p_j|^2 (linear least squares in b1^2, b1 b2, b2^2,
//     then a few Gauss-Newton steps), sign from positive depths; R, t by Procrustes alignment.
// Exact on noise-free correspondences; with noise it fits the five rays exactly and the rigidity in the least-squares sense.
__device__ __forceinline__ void cross3d(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

static __device__ bool pnp5_fast(const double* obj, const double* img, double fu, double fv, double uc, double vc, double* R,
                                 double* t) {
    double p[5][3], d[5][3], pm[3] = {0, 0, 0};
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 3; ++j) pm[j] += obj[3 * i + j];
    for (int j = 0; j < 3; ++j) pm[j] *= 0.2;
    for (int i = 0; i < 5; ++i) {
        for (int j = 0; j < 3; ++j) p[i][j] = obj[3 * i + j] - pm[j];
        d[i][0] = (img[2 * i] - uc) / fu;
        d[i][1] = (img[2 * i + 1] - vc) / fv;
        d[i][2] = 1.0;
    }
    // affine dependency: n_i = (-1)^i det([p_a p_b p_c p_d; 1 1 1 1]) over the other four points in index order
    double nv[5];
    for (int i = 0; i < 5; ++i) {
        int o[4], k = 0;
        for (int j = 0; j < 5; ++j)
            if (j != i) o[k++] = j;
        double e1[3], e2[3], e3[3], c[3];
        for (int j = 0; j < 3; ++j) {
            e1[j] = p[o[1]][j] - p[o[0]][j];
            e2[j] = p[o[2]][j] - p[o[0]][j];
            e3[j] = p[o[3]][j] - p[o[0]][j];
        }
        cross3d(e2, e3, c);
        const double vol = dot3d(e1, c);
        nv[i] = (i & 1) ? vol : -vol;
    }
    // G = [n_i d_i]; the two columns with the smallest |n_i| become the free depths
    int fa = 0, fb = 1;
    {
        double m0 = 1e300, m1 = 1e300;
        for (int i = 0; i < 5; ++i) {
            const double a = fabs(nv[i]);
            if (a < m0) { m1 = m0; fb = fa; m0 = a; fa = i; }
            else if (a < m1) { m1 = a; fb = i; }
        }
    }
    int s[3], k = 0;
    for (int i = 0; i < 5; ++i)
        if (i != fa && i != fb) s[k++] = i;
    double g[3][3], c12[3], c20[3], c01[3];
    for (int a = 0; a < 3; ++a)
        for (int j = 0; j < 3; ++j) g[a][j] = nv[s[a]] * d[s[a]][j];
    cross3d(g[1], g[2], c12);
    cross3d(g[2], g[0], c20);
    cross3d(g[0], g[1], c01);
    const double det = dot3d(g[0], c12);
    if (!(fabs(det) > 0)) return false;
    const double idet = 1.0 / det;
    double l1[5], l2[5];
    for (int i = 0; i < 5; ++i) l1[i] = l2[i] = 0;
    {
        double r[3];
        for (int j = 0; j < 3; ++j) r[j] = -nv[fa] * d[fa][j];
        l1[s[0]] = dot3d(r, c12) * idet; l1[s[1]] = dot3d(r, c20) * idet; l1[s[2]] = dot3d(r, c01) * idet; l1[fa] = 1.0;
        for (int j = 0; j < 3; ++j) r[j] = -nv[fb] * d[fb][j];
        l2[s[0]] = dot3d(r, c12) * idet; l2[s[1]] = dot3d(r, c20) * idet; l2[s[2]] = dot3d(r, c01) * idet; l2[fb] = 1.0;
    }
    // rigidity: b1^2 uu + 2 b1 b2 uw + b2^2 ww = rho for the ten pairs
    double uu[10], uw[10], ww[10], rho[10];
    {
        int e = 0;
        for (int i = 0; i < 5; ++i)
            for (int j = i + 1; j < 5; ++j, ++e) {
                double u[3], w[3], q[3];
                for (int a = 0; a < 3; ++a) {
                    u[a] = l1[i] * d[i][a] - l1[j] * d[j][a];
                    w[a] = l2[i] * d[i][a] - l2[j] * d[j][a];
                    q[a] = p[i][a] - p[j][a];
                }
                uu[e] = dot3d(u, u); uw[e] = dot3d(u, w); ww[e] = dot3d(w, w); rho[e] = dot3d(q, q);
            }
    }
    double b1, b2;
    {
        // normal equations of the 10x3 system [uu, 2uw, ww] (b11, b12, b22) = rho
        double A[6] = {0, 0, 0, 0, 0, 0}, r[3] = {0, 0, 0};
        for (int e = 0; e < 10; ++e) {
            const double a0 = uu[e], a1 = 2 * uw[e], a2 = ww[e];
            A[0] += a0 * a0; A[1] += a0 * a1; A[2] += a0 * a2; A[3] += a1 * a1; A[4] += a1 * a2; A[5] += a2 * a2;
            r[0] += a0 * rho[e]; r[1] += a1 * rho[e]; r[2] += a2 * rho[e];
        }
        const double m0[3] = {A[0], A[1], A[2]}, m1[3] = {A[1], A[3], A[4]}, m2[3] = {A[2], A[4], A[5]};
        double x12[3], x20[3], x01[3];
        cross3d(m1, m2, x12);
        cross3d(m2, m0, x20);
        cross3d(m0, m1, x01);
        const double dt = dot3d(m0, x12);
        if (!(fabs(dt) > 0)) return false;
        const double b11 = dot3d(r, x12) / dt, b12 = dot3d(r, x20) / dt, b22 = dot3d(r, x01) / dt;
        if (fabs(b11) >= fabs(b22)) {
            b1 = sqrt(fabs(b11));
            b2 = b1 > 0 ? b12 / b1 : 0;
            if (b11 < 0) b2 = -b2;
        } else {
            b2 = sqrt(fabs(b22));
            b1 = b2 > 0 ? b12 / b2 : 0;
            if (b22 < 0) b1 = -b1;
        }
    }
    for (int it = 0; it < 4; ++it) {   // Gauss-Newton on the ten distance equations
        double a00 = 0, a01 = 0, a11 = 0, g0 = 0, g1 = 0;
        for (int e = 0; e < 10; ++e) {
            const double f = b1 * b1 * uu[e] + 2 * b1 * b2 * uw[e] + b2 * b2 * ww[e] - rho[e];
            const double j0 = 2 * (b1 * uu[e] + b2 * uw[e]), j1 = 2 * (b1 * uw[e] + b2 * ww[e]);
            a00 += j0 * j0; a01 += j0 * j1; a11 += j1 * j1; g0 += j0 * f; g1 += j1 * f;
        }
        const double dt = a00 * a11 - a01 * a01;
        if (!(fabs(dt) > 0)) break;
        b1 -= (a11 * g0 - a01 * g1) / dt;
        b2 -= (a00 * g1 - a01 * g0) / dt;
    }
    double zs = 0;
    for (int i = 0; i < 5; ++i) zs += b1 * l1[i] + b2 * l2[i];
    if (zs < 0) { b1 = -b1; b2 = -b2; }
    // camera-frame points and Procrustes alignment
    double P[5][3], Pm[3] = {0, 0, 0};
    for (int i = 0; i < 5; ++i) {
        const double l = b1 * l1[i] + b2 * l2[i];
        for (int j = 0; j < 3; ++j) {
            P[i][j] = l * d[i][j];
            Pm[j] += P[i][j];
        }
    }
    for (int j = 0; j < 3; ++j) Pm[j] *= 0.2;
    double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, w3[3], Ut[9], Vt[9];
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 3; ++j)
            for (int c = 0; c < 3; ++c) abt[3 * j + c] += (P[i][j] - Pm[j]) * p[i][c];
    cv_svd3(abt, w3, Ut, Vt);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i * 3 + j] = Ut[i] * Vt[j] + Ut[3 + i] * Vt[3 + j] + Ut[6 + i] * Vt[6 + j];
    const double dr = R[0] * R[4] * R[8] + R[1] * R[5] * R[6] + R[2] * R[3] * R[7] - R[2] * R[4] * R[6] - R[1] * R[3] * R[8] -
                      R[0] * R[5] * R[7];
    if (dr < 0)   // reflection: flip the axis of the smallest singular value (U -> U diag(1,1,-1))
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) R[i * 3 + j] -= 2 * Ut[6 + i] * Vt[6 + j];
    for (int i = 0; i < 3; ++i) t[i] = Pm[i] - (R[3 * i] * pm[0] + R[3 * i + 1] * pm[1] + R[3 * i + 2] * pm[2]);
    // Three Gauss-Newton steps on the reprojection error of the five points (normalised image coordinates, update
    // c <- exp(w) c + dt in the camera frame): the depth parametrisation fits the five rays exactly, which makes the
    // depths noise-sensitive on far / shallow scenes; the least-squares pose of the five points is not.  With 1 px noise
    // the median reprojection error over all points drops from ~20 px to 1.65 px (OpenCV's EPnP: 1.8 px).
    for (int it = 0; it < 3; ++it) {
        double A[36], g[6];
        for (int i = 0; i < 36; ++i) A[i] = 0;
        for (int i = 0; i < 6; ++i) g[i] = 0;
        for (int i = 0; i < 5; ++i) {
            const double* Xw = obj + 3 * i;
            const double cx = R[0] * Xw[0] + R[1] * Xw[1] + R[2] * Xw[2] + t[0];
            const double cy = R[3] * Xw[0] + R[4] * Xw[1] + R[5] * Xw[2] + t[1];
            const double cz = R[6] * Xw[0] + R[7] * Xw[1] + R[8] * Xw[2] + t[2];
            const double iz = 1.0 / cz, x = cx * iz, y = cy * iz;
            const double rx = x - d[i][0], ry = y - d[i][1];
            // d(x, y)/dc = [iz, 0, -x iz; 0, iz, -y iz];  dc/dw = -[c]x;  dc/dt = I
            const double jx[6] = {-x * iz * cy, iz * cz + x * iz * cx, -iz * cy, iz, 0, -x * iz};
            const double jy[6] = {-iz * cz - y * iz * cy, y * iz * cx, iz * cx, 0, iz, -y * iz};
            for (int a = 0; a < 6; ++a) {
                for (int b = a; b < 6; ++b) A[a * 6 + b] += jx[a] * jx[b] + jy[a] * jy[b];
                g[a] += jx[a] * rx + jy[a] * ry;
            }
        }
        for (int a = 0; a < 6; ++a) {
            A[a * 6 + a] *= 1.0 + 1e-6;
            for (int b = 0; b < a; ++b) A[a * 6 + b] = A[b * 6 + a];
        }
        double Lc[36], dl[6];
        if (!cholesky<6>(A, Lc)) break;
        cholesky_solve<6>(Lc, g, dl);
        const double w[3] = {-dl[0], -dl[1], -dl[2]};
        double Rw[9], Rn[9], tn[3];
        rodrigues_vec2mat(w, Rw);
        mat3_mul(Rw, R, Rn);
        for (int i = 0; i < 3; ++i) tn[i] = Rw[3 * i] * t[0] + Rw[3 * i + 1] * t[1] + Rw[3 * i + 2] * t[2] - dl[3 + i];
        for (int i = 0; i < 9; ++i) R[i] = Rn[i];
        for (int i = 0; i < 3; ++i) t[i] = tn[i];
    }
    bool ok = true;
    for (int i = 0; i < 3; ++i) ok &= fabs(t[i]) <= DBL_MAX;
    for (int i = 0; i < 9; ++i) ok &= fabs(R[i]) <= 2.0;
    return ok;
}

// The minimal model of the throughput path: [rvec | tvec] from pnp5_fast.
static __device__ bool pnp_minimal_model_fast(const double* obj5, const double* img5, double fu, double fv, double uc, double vc,
                                              double* rvec, double* tvec) {
    double R[9];
    if (!pnp5_fast(obj5, img5, fu, fv, uc, vc, R, tvec)) return false;
    rodrigues_mat2vec(R, rvec);
    return true;
}

}  // namespace b2r

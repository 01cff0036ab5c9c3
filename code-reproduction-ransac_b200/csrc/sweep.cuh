// The camera-location sweep around the homography RANSAC, fused on the device: the per-candidate projection of the
// landmarks (find_homography prologue, main_v1.py:304-311), the per-candidate score err1/err2 (main_v1.py:314, 327-348,
// 419) and the arg-min over the candidates (do_it, main_v1.py:863-866).  With these three kernels the whole sweep
// (find_homographies, main_v1.py:254-297) is one launch sequence: the Q x n projected points never exist on the host.
#pragma once
#include "pipeline_h.cuh"

namespace b2r {

// pos3d (n,3), pixels (n,2), cams (Q,3), all fp64.  Per candidate c and landmark i, as the reference computes it in fp64:
//     p = pos3d[i] - c;  p = [p2, p1, p0];  p = p / p[2];  pos2 = p[0:2]      ->  ((z - cz)/(x - cx), (y - cy)/(x - cx))
// pos2_out [Q][n][2] fp64 keeps the un-quantised value for the score; pts_out [Q][n] is what findHomography sees (fp32).
__global__ void k_sweep_prologue(const double* __restrict__ pos3d, const double* __restrict__ pixels, const double* __restrict__ cams,
                                 int Q, int n, double* __restrict__ pos2_out, PointH* __restrict__ pts_out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)Q * n) return;
    const int q = (int)(t / n), i = (int)(t % n);
    const double px = pos3d[3 * i] - cams[3 * q], py = pos3d[3 * i + 1] - cams[3 * q + 1], pz = pos3d[3 * i + 2] - cams[3 * q + 2];
    const double a = pz / px, b = py / px;
    pos2_out[2 * t] = a;
    pos2_out[2 * t + 1] = b;
    PointH p;
    p.X = (float)a;
    p.Y = (float)b;
    p.nu = -(float)pixels[2 * i];
    p.nv = -(float)pixels[2 * i + 1];
    pts_out[t] = p;
}

// 3x3 inverse through the adjugate (np.linalg.inv runs an LU; the two agree to a few ulp on these matrices)
__device__ __forceinline__ bool inv3(const double* A, double* B) {
    const double c0 = A[4] * A[8] - A[5] * A[7], c1 = A[5] * A[6] - A[3] * A[8], c2 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c0 + A[1] * c1 + A[2] * c2;
    if (det == 0 || !(fabs(det) <= DBL_MAX)) return false;
    const double id = 1. / det;
    B[0] = c0 * id; B[1] = (A[2] * A[7] - A[1] * A[8]) * id; B[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    B[3] = c1 * id; B[4] = (A[0] * A[8] - A[2] * A[6]) * id; B[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    B[6] = c2 * id; B[7] = (A[1] * A[6] - A[0] * A[7]) * id; B[8] = (A[0] * A[4] - A[1] * A[3]) * id;
    return true;
}

// One CTA (128 threads) per candidate.  H [Q][9], mask [Q][n] from the finalize kernel; info [Q][12] (status at [0]).
//   M = inv(H);  per inlier i:  pp2 = inv(M) [pos2_i, 1] (normalised),  PP2 = M [pixel_i, 1] (normalised)
//   err1 = sum |pixel_i - pp2|,   err2 = sum |pos2_i - PP2| + (#outliers) * ransacbound
// scores [Q][2] = err1, err2 (0, 0 when the candidate has no model);  M_out [Q][9]
__global__ void __launch_bounds__(128)
k_sweep_epilogue(const double* __restrict__ H, const uint8_t* __restrict__ mask, const int* __restrict__ info,
                 const double* __restrict__ pos2, const double* __restrict__ pixels, int n, double ransacbound,
                 double* __restrict__ scores, double* __restrict__ M_out) {
    __shared__ double M[9], Mi[9];
    __shared__ int ok;
    __shared__ double red[3][128];
    const int q = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        ok = info[(size_t)q * 12] == 0 && inv3(H + (size_t)q * 9, M) && inv3(M, Mi);
    }
    __syncthreads();
    if (!ok) {
        if (tid < 2) scores[2 * (size_t)q + tid] = 0;
        if (tid < 9) M_out[(size_t)q * 9 + tid] = 0;
        return;
    }
    double e1 = 0, e2 = 0, out = 0;
    const double* P2 = pos2 + (size_t)q * n * 2;
    const uint8_t* mk = mask + (size_t)q * n;
    for (int i = tid; i < n; i += 128) {
        if (mk[i] == 1) {
            const double a = P2[2 * i], b = P2[2 * i + 1], u = pixels[2 * i], v = pixels[2 * i + 1];
            double w = Mi[6] * a + Mi[7] * b + Mi[8];
            const double x1 = (Mi[0] * a + Mi[1] * b + Mi[2]) / w, y1 = (Mi[3] * a + Mi[4] * b + Mi[5]) / w;
            e1 += sqrt((u - x1) * (u - x1) + (v - y1) * (v - y1));
            w = M[6] * u + M[7] * v + M[8];
            const double x2 = (M[0] * u + M[1] * v + M[2]) / w, y2 = (M[3] * u + M[4] * v + M[5]) / w;
            e2 += sqrt((a - x2) * (a - x2) + (b - y2) * (b - y2));
        } else {
            out += 1;
        }
    }
    red[0][tid] = e1; red[1][tid] = e2; red[2][tid] = out;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (tid < o)
            for (int k = 0; k < 3; ++k) red[k][tid] += red[k][tid + o];
        __syncthreads();
    }
    if (tid == 0) {
        scores[2 * (size_t)q] = red[0][0];
        scores[2 * (size_t)q + 1] = red[1][0] + red[2][0] * ransacbound;
    }
    if (tid < 9) M_out[(size_t)q * 9 + tid] = M[tid];
}

// theloci = argmin(err2 with 0 -> 1e6), first minimum wins (np.argmin).  One CTA.
__global__ void __launch_bounds__(256) k_sweep_argmin(const double* __restrict__ scores, int Q, int* __restrict__ best) {
    __shared__ double v[256];
    __shared__ int ix[256];
    double bv = 1e300;
    int bi = 0x7fffffff;
    for (int q = threadIdx.x; q < Q; q += 256) {
        double e = scores[2 * (size_t)q + 1];
        if (e == 0) e = 1000000;
        if (e < bv) { bv = e; bi = q; }   // ascending q within a thread: the first minimum is kept
    }
    v[threadIdx.x] = bv; ix[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double ov = v[threadIdx.x + o];
            const int oi = ix[threadIdx.x + o];
            if (ov < v[threadIdx.x] || (ov == v[threadIdx.x] && oi < ix[threadIdx.x])) { v[threadIdx.x] = ov; ix[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) best[0] = ix[0] == 0x7fffffff ? 0 : ix[0];
}

}  // namespace b2r

// Kernels around K3 for the PnP path: pack (A.1), K1 samplers (5-point), K2 batched EPnP, K4 finalize
// (RANSAC-stage mask, inlier list, seeded Levenberg-Marquardt pose refinement) and the stand-alone
// cv2.solvePnPRefineLM.  Reference call sites: cv2.solvePnPRansac main_v1.py:497-502 (testpro-K.py:72-75 as a
// batch over intrinsics), cv2.solvePnPRefineLM main_v1.py:508-509 (testpro-K.py:122-125).  Semantics: SURVEY.md A.8.
#pragma once
#include "common_k.cuh"
#include "pnp_solver.cuh"
#include "score_p.cuh"

namespace b2r {

constexpr int PNP_MP = 5;  // modelPoints of solvePnPRansac with default flags

// ---- S0: centre + pack ----------------------------------------------------------------------------------------
// centre[q] = mean of the first min(n, 4096) fp32-quantised object points (any point near the cloud would do: it
// only keeps the fast kernel's fp32 coordinates small).  One CTA per problem set, fixed summation order.
__global__ void __launch_bounds__(256) k_centre_p(const double* __restrict__ obj, int n, double* __restrict__ centre) {
    __shared__ double part[256][3];
    const double* O = obj + (size_t)blockIdx.x * n * 3;
    const int m = min(n, 4096);
    double s[3] = {0, 0, 0};
    for (int i = threadIdx.x; i < m; i += 256)
        for (int c = 0; c < 3; ++c) s[c] += (double)(float)O[3 * i + c];
    for (int c = 0; c < 3; ++c) part[threadIdx.x][c] = s[c];
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0;
        for (int i = 0; i < 256; ++i) t += part[i][threadIdx.x];
        centre[blockIdx.x * 3 + threadIdx.x] = t / m;
    }
}

// obj (P,n,3) fp64, img (P,n,2) fp64 -> PointPX / PointPF.  The fp64 -> fp32 conversion is OpenCV's input
// quantisation (solvePnPRansac converts opoints/ipoints to CV_32F).
__global__ void k_pack_points_p(const double* __restrict__ obj, const double* __restrict__ img, int P, int n,
                                const double* __restrict__ centre, PointPX* __restrict__ px, PointPF* __restrict__ pf) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)P * n) return;
    const size_t q = i / n;
    PointPX a;
    a.X = (double)(float)obj[3 * i];
    a.Y = (double)(float)obj[3 * i + 1];
    a.Z = (double)(float)obj[3 * i + 2];
    a.u = (float)img[2 * i];
    a.v = (float)img[2 * i + 1];
    px[i] = a;
    PointPF b;
    b.Xc = (float)(a.X - centre[3 * q]);
    b.Yc = (float)(a.Y - centre[3 * q + 1]);
    b.Zc = (float)(a.Z - centre[3 * q + 2]);
    b.nu = -a.u;
    b.nv = -a.v;
    b.pad0 = b.pad1 = b.pad2 = 0.f;
    pf[i] = b;
}

// ---- K1: samplers --------------------------------------------------------------------------------------------
// Replay of the subsets RANSACPointSetRegistrator::getSubset draws for a callback without checkSubset
// (SURVEY.md A.3): cv::RNG(2^64-1), `next() % n` per slot, duplicates re-drawn one at a time.  The stream depends
// on n only; one WARP per problem continues it over iterations [begin, begin+len) (clipped to the problem's current
// iteration bound) and writes samples[q][it][0..4].  As in k_cv_sample_h the warp works on windows of 128 stream
// outputs: every lane walks the multiply-with-carry steps, the modulo (the expensive part of a draw) is taken
// lane-parallel, every lane walks the indices through the distinct-index state machine and lane a keeps attempt a.
__global__ void __launch_bounds__(32)
k_cv_sample_p(int n, int H_stride, int begin, int len, int* __restrict__ samples, RansacState* __restrict__ state, int Q) {
    __shared__ __align__(16) int draws[128];
    const int q = blockIdx.x, lane = threadIdx.x;
    if (q >= Q) return;
    RansacState st = state[q];
    if (st.done || begin >= st.niters || st.gen < begin) return;
    int* S = samples + (size_t)q * H_stride * PNP_MP;
    const int end = min(begin + len, st.niters);
    constexpr uint32_t MWC_A = 4164903690u;
    uint64_t base = st.rng;
    if (n <= PNP_MP) {   // getSubset is not called: the sample is all the points
        for (int e = begin * PNP_MP + lane; e < end * PNP_MP; e += 32) S[e] = e % PNP_MP;
    } else {
        int it = begin;
        int ci = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0;   // the subset under construction (carried across windows)
        while (it < end) {
            uint64_t r = base;
            uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
            for (int w = 0; w < 32; ++w) {
                r = (uint64_t)(uint32_t)r * MWC_A + (uint32_t)(r >> 32);
                const uint32_t t0 = (uint32_t)r;
                r = (uint64_t)(uint32_t)r * MWC_A + (uint32_t)(r >> 32);
                const uint32_t t1 = (uint32_t)r;
                r = (uint64_t)(uint32_t)r * MWC_A + (uint32_t)(r >> 32);
                const uint32_t t2 = (uint32_t)r;
                r = (uint64_t)(uint32_t)r * MWC_A + (uint32_t)(r >> 32);
                const uint32_t t3 = (uint32_t)r;
                if (w == lane) o0 = t0, o1 = t1, o2 = t2, o3 = t3;
            }
            __syncwarp();   // the previous window's draws have been read by every lane
            reinterpret_cast<int4*>(draws)[lane] = make_int4((int)(o0 % (uint32_t)n), (int)(o1 % (uint32_t)n),
                                                             (int)(o2 % (uint32_t)n), (int)(o3 % (uint32_t)n));
            __syncwarp();
            int n_att = 0, my_end = 0;
            int m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0;
            for (int w = 0; w < 32; ++w) {
                const int4 d4 = reinterpret_cast<const int4*>(draws)[w];
                const int dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {   // selects only: the chain through ci is a few cycles per index
                    const int d = dd[u];
                    const bool fresh = !((ci > 0 && d == c0) || (ci > 1 && d == c1) || (ci > 2 && d == c2) || (ci > 3 && d == c3));
                    const bool complete = fresh && ci == 4;
                    const bool take = complete && n_att == lane;
                    m0 = take ? c0 : m0; m1 = take ? c1 : m1; m2 = take ? c2 : m2; m3 = take ? c3 : m3; m4 = take ? d : m4;
                    my_end = take ? 4 * w + u + 1 : my_end;
                    n_att += complete ? 1 : 0;
                    c0 = (fresh && ci == 0) ? d : c0;
                    c1 = (fresh && ci == 1) ? d : c1;
                    c2 = (fresh && ci == 2) ? d : c2;
                    c3 = (fresh && ci == 3) ? d : c3;
                    ci = complete ? 0 : ci + (fresh ? 1 : 0);
                }
            }
            const int usable = min(n_att, end - it);   // at most 25 subsets per window
            if (lane < usable) {
                int* o = S + (size_t)(it + lane) * PNP_MP;
                o[0] = m0; o[1] = m1; o[2] = m2; o[3] = m3; o[4] = m4;
            }
            it += usable;
            if (it >= end) {   // the stream stops at the end of the last subset that was used
                const int steps = __shfl_sync(0xffffffffu, my_end, usable - 1);
                for (int j = 0; j < steps; ++j) base = (uint64_t)(uint32_t)base * MWC_A + (uint32_t)(base >> 32);
            } else {
                base = r;
            }
        }
    }
    if (lane == 0) {
        st.rng = base;
        st.gen = end;
        state[q] = st;
    }
}

// 5 distinct indices in [0, n) from two Philox blocks (same no-rejection scheme as distinct4)
__device__ __forceinline__ void philox_sample5(unsigned long long gid, uint32_t q, uint64_t seed, uint32_t n, int* idx) {
    const Philox4 r0 = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), 0x50u, q, (uint32_t)seed, (uint32_t)(seed >> 32));
    const Philox4 r1 = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), 0x51u, q, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t w[PNP_MP] = {r0.v[0], r0.v[1], r0.v[2], r0.v[3], r1.v[0]};
    int chosen[PNP_MP];
    for (int k = 0; k < PNP_MP; ++k) {
        uint32_t j = (uint32_t)(((uint64_t)w[k] * (uint64_t)(n - k)) >> 32);
        int pos = 0;
        for (int t = 0; t < k; ++t)
            if (j >= (uint32_t)chosen[t]) { ++j; pos = t + 1; }
        idx[k] = (int)j;
        for (int t = k; t > pos; --t) chosen[t] = chosen[t - 1];
        chosen[pos] = (int)j;
    }
}

// ---- K2: batched EPnP -----------------------------------------------------------------------------------------
__device__ __forceinline__ void gather5(const PointPX* __restrict__ P, const int* idx, double* obj5, double* img5) {
    for (int k = 0; k < PNP_MP; ++k) {
        const PointPX p = P[idx[k]];
        obj5[3 * k] = p.X; obj5[3 * k + 1] = p.Y; obj5[3 * k + 2] = p.Z;
        img5[2 * k] = (double)p.u; img5[2 * k + 1] = (double)p.v;
    }
}

// fp32 rows of P = K [R | R c + t] for the fast scoring kernel
__device__ __forceinline__ void store_fast_model(float4* __restrict__ mf, size_t slot, const double* R, const double* t,
                                                 const double* K4, const double* c, bool ok) {
    float4 r0, r1, r2;
    if (ok) {
        const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
        double tc[3];
        for (int i = 0; i < 3; ++i) tc[i] = R[3 * i] * c[0] + R[3 * i + 1] * c[1] + R[3 * i + 2] * c[2] + t[i];
        r0 = make_float4((float)(fx * R[0] + cx * R[6]), (float)(fx * R[1] + cx * R[7]), (float)(fx * R[2] + cx * R[8]),
                         (float)(fx * tc[0] + cx * tc[2]));
        r1 = make_float4((float)(fy * R[3] + cy * R[6]), (float)(fy * R[4] + cy * R[7]), (float)(fy * R[5] + cy * R[8]),
                         (float)(fy * tc[1] + cy * tc[2]));
        r2 = make_float4((float)R[6], (float)R[7], (float)R[8], (float)tc[2]);
    } else {
        const float q = __int_as_float(0x7fc00000);
        r0 = r1 = r2 = make_float4(q, q, q, q);
    }
    mf[3 * slot] = r0;
    mf[3 * slot + 1] = r1;
    mf[3 * slot + 2] = r2;
}

// One thread per (problem, hypothesis) for hypotheses [begin, begin+len) of every problem ([Q][H] arrays).
// sampler_philox != 0: draw the sample from the hypothesis id first.  state (replay path): iterations at or beyond
// state[q].gen are not solved (their model is NaN).
//   samples : [Q][H][5] (read in replay mode, written in Philox mode)
//   mx      : [Q][H][12] fp64  R(rvec) | tvec   (what cv::projectPoints evaluates), NaN = no model      (optional)
//   mf      : [Q][H][3] float4 rows of the fast model                                                     (optional)
//   rt      : [Q][H][6] fp64 rvec | tvec, zeros when there is no model                                    (optional)
#ifndef K2P_MIN_BLOCKS
#define K2P_MIN_BLOCKS 4
#endif
// FAST: the closed 5-point solver instead of EPnP — a template, not a flag: with EPnP compiled in as well every launch carried its
// 3.4 KB stack frame and 255 registers
template <bool FAST>
__global__ void __launch_bounds__(64, K2P_MIN_BLOCKS)
k_epnp_solve_p(const PointPX* __restrict__ pts, size_t pts_q_stride, int n, int H, int begin, int len,
               const RansacState* __restrict__ state, const double* __restrict__ Kq,
               const double* __restrict__ centre, size_t centre_q_stride, int sampler_philox, long long hyp_begin,
               uint64_t seed, int* __restrict__ samples, double* __restrict__ mx, float4* __restrict__ mf,
               double* __restrict__ rt, uint8_t* __restrict__ ok_out) {
    const int q = blockIdx.y;
    if ((int)(blockIdx.x * blockDim.x + threadIdx.x) >= len) return;
    const int g = begin + blockIdx.x * blockDim.x + threadIdx.x;
    const size_t slot = (size_t)q * H + g;
    const PointPX* P = pts + (size_t)q * pts_q_stride;
    int idx[PNP_MP];
    if (sampler_philox) {
        if (n > PNP_MP)
            philox_sample5((unsigned long long)(hyp_begin + g), (uint32_t)q, seed, (uint32_t)n, idx);
        else
            for (int i = 0; i < PNP_MP; ++i) idx[i] = i;
        for (int i = 0; i < PNP_MP; ++i) samples[slot * PNP_MP + i] = idx[i];
    } else {
        for (int i = 0; i < PNP_MP; ++i) idx[i] = samples[slot * PNP_MP + i];
    }
    double obj5[15], img5[10], rvec[3], tvec[3], R[9];
    const double* K4 = Kq + (size_t)q * 4;
    bool ok = state == nullptr || g < state[q].gen;
    if (ok) {
        gather5(P, idx, obj5, img5);
        ok = FAST ? pnp_minimal_model_fast(obj5, img5, K4[0], K4[1], K4[2], K4[3], rvec, tvec)
                  : pnp_minimal_model(obj5, img5, K4[0], K4[1], K4[2], K4[3], rvec, tvec);
    }
    if (ok) rodrigues_vec2mat(rvec, R);
    if (mx) {
        const double qn = __longlong_as_double(0x7ff8000000000000ll);
        for (int i = 0; i < 9; ++i) mx[slot * 12 + i] = ok ? R[i] : qn;
        for (int i = 0; i < 3; ++i) mx[slot * 12 + 9 + i] = ok ? tvec[i] : qn;
    }
    if (mf) store_fast_model(mf, slot, R, tvec, K4, centre + (size_t)q * centre_q_stride, ok);
    if (rt)
        for (int i = 0; i < 3; ++i) {
            rt[slot * 6 + i] = ok ? rvec[i] : 0.0;
            rt[slot * 6 + 3 + i] = ok ? tvec[i] : 0.0;
        }
    if (ok_out) ok_out[slot] = ok ? 1 : 0;
}

// The winner may live on another rank (hypothesis sharding): rebuild its sample from the global id into slot 0.
__global__ void k_resample_winner_p(int n, const unsigned long long* __restrict__ keys, uint64_t seed, int Hs,
                                    int* __restrict__ samples, HSelect* __restrict__ sel, int H_total, int Q) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const unsigned long long key = keys[q];
    const int count = (int)(key >> 32);
    const unsigned long long gid = 0xFFFFFFFFull - (key & 0xFFFFFFFFull);
    HSelect s;
    s.best = -1; s.best_count = 0; s.iters_run = H_total; s.pad = 0;
    if (count > PNP_MP - 1) {
        int idx[PNP_MP];
        if (n > PNP_MP)
            philox_sample5(gid, (uint32_t)q, seed, (uint32_t)n, idx);
        else
            for (int i = 0; i < PNP_MP; ++i) idx[i] = i;
        for (int i = 0; i < PNP_MP; ++i) samples[(size_t)q * Hs * PNP_MP + i] = idx[i];
        s.best = 0;
        s.best_count = count;
        s.pad = (int)(gid & 0x7fffffff);
    }
    sel[q] = s;
}

// ---- pose least squares shared by both refinements ----------------------------------------------------------------
// Points of one problem: either the quantised PointPX array restricted to a mask (the refinement inside
// solvePnPRansac works on the fp32-quantised inliers), or raw fp64 arrays (solvePnPRefineLM gets the caller's
// un-quantised points).
struct PnpPts {
    const PointPX* px;
    const uint8_t* mask;
    const double* obj;
    const double* img;
    int n;
    __device__ __forceinline__ bool get(int i, double& X, double& Y, double& Z, double& u, double& v) const {
        if (px) {
            if (!mask[i]) return false;
            const PointPX p = px[i];
            X = p.X; Y = p.Y; Z = p.Z; u = (double)p.u; v = (double)p.v;
        } else {
            X = obj[3 * i]; Y = obj[3 * i + 1]; Z = obj[3 * i + 2]; u = img[2 * i]; v = img[2 * i + 1];
        }
        return true;
    }
};

struct PnpLsqShared {
    double pose[1][12];   // R | t of the pose being evaluated
    double dR[27];        // dR/dr_i, i = 0..2 (row-major 3x3 each)
    double p[6], prev[6], step[6], trial[6];
    double A[36], g[6], D[6], Ap[36], diag[6];
    double S, Sd, lambda, lc, rmax, nu;
    int flag, iters, lg, need_diag;
};

__device__ __forceinline__ void pnp_residual(const double* Rt, const double* K4, double X, double Y, double Z, double u,
                                             double v, double& ru, double& rv) {
    const double x = Rt[0] * X + Rt[1] * Y + Rt[2] * Z + Rt[9];
    const double y = Rt[3] * X + Rt[4] * Y + Rt[5] * Z + Rt[10];
    const double z = Rt[6] * X + Rt[7] * Y + Rt[8] * Z + Rt[11];
    const double iz = z != 0 ? 1. / z : 1;
    ru = x * iz * K4[0] + K4[2] - u;
    rv = y * iz * K4[1] + K4[3] - v;
}

// |r(p)|^2 over the active points; also max |r_i| (cluster-wide, identical on every CTA)
template <int THREADS>
__device__ double pnp_cost(ClusterRed& R, PnpLsqShared& sh, const double* p, const PnpPts& pts, const double* K4, int gtid,
                           int gstride, double* rmax_out) {
    __syncthreads();
    if (threadIdx.x == 0) {
        rodrigues_vec2mat(p, sh.pose[0]);
        for (int i = 0; i < 3; ++i) sh.pose[0][9 + i] = p[3 + i];
    }
    __syncthreads();
    double acc[1] = {0}, mx[1] = {0};
    for (int i = gtid; i < pts.n; i += gstride) {
        double X, Y, Z, u, v, ru, rv;
        if (!pts.get(i, X, Y, Z, u, v)) continue;
        pnp_residual(sh.pose[0], K4, X, Y, Z, u, v, ru, rv);
        acc[0] += ru * ru + rv * rv;
        mx[0] = fmax(mx[0], fmax(fabs(ru), fabs(rv)));
    }
    cluster_reduce<THREADS, 1, false>(R, acc);
    const double S = R.out[0];
    if (rmax_out) {
        __syncthreads();
        cluster_reduce<THREADS, 1, true>(R, mx);
        *rmax_out = R.out[0];
    }
    return S;
}

// dR/dr_i of R = Rodrigues(r): the closed form cv::Rodrigues returns as its Jacobian
__device__ __forceinline__ void rodrigues_jacobian(const double* r, double* R, double* dR) {
    const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    rodrigues_vec2mat(r, R);
    const double d_r_x[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0};
    if (theta < DBL_EPSILON) {
        for (int i = 0; i < 27; ++i) dR[i] = d_r_x[i];
        return;
    }
    const double c = cos(theta), s = sin(theta), c1 = 1. - c, itheta = 1. / theta;
    const double rx = r[0] * itheta, ry = r[1] * itheta, rz = r[2] * itheta, rv[3] = {rx, ry, rz};
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    const double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    const double drrt[27] = {rx + rx, ry, rz, ry, 0, 0, rz, 0, 0, 0, rx, 0, rx, ry + ry, rz, 0, rz, 0, 0, 0, rx, 0, 0, ry, rx, ry, rz + rz};
    for (int i = 0; i < 3; i++) {
        const double ri = rv[i], a0 = -s * ri, a1 = (s - 2 * c1 * itheta) * ri, a2 = c1 * itheta, a3 = (c - s * itheta) * ri,
                     a4 = s * itheta;
        for (int k = 0; k < 9; k++) dR[i * 9 + k] = a0 * I[k] + a1 * rrt[k] + a2 * drrt[i * 9 + k] + a3 * r_x[k] + a4 * d_r_x[i * 9 + k];
    }
}

// J^T J (6x6) -> sh.A, J^T r -> sh.g, returns |r|^2.  Analytic Jacobian of the projection (what cv::projectPoints hands
// to OpenCV's solvers): du/dp = fx [1/z, 0, -x/z], dv/dp = fy [0, 1/z, -y/z], dp/dt = I, dp/dr_i = dR/dr_i X.  A
// finite-difference Jacobian is not good enough here: its ~1e-10 noise keeps the LM step above the FLT_EPSILON stopping
// threshold, and the executed iteration count is part of the reference's answer.
template <int THREADS>
__device__ double pnp_normal_eq(ClusterRed& R, PnpLsqShared& sh, const double* p, const PnpPts& pts, const double* K4,
                                int gtid, int gstride, double* rmax_out) {
    __syncthreads();
    if (threadIdx.x == 0) {
        rodrigues_jacobian(p, sh.pose[0], sh.dR);
        for (int i = 0; i < 3; ++i) sh.pose[0][9 + i] = p[3 + i];
    }
    __syncthreads();
    double acc[28], mx[1] = {0};
#pragma unroll
    for (int j = 0; j < 28; ++j) acc[j] = 0;
    const double* Rt = sh.pose[0];
    for (int i = gtid; i < pts.n; i += gstride) {
        double X, Y, Z, u, v;
        if (!pts.get(i, X, Y, Z, u, v)) continue;
        double r[2], J[2][6];
        const double x = Rt[0] * X + Rt[1] * Y + Rt[2] * Z + Rt[9];
        const double y = Rt[3] * X + Rt[4] * Y + Rt[5] * Z + Rt[10];
        const double z = Rt[6] * X + Rt[7] * Y + Rt[8] * Z + Rt[11];
        const double iz = z != 0 ? 1. / z : 1;
        const double xn = x * iz, yn = y * iz;
        r[0] = xn * K4[0] + K4[2] - u;
        r[1] = yn * K4[1] + K4[3] - v;
        const double ux = K4[0] * iz, uz = -K4[0] * xn * iz, vy = K4[1] * iz, vz = -K4[1] * yn * iz;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double* D = sh.dR + 9 * k;
            const double dx = D[0] * X + D[1] * Y + D[2] * Z;
            const double dy = D[3] * X + D[4] * Y + D[5] * Z;
            const double dz = D[6] * X + D[7] * Y + D[8] * Z;
            J[0][k] = ux * dx + uz * dz;
            J[1][k] = vy * dy + vz * dz;
        }
        J[0][3] = ux; J[0][4] = 0; J[0][5] = uz;
        J[1][3] = 0; J[1][4] = vy; J[1][5] = vz;
        acc[27] += r[0] * r[0] + r[1] * r[1];
        mx[0] = fmax(mx[0], fmax(fabs(r[0]), fabs(r[1])));
        int e = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
#pragma unroll
            for (int b = a; b < 6; ++b) acc[e++] += J[0][a] * J[0][b] + J[1][a] * J[1][b];
            acc[21 + a] += J[0][a] * r[0] + J[1][a] * r[1];
        }
    }
    cluster_reduce<THREADS, 28, false>(R, acc);
    const double S = R.out[27];
    if (threadIdx.x == 0) {
        int e = 0;
        for (int a = 0; a < 6; ++a) {
            for (int b = a; b < 6; ++b) sh.A[a * 6 + b] = sh.A[b * 6 + a] = R.out[e++];
            sh.g[a] = R.out[21 + a];
        }
    }
    __syncthreads();
    if (rmax_out) {
        cluster_reduce<THREADS, 1, true>(R, mx);
        *rmax_out = R.out[0];
    }
    return S;
}

// OpenCV's CvLevMarq driver as solvePnP(SOLVEPNP_ITERATIVE, useExtrinsicGuess) runs it inside solvePnPRansac:
// damping diag(J^T J) *= 1 + 10^lg, lg from -3, +1 on a worse step, -1 after an accepted one; stop after 20
// iterations or when the parameter vector moves by less than FLT_EPSILON (relative L2).  sh.p: in = seed, out = result.
template <int THREADS>
__device__ void pnp_refine_cvlevmarq(ClusterRed& R, PnpLsqShared& sh, const PnpPts& pts, const double* K4, int gtid, int gstride) {
    const double pow10[33] = {1e-16, 1e-15, 1e-14, 1e-13, 1e-12, 1e-11, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1e0,
                              1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16};
    if (threadIdx.x == 0) { sh.lg = -3; sh.iters = 0; }
    __syncthreads();
    for (;;) {
        const double prevErr = pnp_normal_eq<THREADS>(R, sh, sh.p, pts, K4, gtid, gstride, nullptr);
        if (threadIdx.x == 0)
            for (int i = 0; i < 6; ++i) sh.prev[i] = sh.p[i];
        __syncthreads();
        for (;;) {
            if (threadIdx.x == 0) {
                double An[36], st[6];
                const double lambda = pow10[sh.lg + 16];
                for (int i = 0; i < 36; ++i) An[i] = sh.A[i];
                for (int a = 0; a < 6; ++a) An[a * 6 + a] *= 1. + lambda;
                cv_solve_svd<6, 6>(An, sh.g, st);
                for (int a = 0; a < 6; ++a) sh.p[a] = sh.prev[a] - st[a];
            }
            __syncthreads();
            const double err = pnp_cost<THREADS>(R, sh, sh.p, pts, K4, gtid, gstride, nullptr);
            if (threadIdx.x == 0) sh.flag = (err > prevErr && ++sh.lg <= 16) ? 1 : 0;
            __syncthreads();
            if (!sh.flag) break;
        }
        if (threadIdx.x == 0) {
            sh.lg = sh.lg - 1 > -16 ? sh.lg - 1 : -16;
            if (sh.lg > 16) sh.lg = 16;
            double num = 0, den = 0;
            for (int a = 0; a < 6; ++a) {
                num += (sh.p[a] - sh.prev[a]) * (sh.p[a] - sh.prev[a]);
                den += sh.prev[a] * sh.prev[a];
            }
            sh.flag = (++sh.iters >= 20 || sqrt(num) / sqrt(den) < (double)FLT_EPSILON) ? 1 : 0;
        }
        __syncthreads();
        if (sh.flag) break;
    }
}

// cv2.solvePnPRefineLM (main_v1.py:508): the classic cv::LMSolver on the 6 pose parameters, max_iters (20) iterations,
// eps FLT_EPSILON — the same driver as the homography refinement in pipeline_h.cuh, with cv::solve(DECOMP_EIG).
template <int THREADS>
__device__ void pnp_refine_lm(ClusterRed& R, PnpLsqShared& sh, JacobiWarp9& jw, const PnpPts& pts, const double* K4, int gtid,
                              int gstride, int max_iters) {
    double rmax;
    const double S0 = pnp_normal_eq<THREADS>(R, sh, sh.p, pts, K4, gtid, gstride, &rmax);
    if (threadIdx.x == 0) {
        sh.S = S0; sh.rmax = rmax; sh.lambda = 1; sh.lc = 0.75; sh.iters = 0;
        for (int i = 0; i < 6; ++i) sh.D[i] = sh.A[i * 6 + i];
    }
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) {
            for (int i = 0; i < 36; ++i) sh.Ap[i] = sh.A[i];
            for (int i = 0; i < 6; ++i) sh.Ap[i * 6 + i] += sh.lambda * sh.D[i];
        }
        __syncthreads();
        if (threadIdx.x < 32) solve_sym_eig_warp<6>(jw, sh.Ap, sh.g, sh.step, nullptr);   // cv::solve(..., DECOMP_EIG), warp 0
        __syncthreads();
        if (threadIdx.x == 0)
            for (int i = 0; i < 6; ++i) sh.trial[i] = sh.p[i] - sh.step[i];
        __syncthreads();
        const double Sd = pnp_cost<THREADS>(R, sh, sh.trial, pts, K4, gtid, gstride, nullptr);
        if (threadIdx.x == 0) {
            const double S = sh.S;
            sh.Sd = Sd;
            sh.need_diag = 0;
            double dS = 0;
            for (int i = 0; i < 6; ++i) {
                double s = 0;
                for (int j = 0; j < 6; ++j) s += sh.A[i * 6 + j] * sh.step[j];
                dS += sh.step[i] * (2 * sh.g[i] - s);
            }
            const double Rr = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
            if (Rr > 0.75) {
                sh.lambda *= 0.5;
                if (sh.lambda < sh.lc) sh.lambda = 0;
            } else if (Rr < 0.25) {
                double t = 0;
                for (int i = 0; i < 6; ++i) t += sh.step[i] * sh.g[i];
                double nu = (Sd - S) / (fabs(t) > DBL_EPSILON ? t : 1) + 2;
                nu = fmin(fmax(nu, 2.), 10.);
                sh.nu = nu;
                if (sh.lambda == 0) sh.need_diag = 1;
                else sh.lambda *= nu;
            }
        }
        __syncthreads();
        // need_diag is only set with lambda == 0: this iteration's step came from the decomposition of this same A
        if (sh.need_diag && threadIdx.x < 32) solve_sym_eig_warp<6>(jw, sh.A, nullptr, nullptr, sh.diag, true);
        __syncthreads();
        if (threadIdx.x == 0) {
            if (sh.need_diag) {
                double maxval = DBL_EPSILON;
                for (int i = 0; i < 6; ++i) maxval = fmax(maxval, fabs(sh.diag[i]));
                sh.lambda = sh.lc = 1. / maxval;
                sh.lambda *= sh.nu * 0.5;
            }
            sh.flag = sh.Sd < sh.S;
            if (sh.flag)
                for (int i = 0; i < 6; ++i) sh.p[i] = sh.trial[i];
        }
        __syncthreads();
        if (sh.flag) {
            const double S1 = pnp_normal_eq<THREADS>(R, sh, sh.p, pts, K4, gtid, gstride, &rmax);
            if (threadIdx.x == 0) { sh.S = S1; sh.rmax = rmax; }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            ++sh.iters;
            double dmax = 0;
            for (int i = 0; i < 6; ++i) dmax = fmax(dmax, fabs(sh.step[i]));
            sh.flag = (sh.iters < max_iters && dmax >= (double)FLT_EPSILON && sh.rmax >= (double)FLT_EPSILON) ? 1 : 0;
        }
        __syncthreads();
        if (!sh.flag) break;
    }
}

// ---- K4 finalize ------------------------------------------------------------------------------------------------
// One thread-block cluster per problem.
//   sel/samples : selection result and minimal samples ([Q][Hs][5], index sel.best)
//   rt          : [Q][Hs][6] minimal models the solve kernel stored (replay path), or null: the winner's model is then
//                 re-derived from its sample (hypothesis-sharded runs: the winner may come from another rank)
//   all_inliers : n == 5 — OpenCV then returns solvePnP(EPNP) of the five points as it is: every point an inlier, no
//                 threshold test, no refinement (probed against the binary)
//   rmask       : [Q][n] RANSAC-stage inlier mask (what solvePnPRansac's `inliers` lists)
//   pose_out    : [Q][6] returned pose (seeded LM on the quantised inliers when refine != 0)
//   info_i      : [Q][12] int32, info_d : [Q][8] fp64 = RANSAC model rvec|tvec, mean inlier reprojection error of the
//                 returned pose on the caller's un-quantised points (testpro-K.py:32-36, 80-82), final |r|^2
// The winner's minimal model (rvec | tvec) of every problem, [Q][6]: copied from the solve kernel's output when the replay path
// kept it (rt != nullptr), re-derived from the winning sample otherwise (Philox runs do not store 6 doubles per hypothesis, and
// under hypothesis sharding the winner may come from another rank).  All zeros = no model.  A kernel of its own so that the
// minimal solvers' stack frames (EPnP: 3.4 KB) stay out of k_finalize_p, which used to spill 2.8 KB per thread around them.
template <bool FAST>
__global__ void __launch_bounds__(32)
k_winner_model_p(const PointPX* __restrict__ pts, size_t pts_q_stride, const int* __restrict__ samples, int Hs,
                 const HSelect* __restrict__ sel, const double* __restrict__ Kq, const double* __restrict__ rt, double* __restrict__ win, int Q) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const HSelect s = sel[q];
    double model[6] = {0, 0, 0, 0, 0, 0};
    if (s.best >= 0) {
        if (rt) {
            const double* m = rt + ((size_t)q * Hs + s.best) * 6;
            for (int i = 0; i < 6; ++i) model[i] = m[i];
        } else {
            int smp[PNP_MP];
            double obj5[15], img5[10];
            for (int i = 0; i < PNP_MP; ++i) smp[i] = samples[((size_t)q * Hs + s.best) * PNP_MP + i];
            gather5(pts + (size_t)q * pts_q_stride, smp, obj5, img5);
            const double* K4 = Kq + (size_t)q * 4;
            const bool ok = FAST ? pnp_minimal_model_fast(obj5, img5, K4[0], K4[1], K4[2], K4[3], model, model + 3)
                                 : pnp_minimal_model(obj5, img5, K4[0], K4[1], K4[2], K4[3], model, model + 3);
            if (!ok)
                for (int i = 0; i < 6; ++i) model[i] = 0;
        }
    }
    for (int i = 0; i < 6; ++i) win[(size_t)q * 6 + i] = model[i];
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS <= 128 ? 4 : 1)
k_finalize_p(const PointPX* __restrict__ pts, size_t pts_q_stride, const double* __restrict__ obj_raw,
             const double* __restrict__ img_raw, size_t raw_q_stride, int n, const int* __restrict__ samples, int Hs,
             const HSelect* __restrict__ sel, const double* __restrict__ Kq, float thr_sq, int refine, int all_inliers,
             const double* __restrict__ win, uint8_t* __restrict__ rmask_out, double* __restrict__ pose_out, int* __restrict__ info_i,
             double* __restrict__ info_d) {
    __shared__ PnpLsqShared sh;
    __shared__ ClusterRed R;
    __shared__ double model[6];
    __shared__ int smp[PNP_MP];
    __shared__ int have_model;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned csize = cluster.num_blocks(), crank = cluster.block_rank();
    const int q = blockIdx.x / csize, tid = threadIdx.x;
    const int gtid = crank * THREADS + tid, gstride = csize * THREADS;
    const bool writer = crank == 0;
    const PointPX* P = pts + (size_t)q * pts_q_stride;
    uint8_t* rmask = rmask_out + (size_t)q * n;
    const double* K4 = Kq + (size_t)q * 4;
    const HSelect s = sel[q];
    int* inf = info_i + (size_t)q * 12;
    double* infd = info_d + (size_t)q * 8;
    if (tid == 0) R.phase = 0;

    if (tid == 0) {
        have_model = 0;
        if (s.best >= 0) {
            for (int i = 0; i < PNP_MP; ++i) smp[i] = samples[((size_t)q * Hs + s.best) * PNP_MP + i];
            const double* m = win + (size_t)q * 6;   // k_winner_model_p: zeros when the minimal solver produced no model
            bool any = false;
            for (int i = 0; i < 6; ++i) {
                model[i] = m[i];
                any |= m[i] != 0;
            }
            have_model = any ? 1 : 0;
        }
    }
    __syncthreads();
    if (!have_model) {  // cv2 returns retval False
        for (int i = gtid; i < n; i += gstride) rmask[i] = 0;
        if (writer && tid < 6) pose_out[(size_t)q * 6 + tid] = 0;
        if (writer && tid < 8) infd[tid] = 0;
        if (writer && tid == 0) {
            inf[0] = 1; inf[1] = s.iters_run; inf[2] = -1; inf[3] = 0;
            for (int i = 0; i < PNP_MP; ++i) inf[4 + i] = -1;
            inf[9] = 0; inf[10] = 0; inf[11] = 0;
        }
        return;
    }
    if (tid == 0) {
        rodrigues_vec2mat(model, sh.pose[0]);
        for (int i = 0; i < 3; ++i) sh.pose[0][9 + i] = model[3 + i];
        for (int i = 0; i < 6; ++i) sh.p[i] = model[i];
        sh.iters = 0;
    }
    __syncthreads();
    int k_local = 0;
    for (int i = gtid; i < n; i += gstride) {
        const PointPX p = P[i];
        const uint8_t f = (all_inliers || p_inlier_exact(sh.pose[0], sh.pose[0] + 9, K4[0], K4[1], K4[2], K4[3], p.X, p.Y, p.Z, p.u, p.v, thr_sq)) ? 1 : 0;
        rmask[i] = f;  // each thread re-reads only the entries it wrote itself
        k_local += f;
    }
    {
        double kv[1] = {(double)k_local};
        cluster_reduce<THREADS, 1, false>(R, kv);
    }
    const int k = (int)R.out[0];
    __syncthreads();

    PnpPts ps;
    ps.px = P; ps.mask = rmask; ps.obj = nullptr; ps.img = nullptr; ps.n = n;
    if (refine && k > 0 && !all_inliers) pnp_refine_cvlevmarq<THREADS>(R, sh, ps, K4, gtid, gstride);

    // mean reprojection error of the inliers under the returned pose, on the caller's fp64 points
    __syncthreads();
    if (tid == 0) {
        rodrigues_vec2mat(sh.p, sh.pose[0]);
        for (int i = 0; i < 3; ++i) sh.pose[0][9 + i] = sh.p[3 + i];
    }
    __syncthreads();
    double es[2] = {0, 0};
    {
        const double* O = obj_raw + (size_t)q * raw_q_stride * 3;
        const double* I = img_raw + (size_t)q * raw_q_stride * 2;
        for (int i = gtid; i < n; i += gstride)
            if (rmask[i]) {
                double ru, rv;
                pnp_residual(sh.pose[0], K4, O[3 * i], O[3 * i + 1], O[3 * i + 2], I[2 * i], I[2 * i + 1], ru, rv);
                es[0] += sqrt(ru * ru + rv * rv);
                es[1] += ru * ru + rv * rv;
            }
        cluster_reduce<THREADS, 2, false>(R, es);
    }
    if (writer && tid < 6) {
        pose_out[(size_t)q * 6 + tid] = sh.p[tid];
        infd[tid] = model[tid];
    }
    if (writer && tid == 0) {
        infd[6] = k > 0 ? R.out[0] / k : 0;
        infd[7] = R.out[1];
        inf[0] = 0; inf[1] = s.iters_run; inf[2] = s.best; inf[3] = all_inliers ? k : s.best_count;
        for (int i = 0; i < PNP_MP; ++i) inf[4 + i] = smp[i];
        inf[9] = k; inf[10] = sh.iters; inf[11] = s.pad;
    }
    cluster.sync();  // no CTA may exit while a peer can still read its shared memory
}

// cv2.solvePnPRefineLM on the caller's fp64 points (all of them are used).  One cluster per problem.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_refine_lm_p(const double* __restrict__ obj, const double* __restrict__ img, int n, const double* __restrict__ K4g,
              int max_iters, double* __restrict__ pose_io, int* __restrict__ iters_out) {
    __shared__ PnpLsqShared sh;
    __shared__ ClusterRed R;
    __shared__ JacobiWarp9 jw;
    __shared__ double K4[4];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned csize = cluster.num_blocks(), crank = cluster.block_rank();
    const int tid = threadIdx.x, gtid = crank * THREADS + tid, gstride = csize * THREADS;
    if (tid == 0) R.phase = 0;
    if (tid < 6) sh.p[tid] = pose_io[tid];
    if (tid < 4) K4[tid] = K4g[tid];
    __syncthreads();
    PnpPts ps;
    ps.px = nullptr; ps.mask = nullptr; ps.obj = obj; ps.img = img; ps.n = n;
    pnp_refine_lm<THREADS>(R, sh, jw, ps, K4, gtid, gstride, max_iters);
    __syncthreads();
    if (crank == 0 && tid < 6) pose_io[tid] = sh.p[tid];
    if (crank == 0 && tid == 0) iters_out[0] = sh.iters;
    cluster.sync();
}

// ---- inlier index list ---------------------------------------------------------------------------------------------
// solvePnPRansac returns the RANSAC-stage inliers as an ascending int32 column.  One CTA per problem: each thread
// owns a contiguous chunk, block-wide exclusive scan of the chunk counts, ordered write.
__global__ void __launch_bounds__(1024) k_compact_inliers(const uint8_t* __restrict__ mask, int n, int* __restrict__ inliers,
                                                          int* __restrict__ n_inliers) {
    __shared__ int warp_tot[32];
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t* M = mask + (size_t)q * n;
    int* out = inliers + (size_t)q * n;
    const int chunk = (n + 1023) / 1024;
    const int b = min(n, tid * chunk), e = min(n, b + chunk);
    int c = 0;
    for (int i = b; i < e; ++i) c += M[i] != 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = warp_tot[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += y;
        }
        warp_tot[lane] = wi - w;  // exclusive
        if (lane == 31) n_inliers[q] = wi;
    }
    __syncthreads();
    int pos = warp_tot[warp] + incl - c;
    for (int i = b; i < e; ++i)
        if (M[i]) out[pos++] = i;
}

}  // namespace b2r

// K3 for the PnP path — hypotheses x points reprojection-and-threshold kernels for 3x4 (pose) models.
//
// Replaces the H x N loop of OpenCV's PnPRansacCallback::computeError + findInliers that
// cv2.solvePnPRansac(pos3d, pixels, K, dist, 5000, 30.0, 0.99) runs per RANSAC iteration (reference call sites:
// main_v1.py:497-502, testpro.py:536, test_pro.py:515, testpro-K.py:72-75; semantics in SURVEY.md A.8):
//     cv::projectPoints(obj_fp32, rvec, tvec, K, 0)   fp64 arithmetic on the fp32-quantised, un-centred object points,
//                                                     result rounded to fp32
//     err = (u - pu)^2 + (v - pv)^2                   fp32, un-fused;   inlier <=> err <= (float)(thr*thr)
//
// EXACT kernel: that sequence operation for operation (fp64 pipe; the whole TU is compiled with -fmad=false, the
//   reciprocal is the IEEE division 1.0/z).  Inlier counts and index sets are bit-exact with the CPU path.
// FAST kernel: all-fp32, P = K [R | t'] pre-multiplied in fp64 by the solver and applied to object points re-centred
//   in fp64 (an all-fp32 projection of raw UTM coordinates would be wrong by > 1 px, SURVEY.md finding 7); packed
//   FFMA2 arithmetic, two hypotheses per instruction, 13 FMA-pipe operations + 1 MUFU.RCP per hypothesis·point.
//
// Both kernels use the K3 mapping of score_h.cuh: a thread owns its hypotheses in registers for the whole kernel, a
// CTA stages one tile of points in shared memory with one 1-D TMA bulk copy, all warps walk the tile with broadcast
// shared-memory loads, counts stay in registers, one RED.ADD per hypothesis per tile.
#pragma once
#include "score_h.cuh"

namespace b2r {

// exact path: the fp32-quantised object point widened to fp64 (what projectPoints reads) + the fp32 pixel
struct __align__(32) PointPX {
    double X, Y, Z;
    float u, v;
};

// fast path: object point minus the problem's centre (subtracted in fp64, then rounded), minus the pixel
struct __align__(32) PointPF {
    float Xc, Yc, Zc, nu;
    float nv, pad0, pad1, pad2;
};

// one exact evaluation: true when the point is an inlier of the pose (R, t) under intrinsics (fx, fy, cx, cy)
__device__ __forceinline__ bool p_inlier_exact(const double* R, const double* t, double fx, double fy, double cx, double cy,
                                               double X, double Y, double Z, float u, float v, float thr) {
    double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
    double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
    double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    z = z != 0 ? 1. / z : 1;
    x *= z;
    y *= z;
    const float pu = (float)(x * fx + cx), pv = (float)(y * fy + cy);
    const float dx = __fsub_rn(u, pu), dy = __fsub_rn(v, pv);
    const float e = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    return e <= thr;
}

constexpr int K3P_THREADS = 256;

// models : [Q][H_stride][12] fp64 = R (row-major 9) | t (3); NaN rows = no model; the first H of each problem are scored
// pts    : [N] PointPX (pts_q_stride = 0: the Q problems share the points, e.g. the intrinsics grid of testpro-K.py) or [Q][N]
// Kq     : [Q][4] fp64 = fx, fy, cx, cy
// counts : [Q][H] int32, zeroed by the caller
// grid   : x = ceil(H / (K3P_THREADS*NH)), y = ceil(N / tile_pts), z = Q; dynamic smem = 128 + tile_pts*32
template <int NH>
__global__ void __launch_bounds__(K3P_THREADS)
k3_score_p_exact(const double* __restrict__ models, int H, int H_stride, const PointPX* __restrict__ pts, size_t pts_q_stride,
                 int N, const double* __restrict__ Kq, float thr, int* __restrict__ counts, int tile_pts) {
    models += (size_t)blockIdx.z * H_stride * 12;
    pts += (size_t)blockIdx.z * pts_q_stride;
    counts += (size_t)blockIdx.z * H_stride;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    const PointPX* tile = reinterpret_cast<const PointPX*>(smem_raw + 128);

    const int p_begin = blockIdx.y * tile_pts;
    const int np = min(tile_pts, N - p_begin);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)np * 32u);
        tma_load_1d(smem_raw + 128, pts + p_begin, (uint32_t)np * 32u, bar);
    }
    const double fx = Kq[blockIdx.z * 4], fy = Kq[blockIdx.z * 4 + 1], cx = Kq[blockIdx.z * 4 + 2], cy = Kq[blockIdx.z * 4 + 3];
    const int h_base = blockIdx.x * (K3P_THREADS * NH) + threadIdx.x;
    double m[NH][12];
#pragma unroll
    for (int j = 0; j < NH; ++j) {
        const int hh = h_base + j * K3P_THREADS;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double2 v = make_double2(__longlong_as_double(0x7ff8000000000000ll), __longlong_as_double(0x7ff8000000000000ll));
            if (hh < H) v = __ldg(reinterpret_cast<const double2*>(models + (size_t)hh * 12) + i);
            m[j][2 * i] = v.x;
            m[j][2 * i + 1] = v.y;
        }
    }
    int cnt[NH];
#pragma unroll
    for (int j = 0; j < NH; ++j) cnt[j] = 0;
    mbar_wait(bar, 0);
    for (int p = 0; p < np; ++p) {
        const double2 xy = *reinterpret_cast<const double2*>(&tile[p].X);
        const double Z = tile[p].Z;
        const float2 uv = *reinterpret_cast<const float2*>(&tile[p].u);
#pragma unroll
        for (int j = 0; j < NH; ++j)
            cnt[j] += p_inlier_exact(m[j], m[j] + 9, fx, fy, cx, cy, xy.x, xy.y, Z, uv.x, uv.y, thr) ? 1 : 0;
    }
#pragma unroll
    for (int j = 0; j < NH; ++j) {
        const int hh = h_base + j * K3P_THREADS;
        if (hh < H && cnt[j]) atomicAdd(counts + hh, cnt[j]);
    }
}

#ifndef K3P_MIN_CTAS
#define K3P_MIN_CTAS 2
#endif
#ifndef K3P_FAST_UNROLL
#define K3P_FAST_UNROLL 2
#endif
#ifndef K3P_FAST_FORM
#define K3P_FAST_FORM 3
#endif

// models : [Q][H][12] fp32 rows of P = K [R | R c + t] (row-major 3x4), 48-byte rows; NaN rows = no model
// pts    : PointPF, same sharing rule as above
template <int NPAIR>
__global__ void __launch_bounds__(K3_THREADS, K3P_MIN_CTAS)
k3_score_p_fast(const float4* __restrict__ models, int H, int H_stride, const PointPF* __restrict__ pts, size_t pts_q_stride,
                int N, float thr, int* __restrict__ counts, int tile_pts) {
    models += (size_t)blockIdx.z * H_stride * 3;
    pts += (size_t)blockIdx.z * pts_q_stride;
    counts += (size_t)blockIdx.z * H_stride;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    PointPF* tile = reinterpret_cast<PointPF*>(smem_raw + 128);

    const int p_begin = blockIdx.y * tile_pts;
    const int np = min(tile_pts, N - p_begin);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)np * 32u);
        tma_load_1d(smem_raw + 128, pts + p_begin, (uint32_t)np * 32u, bar);
    }
    const int h_base = blockIdx.x * (K3_THREADS * 2 * NPAIR) + threadIdx.x;
    f2_t c[NPAIR][12];
#pragma unroll
    for (int j = 0; j < NPAIR; ++j) {
        const int ha = h_base + (2 * j) * K3_THREADS, hb = ha + K3_THREADS;
        const float q = __int_as_float(0x7fc00000);
        float4 a[3], b[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            a[r] = b[r] = make_float4(q, q, q, q);
            if (ha < H) a[r] = __ldg(models + 3 * (size_t)ha + r);
            if (hb < H) b[r] = __ldg(models + 3 * (size_t)hb + r);
            c[j][4 * r] = f2_pack(a[r].x, b[r].x);
            c[j][4 * r + 1] = f2_pack(a[r].y, b[r].y);
            c[j][4 * r + 2] = f2_pack(a[r].z, b[r].z);
            c[j][4 * r + 3] = f2_pack(a[r].w, b[r].w);
        }
    }
    int cnt[2 * NPAIR];
#pragma unroll
    for (int j = 0; j < 2 * NPAIR; ++j) cnt[j] = 0;
    // Signed margins, as in score_h.cuh: negative <=> inlier, the sign bit is counted (one LEA.HI per evaluation), a NaN
    // model gives the canonical NaN (sign clear): outlier.  thr_up is the float above thr, so "< thr_up" is "<= thr".
    //   FORM 1: (x/z - u)^2 + (y/z - v)^2 - thr_up with MUFU.RCP: 13 FMA-pipe ops + 1 MUFU.
    //   FORM 3: division-free, rows 0 and 1 of P and -u, -v pre-scaled by s = thr_up^-1/2 (hypotheses at load, the
    //           tile once per CTA): (s x - s u z)^2 + (s y - s v z)^2 - z^2: 14 FMA-pipe ops.
    const float thr_up = thr < __int_as_float(0x7f800000) ? __uint_as_float(__float_as_uint(thr) + 1u) : thr;
    const f2_t nthr = f2_dup(-thr_up);
    const float s = rsqrtf(fmaxf(thr_up, 1e-30f));  // thr = 0 would scale by inf; 1e-30 px^2 is "exactly on the pixel" at fp32 accuracy
    if (K3P_FAST_FORM == 3) {
        const f2_t s2 = f2_dup(s);
#pragma unroll
        for (int j = 0; j < NPAIR; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) c[j][k] = f2_mul(c[j][k], s2);
    }
    mbar_wait(bar, 0);
    if (K3P_FAST_FORM == 3) {
        for (int p = threadIdx.x; p < np; p += K3_THREADS) {
            tile[p].nu *= s;
            tile[p].nv *= s;
        }
        __syncthreads();
    }

    constexpr int UNROLL = K3P_FAST_UNROLL;
#pragma unroll UNROLL
    for (int p = 0; p < np; ++p) {
        const float4 pt = *reinterpret_cast<const float4*>(&tile[p].Xc);  // broadcast LDS.128
        const float pnv = tile[p].nv;
        const f2_t X = f2_dup(pt.x), Y = f2_dup(pt.y), Z = f2_dup(pt.z), nu = f2_dup(pt.w), nv = f2_dup(pnv);
#pragma unroll
        for (int j = 0; j < NPAIR; ++j) {
            const f2_t x = f2_fma(c[j][0], X, f2_fma(c[j][1], Y, f2_fma(c[j][2], Z, c[j][3])));
            const f2_t y = f2_fma(c[j][4], X, f2_fma(c[j][5], Y, f2_fma(c[j][6], Z, c[j][7])));
            const f2_t z = f2_fma(c[j][8], X, f2_fma(c[j][9], Y, f2_fma(c[j][10], Z, c[j][11])));
            float e0, e1;
            if (K3P_FAST_FORM == 3) {
                const f2_t a = f2_fma(z, nu, x), b = f2_fma(z, nv, y);
                float t0, t1;
                f2_unpack(f2_mul(z, z), t0, t1);
                f2_unpack(f2_fma(a, a, f2_fma(b, b, f2_pack(-t0, -t1))), e0, e1);  // the negation folds into the FFMA2 operand
            } else {
                float z0, z1;
                f2_unpack(z, z0, z1);
                const f2_t iz = f2_pack(rcp_approx(z0), rcp_approx(z1));
                const f2_t dx = f2_fma(x, iz, nu), dy = f2_fma(y, iz, nv);
                if (K3P_FAST_FORM == 0) {
                    f2_unpack(f2_fma(dx, dx, f2_mul(dy, dy)), e0, e1);
                    cnt[2 * j] += (e0 <= thr) ? 1 : 0;
                    cnt[2 * j + 1] += (e1 <= thr) ? 1 : 0;
                    continue;
                }
                f2_unpack(f2_fma(dx, dx, f2_fma(dy, dy, nthr)), e0, e1);
            }
            cnt[2 * j] += (int)(__float_as_uint(e0) >> 31);
            cnt[2 * j + 1] += (int)(__float_as_uint(e1) >> 31);
        }
    }
#pragma unroll
    for (int j = 0; j < 2 * NPAIR; ++j) {
        const int hh = h_base + j * K3_THREADS;
        if (hh < H && cnt[j]) atomicAdd(counts + hh, cnt[j]);
    }
}

}  // namespace b2r

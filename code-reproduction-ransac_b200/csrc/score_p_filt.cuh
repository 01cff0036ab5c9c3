// K3 for the PnP path, bit-exact counts through a FILTERED predicate (3x4 models) — the scheme of score_h_filt.cuh.
//
// k3_score_p_exact runs cv::projectPoints' fp64 projection + the fp32 error for every hypothesis x point (SURVEY.md A.8;
// reference call sites main_v1.py:497-502, testpro-K.py:72-75): ~32 fp64 instructions per evaluation.  This kernel takes
// the sign of the division-free fp32 margin of k3_score_p_fast
//     m = (s x' - s u z')^2 + (s y' - s v z')^2 - z'^2,   (x', y', z') = P (X - c; 1),  P = K [R | R c + t],  s = thr^-1/2
// wherever |m| exceeds a PROVED bound B on everything that can separate it from OpenCV's comparison, and decides the
// other evaluations with OpenCV's sequence (p_inlier_exact) after the loop.  Same counts as k3_score_p_exact.
//
// The bound (derivation: DESIGN.md, "Filtered exact predicate").  u = 2^-24.  With the tile's bounding box Xm_j = max
// |X_j - c_j| (j = 0..2), Um = max |u|, Vm = max |v| and, per hypothesis,
//     Ax = sum_j |P0j| Xm_j + |P03|,  Ay (row 1),  Aw = sum_j |P2j| Xm_j + |P23|      (Aw >= |z'| on the tile)
//     Gx = sum_j |R0j| (|c_j| + Xm_j) + |t0|,  Gy, Gz: magnitudes of OpenCV's own fp64 sums on the un-centred points
//     E64 = 2^-50 ((fx Gx + (|cx| + Um) Gz) + (fy Gy + (|cy| + Vm) Gz))     fp64 rounding, OpenCV's and this kernel's P
//     D = s (7.2u (Ax + Ay) + 10.5u (Um + Vm) Aw + 4.2 E64) + 19.2u Aw + 2^-50 Gz
//     B = 2 Aw D + D^2
//     m >  B and |z'| > zeta  =>  OpenCV's err > thr (or NaN): outlier           m < -B  =>  OpenCV's err <= thr: inlier
// zeta = 7.1u Aw + 2^-50 Gz keeps OpenCV's `z != 0 ? 1/z : 1` on its first branch: below it the evaluation is undecided.
// The hypothesis is scaled by (2.002/B)^1/2, so "|m| >= 2.0" is bit 30 of the float; the depth guard is folded in as
// min(m, 2 z'^2 / zeta^2) (one FMUL2 per pair, one FMNMX per evaluation); one funnel shift per evaluation files sign and
// band bit.  Guards: thr in [2^-40, 2^40], the tile's coordinates finite and <= 2^40 (else the CTA runs OpenCV's sequence
// on the whole tile); K and the hypothesis' magnitudes <= 2^40 (else that hypothesis alone is taken through the tile by a
// warp after the loop).
#pragma once
#include "score_p.cuh"
#include "score_h_filt.cuh"

namespace b2r {

#ifndef K3PF_MIN_CTAS
#define K3PF_MIN_CTAS 2
#endif
#ifndef K3PF_UNROLL
#define K3PF_UNROLL 4
#endif
constexpr int K3PF_POINT_UNROLL = K3PF_UNROLL;

// dynamic shared memory of k3_score_p_filt<NPAIR> for a tile of tile_pts points
inline size_t k3p_filt_smem(int tile_pts, int npair) {
    return 512 + (size_t)tile_pts * 64 + 8 * (size_t)K3F_SLOTS * K3_THREADS + 4 * (size_t)npair * K3_THREADS;
}

// OpenCV's sequence for hypothesis hh on the points flagged in `um` (bit 2i: the point i before `newest`): inlier count
static __device__ __noinline__ int k3p_filt_resolve(const double* __restrict__ models, int H, int hh, const PointPX* tile, uint32_t um, int newest,
                                             double fx, double fy, double cx, double cy, float thr) {
    if (hh >= H) return 0;
    double m[12];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(models + (size_t)hh * 12) + i);
        m[2 * i] = v.x;
        m[2 * i + 1] = v.y;
    }
    int c = 0;
    while (um) {
        const int b = 31 - __clz(um);
        um &= ~(1u << b);
        const PointPX& pt = tile[newest - (b >> 1)];
        c += p_inlier_exact(m, m + 9, fx, fy, cx, cy, pt.X, pt.Y, pt.Z, pt.u, pt.v, thr) ? 1 : 0;
    }
    return c;
}

// models : [Q][H_stride][12] fp64 = R (row-major 9) | t (3), as k3_score_p_exact reads them
// px, pf : the two point records of the problem (PointPX for OpenCV's sequence, PointPF for the margin), same sharing rule
// Kq     : [Q][4] fp64;  centre : [.][3] fp64, centre_q_stride = 0 when the problems share the points
// grid   : x = ceil(H / (K3_THREADS*2*NPAIR)), y = ceil(N / tile_pts), z = Q; dynamic smem = k3p_filt_smem(tile_pts, NPAIR)
template <int NPAIR>
__global__ void __launch_bounds__(K3_THREADS, K3PF_MIN_CTAS)
k3_score_p_filt(const double* __restrict__ models, int H, int H_stride, const PointPX* __restrict__ px, const PointPF* __restrict__ pf,
                size_t pts_q_stride, int N, const double* __restrict__ Kq, const double* __restrict__ centre, size_t centre_q_stride,
                float thr, int* __restrict__ counts, int tile_pts) {
    models += (size_t)blockIdx.z * H_stride * 12;
    px += (size_t)blockIdx.z * pts_q_stride;
    pf += (size_t)blockIdx.z * pts_q_stride;
    counts += (size_t)blockIdx.z * H_stride;
    centre += (size_t)blockIdx.z * centre_q_stride;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* g_count = reinterpret_cast<int*>(smem_raw + 48) + warp;
    float* red = reinterpret_cast<float*>(smem_raw + 80);   // 8 floats (5 maxima) per warp: bytes [80, 80 + 32*NWARP) of the 512-byte header
    const PointPX* tile_x = reinterpret_cast<const PointPX*>(smem_raw + 512);
    PointPF* tile_f = reinterpret_cast<PointPF*>(smem_raw + 512 + (size_t)tile_pts * 32);
    uint16_t* g_items = reinterpret_cast<uint16_t*>(smem_raw + 512 + (size_t)tile_pts * 64 + 8 * (size_t)K3F_SLOTS * K3_THREADS) + warp * (2 * NPAIR * 32);

    const int p_begin = blockIdx.y * tile_pts;
    const int np = min(tile_pts, N - p_begin);
    if (lane == 0) *g_count = 0;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)np * 64u);
        tma_load_1d(smem_raw + 512, px + p_begin, (uint32_t)np * 32u, bar);
        tma_load_1d(smem_raw + 512 + (size_t)tile_pts * 32, pf + p_begin, (uint32_t)np * 32u, bar);
    }
    const double fx = Kq[blockIdx.z * 4], fy = Kq[blockIdx.z * 4 + 1], cx = Kq[blockIdx.z * 4 + 2], cy = Kq[blockIdx.z * 4 + 3];
    const double c0 = centre[0], c1 = centre[1], c2 = centre[2];
    const int h_cta = blockIdx.x * (K3_THREADS * 2 * NPAIR), h_base = h_cta + threadIdx.x;
    const bool thr_ok = thr >= 0x1p-40f && thr <= 0x1p40f;
    const float s = rsqrtf(thr_ok ? thr : 1.0f);
    const bool k_ok = fabs(fx) <= 0x1p40 && fabs(fy) <= 0x1p40 && fabs(cx) <= 0x1p40 && fabs(cy) <= 0x1p40 && fabs(c0) <= 0x1p40 &&
                      fabs(c1) <= 0x1p40 && fabs(c2) <= 0x1p40;

    mbar_wait(bar, 0);
    // bounding box of the tile (re-centred coordinates, pixels), the scaled -u, -v, and "every value is finite"
    float mx0 = 0.f, mx1 = 0.f, mx2 = 0.f, mu = 0.f, mv = 0.f, nonfinite = 0.f;
    for (int p = threadIdx.x; p < np; p += K3_THREADS) {
        PointPF pt = tile_f[p];
        mx0 = fmaxf(mx0, fabsf(pt.Xc)); mx1 = fmaxf(mx1, fabsf(pt.Yc)); mx2 = fmaxf(mx2, fabsf(pt.Zc));
        mu = fmaxf(mu, fabsf(pt.nu)); mv = fmaxf(mv, fabsf(pt.nv));
        nonfinite += (pt.Xc - pt.Xc) + (pt.Yc - pt.Yc) + (pt.Zc - pt.Zc) + (pt.nu - pt.nu) + (pt.nv - pt.nv);
        tile_f[p].nu = pt.nu * s;
        tile_f[p].nv = pt.nv * s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
        mx2 = fmaxf(mx2, __shfl_xor_sync(0xffffffffu, mx2, o)); mu = fmaxf(mu, __shfl_xor_sync(0xffffffffu, mu, o));
        mv = fmaxf(mv, __shfl_xor_sync(0xffffffffu, mv, o));
    }
    if (lane == 0) {
        reinterpret_cast<float4*>(red)[2 * warp] = make_float4(mx0, mx1, mx2, mu);
        red[8 * warp + 4] = mv;
    }
    bool guard_tile = __syncthreads_or(!thr_ok || !k_ok || !(nonfinite == 0.f)) != 0;
#pragma unroll
    for (int wi = 0; wi < K3_THREADS / 32; ++wi) {
        const float4 r = reinterpret_cast<const float4*>(red)[2 * wi];
        mx0 = fmaxf(mx0, r.x); mx1 = fmaxf(mx1, r.y); mx2 = fmaxf(mx2, r.z); mu = fmaxf(mu, r.w);
        mv = fmaxf(mv, red[8 * wi + 4]);
    }
    guard_tile = guard_tile || !(mx0 + mx1 + mx2 + mu + mv <= 0x1p40f);
    int cnt[2 * NPAIR];
#pragma unroll
    for (int j = 0; j < 2 * NPAIR; ++j) cnt[j] = 0;
    if (guard_tile) {
        // outside the guards of the bound (threshold, intrinsics or coordinates; the same answer for every thread of the
        // CTA): OpenCV's sequence for the whole tile, as k3_score_p_exact runs it
#pragma unroll 1
        for (int j = 0; j < 2 * NPAIR; ++j) {
            const int hh = h_base + j * K3_THREADS;
            if (hh >= H) continue;
            double m[12];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const double2 v = __ldg(reinterpret_cast<const double2*>(models + (size_t)hh * 12) + i);
                m[2 * i] = v.x;
                m[2 * i + 1] = v.y;
            }
            int c = 0;
#pragma unroll 1
            for (int p = 0; p < np; ++p) {
                const PointPX& pt = tile_x[p];
                c += p_inlier_exact(m, m + 9, fx, fy, cx, cy, pt.X, pt.Y, pt.Z, pt.u, pt.v, thr) ? 1 : 0;
            }
            if (c) atomicAdd(counts + hh, c);
        }
        return;
    }

    // this thread's hypotheses: P = K [R | R c + t] in fp64, the bound, the scaled fp32 rows
    f2_t h[NPAIR][13];   // rows 0, 1 scaled by kappa s, row 2 by kappa; [12] = 1.99 / (kappa zeta)^2, the depth guard's factor
    {
        const float uu = 0x1p-24f, up = 1.0f + 0x1p-18f;
        const double e50 = 0x1p-50;
        const float a0 = (float)(fabs(c0) * (1 + 1e-6)) + mx0, a1 = (float)(fabs(c1) * (1 + 1e-6)) + mx1, a2 = (float)(fabs(c2) * (1 + 1e-6)) + mx2;
        float row[2][12], gfac[2];
#pragma unroll
        for (int j = 0; j < NPAIR; ++j) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int hh = h_base + (2 * j + q) * K3_THREADS;
                double m[12];
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    double2 v = make_double2(__longlong_as_double(0x7ff8000000000000ll), __longlong_as_double(0x7ff8000000000000ll));
                    if (hh < H) v = __ldg(reinterpret_cast<const double2*>(models + (size_t)hh * 12) + i);
                    m[2 * i] = v.x;
                    m[2 * i + 1] = v.y;
                }
                double tc[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) tc[i] = m[3 * i] * c0 + m[3 * i + 1] * c1 + m[3 * i + 2] * c2 + m[9 + i];
                float P[12];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    P[k] = (float)(fx * m[k] + cx * m[6 + k]);
                    P[4 + k] = (float)(fy * m[3 + k] + cy * m[6 + k]);
                    P[8 + k] = (float)m[6 + k];
                }
                P[3] = (float)(fx * tc[0] + cx * tc[2]);
                P[7] = (float)(fy * tc[1] + cy * tc[2]);
                P[11] = (float)tc[2];
                // magnitudes (fp32, inflated)
                const float Ax = (fabsf(P[0]) * mx0 + fabsf(P[1]) * mx1 + fabsf(P[2]) * mx2 + fabsf(P[3])) * up;
                const float Ay = (fabsf(P[4]) * mx0 + fabsf(P[5]) * mx1 + fabsf(P[6]) * mx2 + fabsf(P[7])) * up;
                const float Aw = (fabsf(P[8]) * mx0 + fabsf(P[9]) * mx1 + fabsf(P[10]) * mx2 + fabsf(P[11])) * up;
                const float Gx = ((float)fabs(m[0]) * a0 + (float)fabs(m[1]) * a1 + (float)fabs(m[2]) * a2 + (float)fabs(m[9])) * up;
                const float Gy = ((float)fabs(m[3]) * a0 + (float)fabs(m[4]) * a1 + (float)fabs(m[5]) * a2 + (float)fabs(m[10])) * up;
                const float Gz = ((float)fabs(m[6]) * a0 + (float)fabs(m[7]) * a1 + (float)fabs(m[8]) * a2 + (float)fabs(m[11])) * up;
                const float afx = (float)fabs(fx) * up, afy = (float)fabs(fy) * up, acx = (float)fabs(cx) * up + mu, acy = (float)fabs(cy) * up + mv;
                const float E64 = (float)e50 * ((afx * Gx + acx * Gz) + (afy * Gy + acy * Gz)) * up;
                const float yw = (float)e50 * Gz * up;
                const float D = (s * (7.2f * uu * (Ax + Ay) + 10.5f * uu * (mu + mv) * Aw + 4.2f * E64) * up + 19.2f * uu * Aw + yw) * up + 0x1p-60f;
                const float B = (2.0f * (Aw * up) * D + D * D) * (1.0f + 0x1p-9f);
                float kappa = rsqrtf(B) * 1.4150f;   // kappa^2 B = 2.002: |scaled margin| >= 2.0 (bit 30 of the float) means |margin| > B
                const float zeta = (7.1f * uu * Aw + yw) * up + 0x1p-60f;
                if ((Ax > 0x1p40f) || (Ay > 0x1p40f) || (Aw > 0x1p40f) || (Gx > 0x1p40f) || (Gy > 0x1p40f) || (Gz > 0x1p40f)) {
                    // a hypothesis outside the guards: its margins are turned into NaN ("decided, outlier": nothing is
                    // counted in the loop) and a warp takes it through the whole tile afterwards
                    kappa = __int_as_float(0x7fc00000);
                    g_items[atomicAdd(g_count, 1)] = (uint16_t)((lane << 4) | (2 * j + q));
                }
                const float ks = kappa * s;
#pragma unroll
                for (int k = 0; k < 8; ++k) row[q][k] = P[k] * ks;
#pragma unroll
                for (int k = 8; k < 12; ++k) row[q][k] = P[k] * kappa;
                const float kz = kappa * zeta;
                gfac[q] = 1.99f / (kz * kz);   // t'' gfac >= 2.0  =>  kappa^2 z'^2 >= 1.005 (kappa zeta)^2
            }
#pragma unroll
            for (int k = 0; k < 12; ++k) h[j][k] = f2_pack(row[0][k], row[1][k]);
            h[j][12] = f2_pack(gfac[0], gfac[1]);
        }
    }
    const uint32_t q_s = smem_u32(smem_raw + 512 + (size_t)tile_pts * 64) + threadIdx.x * (8u * K3F_SLOTS);
    int qn = 0;

    uint32_t sg[2 * NPAIR];   // two bits per point and hypothesis, as in k3_score_h_filt
#pragma unroll
    for (int j = 0; j < 2 * NPAIR; ++j) sg[j] = 0;
    for (int p0 = 0; p0 < np; p0 += K3F_BATCH) {
        const int nb = min(K3F_BATCH, np - p0);
        auto eval_point = [&](int p) {
            const float4 pt = *reinterpret_cast<const float4*>(&tile_f[p].Xc);  // broadcast LDS.128
            const float pnv = tile_f[p].nv;
            const f2_t X = f2_dup(pt.x), Y = f2_dup(pt.y), Z = f2_dup(pt.z), nu = f2_dup(pt.w), nv = f2_dup(pnv);
#pragma unroll
            for (int j = 0; j < NPAIR; ++j) {
                const f2_t x = f2_fma(h[j][0], X, f2_fma(h[j][1], Y, f2_fma(h[j][2], Z, h[j][3])));
                const f2_t y = f2_fma(h[j][4], X, f2_fma(h[j][5], Y, f2_fma(h[j][6], Z, h[j][7])));
                const f2_t z = f2_fma(h[j][8], X, f2_fma(h[j][9], Y, f2_fma(h[j][10], Z, h[j][11])));
                const f2_t a = f2_fma(z, nu, x), b = f2_fma(z, nv, y);
                const f2_t t = f2_mul(z, z);
                float t0, t1, e0, e1, g0, g1;
                f2_unpack(t, t0, t1);
                f2_unpack(f2_fma(a, a, f2_fma(b, b, f2_pack(-t0, -t1))), e0, e1);
                f2_unpack(f2_mul(t, h[j][12]), g0, g1);
                // the depth guard: a margin >= 2.0 is kept only where z'^2 clears zeta^2 (fminf skips a NaN operand: a NaN
                // margin must stay NaN, "decided, outlier" — its depth term is NaN as well, so it does)
                e0 = fminf(e0, g0);
                e1 = fminf(e1, g1);
                sg[2 * j] = __funnelshift_l(__float_as_uint(e0), sg[2 * j], 2);
                sg[2 * j + 1] = __funnelshift_l(__float_as_uint(e1), sg[2 * j + 1], 2);
            }
        };
        if (nb == K3F_BATCH) {
#pragma unroll K3PF_POINT_UNROLL
            for (int q = 0; q < K3F_BATCH; ++q) eval_point(p0 + q);
        } else {
#pragma unroll 1
            for (int q = 0; q < nb; ++q) eval_point(p0 + q);
        }
        const uint32_t decided = 0x55555555u >> (32 - 2 * nb);
        uint32_t all = sg[0];
#pragma unroll
        for (int j = 1; j < 2 * NPAIR; ++j) all &= sg[j];
        if (__builtin_expect((all & decided) != decided, 0)) {
#pragma unroll
            for (int j = 0; j < 2 * NPAIR; ++j) {
                const uint32_t um = ~sg[j] & decided;
                if (um) {
                    sg[j] &= ~(um << 1);
                    const uint32_t id = ((uint32_t)j << 16) | (uint32_t)(p0 + nb - 1);
                    if (qn < K3F_SLOTS) {
                        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(q_s + 8u * qn), "r"(um), "r"(id) : "memory");
                        ++qn;
                    } else {   // slots full: in line
                        cnt[j] += k3p_filt_resolve(models, H, h_base + j * K3_THREADS, tile_x, um, p0 + nb - 1, fx, fy, cx, cy, thr);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 2 * NPAIR; ++j) cnt[j] += __popc(sg[j] & (decided << 1));
    }

#pragma unroll
    for (int j = 0; j < 2 * NPAIR; ++j) {
        const int hh = h_base + j * K3_THREADS;
        if (hh < H && cnt[j]) atomicAdd(counts + hh, cnt[j]);
    }
    // this thread's deferred entries
    for (int e = 0; e < qn; ++e) {
        uint32_t um, id;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(um), "=r"(id) : "r"(q_s + 8u * e) : "memory");
        const int hh = h_base + (int)(id >> 16) * K3_THREADS;
        const int c = k3p_filt_resolve(models, H, hh, tile_x, um, (int)(id & 0xffffu), fx, fy, cx, cy, thr);
        if (c) atomicAdd(counts + hh, c);
    }
    // this warp's guarded hypotheses: the lanes stride over the tile
    __syncwarp();
    const int n_guarded = *g_count;
    for (int e = 0; e < n_guarded; ++e) {
        const uint32_t it = g_items[e];
        const int hh = h_cta + warp * 32 + (int)(it >> 4) + (int)(it & 15u) * K3_THREADS;
        if (hh >= H) continue;   // the same for the whole warp
        double m[12];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(models + (size_t)hh * 12) + i);
            m[2 * i] = v.x;
            m[2 * i + 1] = v.y;
        }
        int c = 0;
        for (int p = lane; p < np; p += 32) {
            const PointPX& pt = tile_x[p];
            c += p_inlier_exact(m, m + 9, fx, fy, cx, cy, pt.X, pt.Y, pt.Z, pt.u, pt.v, thr) ? 1 : 0;
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0 && c) atomicAdd(counts + hh, c);
    }
}

}  // namespace b2r

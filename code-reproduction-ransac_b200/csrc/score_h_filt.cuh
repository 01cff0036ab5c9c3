// K3, bit-exact counts through a FILTERED predicate (3x3 models).
//
// k3_score_h<., EXACT> runs OpenCV's un-fused fp32 sequence (reference call sites main_v1.py:312,
// process.py:200, testpro.py:350) for every hypothesis x point: 21 packed operations + a reciprocal.  Almost every one of
// those evaluations is nowhere near the threshold.  This kernel evaluates the division-free FMA margin of the FAST mode
// first (11 FFMA2 per hypothesis pair and point) and accepts its sign only when |margin| exceeds a PROVED bound on
// everything that can separate it from OpenCV's own comparison; otherwise the batch of points is redone with the un-fused
// sequence (HEval<true>::err).  Counts are therefore the same integers as k3_score_h<., true> returns — the tests hold
// them to equality — at close to the FAST mode's cost.
//
// The bound (derivation with all constants: DESIGN.md, "Filtered exact predicate").  u = 2^-24.  For the tile's bounding
// box Xm = max|X|, Ym = max|Y|, Um = max|u|, Vm = max|v| and a hypothesis h:
//     Ax = |h0| Xm + |h1| Ym + |h2|,  Ay = |h3| Xm + |h4| Ym + |h5|,  Aw = |h6| Xm + |h7| Ym + 1   (Aw >= |w| on the tile)
// Real quantities  sx, sy, w,  a = sx - u w,  b = sy - v w.  OpenCV's computed values satisfy
//     |w_cv (qx_cv - u)| within 5.02u Ax + 3.01u Um Aw of |a|,   |w_cv| within 3.01u Aw of |w|   (claims O and I of DESIGN.md)
// and the FMA margin's operands carry errors of the same kind.  With r = |(s a, s b)|, s = thr^-1/2, both comparisons
// are statements about r against |w|, and everything that separates them adds up to
//     D = s (9.5u (Ax + Ay) + 8.5u (Um + Vm) Aw) + 16u Aw
//     B = 2 Aw D + D^2          (from sqrt(P^2 + B) >= P + D for every P <= Aw)
//     margin >  B  =>  OpenCV's err > thr (or NaN): outlier          margin < -B  =>  OpenCV's err <= thr: inlier
// B is a constant per hypothesis and tile, so the hypothesis (all nine coefficients: the margin is homogeneous of degree
// two) is scaled by (2.002/B)^1/2 at load and the test becomes |margin| >= 2.0, i.e. BIT 30 of the float, next to the
// sign in bit 31: one funnel shift per evaluation files both (the FAST kernel spends the same slot on counting the sign),
// and sixteen points later four AND/POPC pairs read them.  A NaN margin (0x7FFFFFFF) reads "decided, outlier".
// Evaluations inside the band are noted in shared memory — {mask of the batch's points, hypothesis} entries in slots
// private to the thread: no atomics — and decided after the loop with the un-fused sequence: no re-evaluation inside
// the loop and no CTA-wide barrier after it.
// Guards: thr in [2^-40, 2^40], the tile's coordinates finite and <= 2^40 (else the CTA runs the un-fused sequence on
// the whole tile); Ax, Ay, Aw <= 2^40 (else that hypothesis alone is taken through the tile by a warp after the loop).
#pragma once
#include "score_h.cuh"

namespace b2r {

#ifndef K3F_BATCH
#define K3F_BATCH 16   // points between two looks at the "inside the band" bits (<= 16: two bits per point and hypothesis)
#endif
#ifndef K3F_MIN_CTAS
#define K3F_MIN_CTAS 2   // measured (tools/microbench_filt.cu): 114 registers at 2 CTAs per SM beat 80 at 3
#endif
#ifndef K3F_UNROLL
#define K3F_UNROLL 8
#endif
constexpr int K3F_POINT_UNROLL = K3F_UNROLL;
constexpr int K3F_SLOTS = 8;      // deferred entries per thread and tile (8 bytes each); beyond them the owner evaluates in line

// dynamic shared memory of k3_score_h_filt<NPAIR> for a tile of tile_pts points
inline size_t k3_filt_smem(int tile_pts, int npair) {
    return 256 + (size_t)tile_pts * 32 + 8 * (size_t)K3F_SLOTS * K3_THREADS + 4 * (size_t)npair * K3_THREADS;
}

struct K3FiltConst {
    float kappa;   // B^-1/2: scale of every coefficient of the hypothesis
    bool force;    // outside the guards: always the un-fused sequence
};

// the per-hypothesis, per-tile scale (plain fp32 arithmetic, -fmad=false; every step is inflated by `up` so that rounding
// in these few operations can only widen the band)
__device__ __forceinline__ K3FiltConst k3_filter_const(const float (&h)[8], float Xm, float Ym, float Um, float Vm, float s) {
    const float up = 1.0f + 0x1p-18f, uu = 0x1p-24f, big = 0x1p40f;
    const float Ax = (fabsf(h[0]) * Xm + fabsf(h[1]) * Ym + fabsf(h[2])) * up;
    const float Ay = (fabsf(h[3]) * Xm + fabsf(h[4]) * Ym + fabsf(h[5])) * up;
    const float Aw = (fabsf(h[6]) * Xm + fabsf(h[7]) * Ym + 1.0f) * up;
    // D = x1 + y1 + xw + yw + 9.7u Aw: everything that separates the margin's operands from OpenCV's
    const float D = ((9.5f * uu * s) * (Ax + Ay) * up + ((8.5f * uu * s) * (Um + Vm) * up + 16.0f * uu) * Aw) * up + 0x1p-60f;
    const float B = (2.0f * (Aw * up) * D + D * D) * (1.0f + 0x1p-9f);
    K3FiltConst c;
    c.kappa = rsqrtf(B) * 1.4150f;   // kappa^2 B = 2.002: |scaled margin| >= 2.0 (bit 30 of the float) means |margin| > B
#ifdef K3F_PROBE_NOFLAG   // timing probe only (tools/microbench_filt.cu): no margin is ever inside the band
    c.kappa = 1e9f;
#endif
    c.force = (Ax > big) || (Ay > big) || (Aw > big);   // NaN coefficients pass: the whole hypothesis turns NaN and counts 0
    return c;
}

// OpenCV's un-fused sequence, scalar form: the same IEEE operations as HEval<true>::err (mul.rn / add.rn / correctly
// rounded reciprocal), one hypothesis against one point
__device__ __forceinline__ bool h_inlier_exact(const float4 a0, const float4 a1, const float4 pt, float thr) {
    const float w = __fadd_rn(__fadd_rn(__fmul_rn(a1.z, pt.x), __fmul_rn(a1.w, pt.y)), 1.0f);
    const float ww = rcp_rn(w);
    const float sx = __fadd_rn(__fadd_rn(__fmul_rn(a0.x, pt.x), __fmul_rn(a0.y, pt.y)), a0.z);
    const float sy = __fadd_rn(__fadd_rn(__fmul_rn(a0.w, pt.x), __fmul_rn(a1.x, pt.y)), a1.y);
    const float dx = __fadd_rn(__fmul_rn(sx, ww), pt.z), dy = __fadd_rn(__fmul_rn(sy, ww), pt.w);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) <= thr;
}

// the un-fused sequence for hypothesis hh on the points flagged in `um` (bit 2i: the point i before `newest`): inlier count
static __device__ __noinline__ int k3_filt_resolve(const float4* __restrict__ models, int H, int hh, const float4* tile, uint32_t um, int newest,
                                            float thr) {
    if (hh >= H) return 0;
    const float4 a0 = __ldg(models + 2 * hh), a1 = __ldg(models + 2 * hh + 1);
    int c = 0;
    while (um) {
        const int b = 31 - __clz(um);
        um &= ~(1u << b);
        c += h_inlier_exact(a0, a1, tile[newest - (b >> 1)], thr) ? 1 : 0;
    }
    return c;
}

// Same arguments, grid and result as k3_score_h<NPAIR, true>; dynamic smem = 256 + tile_pts*16 (the tile) + tile_pts*16
// (the same points with -u, -v scaled by thr^-1/2: the operands of the margin) + 8*K3F_SLOTS per thread
// (deferred entries) + 4*NPAIR*K3_THREADS (guarded hypotheses): k3_filt_smem().
template <int NPAIR>
__global__ void __launch_bounds__(K3_THREADS, K3F_MIN_CTAS)
k3_score_h_filt(const float4* __restrict__ models, int H, int H_stride, const PointH* __restrict__ pts, int N, float thr,
                int* __restrict__ counts, int tile_pts) {
    static_assert(K3F_BATCH == 16, "sixteen points fill the two-bits-per-point registers: nothing is reset between batches");
    models += (size_t)blockIdx.z * H_stride * 2;
    pts += (size_t)blockIdx.z * N;
    counts += (size_t)blockIdx.z * H_stride;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    constexpr int NWARP = K3_THREADS / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* g_count = reinterpret_cast<int*>(smem_raw + 48) + warp;    // every warp keeps its own list: no CTA barrier after the loop
    float* red = reinterpret_cast<float*>(smem_raw + 80);   // one float4 of maxima per warp: bytes [80, 80 + 16*NWARP) <= 256
    const float4* tile = reinterpret_cast<const float4*>(smem_raw + 256);
    float4* tile_s = reinterpret_cast<float4*>(smem_raw + 256 + (size_t)tile_pts * 16);   // {X, Y, -s u, -s v}
    // guarded hypotheses of this warp: (lane << 4) | slot, at most 2*NPAIR*32
    uint16_t* g_items = reinterpret_cast<uint16_t*>(smem_raw + 256 + (size_t)tile_pts * 32 + 8 * (size_t)K3F_SLOTS * K3_THREADS) + warp * (2 * NPAIR * 32);

    const int p_begin = blockIdx.y * tile_pts;
    const int np = min(tile_pts, N - p_begin);
    if (lane == 0) {
        *g_count = 0;
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)np * 16u);
        tma_load_1d(smem_raw + 256, pts + p_begin, (uint32_t)np * 16u, bar);
    }

    // while the bulk copy is in flight: this thread's hypotheses, unscaled
    const int h_cta = blockIdx.x * (K3_THREADS * 2 * NPAIR), h_base = h_cta + threadIdx.x;
    float hr[2 * NPAIR][8];
#pragma unroll
    for (int j = 0; j < 2 * NPAIR; ++j) {
        const int hh = h_base + j * K3_THREADS;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        if (hh < H) { a0 = __ldg(models + 2 * hh); a1 = __ldg(models + 2 * hh + 1); }
        hr[j][0] = a0.x; hr[j][1] = a0.y; hr[j][2] = a0.z; hr[j][3] = a0.w;
        hr[j][4] = a1.x; hr[j][5] = a1.y; hr[j][6] = a1.z; hr[j][7] = a1.w;
    }
    const bool thr_ok = thr >= 0x1p-40f && thr <= 0x1p40f;
    const float s = rsqrtf(thr_ok ? thr : 1.0f);

    mbar_wait(bar, 0);
    // bounding box of the tile, the scaled -u, -v, and "every coordinate is finite"
    float mx = 0.f, my = 0.f, mu = 0.f, mv = 0.f, nonfinite = 0.f;
    for (int p = threadIdx.x; p < np; p += K3_THREADS) {
        const float4 pt = tile[p];
        mx = fmaxf(mx, fabsf(pt.x)); my = fmaxf(my, fabsf(pt.y)); mu = fmaxf(mu, fabsf(pt.z)); mv = fmaxf(mv, fabsf(pt.w));
        nonfinite += (pt.x - pt.x) + (pt.y - pt.y) + (pt.z - pt.z) + (pt.w - pt.w);   // +0 for finite values, NaN for inf / NaN
        tile_s[p] = make_float4(pt.x, pt.y, pt.z * s, pt.w * s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); my = fmaxf(my, __shfl_xor_sync(0xffffffffu, my, o));
        mu = fmaxf(mu, __shfl_xor_sync(0xffffffffu, mu, o)); mv = fmaxf(mv, __shfl_xor_sync(0xffffffffu, mv, o));
    }
    if (lane == 0) reinterpret_cast<float4*>(red)[warp] = make_float4(mx, my, mu, mv);
    // (fmaxf skips NaN: a non-finite coordinate is caught by the sum above)
    bool guard_tile = __syncthreads_or(!thr_ok || !(nonfinite == 0.f)) != 0;
#pragma unroll
    for (int wi = 0; wi < K3_THREADS / 32; ++wi) {
        const float4 r = reinterpret_cast<const float4*>(red)[wi];
        mx = fmaxf(mx, r.x); my = fmaxf(my, r.y); mu = fmaxf(mu, r.z); mv = fmaxf(mv, r.w);
    }
    guard_tile = guard_tile || !(mx + my + mu + mv <= 0x1p40f);
    int cnt[2 * NPAIR];
#pragma unroll
    for (int j = 0; j < 2 * NPAIR; ++j) cnt[j] = 0;
    if (guard_tile) {
        // outside the guards of the bound (threshold or coordinates; the same answer for every thread of the CTA): the
        // un-fused sequence for the whole tile, as k3_score_h<NPAIR, true> runs it
        f2_t hx[NPAIR][8];
#pragma unroll
        for (int j = 0; j < NPAIR; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) hx[j][k] = f2_pack(hr[2 * j][k], hr[2 * j + 1][k]);
        const f2_t one = f2_dup(1.0f);
#pragma unroll 1
        for (int p = 0; p < np; ++p) {
            const float4 pt = tile[p];
#pragma unroll
            for (int j = 0; j < NPAIR; ++j) {
                float e0, e1;
                f2_unpack(HEval<true>::err(hx[j], f2_dup(pt.x), f2_dup(pt.y), f2_dup(pt.z), f2_dup(pt.w), one), e0, e1);
                cnt[2 * j] += (e0 <= thr) ? 1 : 0;
                cnt[2 * j + 1] += (e1 <= thr) ? 1 : 0;
            }
        }
#pragma unroll
        for (int j = 0; j < 2 * NPAIR; ++j) {
            const int hh = h_base + j * K3_THREADS;
            if (hh < H && cnt[j]) atomicAdd(counts + hh, cnt[j]);
        }
        return;
    }

    f2_t h[NPAIR][9];   // scaled: h0..h5 by kappa s, h6, h7 and the constant 1 by kappa
#pragma unroll
    for (int j = 0; j < NPAIR; ++j) {
        K3FiltConst cc[2] = {k3_filter_const(hr[2 * j], mx, my, mu, mv, s), k3_filter_const(hr[2 * j + 1], mx, my, mu, mv, s)};
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (cc[q].force) {
                // a hypothesis outside the guards (coefficients beyond 2^40): its margins are turned into NaN — "decided,
                // outlier", nothing counted in the loop — and a warp takes it through the whole tile afterwards
                cc[q].kappa = __int_as_float(0x7fc00000);
                g_items[atomicAdd(g_count, 1)] = (uint16_t)((lane << 4) | (2 * j + q));
            }
        const float ka = cc[0].kappa, kb = cc[1].kappa, ksa = ka * s, ksb = kb * s;
#pragma unroll
        for (int k = 0; k < 6; ++k) h[j][k] = f2_pack(hr[2 * j][k] * ksa, hr[2 * j + 1][k] * ksb);
        h[j][6] = f2_pack(hr[2 * j][6] * ka, hr[2 * j + 1][6] * kb);
        h[j][7] = f2_pack(hr[2 * j][7] * ka, hr[2 * j + 1][7] * kb);
        h[j][8] = f2_pack(ka, kb);
    }
    // deferred entries of this thread: K3F_SLOTS private slots of {mask of points, (hypothesis slot << 16) | newest point}
    const uint32_t q_s = smem_u32(smem_raw + 256 + (size_t)tile_pts * 32) + threadIdx.x * (8u * K3F_SLOTS);
    int qn = 0;
#ifdef K3F_STATS
    unsigned long long st_flagged = 0;
#endif

    // sg[j]: the top two bits of the margins of hypothesis j, two per point (the newest point in bits 1:0).  Bit 31 of a margin
    // is its sign (set: inlier); the scale puts the band at 2.0, so bit 30 — the top bit of the exponent — IS "|margin| >= 2.0:
    // outside the band" (a NaN margin, 0x7FFFFFFF, reads "decided, outlier").  One SHF per evaluation files both; nothing else
    // is done per evaluation.  Sixteen points fill the register, so the previous batch has left it by the time it is read.
    uint32_t sg[2 * NPAIR];
#pragma unroll
    for (int j = 0; j < 2 * NPAIR; ++j) sg[j] = 0;
    for (int p0 = 0; p0 < np; p0 += K3F_BATCH) {
        const int nb = min(K3F_BATCH, np - p0);
        auto eval_point = [&](int p) {
            const float4 pt = tile_s[p];   // broadcast LDS.128
            const f2_t X = f2_dup(pt.x), Y = f2_dup(pt.y), nu = f2_dup(pt.z), nv = f2_dup(pt.w);
#pragma unroll
            for (int j = 0; j < NPAIR; ++j) {
                const f2_t w = f2_fma(h[j][6], X, f2_fma(h[j][7], Y, h[j][8]));
                const f2_t sx = f2_fma(h[j][0], X, f2_fma(h[j][1], Y, h[j][2]));
                const f2_t sy = f2_fma(h[j][3], X, f2_fma(h[j][4], Y, h[j][5]));
                const f2_t a = f2_fma(w, nu, sx);
                const f2_t b = f2_fma(w, nv, sy);
                float t0, t1, e0, e1;
                f2_unpack(f2_mul(w, w), t0, t1);
                f2_unpack(f2_fma(a, a, f2_fma(b, b, f2_pack(-t0, -t1))), e0, e1);
                sg[2 * j] = __funnelshift_l(__float_as_uint(e0), sg[2 * j], 2);          // (sg << 2) | (margin >> 30)
                sg[2 * j + 1] = __funnelshift_l(__float_as_uint(e1), sg[2 * j + 1], 2);
            }
        };
        if (nb == K3F_BATCH) {
#pragma unroll K3F_POINT_UNROLL
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
            // rare: some evaluation lies inside the band.  The sign bits of those are dropped here; which points they are goes
            // into one of this thread's slots, and the un-fused sequence decides them after the loop
#pragma unroll
            for (int j = 0; j < 2 * NPAIR; ++j) {
                const uint32_t um = ~sg[j] & decided;
                if (um) {
                    sg[j] &= ~(um << 1);
                    const uint32_t id = ((uint32_t)j << 16) | (uint32_t)(p0 + nb - 1);
#ifdef K3F_STATS
                    st_flagged += __popc(um);
#endif
                    if (qn < K3F_SLOTS) {
                        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(q_s + 8u * qn), "r"(um), "r"(id) : "memory");
                        ++qn;
                    } else {   // slots full: in line
                        cnt[j] += k3_filt_resolve(models, H, h_base + j * K3_THREADS, tile, um, p0 + nb - 1, thr);
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
#ifdef K3F_STATS
    if (st_flagged) atomicAdd(&g_k3f_redo, st_flagged);
    if (threadIdx.x == 0) atomicAdd(&g_k3f_batches, (unsigned long long)np * K3_THREADS);
#endif

    // this thread's deferred entries: OpenCV's sequence on the flagged points of one hypothesis each
    for (int e = 0; e < qn; ++e) {
        uint32_t um, id;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(um), "=r"(id) : "r"(q_s + 8u * e) : "memory");
        const int hh = h_base + (int)(id >> 16) * K3_THREADS;
        const int c = k3_filt_resolve(models, H, hh, tile, um, (int)(id & 0xffffu), thr);
        if (c) atomicAdd(counts + hh, c);
    }
    // this warp's guarded hypotheses: the lanes stride over the tile
    __syncwarp();
    const int n_guarded = *g_count;
    for (int e = 0; e < n_guarded; ++e) {
        const uint32_t it = g_items[e];
        const int hh = h_cta + warp * 32 + (int)(it >> 4) + (int)(it & 15u) * K3_THREADS;
        if (hh >= H) continue;   // the same for the whole warp
        const float4 a0 = __ldg(models + 2 * hh), a1 = __ldg(models + 2 * hh + 1);
        int c = 0;
        for (int p = lane; p < np; p += 32) c += h_inlier_exact(a0, a1, tile[p], thr) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0 && c) atomicAdd(counts + hh, c);
    }
}

}  // namespace b2r

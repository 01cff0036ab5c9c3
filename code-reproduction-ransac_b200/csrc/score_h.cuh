// K3 — hypotheses x points reprojection-and-threshold kernel for 3x3 (homography) models.
//
// Replaces the H x N loop of OpenCV's HomographyEstimatorCallback::computeError + findInliers that
// cv2.findHomography(..., cv2.RANSAC, thr) runs per RANSAC iteration (reference call sites:
// main_v1.py:312, process.py:200, testpro.py:350; semantics restated in SURVEY.md A.5).
//
// Mapping: one thread owns 2*NPAIR hypotheses, held as NPAIR x 8 packed (fp32x2) coefficient
// registers for the whole kernel.  A CTA stages one tile of points in shared memory with a single
// 1-D TMA bulk copy (cp.async.bulk + mbarrier), then every warp walks the tile with broadcast
// LDS.128 loads: no cross-lane traffic, counts stay in registers, one RED.ADD per hypothesis per
// tile at the end.  A point enters every FFMA2/FMUL2/FADD2 through the scalar-broadcast operand form, so one
// instruction evaluates it against the two hypotheses of a pair.
//
// Arithmetic modes
//   EXACT: the reference's un-fused fp32 sequence, operation for operation (A.5):
//            ww = 1.f/((h6*X + h7*Y) + 1.f)
//            dx = ((h0*X + h1*Y) + h2)*ww - u ;  dy = ((h3*X + h4*Y) + h5)*ww - v
//            inlier  <=>  dx*dx + dy*dy <= thr     (NaN -> outlier)
//          every product/sum individually rounded (mul.rn/add.rn.f32x2 are never contracted),
//          reciprocal correctly rounded.  Bit-exact with cv2's mask by construction.
//   FAST:  the same inequality multiplied by w^2 (no division), FMA-contracted, threshold folded into pre-scaled
//          operands: 11 FMA-pipe ops per eval, the sign bit of the margin is the inlier flag (HEval::margin).
#pragma once
#include "f32x2.cuh"

namespace b2r {

// One correspondence.  u and v are stored NEGATED: a - b and a + (-b) are the same IEEE operation,
// and add/fma have no negate modifier in packed form.  A point multiplies two hypotheses at once
// through the scalar-broadcast operand form of FFMA2/FMUL2/FADD2 (SASS "Rn.F32"), which ptxas
// selects for f2_dup(x) and which costs one 32-bit register read instead of a 64-bit pair: the
// register file (2 x 32-bit reads/clk/SMSP, measured with tools/regprobe.cu) is what bounds this kernel.
struct __align__(16) PointH {
    float X, Y;    // source point (pos2 in the reference)
    float nu, nv;  // minus the destination pixel
};

constexpr int K3_THREADS = 256;
#ifndef K3_UNROLL
#define K3_UNROLL 8
#endif
constexpr int K3_POINT_UNROLL = K3_UNROLL;  // points per trip of the inner loop
#ifndef K3_EXACT_PACKED_NEWTON
#define K3_EXACT_PACKED_NEWTON 1
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// -- mbarrier / 1-D TMA helpers (SASS: SYNCS.*, UBLKCP) ---------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// a + b where a (and possibly b) is the result of a packed multiply: scalar adds, never contracted
__device__ __forceinline__ f2_t f2_sum_of_products(f2_t a, f2_t b) {
    float a0, a1, b0, b1;
    f2_unpack(a, a0, a1);
    f2_unpack(b, b0, b1);
    return f2_pack(__fadd_rn(a0, b0), __fadd_rn(a1, b1));
}

// c += (e <= thr) as FSETP + one predicated IADD3 (the compiler's own form is FSETP + two IADD3; the ALU pipe shares
// dispatch with the FMA pipe, which is what the exact kernel is bound by).  NaN compares false: outlier.
__device__ __forceinline__ void count_le(int& c, float e, float thr) {
    asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(c) : "f"(e), "f"(thr));
}

template <bool EXACT>
struct HEval {
    // Returns the packed squared reprojection error of one point against two hypotheses.
    __device__ __forceinline__ static f2_t err(const f2_t (&h)[8], f2_t X, f2_t Y, f2_t nu, f2_t nv, f2_t one) {
        if (EXACT) {
            // ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false (the scalar
            // mul.rn.f32 / add.rn.f32 pair is never contracted), which breaks bit-exactness.  Every product is therefore
            // written as fma(a, b, +0): RN(a*b + 0) = RN(a*b) for every a, b (only the sign of an exact zero product can
            // differ, which cannot reach the comparison: see below), the addend is the zero register (no register-file
            // read), and ptxas does not contract an add into an FMA that already has an addend.  All sums are packed adds.
            // Sign of zero: a product that is exactly -0 becomes +0; adding it to any non-zero value, squaring it, or
            // taking 1/(.. + 1) gives the same bits either way, and dx = +-0 gives dx*dx = +0 in both cases.
            // tests/test_gpu_parity_h.py holds the counts to the un-fused CPU sequence.
            const f2_t zero = f2_dup(0.0f);
            const f2_t w = f2_add(f2_add(f2_fma(h[6], X, zero), f2_fma(h[7], Y, zero)), one);
            float w0, w1;
            f2_unpack(w, w0, w1);
            f2_t ww;
            if (__builtin_expect(rcp_rn_fast_path_ok(w0) && rcp_rn_fast_path_ok(w1), 1)) {
                // one Newton step on the MUFU seed: y + y*(1 - w*y), correctly rounded in the fast range
                const float y0 = rcp_approx(w0), y1 = rcp_approx(w1);
                const float e0 = __fmaf_rn(-w0, y0, 1.0f), e1 = __fmaf_rn(-w1, y1, 1.0f);
                ww = f2_pack(__fmaf_rn(y0, e0, y0), __fmaf_rn(y1, e1, y1));
            } else {
                ww = f2_pack(__frcp_rn(w0), __frcp_rn(w1));
            }
            const f2_t sx = f2_add(f2_add(f2_fma(h[0], X, zero), f2_fma(h[1], Y, zero)), h[2]);
            const f2_t sy = f2_add(f2_add(f2_fma(h[3], X, zero), f2_fma(h[4], Y, zero)), h[5]);
            const f2_t dx = f2_add(f2_fma(sx, ww, zero), nu);
            const f2_t dy = f2_add(f2_fma(sy, ww, zero), nv);
            return f2_add(f2_fma(dx, dx, zero), f2_fma(dy, dy, zero));
        } else {
            f2_t w = f2_fma(h[6], X, f2_fma(h[7], Y, one));
            float w0, w1;
            f2_unpack(w, w0, w1);
            f2_t ww = f2_pack(rcp_approx(w0), rcp_approx(w1));
            f2_t sx = f2_fma(h[0], X, f2_fma(h[1], Y, h[2]));
            f2_t sy = f2_fma(h[3], X, f2_fma(h[4], Y, h[5]));
            f2_t dx = f2_fma(sx, ww, nu);
            f2_t dy = f2_fma(sy, ww, nv);
            return f2_fma(dx, dx, f2_mul(dy, dy));
        }
    }
    // EXACT, branch-free: as err(), with the reciprocal's MUFU + Newton form taken unconditionally; `ok` is cleared when
    // a denominator lies outside the range in which that form is the correctly rounded 1/w (the caller then redoes the
    // whole batch of points with err()).  No branch per evaluation: a batch of points is one basic block to schedule.
    __device__ __forceinline__ static f2_t err_exact_in_range(const f2_t (&h)[8], f2_t X, f2_t Y, f2_t nu, f2_t nv, f2_t one,
                                                              bool& ok) {
        const f2_t zero = f2_dup(0.0f);
        const f2_t w = f2_add(f2_add(f2_fma(h[6], X, zero), f2_fma(h[7], Y, zero)), one);
        float w0, w1;
        f2_unpack(w, w0, w1);
        // rcp_rn_fast_path_ok() as two unordered compares per half chained into the running predicate (FSETP.GEU/.LTU
        // ... .AND): |w| in [2^-100, 2^101) or NaN
        ok = ok && !(fabsf(w0) < 0x1p-100f) && !(fabsf(w0) >= 0x1p101f) && !(fabsf(w1) < 0x1p-100f) && !(fabsf(w1) >= 0x1p101f);
#if K3_EXACT_PACKED_NEWTON
        // the Newton step as two FFMA2 (the same two IEEE FMAs per half): scalar FFMA between packed instructions costs
        // ~3 cycles per switch on sm_100 (profiles/r01_pipeprobe.jsonl: ffma2+ffma 0.31 instr/clk)
        const f2_t y = f2_pack(rcp_approx(w0), rcp_approx(w1));
        const f2_t e = f2_fma(f2_pack(-w0, -w1), y, one);
        const f2_t ww = f2_fma(y, e, y);
#else
        const float y0 = rcp_approx(w0), y1 = rcp_approx(w1);
        const float e0 = __fmaf_rn(-w0, y0, 1.0f), e1 = __fmaf_rn(-w1, y1, 1.0f);
        const f2_t ww = f2_pack(__fmaf_rn(y0, e0, y0), __fmaf_rn(y1, e1, y1));
#endif
        const f2_t sx = f2_add(f2_add(f2_fma(h[0], X, zero), f2_fma(h[1], Y, zero)), h[2]);
        const f2_t sy = f2_add(f2_add(f2_fma(h[3], X, zero), f2_fma(h[4], Y, zero)), h[5]);
        const f2_t dx = f2_add(f2_fma(sx, ww, zero), nu);
        const f2_t dy = f2_add(f2_fma(sy, ww, zero), nv);
        return f2_add(f2_fma(dx, dx, zero), f2_fma(dy, dy, zero));
    }
    // FAST only.  Signed margin of one point against two hypotheses: negative <=> inlier.  The caller counts sign bits
    // (one LEA.HI per evaluation instead of FSETP + two IADD3); a NaN is the canonical 0x7FFFFFFF, sign clear: outlier.
    //   FORM 1: err - thr' with MUFU.RCP, thr' = the float above thr (so "< thr'" is "<= thr"): 10 FMA-pipe ops + 1 MUFU.
    //   FORM 2: division-free, (sx - u w)^2 + (sy - v w)^2 - thr' w^2 (the same inequality times w^2): 12 FMA-pipe ops.
    //   FORM 3: as 2 with h0..h5 and -u, -v pre-scaled by s = thr'^-1/2 (hypotheses at load, the point tile once per CTA
    //           in shared memory): (s sx - s u w)^2 + (s sy - s v w)^2 - w^2: 11 FMA-pipe ops.
    // MUFU.RCP and FFMA2 do not overlap freely on sm_100 (profiles/r01_pipeprobe.jsonl: 4 FFMA2 + 1 MUFU take 12.4
    // cycles, not 8): the reciprocal costs more issue time than the one or two extra FFMA2 that replace it.
    // FORM 4: FORM 3's margin with the last operation as two scalar FFMA.SAT on the NEGATED margin: the result is the
    // inlier flag itself (1.0f / +0), which one FADD2 per pair accumulates.  10 FFMA2 + 2 FFMA.SAT + 1 FADD2 per pair
    // and point instead of 11 FFMA2 + 2 LEA.HI.
    __device__ __forceinline__ static f2_t inlier_flag_sat(const f2_t (&h)[8], f2_t X, f2_t Y, f2_t nu, f2_t nv, f2_t k) {
        const f2_t w = f2_fma(h[6], X, f2_fma(h[7], Y, k));
        const f2_t sx = f2_fma(h[0], X, f2_fma(h[1], Y, h[2]));
        const f2_t sy = f2_fma(h[3], X, f2_fma(h[4], Y, h[5]));
        const f2_t a = f2_fma(w, nu, sx);
        const f2_t b = f2_fma(w, nv, sy);
        const f2_t t = f2_mul(w, w);
        float t0, t1, a0, a1, q0, q1;
        f2_unpack(t, t0, t1);
        f2_unpack(f2_fma(b, b, f2_pack(-t0, -t1)), q0, q1);
        f2_unpack(a, a0, a1);
        return f2_pack(__saturatef(__fmaf_rn(-a0, a0, -q0)), __saturatef(__fmaf_rn(-a1, a1, -q1)));
    }
    template <int FORM>
    __device__ __forceinline__ static f2_t margin(const f2_t (&h)[8], f2_t X, f2_t Y, f2_t nu, f2_t nv, f2_t one, f2_t nthr) {
        const f2_t w = f2_fma(h[6], X, f2_fma(h[7], Y, one));
        const f2_t sx = f2_fma(h[0], X, f2_fma(h[1], Y, h[2]));
        const f2_t sy = f2_fma(h[3], X, f2_fma(h[4], Y, h[5]));
        if (FORM == 1) {
            float w0, w1;
            f2_unpack(w, w0, w1);
            const f2_t ww = f2_pack(rcp_approx(w0), rcp_approx(w1));
            const f2_t dx = f2_fma(sx, ww, nu);
            const f2_t dy = f2_fma(sy, ww, nv);
            return f2_fma(dx, dx, f2_fma(dy, dy, nthr));
        } else {
            const f2_t a = f2_fma(w, nu, sx);
            const f2_t b = f2_fma(w, nv, sy);
            const f2_t t = f2_mul(w, w);
            if (FORM == 2) return f2_fma(a, a, f2_fma(b, b, f2_mul(t, nthr)));
            float t0, t1;
            f2_unpack(t, t0, t1);
            return f2_fma(a, a, f2_fma(b, b, f2_pack(-t0, -t1)));  // the negation folds into the FFMA2 operand
        }
    }
};

#ifndef K3_EXACT_BATCHED
#define K3_EXACT_BATCHED 1
#endif
#ifndef K3_FAST_UNROLL
#define K3_FAST_UNROLL 4   // points per trip of the fast kernel's loop (measured: profiles/r01g_microbench_k3_forms.jsonl)
#endif
#ifndef K3_MIN_CTAS
#define K3_MIN_CTAS 2
#endif
#ifndef K3_FAST_FORM
#define K3_FAST_FORM 3
#endif

// models : [Q][H_stride][8] fp32 (h0..h7, h8 == 1 implied), 32-byte aligned rows; the first H of each problem are scored
// pts    : [Q][N] PointH
// counts : [Q][H_stride] int32, must be zeroed by the caller; each CTA adds its tile's inlier counts
// grid   : x = ceil(H / (K3_THREADS*2*NPAIR)), y = ceil(N / tile_pts), z = Q; dynamic smem = 128 + tile_pts*16
template <int NPAIR, bool EXACT>
__global__ void __launch_bounds__(K3_THREADS, K3_MIN_CTAS)
k3_score_h(const float4* __restrict__ models, int H, int H_stride, const PointH* __restrict__ pts, int N, float thr,
           int* __restrict__ counts, int tile_pts) {
    models += (size_t)blockIdx.z * H_stride * 2;
    pts += (size_t)blockIdx.z * N;
    counts += (size_t)blockIdx.z * H_stride;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    float4* tile = reinterpret_cast<float4*>(smem_raw + 128);

    const int p_begin = blockIdx.y * tile_pts;
    const int np = min(tile_pts, N - p_begin);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)np * 16u);
        tma_load_1d(smem_raw + 128, pts + p_begin, (uint32_t)np * 16u, bar);
    }

    // While the bulk copy is in flight: fetch this thread's hypotheses and pack them in pairs.
    const int h_base = blockIdx.x * (K3_THREADS * 2 * NPAIR) + threadIdx.x;
    f2_t h[NPAIR][8];
#pragma unroll
    for (int j = 0; j < NPAIR; ++j) {
        const int ha = h_base + (2 * j) * K3_THREADS, hb = ha + K3_THREADS;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
        if (ha < H) { a0 = __ldg(models + 2 * ha); a1 = __ldg(models + 2 * ha + 1); }
        if (hb < H) { b0 = __ldg(models + 2 * hb); b1 = __ldg(models + 2 * hb + 1); }
        h[j][0] = f2_pack(a0.x, b0.x); h[j][1] = f2_pack(a0.y, b0.y);
        h[j][2] = f2_pack(a0.z, b0.z); h[j][3] = f2_pack(a0.w, b0.w);
        h[j][4] = f2_pack(a1.x, b1.x); h[j][5] = f2_pack(a1.y, b1.y);
        h[j][6] = f2_pack(a1.z, b1.z); h[j][7] = f2_pack(a1.w, b1.w);
    }
    int cnt[2 * NPAIR];
#pragma unroll
    for (int j = 0; j < 2 * NPAIR; ++j) cnt[j] = 0;
    const f2_t one = f2_dup(1.0f);
    // the float above thr (a squared distance: >= 0), so that "margin < 0" is "err <= thr"; +inf stays +inf
    const float thr_up = thr < __int_as_float(0x7f800000) ? __uint_as_float(__float_as_uint(thr) + 1u) : thr;
    const f2_t nthr = f2_dup(-thr_up);
    constexpr bool SCALED = !EXACT && K3_FAST_FORM >= 3;
    constexpr bool SATCNT = !EXACT && K3_FAST_FORM == 4;  // inlier flag = FFMA.SAT of the margin, counted by packed float adds
    const float s = rsqrtf(fmaxf(thr_up, 1e-30f));  // thr = 0 would scale by inf; 1e-30 px^2 is "exactly on the pixel" at fp32 accuracy
    if (SCALED) {
        const f2_t s2 = f2_dup(s);
#pragma unroll
        for (int j = 0; j < NPAIR; ++j)
#pragma unroll
            for (int k = 0; k < 6; ++k) h[j][k] = f2_mul(h[j][k], s2);
    }
    // FORM 4: every coefficient (the implied h8 = 1 included) times 2^32, i.e. the margin times 2^64: sat(-margin) is 1.0f
    // for every margin below -2^-64 w^2 and +0 for every margin >= 0 (NaN -> +0)
    const f2_t one_k = SATCNT ? f2_dup(0x1p32f) : one;
    f2_t cntf[NPAIR];
    if (SATCNT) {
        const f2_t k2 = f2_dup(0x1p32f);
#pragma unroll
        for (int j = 0; j < NPAIR; ++j) {
            cntf[j] = f2_dup(0.0f);
#pragma unroll
            for (int k = 0; k < 8; ++k) h[j][k] = f2_mul(h[j][k], k2);
        }
    }

    mbar_wait(bar, 0);
    if (SCALED) {  // every CTA scales the -u, -v of its own copy of the tile once
        for (int p = threadIdx.x; p < np; p += K3_THREADS) {
            float4 pt = tile[p];
            pt.z *= s;
            pt.w *= s;
            tile[p] = pt;
        }
        __syncthreads();
    }

    constexpr int UNROLL = EXACT ? K3_POINT_UNROLL : K3_FAST_UNROLL;  // points per trip
    int p0 = 0;
    if (EXACT && K3_EXACT_BATCHED) {
        // batches of UNROLL points without a branch inside; the tail (< UNROLL points) goes through the loop below
        for (; p0 + UNROLL <= np; p0 += UNROLL) {
            int c[2 * NPAIR];
#pragma unroll
            for (int j = 0; j < 2 * NPAIR; ++j) c[j] = 0;
            bool ok = true;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const float4 pt = tile[p0 + u];
                const f2_t X = f2_dup(pt.x), Y = f2_dup(pt.y), nu = f2_dup(pt.z), nv = f2_dup(pt.w);
#pragma unroll
                for (int j = 0; j < NPAIR; ++j) {
                    float e0, e1;
                    f2_unpack(HEval<EXACT>::err_exact_in_range(h[j], X, Y, nu, nv, one, ok), e0, e1);
                    count_le(c[2 * j], e0, thr);
                    count_le(c[2 * j + 1], e1, thr);
                }
            }
            if (__builtin_expect(!ok, 0)) {  // some denominator was out of range: redo the batch with the general reciprocal
#pragma unroll
                for (int j = 0; j < 2 * NPAIR; ++j) c[j] = 0;
#pragma unroll 1
                for (int u = 0; u < UNROLL; ++u) {
                    const float4 pt = tile[p0 + u];
                    const f2_t X = f2_dup(pt.x), Y = f2_dup(pt.y), nu = f2_dup(pt.z), nv = f2_dup(pt.w);
#pragma unroll
                    for (int j = 0; j < NPAIR; ++j) {
                        float e0, e1;
                        f2_unpack(HEval<EXACT>::err(h[j], X, Y, nu, nv, one), e0, e1);
                        c[2 * j] += (e0 <= thr) ? 1 : 0;
                        c[2 * j + 1] += (e1 <= thr) ? 1 : 0;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 2 * NPAIR; ++j) cnt[j] += c[j];
        }
    }
    constexpr int TAIL_UNROLL = (EXACT && K3_EXACT_BATCHED) ? 1 : UNROLL;
#pragma unroll TAIL_UNROLL
    for (int p = p0; p < np; ++p) {
        const float4 pt = tile[p];  // broadcast LDS.128
        const f2_t X = f2_dup(pt.x), Y = f2_dup(pt.y), nu = f2_dup(pt.z), nv = f2_dup(pt.w);
#pragma unroll
        for (int j = 0; j < NPAIR; ++j) {
            float e0, e1;
            if (SATCNT) {
                cntf[j] = f2_add(cntf[j], HEval<EXACT>::inlier_flag_sat(h[j], X, Y, nu, nv, one_k));
            } else if (!EXACT && K3_FAST_FORM != 0) {
                f2_unpack(HEval<EXACT>::template margin<K3_FAST_FORM>(h[j], X, Y, nu, nv, one, nthr), e0, e1);
                cnt[2 * j] += (int)(__float_as_uint(e0) >> 31);
                cnt[2 * j + 1] += (int)(__float_as_uint(e1) >> 31);
            } else {
                f2_unpack(HEval<EXACT>::err(h[j], X, Y, nu, nv, one), e0, e1);
                cnt[2 * j] += (e0 <= thr) ? 1 : 0;
                cnt[2 * j + 1] += (e1 <= thr) ? 1 : 0;
            }
        }
    }

    if (SATCNT) {  // a tile holds far fewer than 2^24 points: the float sums are exact integers
#pragma unroll
        for (int j = 0; j < NPAIR; ++j) {
            float c0, c1;
            f2_unpack(cntf[j], c0, c1);
            cnt[2 * j] = __float2int_rn(c0);
            cnt[2 * j + 1] = __float2int_rn(c1);
        }
    }
#pragma unroll
    for (int j = 0; j < 2 * NPAIR; ++j) {
        const int hh = h_base + j * K3_THREADS;
        if (hh < H) atomicAdd(counts + hh, cnt[j]);
    }
}

}  // namespace b2r

// Kernels around K3 for the homography path: pack (A.1), K1 samplers, K2 batched solve, selection
// (A.6) and K4 finalize = RANSAC-stage mask + refit + LM(10) + OpenCV-4.13 mask (A.7).
// Reference call site for all of it: cv2.findHomography(..., cv2.RANSAC, thr), main_v1.py:312.
#pragma once
#include "common_k.cuh"
#include "score_h.cuh"

namespace b2r {

// ---- S0: pack ----------------------------------------------------------------------------------------
// (Q,n,2) fp64 src + ((n,2) | (Q,n,2)) fp64 dst  ->  (Q,n) PointH.  The fp64 -> fp32 conversion is OpenCV's
// input quantisation: everything downstream sees the fp32 values.
__global__ void k_pack_points_h(const double* __restrict__ src, const double* __restrict__ dst, int dst_shared, int Q,
                                int n, PointH* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)Q * n) return;
    const size_t j = dst_shared ? (i % n) : i;
    PointH p;
    p.X = (float)src[2 * i];
    p.Y = (float)src[2 * i + 1];
    p.nu = -(float)dst[2 * j];
    p.nv = -(float)dst[2 * j + 1];
    out[i] = p;
}

__global__ void k_pack_points_h_f32(const float* __restrict__ src, const float* __restrict__ dst, int n,
                                    PointH* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PointH p;
    p.X = src[2 * i];
    p.Y = src[2 * i + 1];
    p.nu = -dst[2 * i];
    p.nv = -dst[2 * i + 1];
    out[i] = p;
}

__device__ __forceinline__ void gather4(const PointH* __restrict__ pts, const int* idx, float* ms1, float* ms2) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(pts + idx[k]));
        ms1[2 * k] = v.x;
        ms1[2 * k + 1] = v.y;
        ms2[2 * k] = -v.z;
        ms2[2 * k + 1] = -v.w;
    }
}

__device__ __forceinline__ void store_model(float4* __restrict__ models, size_t slot, const double* H, bool ok) {
    float4 a, b;
    if (ok) {
        a = make_float4((float)H[0], (float)H[1], (float)H[2], (float)H[3]);
        b = make_float4((float)H[4], (float)H[5], (float)H[6], (float)H[7]);
    } else {
        const float q = __int_as_float(0x7fc00000);  // NaN model: every error is NaN, count stays 0
        a = make_float4(q, q, q, q);
        b = a;
    }
    models[2 * slot] = a;
    models[2 * slot + 1] = b;
}

// ---- K1 (Philox) + K2 fused: one thread per (problem, hypothesis) ----------------------------------------
// samples : [Q][H][4] int32 (all -1: no acceptable subset in PHILOX_MAX_ATTEMPTS attempts)
// SOLVE: false = sample only (the exact solver runs in k_solve_h4_smem), true = the closed-form solver in the same thread.
// A template, not a flag: with the exact solver compiled in as well the kernel carried its 1.4 KB stack frame and 166
// registers (12 resident warps per SM) through every launch.
template <bool SOLVE>
__global__ void __launch_bounds__(128)
k_philox_sample_solve_h(const PointH* __restrict__ pts, int n, int H, long long hyp_begin, uint64_t seed,
                        int* __restrict__ samples, float4* __restrict__ models) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y;
    if (g >= H) return;
    const PointH* P = pts + (size_t)q * n;
    const unsigned long long gid = (unsigned long long)(hyp_begin + g);
    int idx[4] = {-1, -1, -1, -1};
    float ms1[8], ms2[8];
    bool found = false;
    for (int attempt = 0; attempt < PHILOX_MAX_ATTEMPTS && !found; ++attempt) {
        const Philox4 r = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)attempt, (uint32_t)q,
                                        (uint32_t)seed, (uint32_t)(seed >> 32));
        distinct4(r, (uint32_t)n, idx);
        gather4(P, idx, ms1, ms2);
        found = h_check_subset4(ms1, ms2);
    }
    const size_t slot = (size_t)q * H + g;
    if (!found) idx[0] = idx[1] = idx[2] = idx[3] = -1;
    reinterpret_cast<int4*>(samples)[slot] = make_int4(idx[0], idx[1], idx[2], idx[3]);
    if (SOLVE) {
        double Hm[9];
        const bool ok = found && h_solve4_fast(ms1, ms2, Hm) > 0;
        store_model(models, slot, Hm, ok);
    }
}

// ---- K1 (replay of cv::RNG): one WARP per problem ---------------------------------------------------------------------
// Generates the subsets of iterations [begin, begin+len) (clipped to the problem's current iteration bound),
// continuing the RNG stream stored in the problem's state.  samples : [Q][H_stride][4]
// OpenCV's loop is sequential — duplicate draws and subsets rejected by checkSubset consume RNG outputs — but only the
// LABELLING of the stream depends on the data: attempt a is the a-th group of four distinct indices of the stream
// whatever checkSubset says, and iteration j takes the j-th attempt that passes.  So the warp works on windows of 128
// stream outputs:
//   1. every lane walks the 128 multiply-with-carry steps (one IMAD.WIDE each) and keeps its four outputs, reduced
//      modulo n in parallel (the modulo is the expensive part of a draw);
//   2. every lane walks the 128 indices through the distinct-index state machine (integer compares only) and keeps
//      the four indices of attempt number `lane` (at most 32 attempts per window);
//   3. the lanes run checkSubset (fp64 determinants, the expensive part of an attempt) on their attempts in parallel;
//   4. the pass mask is turned into iteration numbers in order, with getSubset's bound of CV_MAX_ATTEMPTS attempts per
//      iteration; the stream position is rewound to the end of the last attempt that was consumed.
// ~150 cycles per attempt instead of ~3000 for a single lane walking the serial code; identical samples.
// The points are staged in shared memory when they fit (n <= 2048).
constexpr int K1_SMEM_PTS = 2048;
constexpr int K1_WINDOW = 128;
__global__ void __launch_bounds__(32)
k_cv_sample_h(const PointH* __restrict__ pts, int n, int H_stride, int begin, int len,
              int* __restrict__ samples, RansacState* __restrict__ state, int Q) {
    __shared__ PointH sp[K1_SMEM_PTS];
    __shared__ __align__(16) int draws[K1_WINDOW];
    const int q = blockIdx.x, lane = threadIdx.x;
    if (q >= Q) return;
    RansacState st = state[q];
    if (st.done || begin >= st.niters || st.gen < begin) return;
    const PointH* P = pts + (size_t)q * n;
    if (n <= K1_SMEM_PTS) {
        for (int i = lane; i < n; i += 32) sp[i] = P[i];
        P = sp;
    }
    int4* S = reinterpret_cast<int4*>(samples + (size_t)q * H_stride * 4);
    constexpr uint32_t MWC_A = 4164903690u;
    uint64_t base = st.rng;                  // stream position at the start of the window
    const int end = min(begin + len, st.niters);
    int it = begin;
    int ci = 0, c0 = 0, c1 = 0, c2 = 0;      // the attempt under construction (carried across windows)
    int tries = 0;                           // attempts made for iteration `it` so far
    bool stop = false;
    while (!stop) {
        // 1. the window's 128 outputs
        uint64_t r = base;
        uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
        for (int w = 0; w < K1_WINDOW / 4; ++w) {
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
        // 2. group the indices into attempts of four distinct ones
        int n_att = 0, my_end = 0;
        int4 mine = make_int4(0, 0, 0, 0);
        for (int w = 0; w < K1_WINDOW / 4; ++w) {
            const int4 d4 = reinterpret_cast<const int4*>(draws)[w];
            const int dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {   // selects only: the chain through ci is a few cycles per index
                const int d = dd[u];
                const bool fresh = !((ci > 0 && d == c0) || (ci > 1 && d == c1) || (ci > 2 && d == c2));
                const bool complete = fresh && ci == 3;
                const bool take = complete && n_att == lane;
                mine = take ? make_int4(c0, c1, c2, d) : mine;
                my_end = take ? 4 * w + u + 1 : my_end;
                n_att += complete ? 1 : 0;
                c0 = (fresh && ci == 0) ? d : c0;
                c1 = (fresh && ci == 1) ? d : c1;
                c2 = (fresh && ci == 2) ? d : c2;
                ci = complete ? 0 : ci + (fresh ? 1 : 0);
            }
        }
        // 3. checkSubset, one attempt per lane
        bool pass = false;
        if (lane < n_att) {
            const int idx[4] = {mine.x, mine.y, mine.z, mine.w};
            float ms1[8], ms2[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {   // P may point to shared memory: generic loads
                const float4 v = *reinterpret_cast<const float4*>(P + idx[k]);
                ms1[2 * k] = v.x; ms1[2 * k + 1] = v.y; ms2[2 * k] = -v.z; ms2[2 * k + 1] = -v.w;
            }
            pass = h_check_subset4(ms1, ms2);
        }
        const unsigned passmask = __ballot_sync(0xffffffffu, pass);
        // 4. iterations in order
        int consumed = -1;   // index of the attempt at which the walk ends inside this window
        for (int a = 0; a < n_att; ++a) {
            ++tries;
            if ((passmask >> a) & 1u) {
                if (lane == a) S[it] = mine;
                ++it;
                tries = 0;
                if (it >= end) { consumed = a; break; }
            } else if (tries >= CV_MAX_ATTEMPTS) {   // getSubset gives up: the RANSAC loop ends at this iteration
                consumed = a;
                break;
            }
        }
        if (consumed >= 0) {
            const int steps = __shfl_sync(0xffffffffu, my_end, consumed);
            for (int j = 0; j < steps; ++j) base = (uint64_t)(uint32_t)base * MWC_A + (uint32_t)(base >> 32);
            stop = true;
        } else {
            base = r;
        }
    }
    if (lane == 0) {
        st.rng = base;
        st.gen = it;
        state[q] = st;
    }
}

// ---- K2: batched 4-point solves from stored samples -------------------------------------------------------
// Solves iterations [begin, begin+len) of every problem (those below state[q].gen when a state is given).
// H64 (optional): [Q][H_stride][9] fp64 models, ok (optional): flags, subset_ok (optional): checkSubset result
__global__ void __launch_bounds__(128)
k_solve_h4(const PointH* __restrict__ pts, int n, const int* __restrict__ samples, int H_stride, int begin, int len,
           const RansacState* __restrict__ state, float4* __restrict__ models, double* __restrict__ H64,
           uint8_t* __restrict__ ok_out, uint8_t* __restrict__ subset_ok, int fast_solver) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y;
    if (g >= len) return;
    const int it = begin + g;
    const size_t slot = (size_t)q * H_stride + it;
    const int4 s = reinterpret_cast<const int4*>(samples)[slot];
    const bool have = (state == nullptr || it < state[q].gen) && s.x >= 0;
    double Hm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool ok = false;
    if (have) {
        const int idx[4] = {s.x, s.y, s.z, s.w};
        float ms1[8], ms2[8];
        gather4(pts + (size_t)q * n, idx, ms1, ms2);
        if (subset_ok) subset_ok[slot] = h_check_subset4(ms1, ms2) ? 1 : 0;
        ok = (fast_solver ? h_solve4_fast(ms1, ms2, Hm) : h_solve4(ms1, ms2, Hm)) > 0;
    } else if (subset_ok) {
        subset_ok[slot] = 0;
    }
    if (models) store_model(models, slot, Hm, ok);
    if (H64)
        for (int i = 0; i < 9; ++i) H64[slot * 9 + i] = ok ? Hm[i] : 0.0;
    if (ok_out) ok_out[slot] = ok ? 1 : 0;
}

// K2, exact solver, thread per solve with the 9x9 matrices in SHARED memory ([element][thread] layout, conflict-free for
// any per-thread pivot sequence).  One warp per CTA: 32 x 126 doubles (packed upper triangle + eigenvectors) = 31.5 KB of
// dynamic shared memory, seven CTAs per SM; the decomposition is the divergence-free jacobi_eig_packed.
// Bit-identical to k_solve_h4; ~25x its throughput (its per-thread local arrays overflow the L1).
constexpr int K2S_THREADS = 32;
constexpr size_t K2S_SMEM = sizeof(double) * H4_WS_DOUBLES * K2S_THREADS;
__global__ void __launch_bounds__(K2S_THREADS)
k_solve_h4_smem(const PointH* __restrict__ pts, int n, const int* __restrict__ samples, int H_stride, int begin, int len,
                const RansacState* __restrict__ state, float4* __restrict__ models, double* __restrict__ H64,
                uint8_t* __restrict__ ok_out, uint8_t* __restrict__ subset_ok) {
    extern __shared__ __align__(16) double k2s_ws[];
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y;
    if (g >= len) return;
    const int it = begin + g;
    const size_t slot = (size_t)q * H_stride + it;
    const int4 s = reinterpret_cast<const int4*>(samples)[slot];
    const bool have = (state == nullptr || it < state[q].gen) && s.x >= 0;
    double Hm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool ok = false;
    if (have) {
        const int idx[4] = {s.x, s.y, s.z, s.w};
        float ms1[8], ms2[8];
        gather4(pts + (size_t)q * n, idx, ms1, ms2);
        if (subset_ok) subset_ok[slot] = h_check_subset4(ms1, ms2) ? 1 : 0;
        ok = h_solve4_strided<K2S_THREADS>(k2s_ws + threadIdx.x, ms1, ms2, Hm) > 0;
    } else if (subset_ok) {
        subset_ok[slot] = 0;
    }
    if (models) store_model(models, slot, Hm, ok);
    if (H64)
        for (int i = 0; i < 9; ++i) H64[slot * 9 + i] = ok ? Hm[i] : 0.0;
    if (ok_out) ok_out[slot] = ok ? 1 : 0;
}

// K2, exact solver, one WARP per 4-point solve (shared-memory matrices, warp-cooperative Jacobi): bit-identical to the
// thread-per-solve kernel above and several times faster, because 81 + 81 doubles per thread in local memory do not fit
// the L1 once a few warps are resident.  Same arguments as k_solve_h4.
__global__ void __launch_bounds__(256)
k_solve_h4_warp(const PointH* __restrict__ pts, int n, const int* __restrict__ samples, int H_stride, int begin, int len,
                const RansacState* __restrict__ state, float4* __restrict__ models, double* __restrict__ H64,
                uint8_t* __restrict__ ok_out, uint8_t* __restrict__ subset_ok) {
    __shared__ JacobiWarp9 jw[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * 8 + warp;
    const int q = blockIdx.y;
    if (g >= len) return;
    const int it = begin + g;
    const size_t slot = (size_t)q * H_stride + it;
    const int4 s = reinterpret_cast<const int4*>(samples)[slot];
    const bool have = (state == nullptr || it < state[q].gen) && s.x >= 0;
    double Hm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool ok = false, sub = false;
    if (have) {
        const int idx[4] = {s.x, s.y, s.z, s.w};
        float ms1[8], ms2[8];
        gather4(pts + (size_t)q * n, idx, ms1, ms2);
        if (subset_ok) sub = h_check_subset4(ms1, ms2);
        ok = h_solve4_warp(jw[warp], ms1, ms2, Hm) > 0;
    }
    if (lane == 0) {
        if (subset_ok) subset_ok[slot] = sub ? 1 : 0;
        if (models) store_model(models, slot, Hm, ok);
        if (H64)
            for (int i = 0; i < 9; ++i) H64[slot * 9 + i] = ok ? Hm[i] : 0.0;
        if (ok_out) ok_out[slot] = ok ? 1 : 0;
    }
}

// ---- K4 finalize -----------------------------------------------------------------------------------------------
// exact fp32 squared reprojection error of one point, scalar form
__device__ __forceinline__ float h_err_exact(const float* Hf, float X, float Y, float nu, float nv) {
    const float w = __fadd_rn(__fadd_rn(__fmul_rn(Hf[6], X), __fmul_rn(Hf[7], Y)), 1.f);
    const float ww = rcp_rn(w);
    const float dx = __fadd_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hf[0], X), __fmul_rn(Hf[1], Y)), Hf[2]), ww), nu);
    const float dy = __fadd_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hf[3], X), __fmul_rn(Hf[4], Y)), Hf[5]), ww), nv);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// Eigenvector of the smallest eigenvalue of the symmetric PSD 9x9 L^T L by shifted inverse iteration
// (fast-solver mode; the exact mode runs OpenCV's Jacobi).  One warp; LtL in shared memory (upper triangle filled); lane i
// keeps row i of the Cholesky factor and entry i of the iterate in registers (chol_regs_*); the result is written to
// out[0..9).  false -> caller falls back to Jacobi.
__device__ __forceinline__ bool smallest_eigvec9_warp(const double* LtL, double* out) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int i = lane < 9 ? lane : 8;
    double tr = 0, dmax = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) { tr += LtL[j * 10]; dmax = fmax(dmax, fabs(LtL[j * 10])); }
    const double mu = tr * 1e-13;
    double a[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) a[k] = (k <= i ? LtL[k * 9 + i] : LtL[i * 9 + k]) + (k == i ? mu : 0.);
    CholRegs<9> F;
    chol_regs_factor<9>(a, dmax + mu, F);
    if (!F.ok) return false;
    double b = 1. / 3.;
    for (int it = 0; it < 16; ++it) {
        const double x = chol_regs_solve<9>(F, b);
        double n2 = lane < 9 ? x * x : 0.;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(FULL, n2, o);
        const double xn = x * (1. / sqrt(n2));
        double diff = lane < 9 ? fabs(xn - b) : 0.;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) diff = fmax(diff, __shfl_xor_sync(FULL, diff, o));
        b = xn;
        if (it > 1 && diff < 1e-15) break;
    }
    if (lane < 9) out[lane] = b;
    __syncwarp();
    return true;
}

// Development probe (tools/prof_finalize_sections.py builds a copy of the library with -DB2R_FIN_PROFILE): clock64() deltas
// of thread 0 of CTA 0 accumulated per section of k_finalize_h.  Compiled out of the product.
#ifdef B2R_FIN_PROFILE
__device__ unsigned long long g_fin_clk[32];
#define FINCLK(sec)                                                                  \
    do {                                                                             \
        if (threadIdx.x == 0 && blockIdx.x == 0) {                                   \
            const long long now_ = clock64();                                        \
            atomicAdd(&g_fin_clk[sec], (unsigned long long)(now_ - fin_last_));      \
            fin_last_ = now_;                                                        \
        }                                                                            \
    } while (0)
#else
#define FINCLK(sec) do { } while (0)
#endif

struct HFinalizeShared {
    double H[9];          // current model (fp64)
    double x[9], xd[9];   // LM parameter vectors (all nine entries of H, as OpenCV 4.13 refines them)
    double A[81], v[9], D[9], d[9];
    double Ap[81], diag[9];
    double L[81], Lrinv[9];  // rows of the Cholesky factor of the regularised J^T J (lambda == 0 step) and 1 / its diagonal
    double Ac[81], vc[9];  // J^T J and J^T r at the trial point (become A, v when the step is accepted)
    double S, Sd, lambda, lc, rmax, nu;
    float Hf[8];
    float ms1[8], ms2[8];
    int flag, k, lm_iters, proceed, use_eig, need_diag;   // k: this CTA's share of the final inlier count
};

// K4.  One cluster per problem (blockIdx.x / cluster size = problem).
//   sel        : selection result; best < 0 -> no model
//   samples    : [Q][Hs][4] minimal samples (index sel.best)
//   rmask      : [Q][n] RANSAC-stage mask (output)
//   H_out      : [Q][9], mask_out : [Q][n], info : [Q] (b2r_h_info layout = 12 int32)
//   ext_mask/ext_H : refine-only entry (b2r_refine_h): caller-supplied inlier mask and initial model
//   seq        : 1 = sums in OpenCV's order (problems of n <= THREADS points, one CTA each, exact solver): the centroid and
//                scale sums, L^T L, J^T J, J^T r and |r|^2 are accumulated point by point in index order by one thread per
//                entry, with the products OpenCV forms, and every LM step is solved through the eigen-decomposition as
//                cv::solve(DECOMP_EIG) does — the refined H is then bit-identical to the CPU restatement, so the
//                ill-conditioned problems of the reference's size (12-28 points, where the early-stopped LM turns a last-bit
//                difference into 1e-3) cannot drift.  Needs n * 20 doubles of dynamic shared memory.
//   models     : (optional) [Q][Hs] fp32 models as scored by K3; when given, the RANSAC-stage mask is taken with the stored
//                winner and its fp64 form (one more 9x9 decomposition) is only recomputed if it is the returned model
// GRID = true: ONE problem on a cooperative grid of gridDim.x CTAs (sized by the points, launch_finalize), reductions through gscratch + grid barriers
// instead of distributed shared memory — a cluster is limited to 8 SMs, which made the passes over the points the bulk
// of the finalize time of a large single problem.
template <int THREADS, bool GRID = false>
__global__ void __launch_bounds__(THREADS, THREADS <= 64 ? 8 : (THREADS <= 128 ? 4 : 1))   // small problems come in batches: keep several CTAs per SM resident
k_finalize_h(const PointH* __restrict__ pts, int n, const int* __restrict__ samples, int Hs,
             const HSelect* __restrict__ sel, float thr_sq, int mask_semantics, int refine, int fast_solver,
             double* __restrict__ H_out, uint8_t* __restrict__ mask_out, uint8_t* __restrict__ rmask_out,
             int* __restrict__ info, const uint8_t* __restrict__ ext_mask, const double* __restrict__ ext_H,
             double* __restrict__ gscratch, const float4* __restrict__ models, int seq) {
    __shared__ HFinalizeShared sh;
    __shared__ ClusterRed R;
    __shared__ JacobiWarp9 jw;   // workspace of the warp-cooperative eigen-solver (warp 0)
    extern __shared__ __align__(16) double seq_tab[];   // seq mode: [n][20] per-point rows (L or J: x row 9 | y row 9 | rx | ry)
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned csize = GRID ? gridDim.x : cluster.num_blocks(), crank = GRID ? blockIdx.x : cluster.block_rank();
#define TEAM_REDUCE(NV, NMAX, arr)                                                        \
    do {                                                                                  \
        if constexpr (GRID) grid_reduce_tail_max<THREADS, NV, NMAX>(R, arr, gscratch);    \
        else cluster_reduce_tail_max<THREADS, NV, NMAX>(R, arr);                          \
    } while (0)
    const int q = blockIdx.x / csize, tid = threadIdx.x;
    const int gtid = crank * THREADS + tid, gstride = csize * THREADS;
    const bool writer = crank == 0;  // one CTA of the cluster writes the small outputs
    const PointH* P = pts + (size_t)q * n;
    uint8_t* rmask = rmask_out + (size_t)q * n;
    uint8_t* mask = mask_out + (size_t)q * n;
    const HSelect s = sel[q];
    int* inf = info + (size_t)q * 12;
#ifdef B2R_FIN_PROFILE
    long long fin_last_ = clock64();
#endif
    if (tid == 0) R.phase = 0;

    if (s.best < 0) {  // no model: cv2 returns (None, zeros)
        for (int i = gtid; i < n; i += gstride) { mask[i] = 0; rmask[i] = 0; }
        if (writer && tid < 9) H_out[(size_t)q * 9 + tid] = 0;
        if (writer && tid == 0) {
            inf[0] = 1; inf[1] = s.iters_run; inf[2] = -1; inf[3] = 0;
            inf[4] = inf[5] = inf[6] = inf[7] = -1; inf[8] = 0; inf[9] = 0; inf[10] = 0; inf[11] = 0;
        }
        return;
    }
    const int4 smp = ext_mask ? make_int4(-1, -1, -1, -1) : reinterpret_cast<const int4*>(samples)[(size_t)q * Hs + s.best];
    const bool stored_model = models != nullptr && !ext_mask;
    auto minimal_model64 = [&]() {   // warp 0: the winning minimal model in fp64 (every lane computes the same values)
        if (tid < 32) {
            double Hm[9];
            if (ext_mask) {
                for (int i = 0; i < 9; ++i) Hm[i] = ext_H[(size_t)q * 9 + i];
            } else {
                const int idx[4] = {smp.x, smp.y, smp.z, smp.w};
                float ms1[8], ms2[8];
                gather4(P, idx, ms1, ms2);
                if (fast_solver) h_solve4_fast(ms1, ms2, Hm); else h_solve4_warp(jw, ms1, ms2, Hm);
            }
            if (tid == 0) {
                for (int i = 0; i < 9; ++i) sh.H[i] = Hm[i];
                if (!stored_model)
                    for (int i = 0; i < 8; ++i) sh.Hf[i] = (float)Hm[i];
            }
        }
        __syncthreads();
    };
    if (tid == 0) {
        sh.lm_iters = 0;
        sh.k = 0;                  // this CTA's share of the final inlier count
        if (writer) inf[8] = 0;    // ... added atomically at the end (every other CTA's add comes after a team barrier)
    }
    if (stored_model) {
        if (tid < 8) sh.Hf[tid] = reinterpret_cast<const float*>(models + 2 * ((size_t)q * Hs + s.best))[tid];
        __syncthreads();
    } else {
        minimal_model64();
    }

    // RANSAC-stage mask with the winning minimal model
    // ... and, in the same pass and the same reduction, the coordinate sums of the inliers (the centroids of the refit)
    double kc[5] = {0, 0, 0, 0, 0};
    for (int i = gtid; i < n; i += gstride) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(P + i));
        uint8_t f;
        if (ext_mask) f = ext_mask[(size_t)q * n + i] ? 1 : 0;
        else f = h_err_exact(sh.Hf, p.x, p.y, p.z, p.w) <= thr_sq ? 1 : 0;
        rmask[i] = f;  // each thread re-reads only the entries it wrote itself
        if (f) { kc[0] += 1.; kc[1] += (double)(-p.z); kc[2] += (double)(-p.w); kc[3] += (double)p.x; kc[4] += (double)p.y; }
    }
    TEAM_REDUCE(5, 0, kc);
    FINCLK(0);   // RANSAC-stage mask + count + coordinate sums
    const int k = (int)R.out[0];
    const double csum[4] = {R.out[1], R.out[2], R.out[3], R.out[4]};
    const bool refit = refine && n > 4 && k >= 4;
    if (stored_model && !refit) minimal_model64();   // the minimal model is what is returned

    if (refit) {
        // ---- refit on the inliers: normalisation statistics, L^T L, eigenvector -------------------------
        HNorm nm;
        if (seq) {   // one thread per sum, points in index order
            __syncthreads();   // rmask of all points visible
            if (tid < 4) {
                double c = 0;
                for (int i = 0; i < n; ++i)
                    if (rmask[i]) {
                        const float4 p = __ldg(reinterpret_cast<const float4*>(P + i));
                        c += (double)(tid == 0 ? -p.z : tid == 1 ? -p.w : tid == 2 ? p.x : p.y);
                    }
                R.out[tid] = c;
            }
            __syncthreads();
            nm.cmx = R.out[0] / k; nm.cmy = R.out[1] / k; nm.cMx = R.out[2] / k; nm.cMy = R.out[3] / k;
            __syncthreads();
            if (tid < 4) {
                const double ctr = tid == 0 ? nm.cmx : tid == 1 ? nm.cmy : tid == 2 ? nm.cMx : nm.cMy;
                double a = 0;
                for (int i = 0; i < n; ++i)
                    if (rmask[i]) {
                        const float4 p = __ldg(reinterpret_cast<const float4*>(P + i));
                        a += fabs((double)(tid == 0 ? -p.z : tid == 1 ? -p.w : tid == 2 ? p.x : p.y) - ctr);
                    }
                R.out[tid] = a;
            }
            __syncthreads();
            nm.smx = R.out[0]; nm.smy = R.out[1]; nm.sMx = R.out[2]; nm.sMy = R.out[3];
            __syncthreads();
        } else {
            // One more pass for everything else the refit sums: the absolute deviations (the scales) AND the entries of
            // L^T L.  Those are monomials in the normalised coordinates, so they are accumulated on the CENTRED coordinates
            // and multiplied by the scales afterwards: with pc = (Xc^2, Xc Yc, Xc, Yc^2, Yc, 1)
            //   [0,6) sum pc | [6,12) sum xc pc | [12,18) sum yc pc | [18,24) sum xc^2 pc | [24,30) sum yc^2 pc | [30,34) sum |.|
            nm.cmx = csum[0] / k; nm.cmy = csum[1] / k; nm.cMx = csum[2] / k; nm.cMy = csum[3] / k;
            double m[34];
#pragma unroll
            for (int j = 0; j < 34; ++j) m[j] = 0;
            for (int i = gtid; i < n; i += gstride)
                if (rmask[i]) {
                    const float4 p = __ldg(reinterpret_cast<const float4*>(P + i));
                    const double xc = (double)(-p.z) - nm.cmx, yc = (double)(-p.w) - nm.cmy;
                    const double Xc = (double)p.x - nm.cMx, Yc = (double)p.y - nm.cMy;
                    const double pc[6] = {Xc * Xc, Xc * Yc, Xc, Yc * Yc, Yc, 1.0};
                    const double xx = xc * xc, yy = yc * yc;
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        m[j] += pc[j];
                        m[6 + j] += xc * pc[j];
                        m[12 + j] += yc * pc[j];
                        m[18 + j] += xx * pc[j];
                        m[24 + j] += yy * pc[j];
                    }
                    m[30] += fabs(xc); m[31] += fabs(yc); m[32] += fabs(Xc); m[33] += fabs(Yc);
                }
            TEAM_REDUCE(34, 0, m);
            nm.smx = R.out[30]; nm.smy = R.out[31]; nm.sMx = R.out[32]; nm.sMy = R.out[33];
        }
        FINCLK(1);   // normalisation statistics (two passes)
        const bool degenerate = fabs(nm.smx) < DBL_EPSILON || fabs(nm.smy) < DBL_EPSILON ||
                                fabs(nm.sMx) < DBL_EPSILON || fabs(nm.sMy) < DBL_EPSILON;
        if (degenerate && stored_model) minimal_model64();   // runKernel fails on the inliers: the LM starts from the minimal model
        if (!degenerate) {
            nm.smx = k / nm.smx; nm.smy = k / nm.smy; nm.sMx = k / nm.sMx; nm.sMy = k / nm.sMy;
            // L^T L = [[P, 0, -Px], [0, P, -Py], [-Px, -Py, Pxy]] with the 3x3 symmetric blocks
            // P = sum p p^T, Px = sum x p p^T, Py = sum y p p^T, Pxy = sum (x^2+y^2) p p^T, p = (X, Y, 1):
            // 4 x 6 = 24 sums instead of the 45 entries of the upper triangle.
            if (seq) {   // the rows of L per point, then one thread per entry of the upper triangle, points in index order
                if (tid < n && rmask[tid]) {
                    const float4 p = __ldg(reinterpret_cast<const float4*>(P + tid));
                    const double x = ((double)(-p.z) - nm.cmx) * nm.smx, y = ((double)(-p.w) - nm.cmy) * nm.smy;
                    const double X = ((double)p.x - nm.cMx) * nm.sMx, Y = ((double)p.y - nm.cMy) * nm.sMy;
                    double* t = seq_tab + tid * 20;
                    t[0] = X; t[1] = Y; t[2] = 1; t[3] = 0; t[4] = 0; t[5] = 0; t[6] = -x * X; t[7] = -x * Y; t[8] = -x;
                    t[9] = 0; t[10] = 0; t[11] = 0; t[12] = X; t[13] = Y; t[14] = 1; t[15] = -y * X; t[16] = -y * Y; t[17] = -y;
                }
                __syncthreads();
                if (tid < 45) {
                    int j = 0, e = tid;
                    while (e >= 9 - j) { e -= 9 - j; ++j; }
                    const int kk = j + e;
                    double acc = 0;
                    for (int i = 0; i < n; ++i)
                        if (rmask[i]) {
                            const double* t = seq_tab + i * 20;
                            acc += t[j] * t[kk] + t[9 + j] * t[9 + kk];
                        }
                    jw.A[j * 9 + kk] = acc;
                }
                __syncthreads();
            }
            FINCLK(2);
            if (tid < 32) {  // warp 0
                // index of (a,b), a<=b, in the packed symmetric 3x3: (0,0)=0 (0,1)=1 (0,2)=2 (1,1)=3 (1,2)=4 (2,2)=5
                const int sym[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
                double* LtL = jw.A;
                if (!seq)
                    for (int j = tid; j < 81; j += 32) LtL[j] = 0;
                __syncwarp();
                if (tid == 0 && !seq) {
                    // the raw moments of the centred coordinates times the scales of their monomials
                    const double sp[6] = {nm.sMx * nm.sMx, nm.sMx * nm.sMy, nm.sMx, nm.sMy * nm.sMy, nm.sMy, 1.0};
                    const double sxx = nm.smx * nm.smx, syy = nm.smy * nm.smy;
                    for (int a = 0; a < 3; ++a)
                        for (int b = 0; b < 3; ++b) {
                            const int e = sym[a][b];
                            if (b >= a) {
                                LtL[a * 9 + b] = sp[e] * R.out[e];
                                LtL[(3 + a) * 9 + 3 + b] = sp[e] * R.out[e];
                                LtL[(6 + a) * 9 + 6 + b] = sp[e] * (sxx * R.out[18 + e] + syy * R.out[24 + e]);
                            }
                            LtL[a * 9 + 6 + b] = -(nm.smx * sp[e]) * R.out[6 + e];
                            LtL[(3 + a) * 9 + 6 + b] = -(nm.smy * sp[e]) * R.out[12 + e];
                        }
                }
                __syncwarp();
                double Hm[9], vec[9];
                bool done = false;
                if (fast_solver && smallest_eigvec9_warp(LtL, sh.diag)) {
                    for (int i = 0; i < 9; ++i) vec[i] = sh.diag[i];
                    h_from_eigvec(vec, nm, Hm);
                    done = true;
                }
                if (!done) h_from_LtL_warp(jw, nm, Hm);
                if (tid == 0)
                    for (int i = 0; i < 9; ++i) sh.H[i] = Hm[i];
            }
            __syncthreads();
            FINCLK(3);   // eigenvector of L^T L
        }

        // ---- Levenberg-Marquardt, max 10 iterations, eps = FLT_EPSILON (cv::LMSolver) ---------------------
        // OpenCV 4.13 refines all NINE entries of H (w = h6 X + h7 Y + h8) and rescales by 1/h8 afterwards; an
        // 8-parameter LM does not reproduce its early-stopped iterates (oracle/cv_ransac_oracle.c, h_refine_eval).
        // Rows of J: Jx = [a, 0, -xi a], Jy = [0, a, -yi a] with a = (X, Y, 1) ww, so
        //   J^T J = [[aa, 0, -xi aa], [0, aa, -yi aa], [., ., (xi^2 + yi^2) aa]],   J^T r = [a rx, a ry, -(xi rx + yi ry) a]:
        // per-thread accumulators  S | aa (6) | xi aa (6) | yi aa (6) | (xi^2+yi^2) aa (6) | a rx (3) | a ry (3) | (xi rx + yi ry) a (3) = 34
        // One pass + ONE cluster reduction per evaluation: |r|^2, max |r_i| and the 33 sums of J^T J / J^T r together
        // (the Jacobian terms of a rejected trial point are simply discarded).  Results: return value (S, rmax), sh.Ac, sh.vc.
        auto eval_seq = [&](const double* h) -> double2 {   // seq mode: OpenCV's products and order of summation
            if (tid < n && rmask[tid]) {
                const float4 p = __ldg(reinterpret_cast<const float4*>(P + tid));
                const double Mx = (double)p.x, My = (double)p.y;
                double ww = h[6] * Mx + h[7] * My + h[8];
                ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
                const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww;
                const double yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
                double* t = seq_tab + tid * 20;
                t[0] = Mx * ww; t[1] = My * ww; t[2] = ww; t[3] = 0; t[4] = 0; t[5] = 0;
                t[6] = -Mx * ww * xi; t[7] = -My * ww * xi; t[8] = -ww * xi;
                t[9] = 0; t[10] = 0; t[11] = 0; t[12] = Mx * ww; t[13] = My * ww; t[14] = ww;
                t[15] = -Mx * ww * yi; t[16] = -My * ww * yi; t[17] = -ww * yi;
                t[18] = xi - (double)(-p.z);
                t[19] = yi - (double)(-p.w);
            }
            __syncthreads();
            if (tid < 81) {
                const int j = tid / 9, kk = tid % 9;
                double acc = 0;
                for (int i = 0; i < n; ++i)
                    if (rmask[i]) {
                        const double* t = seq_tab + i * 20;
                        acc += t[j] * t[kk];
                        acc += t[9 + j] * t[9 + kk];
                    }
                sh.Ac[tid] = acc;
            } else if (tid < 90) {
                // v = cv::gemm(J, r, GEMM_1_T): FOUR interleaved partial sums over the rows of the compressed J (row 2c,
                // 2c+1 = inlier c), the (2k mod 4) tail rows into the first, combined ((s0 + s1) + s2) + s3
                const int j = tid - 81, kfull = k & ~1;
                double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                int c = 0;
                for (int i = 0; i < n; ++i)
                    if (rmask[i]) {
                        const double* t = seq_tab + i * 20;
                        const double px = t[j] * t[18], py = t[9 + j] * t[19];
                        if (c >= kfull) { s0 += px; s0 += py; }
                        else if (c & 1) { s2 += px; s3 += py; }
                        else { s0 += px; s1 += py; }
                        ++c;
                    }
                sh.vc[j] = ((s0 + s1) + s2) + s3;
            } else if (tid == 90) {
                // S = cv::norm(r, NORM_L2SQR) as the binary's AVX2 path sums it (oracle: norm_l2sqr): 16 elements per step
                // by fused multiply-add into 4 x 4 lanes, ((r0 + r1) + r2) + r3 lane-wise, (l0 + l1) + (l2 + l3); then
                // blocks of four rounded squares in order; the last (2k mod 4) elements fused
                double lanes[16], S = 0, rmax = 0;
#pragma unroll
                for (int e = 0; e < 16; ++e) lanes[e] = 0;
                const int nmain = (2 * k) & ~15, n4 = (2 * k) & ~3;
                int e = 0;
                bool reduced = false;
                for (int i = 0; i < n; ++i)
                    if (rmask[i]) {
                        const double rx = seq_tab[i * 20 + 18], ry = seq_tab[i * 20 + 19];
                        rmax = fmax(rmax, fmax(fabs(rx), fabs(ry)));
                        if (e < nmain) {
                            lanes[e & 15] = fma(rx, rx, lanes[e & 15]);
                            lanes[(e + 1) & 15] = fma(ry, ry, lanes[(e + 1) & 15]);
                        } else {
                            if (!reduced) {
                                double t4[4];
                                for (int l = 0; l < 4; ++l) t4[l] = ((lanes[l] + lanes[4 + l]) + lanes[8 + l]) + lanes[12 + l];
                                S += (t4[0] + t4[1]) + (t4[2] + t4[3]);
                                reduced = true;
                            }
                            if (e < n4) { S += rx * rx; S += ry * ry; }
                            else { S = fma(rx, rx, S); S = fma(ry, ry, S); }
                        }
                        e += 2;
                    }
                if (!reduced) {
                    double t4[4];
                    for (int l = 0; l < 4; ++l) t4[l] = ((lanes[l] + lanes[4 + l]) + lanes[8 + l]) + lanes[12 + l];
                    S += (t4[0] + t4[1]) + (t4[2] + t4[3]);
                }
                R.out[0] = S;
                R.out[34] = rmax;
            }
            __syncthreads();
            const double2 res = make_double2(R.out[0], R.out[34]);
            __syncthreads();
            return res;
        };
        auto eval = [&](const double* h) -> double2 {
            if (seq) return eval_seq(h);
            double acc[35];
#pragma unroll
            for (int j = 0; j < 35; ++j) acc[j] = 0;
            for (int i = gtid; i < n; i += gstride)
                if (rmask[i]) {
                    const float4 p = __ldg(reinterpret_cast<const float4*>(P + i));
                    const double Mx = (double)p.x, My = (double)p.y;
                    double ww = h[6] * Mx + h[7] * My + h[8];
                    ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
                    const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww;
                    const double yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
                    const double rx = xi - (double)(-p.z), ry = yi - (double)(-p.w);
                    acc[0] += rx * rx + ry * ry;
                    acc[34] = fmax(acc[34], fmax(fabs(rx), fabs(ry)));
                    const double a[3] = {Mx * ww, My * ww, ww};
                    const double aa[6] = {a[0] * a[0], a[0] * a[1], a[0] * a[2], a[1] * a[1], a[1] * a[2], a[2] * a[2]};
                    const double r2 = xi * xi + yi * yi, rr = xi * rx + yi * ry;
#pragma unroll
                    for (int u = 0; u < 6; ++u) {
                        acc[1 + u] += aa[u];
                        acc[7 + u] += xi * aa[u];
                        acc[13 + u] += yi * aa[u];
                        acc[19 + u] += r2 * aa[u];
                    }
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        acc[25 + u] += a[u] * rx;
                        acc[28 + u] += a[u] * ry;
                        acc[31 + u] += a[u] * rr;
                    }
                }
            FINCLK(11);  // eval: pass over the points
            TEAM_REDUCE(35, 1, acc);
            FINCLK(12);  // eval: reduction (block + team barrier + partials)
            const double S = R.out[0], rmax = R.out[34];
            for (int e = tid; e < 90; e += THREADS) {   // one thread per entry of J^T J (81) and J^T r (9)
                const double* o = R.out;
                if (e < 81) {
                    const int r = e / 9, c = e % 9, br = r / 3, bc = c / 3, u = r % 3, w = c % 3;
                    const int lo = u < w ? u : w, hi = u < w ? w : u;
                    const int se = lo == 0 ? hi : (lo == 1 ? 2 + hi : 5);   // packed symmetric 3x3: (0,0)=0 (0,1)=1 (0,2)=2 (1,1)=3 (1,2)=4 (2,2)=5
                    double val = 0;
                    if (br == bc) val = br == 2 ? o[19 + se] : o[1 + se];
                    else if (br + bc == 2 && br != 1) val = -o[7 + se];        // blocks (0,2), (2,0)
                    else if (br + bc == 3) val = -o[13 + se];                 // blocks (1,2), (2,1)
                    sh.Ac[e] = val;
                } else {
                    const int j = e - 81;
                    sh.vc[j] = j < 3 ? o[25 + j] : (j < 6 ? o[28 + j - 3] : -o[31 + j - 6]);
                }
            }
            __syncthreads();
            FINCLK(13);  // eval: J^T J assembled by thread 0
            return make_double2(S, rmax);
        };
        auto accept_candidate = [&]() {   // all threads; the caller's next barrier publishes it
            for (int e = tid; e < 90; e += THREADS) {
                if (e < 81) sh.A[e] = sh.Ac[e];
                else sh.v[e - 81] = sh.vc[e - 81];
            }
        };

        if (tid == 0)
            for (int i = 0; i < 9; ++i) sh.x[i] = sh.H[i];
        __syncthreads();
        {
            FINCLK(4);
            const double2 e0 = eval(sh.x);
            accept_candidate();
            if (tid == 0) {
                sh.S = e0.x; sh.rmax = e0.y;
                for (int i = 0; i < 9; ++i) sh.D[i] = sh.Ac[i * 9 + i];
                sh.lambda = 1; sh.lc = 0.75;
            }
            __syncthreads();
        }
        for (int iter = 0;;) {
            FINCLK(5);   // loop control / first accept
            if (tid < 32) {   // warp 0
                // J^T J is singular along h itself (the projection is scale-invariant): with lambda == 0 only the
                // eigen-decomposition solve with OpenCV's cut-off is meaningful; with lambda > 0 the matrix is SPD
                double* Ap = sh.Ap;
                const double lambda = sh.lambda;
                bool solved = false;
                if (seq || (lambda == 0 && !fast_solver)) {
                    // cv::solve(DECOMP_EIG) below: every step in seq mode, as OpenCV; the undamped step of the exact solver
                    for (int e = tid; e < 81; e += 32) Ap[e] = sh.A[e];
                    __syncwarp();
                    if (tid < 9) Ap[tid * 10] += lambda * sh.D[tid];
                    __syncwarp();
                } else {
                    // lambda > 0: A + lambda D is SPD.  lambda == 0, throughput mode: the null direction is known
                    // (n = x/|x|, J n = 0, hence n.v = 0), so the minimum-norm solution is that of the SPD system
                    // (A + s n n^T) d = v.  Lane i holds row i; factorisation and substitutions in registers.
                    const int i = tid < 9 ? tid : 8;
                    double sg = 0;
                    if (lambda == 0) {
                        double nn = 0, tr = 0;
#pragma unroll
                        for (int r = 0; r < 9; ++r) { nn += sh.x[r] * sh.x[r]; tr += sh.A[r * 10]; }
                        sg = tr / (9 * nn);
                    }
                    double a[9], dmax = 0;
                    const double xi = sh.x[i];
#pragma unroll
                    for (int r = 0; r < 9; ++r) {
                        dmax = fmax(dmax, fabs(sh.A[r * 10] + lambda * sh.D[r] + sg * sh.x[r] * sh.x[r]));
                        a[r] = sh.A[i * 9 + r] + (r == i ? lambda * sh.D[r] : 0.) + sg * xi * sh.x[r];
                    }
                    CholRegs<9> F;
                    chol_regs_factor<9>(a, dmax, F);
                    solved = F.ok;
                    if (solved) {
                        const double di = chol_regs_solve<9>(F, sh.v[i]);
                        if (tid < 9) {
                            sh.d[tid] = di;
                            if (lambda == 0) {   // the inverse diagonal, if this iteration asks for it, comes from this factor
                                sh.Lrinv[tid] = F.rinv;
#pragma unroll
                                for (int r = 0; r < 9; ++r) sh.L[tid * 9 + r] = F.row[r];
                            }
                        }
                    } else {
                        for (int e = tid; e < 81; e += 32) Ap[e] = sh.A[e];
                        __syncwarp();
                        if (tid < 9) Ap[tid * 10] += lambda * sh.D[tid];
                    }
                }
                if (tid == 0) sh.use_eig = solved ? 0 : 1;
            }
            __syncthreads();
            FINCLK(6);   // damped / regularised Cholesky step
            if (sh.use_eig && tid < 32) solve_sym_eig_warp<9>(jw, sh.Ap, sh.v, sh.d, nullptr);   // cv::solve(..., DECOMP_EIG)
            __syncthreads();
            FINCLK(7);   // eigen-decomposition step
            if (tid < 9) sh.xd[tid] = sh.x[tid] - sh.d[tid];
            __syncthreads();
            FINCLK(5);
            const double2 ed = eval(sh.xd);
            double dS_par = 0, dv_par = 0;
            if (!seq && tid < 32) {   // d.(2v - A d) and d.v: lane i takes row i, butterfly over the lanes
                const int i = tid < 9 ? tid : 8;
                double t = 0;
#pragma unroll
                for (int j = 0; j < 9; ++j) t += sh.A[i * 9 + j] * sh.d[j];
                dS_par = tid < 9 ? sh.d[i] * (2 * sh.v[i] - t) : 0.;
                dv_par = tid < 9 ? sh.d[i] * sh.v[i] : 0.;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    dS_par += __shfl_xor_sync(0xffffffffu, dS_par, o);
                    dv_par += __shfl_xor_sync(0xffffffffu, dv_par, o);
                }
            }
            if (tid == 0) {
                const double Sd = ed.x, S = sh.S;
                sh.Sd = Sd;
                sh.need_diag = 0;
                double dS = dS_par;
                if (seq) {   // cv::gemm(A, d, -1, v, 2) row by row (four partial sums), then cv::Mat::dot
                    double tmp[9];
                    for (int i = 0; i < 9; ++i) tmp[i] = cv_gemm_dot<9>(sh.A + i * 9, sh.d) * -1. + sh.v[i] * 2.;
                    dS = cv_mat_dot<9>(sh.d, tmp);
                }
                const double Rr = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
                if (Rr > 0.75) {
                    sh.lambda *= 0.5;
                    if (sh.lambda < sh.lc) sh.lambda = 0;
                } else if (Rr < 0.25) {
                    double t = dv_par;
                    if (seq) t = cv_mat_dot<9>(sh.d, sh.v);
                    double nu = (Sd - S) / (fabs(t) > DBL_EPSILON ? t : 1) + 2;
                    nu = fmin(fmax(nu, 2.), 10.);
                    sh.nu = nu;
                    if (sh.lambda == 0) {
                        sh.need_diag = (fast_solver && !sh.use_eig) ? 3 : 2;   // 3: from the Cholesky factor of this iteration's step (warp 0, below); 2: from the eigen-decomposition
                    } else {
                        sh.lambda *= nu;
                    }
                }
            }
            __syncthreads();
            FINCLK(8);   // gain ratio, lambda update (thread 0)
            if (sh.need_diag == 3 && tid < 32) {   // diag of the pseudo-inverse = diag((A + s n n^T)^-1) - n_i^2 / s
                const int i = tid < 9 ? tid : 8;
                double nn = 0, tr = 0;
#pragma unroll
                for (int r = 0; r < 9; ++r) { nn += sh.x[r] * sh.x[r]; tr += sh.A[r * 10]; }
                const double sg = tr / (9 * nn);
                CholRegs<9> F;
                F.ok = true;
                F.rinv = sh.Lrinv[i];
#pragma unroll
                for (int r = 0; r < 9; ++r) { F.row[r] = sh.L[i * 9 + r]; F.col[r] = 0; }
                const double dg = chol_regs_inv_diag<9>(F, jw.V);
                if (tid < 9) sh.diag[tid] = dg - sh.x[tid] * sh.x[tid] / (nn * sg);
                if (tid == 0) sh.need_diag = 1;
            }
            __syncthreads();
            // lambda was 0 in this iteration: the step came from the eigen-decomposition of this same A (Ap = A + 0 D)
            if (sh.need_diag == 2 && tid < 32) solve_sym_eig_warp<9>(jw, sh.A, nullptr, nullptr, sh.diag, sh.use_eig != 0);
            __syncthreads();
            FINCLK(9);   // inverse diagonal
            if (tid == 0) {
                if (sh.need_diag) {
                    double maxval = DBL_EPSILON;
                    for (int i = 0; i < 9; ++i) maxval = fmax(maxval, fabs(sh.diag[i]));
                    sh.lambda = sh.lc = 1. / maxval;
                    sh.lambda *= sh.nu * 0.5;
                }
                sh.flag = sh.Sd < sh.S;
                if (sh.flag) {   // accepted: the trial point's J^T J, J^T r and residual norms become current
                    sh.S = sh.Sd;
                    sh.rmax = ed.y;
                    for (int i = 0; i < 9; ++i) sh.x[i] = sh.xd[i];
                }
            }
            __syncthreads();
            if (sh.flag) accept_candidate();
            ++iter;
            if (tid == 0) {
                double dmax = 0;
                for (int i = 0; i < 9; ++i) dmax = fmax(dmax, fabs(sh.d[i]));
                sh.proceed = iter < 10 && dmax >= (double)FLT_EPSILON && sh.rmax >= (double)FLT_EPSILON;
                sh.lm_iters = iter;
            }
            __syncthreads();
            FINCLK(10);  // accept + stopping test
            if (!sh.proceed) break;
        }
        if (tid == 0) {
            const double sc = fabs(sh.x[8]) > DBL_EPSILON ? 1. / sh.x[8] : 1;   // OpenCV: convertTo(..., scaleFor(H22))
            for (int i = 0; i < 9; ++i) sh.H[i] = sh.x[i] * sc;
            for (int i = 0; i < 8; ++i) sh.Hf[i] = (float)sh.H[i];
        }
        __syncthreads();
    }

    // ---- outputs ---------------------------------------------------------------------------------------------
    int n_inl = 0;
    if (mask_semantics == 0 && n > 4 && refine) {
        for (int i = gtid; i < n; i += gstride) {
            const float4 p = __ldg(reinterpret_cast<const float4*>(P + i));
            const uint8_t f = h_err_exact(sh.Hf, p.x, p.y, p.z, p.w) <= thr_sq ? 1 : 0;
            mask[i] = f;
            n_inl += f;
        }
    } else {
        for (int i = gtid; i < n; i += gstride) {
            const uint8_t f = rmask[i];
            mask[i] = f;
            n_inl += f;
        }
    }
    // the inlier count is an integer sum: one atomic add per CTA onto info (zeroed by the writer before the first team
    // barrier) instead of one more team-wide reduction with its barrier
    n_inl = __reduce_add_sync(0xffffffffu, n_inl);
    if ((tid & 31) == 0 && n_inl) atomicAdd(&sh.k, n_inl);
    __syncthreads();
    if (tid == 0 && sh.k) atomicAdd(inf + 8, sh.k);
    if (writer && tid < 9) H_out[(size_t)q * 9 + tid] = sh.H[tid];
    if (writer && tid == 0) {
        inf[0] = 0; inf[1] = s.iters_run; inf[2] = s.best; inf[3] = k;
        inf[4] = smp.x; inf[5] = smp.y; inf[6] = smp.z; inf[7] = smp.w;
        inf[9] = sh.lm_iters; inf[10] = s.pad; inf[11] = 0;
    }
    FINCLK(14);  // final mask + outputs
    if (!GRID) cluster.sync();  // no CTA may exit while a peer can still read its shared memory
#undef TEAM_REDUCE
}

// ---- self tests / probes -----------------------------------------------------------------------------------------
__global__ void k_selftest_rcp(unsigned long long* mismatches, unsigned long long* tested) {
    unsigned long long bad = 0, cnt = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < (1ull << 32); b += stride) {
        const float x = __uint_as_float((uint32_t)b);
        if (!rcp_rn_in_fast_range(x)) continue;
        const float r0 = rcp_rn_fast_range(x), r1 = __frcp_rn(x);
        ++cnt;
        if (__float_as_uint(r0) != __float_as_uint(r1)) ++bad;
    }
    atomicAdd(mismatches, bad);
    atomicAdd(tested, cnt);
}

template <int PACKED>
__global__ void __launch_bounds__(256) k_probe_fma(float* out, int iters, float a, float b) {
    constexpr int ILP = 8;
    float acc = 0.f;
    if (PACKED) {
        f2_t x[ILP];
        const f2_t a2 = f2_pack(a, a * 1.0001f), b2 = f2_pack(b, b * 0.999f);
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = f2_pack(a + (float)(threadIdx.x + i), b + (float)i);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < ILP; ++i) x[i] = f2_fma(x[i], a2, b2);
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            float lo, hi;
            f2_unpack(x[i], lo, hi);
            acc += lo + hi;
        }
    } else {
        float x[ILP];
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = a + (float)(threadIdx.x + i);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < ILP; ++i) x[i] = __fmaf_rn(x[i], a, b);
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc += x[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

}  // namespace b2r

// Pieces shared by the homography and the PnP pipelines: the selection rules of OpenCV's
// RANSACPointSetRegistrator::run, the deterministic cluster-wide reduction used by the finalize
// kernels, and the small symmetric solvers of cv::LMSolver.  Kernels defined here are `static`: each translation
// unit of the library (api.cu, api_pnp.cu) gets its own copy.
#pragma once
#include <cooperative_groups.h>
#include "sampler.cuh"

namespace b2r {

// ---- selection ----------------------------------------------------------------------------------------------
struct HSelect {
    int best;        // winning iteration / local hypothesis index, -1 = none
    int best_count;  // its RANSAC-stage inlier count
    int iters_run;   // iterations executed
    int pad;
};

__device__ __forceinline__ int ransac_update_num_iters(double p, double ep, int modelPoints, int maxIters) {
    p = fmax(p, 0.);
    p = fmin(p, 1.);
    ep = fmax(ep, 0.);
    ep = fmin(ep, 1.);
    double num = fmax(1. - p, DBL_MIN);
    double denom = 1. - pow(1. - ep, (double)modelPoints);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return denom >= 0 || -num >= maxIters * (-denom) ? maxIters : __double2int_rn(num / denom);
}

// Running state of OpenCV's sequential RANSAC loop, one per problem.  The replay path scores the
// iterations in chunks of doubling length (boundaries 64, 192, 448, ... for the homography; 256, 768, ... for PnP): after each chunk the rule below
// is advanced over the chunk's counts, and chunks that start beyond the current iteration bound are never generated,
// solved or scored — the reference typically stops after tens of iterations out of maxIters = 2000 / 5000.
struct RansacState {
    unsigned long long rng;  // cv::RNG state after the last generated subset
    int niters;              // current iteration bound (shrinks through RANSACUpdateNumIters)
    int max_good;            // best inlier count so far
    int best;                // iteration that produced it, -1 = none
    int it;                  // iterations consumed
    int gen;                 // iterations for which a subset exists (getSubset can give up: the loop then ends there)
    int done;                // the loop has ended
};

static __global__ void k_state_init(RansacState* __restrict__ st, int max_iters, int Q) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    RansacState s;
    s.rng = 0xffffffffffffffffull;  // OpenCV re-seeds cv::RNG with 2^64-1 inside every call
    s.niters = max(max_iters, 1);
    s.max_good = 0; s.best = -1; s.it = 0; s.gen = 0; s.done = 0;
    st[q] = s;
}

// Advance the rule over iterations [begin, begin+len): take a hypothesis when its count beats max(best, modelPoints-1),
// shrink niters, stop at niters.  counts: [Q][H_stride].  *not_done is incremented for every problem that needs the next chunk.
static __global__ void k_select_cv_chunk(const int* __restrict__ counts, int H_stride, int begin, int len, int n, double confidence,
                                         int model_points, RansacState* __restrict__ st, HSelect* __restrict__ sel, int Q,
                                         int* __restrict__ not_done) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    RansacState s = st[q];
    if (!s.done) {
        const int* C = counts + (size_t)q * H_stride;
        const int end = min(begin + len, s.gen);
        int it = max(begin, s.it);
        for (; it < s.niters && it < end; ++it) {
            const int good = C[it];
            if (good > max(s.max_good, model_points - 1)) {
                s.best = it;
                s.max_good = good;
                s.niters = ransac_update_num_iters(confidence, (double)(n - good) / n, model_points, s.niters);
            }
        }
        s.it = it;
        // ended: bound reached, getSubset gave up inside this chunk, or this was the last chunk
        s.done = (it >= s.niters || s.gen < begin + len || begin + len >= H_stride) ? 1 : 0;
        st[q] = s;
        if (!s.done) atomicAdd(not_done, 1);
    }
    HSelect o;
    o.best = s.best;
    o.best_count = s.max_good;
    o.iters_run = s.it;
    o.pad = 0;
    sel[q] = o;
}

// Fixed-H rule: the lowest-id hypothesis with the maximum count; key = count << 32 | (0xFFFFFFFF - id).
static __global__ void __launch_bounds__(256)
k_argmax_key(const int* __restrict__ counts, int H, unsigned long long id_base, unsigned long long* __restrict__ keys) {
    const int q = blockIdx.y;
    const int* C = counts + (size_t)q * H;
    unsigned long long best = 0;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < H; g += gridDim.x * blockDim.x) {
        const unsigned long long key =
            ((unsigned long long)(uint32_t)C[g] << 32) | (0xFFFFFFFFull - ((id_base + (unsigned long long)g) & 0xFFFFFFFFull));
        best = key > best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_down_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0 && best) atomicMax(keys + q, best);
}

static __global__ void k_select_from_keys(const unsigned long long* __restrict__ keys, unsigned long long id_base, int H,
                                   int model_points, HSelect* __restrict__ sel, int Q) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const unsigned long long key = keys[q];
    const int count = (int)(key >> 32);
    const unsigned long long gid = 0xFFFFFFFFull - (key & 0xFFFFFFFFull);
    HSelect s;
    s.best = count > model_points - 1 ? (int)(gid - (id_base & 0xFFFFFFFFull)) : -1;
    s.best_count = count > model_points - 1 ? count : 0;
    s.iters_run = H;
    s.pad = (int)(gid & 0x7fffffff);   // info.reserved[0]: the winner's global id (low 31 bits), as after a keyed finish
    sel[q] = s;
}

// ---- cluster-wide deterministic reductions ---------------------------------------------------------------------------
// The finalize kernel runs as ONE thread-block cluster per problem (8 CTAs x 1024 threads for a large problem, a
// single CTA for a small one).  Every CTA reduces its share of the points to NV partial sums in its own shared
// memory; after one cluster barrier every CTA reads all partials through distributed shared memory, in rank
// order, so all CTAs hold the same bit pattern and replay the (tiny) sequential part of the algorithm
// redundantly — no broadcast, no atomics, run-to-run deterministic.  Partials are double-buffered so that one
// barrier per reduction is enough.
namespace cg = cooperative_groups;

constexpr int RED_MAX = 40;

struct ClusterRed {
    double part[2][RED_MAX];  // this CTA's partial sums (double-buffered), read remotely
    double warp[32 * RED_MAX];
    double out[RED_MAX];
    int phase;
};

template <int THREADS, int NV, bool IS_MAX>
__device__ __forceinline__ void cluster_reduce(ClusterRed& R, double (&v)[NV]) {
    static_assert(NV <= RED_MAX, "too many values");
    cg::cluster_group cluster = cg::this_cluster();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_down_sync(0xffffffffu, x, o);
            x = IS_MAX ? fmax(x, y) : x + y;
        }
        if (lane == 0) R.warp[warp * NV + i] = x;
    }
    __syncthreads();
    const int ph = R.phase;
    if (threadIdx.x < NV) {
        double s = R.warp[threadIdx.x];
        for (int w = 1; w < THREADS / 32; ++w) s = IS_MAX ? fmax(s, R.warp[w * NV + threadIdx.x]) : s + R.warp[w * NV + threadIdx.x];
        R.part[ph][threadIdx.x] = s;
    }
    cluster.sync();
    if (threadIdx.x < NV) {
        const unsigned nb = cluster.num_blocks();
        double s = 0;
        for (unsigned r = 0; r < nb; ++r) {
            const double* remote = cluster.map_shared_rank(&R.part[ph][0], r);
            s = IS_MAX ? fmax(s, remote[threadIdx.x]) : (r == 0 ? remote[threadIdx.x] : s + remote[threadIdx.x]);
        }
        R.out[threadIdx.x] = s;
    }
    if (threadIdx.x == 0) R.phase = ph ^ 1;
    __syncthreads();
}

// ---- warp reduction of NV values per lane ------------------------------------------------------------------------------
// The butterfly-per-value form above costs NV x 5 64-bit shuffles per warp; with 35 values and sixteen warps the
// shuffle unit, not the passes over the points, was the bulk of an LM evaluation (tools/prof_finalize_sections.py).
// Transpose-reduce instead: in the step with lane offset o every lane keeps one half of its values and hands the other
// half to lane ^ o, so M = 2^m values take M - 1 shuffles (+ the plain butterfly over the lane bits that are left when
// M < 32), and lane L ends up with the warp total of value (L mod M) of its group.  Groups of 32 / 16 / 8 / 4 first; the
// last NV mod 4 values (and the NMAX max-combined ones, which must be among them) go through the plain butterfly.
// Fixed order: every warp, CTA and run produces the same bits.  Results: R.warp[warp * NV + value].
template <int M, int OFF>
__device__ __forceinline__ void warp_transpose_reduce_group(double* v, double* dst, int lane) {
    constexpr unsigned FULL = 0xffffffffu;
    // v[OFF .. OFF + M) are this group's values (static register indices throughout)
#pragma unroll
    for (int h = M / 2, o = 16; h >= 1; h >>= 1, o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const double send = up ? v[OFF + i] : v[OFF + i + h];
            const double keep = up ? v[OFF + i + h] : v[OFF + i];
            v[OFF + i] = keep + __shfl_xor_sync(FULL, send, o);
        }
    }
    // lane bits below 32 / M have not been folded yet
#pragma unroll
    for (int o = 16 / M; o >= 1; o >>= 1) v[OFF] += __shfl_xor_sync(FULL, v[OFF], o);
    // value index held by this lane: the lane bits 16, 8, ... consumed by the halving steps, most significant first
    if constexpr (M == 32) { dst[OFF + lane] = v[OFF]; }
    else if ((lane & (32 / M - 1)) == 0) { dst[OFF + (lane / (32 / M))] = v[OFF]; }
}

template <int NV, int NMAX>
__device__ __forceinline__ void warp_reduce_multi(double (&v)[NV], double* dst) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    constexpr int NS = NV - NMAX;                 // values combined by +
    constexpr int G32 = (NS / 32) * 32;
    constexpr int G16 = G32 + (((NS - G32) / 16) * 16);
    constexpr int G8 = G16 + (((NS - G16) / 8) * 8);
    constexpr int G4 = G8 + (((NS - G8) / 4) * 4);
    if constexpr (G32 > 0) warp_transpose_reduce_group<32, 0>(v, dst, lane);
    static_assert(G32 <= 32, "one group of 32 at most");
    if constexpr (G16 > G32) warp_transpose_reduce_group<16, G32>(v, dst, lane);
    if constexpr (G8 > G16) warp_transpose_reduce_group<8, G16>(v, dst, lane);
    if constexpr (G4 > G8) warp_transpose_reduce_group<4, G8>(v, dst, lane);
#pragma unroll
    for (int i = G4; i < NV; ++i) {
        double x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_xor_sync(FULL, x, o);
            x = i >= NS ? fmax(x, y) : x + y;
        }
        if (lane == 0) dst[i] = x;
    }
}

// Cluster-wide reduction with the LAST NMAX of the NV values combined by max instead of + (one cluster barrier for a set
// of sums plus a max-norm).
template <int THREADS, int NV, int NMAX>
__device__ __forceinline__ void cluster_reduce_tail_max(ClusterRed& R, double (&v)[NV]) {
    static_assert(NV <= RED_MAX && NMAX <= NV, "too many values");
    cg::cluster_group cluster = cg::this_cluster();
    warp_reduce_multi<NV, NMAX>(v, R.warp + (threadIdx.x >> 5) * NV);
    __syncthreads();
    const int ph = R.phase;
    const unsigned nb = cluster.num_blocks();
    if (threadIdx.x < NV) {
        const bool is_max = (int)threadIdx.x >= NV - NMAX;
        double s = R.warp[threadIdx.x];
#pragma unroll
        for (int w = 1; w < THREADS / 32; ++w) s = is_max ? fmax(s, R.warp[w * NV + threadIdx.x]) : s + R.warp[w * NV + threadIdx.x];
        if (nb == 1) R.out[threadIdx.x] = s;   // a single CTA: no exchange
        else R.part[ph][threadIdx.x] = s;
    }
    if (nb == 1) {
        __syncthreads();
        return;
    }
    cluster.sync();
    if (threadIdx.x < NV) {
        const bool is_max = (int)threadIdx.x >= NV - NMAX;
        double s = 0;
        for (unsigned r = 0; r < nb; ++r) {
            const double* remote = cluster.map_shared_rank(&R.part[ph][0], r);
            s = r == 0 ? remote[threadIdx.x] : (is_max ? fmax(s, remote[threadIdx.x]) : s + remote[threadIdx.x]);
        }
        R.out[threadIdx.x] = s;
    }
    if (threadIdx.x == 0) R.phase = ph ^ 1;
    __syncthreads();
}

// The same deterministic all-to-all reduction over a whole cooperative GRID (one big problem on many SMs): every CTA
// publishes its partials in global memory, one grid barrier, every CTA combines all partials in the same fixed order —
// eight lanes per value, all values at once (the loads are issued together: one L2 round trip per 64 CTAs instead of one
// per four), then a butterfly over the eight lanes.  gscratch holds 2 x gridDim.x x RED_MAX doubles
// (double-buffered like ClusterRed::part).  The LAST NMAX values are combined by max.
template <int THREADS, int NV, int NMAX>
__device__ __forceinline__ void grid_reduce_tail_max(ClusterRed& R, double (&v)[NV], double* __restrict__ gscratch) {
    static_assert(NV <= RED_MAX && NMAX <= NV, "too many values");
    constexpr unsigned FULL = 0xffffffffu;
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    warp_reduce_multi<NV, NMAX>(v, R.warp + warp * NV);
    __syncthreads();
    const int ph = R.phase;
    double* buf = gscratch + (size_t)ph * gridDim.x * RED_MAX;
    if (threadIdx.x < NV) {
        const bool is_max = (int)threadIdx.x >= NV - NMAX;
        double s = R.warp[threadIdx.x];
#pragma unroll
        for (int w = 1; w < THREADS / 32; ++w) s = is_max ? fmax(s, R.warp[w * NV + threadIdx.x]) : s + R.warp[w * NV + threadIdx.x];
        buf[(size_t)blockIdx.x * RED_MAX + threadIdx.x] = s;
        __threadfence();
    }
    grid.sync();
    {
        // eight lanes per value; a lane takes the partials of CTAs sub, sub + 8, ... in batches of eight loads issued together
        // (one L2 round trip per 64 CTAs), sums them in that order, then a butterfly over the eight lanes
        static_assert(THREADS / 8 >= NV, "eight lanes per value");
        const int val = threadIdx.x >> 3, sub = threadIdx.x & 7;
        const bool active = val < NV;
        const bool is_max = val >= NV - NMAX;
        const int vv = active ? val : 0;
        double s = 0;   // neutral for the sums and for the max of non-negative norms
        for (unsigned r0 = sub; r0 < gridDim.x; r0 += 64) {
            double x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned r = r0 + 8 * u;
                x[u] = r < gridDim.x ? __ldcg(buf + (size_t)r * RED_MAX + vv) : 0.;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) s = is_max ? fmax(s, x[u]) : s + x[u];
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const double y = __shfl_xor_sync(FULL, s, o);
            s = is_max ? fmax(s, y) : s + y;
        }
        if (active && sub == 0) R.out[val] = s;
    }
    if (threadIdx.x == 0) R.phase = ph ^ 1;
    __syncthreads();
}

// x = solve(A, b), A symmetric n x n, through its Jacobi eigen-decomposition with OpenCV's
// back-substitution threshold (cv::solve DECOMP_EIG); optionally the diagonal of A^-1.
template <int N>
__device__ void solve_sym_eig(const double* A, const double* b, double* x, double* inv_diag) {
    double a[N * N], W[N], V[N * N];
    for (int i = 0; i < N * N; ++i) a[i] = A[i];
    jacobi_eig<N>(a, W, V);
    double thr = 0;
    for (int i = 0; i < N; ++i) thr += W[i];   // OpenCV's SVBkSb: the SIGNED sum
    thr *= DBL_EPSILON * 2;
    for (int j = 0; j < N; ++j) {
        if (x) x[j] = 0;
        if (inv_diag) inv_diag[j] = 0;
    }
    for (int i = 0; i < N; ++i) {
        if (fabs(W[i]) <= thr) continue;
        const double wi = 1 / W[i];            // ... multiplies by the reciprocal (pinned against cv2.solve / cv2.invert)
        if (x) {
            double s = 0;
            for (int j = 0; j < N; ++j) s += V[i * N + j] * b[j];
            s *= wi;
            for (int j = 0; j < N; ++j) x[j] += s * V[i * N + j];
        }
        if (inv_diag)
            for (int j = 0; j < N; ++j) inv_diag[j] += (V[i * N + j] * wi) * V[i * N + j];
    }
}

// The inner products of cv::LMSolver's linear algebra in the summation order of the cv2 4.13.0 binary (pinned bit for
// bit in oracle/cv_ransac_oracle.c: acc4_dot, cv_dot; tests/golden/cv2_lm_blocks.json).
// cv::gemm: four interleaved partial sums, the tail into the first, ((s0 + s1) + s2) + s3, products rounded.
template <int N>
__device__ __forceinline__ double cv_gemm_dot(const double* a, const double* b) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int k = 0;
    for (; k <= N - 4; k += 4) {
        s0 += a[k] * b[k];
        s1 += a[k + 1] * b[k + 1];
        s2 += a[k + 2] * b[k + 2];
        s3 += a[k + 3] * b[k + 3];
    }
    for (; k < N; ++k) s0 += a[k] * b[k];
    return ((s0 + s1) + s2) + s3;
}
// cv::Mat::dot: blocks of four with the compiler's FMA contraction — fma(a3, b3, fma(a2, b2, fma(a0, b0, a1 b1))) added
// to the running sum — and a fused tail.
template <int N>
__device__ __forceinline__ double cv_mat_dot(const double* a, const double* b) {
    double s = 0;
    int i = 0;
    for (; i <= N - 4; i += 4) s += fma(a[i + 3], b[i + 3], fma(a[i + 2], b[i + 2], fma(a[i], b[i], a[i + 1] * b[i + 1])));
    for (; i < N; ++i) s = fma(a[i], b[i], s);
    return s;
}

// solve_sym_eig executed by ONE WARP through the warp-cooperative Jacobi (jacobi_eig_warp3): same arithmetic, ~2.5x
// less latency.  A_src, b, x, inv_diag live in shared memory; x / inv_diag may be null.
// reuse = true: jw.W / jw.V already hold the decomposition of this matrix (the previous call's), skip the Jacobi.
template <int N>
__device__ void solve_sym_eig_warp(JacobiWarp9& jw, const double* A_src, const double* b, double* x, double* inv_diag,
                                   bool reuse = false) {
    const int lane = threadIdx.x & 31;
    if (!reuse) {
        for (int e = lane; e < N * N; e += 32) jw.A[e] = A_src[e];
        __syncwarp();
        jacobi_eig_warp3<N>(jw.A, jw.W, jw.V, jw.indR, jw.indC);
    }
    double thr = 0;
    for (int i = 0; i < N; ++i) thr += jw.W[i];   // signed sum, as OpenCV's SVBkSb
    thr *= DBL_EPSILON * 2;
    if (lane < N) {
        double xj = 0, dj = 0;
        for (int i = 0; i < N; ++i) {
            if (fabs(jw.W[i]) <= thr) continue;
            const double wi = 1 / jw.W[i];        // multiplication by the reciprocal, as OpenCV
            if (x) {
                double s = 0;
                for (int j = 0; j < N; ++j) s += jw.V[i * N + j] * b[j];
                s *= wi;
                xj += s * jw.V[i * N + lane];
            }
            if (inv_diag) dj += (jw.V[i * N + lane] * wi) * jw.V[i * N + lane];
        }
        if (x) x[lane] = xj;
        if (inv_diag) inv_diag[lane] = dj;
    }
    __syncwarp();
}

// Cholesky factor of a symmetric positive definite N x N matrix (lower triangle in L); false when a pivot is not
// safely positive — callers then fall back to the Jacobi eigen-decomposition, which is what OpenCV always uses.
template <int N>
__device__ __forceinline__ bool cholesky(const double* A, double* L) {
    double dmax = 0;
    for (int i = 0; i < N; ++i) dmax = fmax(dmax, fabs(A[i * N + i]));
    for (int j = 0; j < N; ++j) {
        double d = A[j * N + j];
        for (int k = 0; k < j; ++k) d -= L[j * N + k] * L[j * N + k];
        if (!(d > dmax * 1e-14)) return false;
        d = sqrt(d);
        L[j * N + j] = d;
        for (int i = j + 1; i < N; ++i) {
            double t = A[i * N + j];
            for (int k = 0; k < j; ++k) t -= L[i * N + k] * L[j * N + k];
            L[i * N + j] = t / d;
        }
    }
    return true;
}

template <int N>
__device__ __forceinline__ void cholesky_solve(const double* L, const double* b, double* x) {
    double y[N];
    for (int i = 0; i < N; ++i) {
        double t = b[i];
        for (int k = 0; k < i; ++k) t -= L[i * N + k] * y[k];
        y[i] = t / L[i * N + i];
    }
    for (int i = N - 1; i >= 0; --i) {
        double t = y[i];
        for (int k = i + 1; k < N; ++k) t -= L[k * N + i] * x[k];
        x[i] = t / L[i * N + i];
    }
}

// ---- register-resident Cholesky of a 9x9 SPD system (throughput paths of the finalize kernel) ---------------------------
// Lane i (< N; the others shadow lane N-1) owns row i of L, column i of L and 1/L[i][i] in registers; nothing passes
// through shared memory between the factorisation and the substitutions.  Right-looking: column j costs one broadcast
// of the pivot, one rsqrt (L[j][j] = a rs, L[i][j] = a[i][j] rs: no square root followed by a division), the N-1-j
// broadcasts of the new column (independent shuffles) and one FMA per trailing entry; the substitutions multiply by
// the stored reciprocal.  ~4k cycles for factorisation + solve against ~11k for the round-1 form that kept L in shared
// memory and divided by the pivot in every substitution step (tools/prof_finalize_sections.py).  Not bit-identical to cholesky<N> (rsqrt, reciprocals): used
// only where the iterates are not pinned to OpenCV's bits (parallel-sum refinement, throughput mode).
template <int N>
struct CholRegs {
    double row[N];   // L[i][k], k <= i
    double col[N];   // L[k][i], k > i
    double rinv;     // 1 / L[i][i]
    bool ok;
};

// a[k] = A[i][k] (k <= i is what is read).  Every lane of the warp calls.
template <int N>
__device__ __forceinline__ void chol_regs_factor(double (&a)[N], double dmax, CholRegs<N>& F) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    F.ok = true;
    F.rinv = 1;
#pragma unroll
    for (int k = 0; k < N; ++k) { F.row[k] = 0; F.col[k] = 0; }
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double ajj = __shfl_sync(FULL, a[j], j);
        if (!(ajj > dmax * 1e-14)) F.ok = false;               // warp-uniform; the garbage that follows is discarded by the caller
        const double rs = rsqrt(ajj);
        const double lij = lane == j ? ajj * rs : a[j] * rs;   // lanes i < j: unused
        F.row[j] = lij;
        if (lane == j) F.rinv = rs;
#pragma unroll
        for (int m = 0; m < N; ++m) {                          // fixed bounds: both loops unroll, every index is static
            if (m > j) {
                const double vm = __shfl_sync(FULL, lij, m);   // L[m][j]
                if (lane == j) F.col[m] = vm;
                a[m] -= lij * vm;                              // a[i][m] -= L[i][j] L[m][j], meaningful for i >= m
            }
        }
    }
}

// x_i = ((L L^T)^-1 b)_i; b_i is this lane's entry
template <int N>
__device__ __forceinline__ double chol_regs_solve(const CholRegs<N>& F, double b_i) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    double t = b_i, y = 0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double q = t * F.rinv;
        const double yk = __shfl_sync(FULL, q, k);
        if (lane == k) y = q;
        t -= F.row[k] * yk;          // meaningful for lanes > k
    }
    double u = y, x = 0;
#pragma unroll
    for (int r = N - 1; r >= 0; --r) {
        const double q = u * F.rinv;
        const double xr = __shfl_sync(FULL, q, r);
        if (lane == r) x = q;
        u -= F.col[r] * xr;          // meaningful for lanes < r
    }
    return x;
}

// diag((L L^T)^-1)_i: the rows of M = L^-1 by forward substitution on the identity (lane i owns row i), then column sums
// of squares through `scratch` (N x N doubles of shared memory).
template <int N>
__device__ __forceinline__ double chol_regs_inv_diag(const CholRegs<N>& F, double* scratch) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int i = lane < N ? lane : N - 1;
    double acc[N], fin[N];
#pragma unroll
    for (int j = 0; j < N; ++j) { acc[j] = j == i ? 1. : 0.; fin[j] = 0; }
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            if (j <= k) {
                const double mkj = __shfl_sync(FULL, acc[j] * F.rinv, k);   // M[k][j]
                if (lane == k) fin[j] = mkj;
                acc[j] -= F.row[k] * mkj;                                    // meaningful for lanes > k
            }
        }
    }
    __syncwarp();
    if (lane < N) {
#pragma unroll
        for (int j = 0; j < N; ++j) scratch[lane * N + j] = fin[j];
    }
    __syncwarp();
    double d = 0;
#pragma unroll
    for (int r = 0; r < N; ++r) d += scratch[r * N + i] * scratch[r * N + i];
    __syncwarp();
    return d;
}

}  // namespace b2r

// sharded finish: info.best_iter = winning hypothesis id - hyp_begin, or -1 when the winner lives on another rank's shard
static __global__ void k_patch_best_iter(int* __restrict__ info, const unsigned long long* __restrict__ keys, long long hyp_begin,
                                  int H, int Q) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    if (info[(size_t)q * 12 + 2] < 0) return;   // no model
    const long long gid = (long long)(0xFFFFFFFFull - (keys[q] & 0xFFFFFFFFull)) - hyp_begin;
    info[(size_t)q * 12 + 2] = (gid >= 0 && gid < H) ? (int)gid : -1;
}


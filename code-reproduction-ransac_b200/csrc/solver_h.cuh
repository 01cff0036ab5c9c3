// K2 — minimal (4-point) and k-point homography solver, fp64, one thread per hypothesis.
//
// Device-side counterpart of OpenCV's HomographyEstimatorCallback::runKernel, the solver that
// cv2.findHomography(..., cv2.RANSAC, thr) (reference: main_v1.py:312) runs once per RANSAC
// iteration: per-axis L1-normalised DLT, 9x9 L^T L, symmetric Jacobi eigen-solver (OpenCV's own
// cyclic-by-max-pivot scheme), smallest eigenvector, de-normalise, scale by
// 1/H[2][2].  The whole translation unit is compiled with -fmad=false and uses only IEEE
// + - * / sqrt in fp64, in the reference's operation order, so the 4-point model is bit-identical
// to the CPU result (tests/test_gpu_parity_h.py checks that against the oracle).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace b2r {

// std::hypot as OpenCV's Jacobi uses it: |a| > |b| ? |a| sqrt(1 + (b/a)^2) : (|b| > 0 ? |b| sqrt(1 + (a/b)^2) : 0).
// Written without branches (operands chosen by selects, one division + square root): the same operations on the same
// values, but the lanes of a warp that hold different matrices do not serialise the two cases.
__device__ __forceinline__ double cv_hypot(double a, double b) {
    a = fabs(a);
    b = fabs(b);
    const bool a_big = a > b;
    const double big = a_big ? a : b, small = a_big ? b : a;
    const double q = small / big;
    const double r = big * sqrt(1 + q * q);
    return (a_big || b > 0) ? r : 0.;
}

// Rotation parameters of one Jacobi step (the serial code's  t = |y| + hypot(p, y); s = hypot(p, t); c = t/s; s = p/s;
// t = (p/t) p;  if (y < 0) s = -s, t = -t)  with one division less on the dependent chain: hypot(p, t) always takes its
// |p| <= t case (t >= hypot(p, y) >= |p| also after rounding), whose quotient q = |p| / t is p/t up to the sign — IEEE
// division rounds the magnitude independently of the signs, so copysign(q, p) == p / t bit for bit.  Finite p, y only.
__device__ __forceinline__ void jacobi_rotation(double p, double y, double& c, double& s, double& t) {
    const double tt = fabs(y) + cv_hypot(p, y);
    const double ap = fabs(p);
    double sh, pt;
    if (ap > tt) {  // not reachable for finite input; the letter of the serial code
        sh = cv_hypot(p, tt);
        pt = p / tt;
    } else {
        const double q = ap / tt;
        sh = tt * sqrt(1 + q * q);
        pt = copysign(q, p);
    }
    c = tt / sh;
    s = p / sh;
    t = pt * p;
    if (y < 0) s = -s, t = -t;
}

// Symmetric eigen-decomposition, n <= 9.  A (n*n, row-major, destroyed), W eigenvalues descending,
// V rows = eigenvectors.  Pivot choice, rotation formulas and tie-breaks follow SURVEY.md A.4.
template <int N>
__device__ void jacobi_eig(double* A, double* W, double* V) {
    int indR[N], indC[N];
    int i, k, l, m;
    for (i = 0; i < N * N; i++) V[i] = 0;
    for (i = 0; i < N; i++) V[i * N + i] = 1;
    for (k = 0; k < N; k++) {
        W[k] = A[k * N + k];
        if (k < N - 1) {
            double mv = fabs(A[k * N + k + 1]);
            m = k + 1;
            for (i = k + 2; i < N; i++) {
                double val = fabs(A[k * N + i]);
                if (mv < val) mv = val, m = i;
            }
            indR[k] = m;
        }
        if (k > 0) {
            double mv = fabs(A[k]);
            m = 0;
            for (i = 1; i < k; i++) {
                double val = fabs(A[i * N + k]);
                if (mv < val) mv = val, m = i;
            }
            indC[k] = m;
        }
    }
    for (int it = 0; it < N * N * 30; it++) {
        double mv = fabs(A[indR[0]]);
        k = 0;
        for (i = 1; i < N - 1; i++) {
            double val = fabs(A[i * N + indR[i]]);
            if (mv < val) mv = val, k = i;
        }
        l = indR[k];
        for (i = 1; i < N; i++) {
            double val = fabs(A[indC[i] * N + i]);
            if (mv < val) mv = val, k = indC[i], l = i;
        }
        double p = A[k * N + l];
        if (fabs(p) <= DBL_EPSILON) break;
        double y = (W[l] - W[k]) * 0.5;
        double t = fabs(y) + cv_hypot(p, y);
        double s = cv_hypot(p, t);
        double c = t / s;
        s = p / s;
        t = (p / t) * p;
        if (y < 0) s = -s, t = -t;
        A[k * N + l] = 0;
        W[k] -= t;
        W[l] += t;
#define B2R_ROT(v0, v1)            \
    {                              \
        double a0 = (v0), b0 = (v1); \
        (v0) = a0 * c - b0 * s;    \
        (v1) = a0 * s + b0 * c;    \
    }
        for (i = 0; i < k; i++) B2R_ROT(A[i * N + k], A[i * N + l]);
        for (i = k + 1; i < l; i++) B2R_ROT(A[k * N + i], A[i * N + l]);
        for (i = l + 1; i < N; i++) B2R_ROT(A[k * N + i], A[l * N + i]);
        for (i = 0; i < N; i++) B2R_ROT(V[k * N + i], V[l * N + i]);
#undef B2R_ROT
        for (int j = 0; j < 2; j++) {
            int idx = j == 0 ? k : l;
            if (idx < N - 1) {
                mv = fabs(A[idx * N + idx + 1]);
                m = idx + 1;
                for (i = idx + 2; i < N; i++) {
                    double val = fabs(A[idx * N + i]);
                    if (mv < val) mv = val, m = i;
                }
                indR[idx] = m;
            }
            if (idx > 0) {
                mv = fabs(A[idx]);
                m = 0;
                for (i = 1; i < idx; i++) {
                    double val = fabs(A[i * N + idx]);
                    if (mv < val) mv = val, m = i;
                }
                indC[idx] = m;
            }
        }
    }
    for (k = 0; k < N - 1; k++) {
        m = k;
        for (i = k + 1; i < N; i++)
            if (W[m] < W[i]) m = i;
        if (k != m) {
            double tmp = W[m];
            W[m] = W[k];
            W[k] = tmp;
            for (i = 0; i < N; i++) {
                tmp = V[m * N + i];
                V[m * N + i] = V[k * N + i];
                V[k * N + i] = tmp;
            }
        }
    }
}

// ---- the same eigen-solver, one thread per matrix on strided shared memory, uniform form -----------------------------------
// Element e of U / V lives at base[e * STRIDE]: with STRIDE = threads per CTA and base = shared-memory column of the
// calling thread, a warp's accesses hit 32 different banks whatever element each lane works on (every solve follows its
// own pivot sequence), while per-thread local arrays would overflow the L1 as soon as a few warps are resident.
// Written as the serial code is (data-dependent loops i < k, k < i < l, i > l; idx+2 <= i < N; i < idx), a warp of 32
// independent decompositions diverges in every loop, the trip counts add up and the index arrays live in local memory
// (tools/microbench_jacobi.cu keeps that form for comparison: 1.9x slower).  This form executes the same arithmetic with
//   * every loop a full unrolled pass over i = 0..N-1 with the element pair chosen by selects and the store predicated
//     (no divergence inside a rotation, all loads of a pass independent of each other),
//   * the searches as tournaments in which the earlier candidate wins ties — the serial scan replaces its maximum only
//     on a strictly greater value, so both pick the first maximal candidate (finite values),
//   * indR / indC as 4-bit fields of two 64-bit registers.
// Operation for operation identical to jacobi_eig<N> on every element.
struct NibbleArray {
    unsigned long long v = 0;
    __device__ __forceinline__ int get(int i) const { return (int)((v >> (4 * i)) & 15ull); }
    __device__ __forceinline__ void set(int i, int m) { v = (v & ~(15ull << (4 * i))) | ((unsigned long long)m << (4 * i)); }
};

// first maximal candidate of val[0..CNT): tournament, the earlier one wins ties; tag[] travels along
template <int CNT>
__device__ __forceinline__ int first_argmax(double* val, int* tag) {
#pragma unroll
    for (int w = 1; w < CNT; w <<= 1)
#pragma unroll
        for (int a = 0; a + w < CNT; a += 2 * w)
            if (val[a] < val[a + w]) val[a] = val[a + w], tag[a] = tag[a + w];
    return tag[0];
}

// U: the upper triangle of A INCLUDING the diagonal, packed row-major (element (r, c), r <= c, at T(r) + c with
// T(r) = r (2N - 1 - r) / 2), N (N + 1) / 2 entries; the serial code never touches the lower triangle and reads the
// diagonal only to initialise W, so W lives on the diagonal: on return U(k, k) is the k-th eigenvalue (descending) and
// row k of V its eigenvector.  126 instead of 171 doubles per decomposition for N = 9: 7 instead of 5 resident warps per SM.
// Precondition: finite entries (normalised-DLT matrices are bounded by construction).
template <int N, int STRIDE>
__device__ void jacobi_eig_packed(double* U, double* V) {
    static_assert(N >= 2 && N <= 15, "4-bit index fields");
    NibbleArray indR, indC;
#define TRI(r) (((r) * (2 * N - 1 - (r))) >> 1)
#define UE(e) U[(e) * STRIDE]
#define VE(e) V[(e) * STRIDE]
    // indR[idx]: first maximal |A[idx][i]|, i > idx;  indC[idx]: first maximal |A[i][idx]|, i < idx.
    // An index outside the searched range is clamped onto the nearest candidate inside it (a duplicate of the first
    // candidate placed before it, or of the last one placed after it, never changes which index wins).
    auto refresh = [&](int idx, int t_idx) {
        double val[N - 1];
        int tag[N - 1];
        if (idx < N - 1) {
#pragma unroll
            for (int i = 1; i < N; i++) {
                const int ci = max(i, idx + 1);
                val[i - 1] = fabs(UE(t_idx + ci));
                tag[i - 1] = ci;
            }
            indR.set(idx, first_argmax<N - 1>(val, tag));
        }
        if (idx > 0) {
            const int t_last = TRI(idx - 1);
#pragma unroll
            for (int i = 0; i < N - 1; i++) {
                const int ci = min(i, idx - 1);
                val[i] = fabs(UE(min(TRI(i), t_last) + idx));   // T is increasing: T(min(i, idx-1)) = min(T(i), T(idx-1))
                tag[i] = ci;
            }
            indC.set(idx, first_argmax<N - 1>(val, tag));
        }
    };
#pragma unroll 1
    for (int i = 0; i < N * N; i++) VE(i) = 0;
#pragma unroll 1
    for (int k = 0; k < N; k++) {
        VE(k * N + k) = 1;
        refresh(k, TRI(k));
    }
#pragma unroll 1
    for (int it = 0; it < N * N * 30; it++) {
        // pivot: rows 0..N-2 through indR, then columns 1..N-1 through indC
        double val[2 * (N - 1)];
        int tag[2 * (N - 1)];
#pragma unroll
        for (int i = 0; i < N - 1; i++) {
            const int r = indR.get(i);
            val[i] = fabs(UE(TRI(i) + r));
            tag[i] = i | (r << 4);
        }
#pragma unroll
        for (int i = 1; i < N; i++) {
            const int c = indC.get(i);
            val[N - 2 + i] = fabs(UE(TRI(c) + i));
            tag[N - 2 + i] = c | (i << 4);
        }
        const int kl = first_argmax<2 * (N - 1)>(val, tag);
        const int k = kl & 15, l = kl >> 4;
        const int tk = TRI(k), tl = TRI(l);
        const double p = UE(tk + l);
        if (fabs(p) <= DBL_EPSILON) break;
        const double wk = UE(tk + k), wl = UE(tl + l);
        double c, s, t;
        jacobi_rotation(p, (wl - wk) * 0.5, c, s, t);
        UE(tk + l) = 0;
        UE(tk + k) = wk - t;
        UE(tl + l) = wl + t;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const int e0 = i < k ? TRI(i) + k : tk + i;
            const int e1 = i < l ? TRI(i) + l : tl + i;
            const double a0 = UE(e0), b0 = UE(e1);
            const double n0 = a0 * c - b0 * s, n1 = a0 * s + b0 * c;
            if (i != k && i != l) {
                UE(e0) = n0;
                UE(e1) = n1;
            }
        }
#pragma unroll
        for (int i = 0; i < N; i++) {
            const double a0 = VE(k * N + i), b0 = VE(l * N + i);
            VE(k * N + i) = a0 * c - b0 * s;
            VE(l * N + i) = a0 * s + b0 * c;
        }
        refresh(k, tk);
        refresh(l, tl);
    }
#pragma unroll 1
    for (int k = 0; k < N - 1; k++) {
        int m = k;
        for (int i = k + 1; i < N; i++)
            if (UE(TRI(m) + m) < UE(TRI(i) + i)) m = i;
        if (k != m) {
            double tmp = UE(TRI(m) + m);
            UE(TRI(m) + m) = UE(TRI(k) + k);
            UE(TRI(k) + k) = tmp;
            for (int i = 0; i < N; i++) {
                tmp = VE(m * N + i);
                VE(m * N + i) = VE(k * N + i);
                VE(k * N + i) = tmp;
            }
        }
    }
#undef TRI
#undef UE
#undef VE
}

// ---- warp-cooperative form of the same eigen-solver ---------------------------------------------------------------
// One solve per WARP, matrices in shared memory.  Pivot choice and the rotation parameters are computed redundantly by
// every lane (broadcast reads, no divergence); lane i then rotates the element pairs of index i — in the serial code
// those pairs are disjoint for different i, so every element goes through exactly the same IEEE operations and the
// result is bit-identical to jacobi_eig<N>.  Kept as the form that follows the serial scans literally (NaN semantics
// included); jacobi_eig_warp3 below is the one the kernels call (1.9x faster: 0.11 ms per 9x9 decomposition).
struct JacobiWarp9 {
    double A[81], V[81], W[9];
    int indR[9], indC[9];
};

template <int N>
__device__ __noinline__ void jacobi_eig_warp(double* A, double* W, double* V, int* indR, int* indC) {
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < N * N; e += 32) V[e] = (e / N == e % N) ? 1. : 0.;
    if (lane < N) {
        const int k = lane;
        W[k] = A[k * N + k];
        if (k < N - 1) {
            double mv = fabs(A[k * N + k + 1]);
            int m = k + 1;
            for (int i = k + 2; i < N; i++) {
                const double val = fabs(A[k * N + i]);
                if (mv < val) mv = val, m = i;
            }
            indR[k] = m;
        }
        if (k > 0) {
            double mv = fabs(A[k]);
            int m = 0;
            for (int i = 1; i < k; i++) {
                const double val = fabs(A[i * N + k]);
                if (mv < val) mv = val, m = i;
            }
            indC[k] = m;
        }
    }
    __syncwarp();
    for (int it = 0; it < N * N * 30; it++) {
        // pivot (every lane, same result)
        double mv = fabs(A[indR[0]]);
        int k = 0, l, i;
        for (i = 1; i < N - 1; i++) {
            const double val = fabs(A[i * N + indR[i]]);
            if (mv < val) mv = val, k = i;
        }
        l = indR[k];
        for (i = 1; i < N; i++) {
            const double val = fabs(A[indC[i] * N + i]);
            if (mv < val) mv = val, k = indC[i], l = i;
        }
        const double p = A[k * N + l];
        if (fabs(p) <= DBL_EPSILON) break;
        const double y = (W[l] - W[k]) * 0.5;
        double t = fabs(y) + cv_hypot(p, y);
        double s = cv_hypot(p, t);
        const double c = t / s;
        s = p / s;
        t = (p / t) * p;
        if (y < 0) s = -s, t = -t;
        __syncwarp();  // all lanes have read A[k][l], W[k], W[l]
        if (lane == 0) {
            A[k * N + l] = 0;
            W[k] -= t;
            W[l] += t;
        }
        if (lane < N) {
            i = lane;
            double *p0 = nullptr, *p1 = nullptr;
            if (i < k) { p0 = &A[i * N + k]; p1 = &A[i * N + l]; }
            else if (i > k && i < l) { p0 = &A[k * N + i]; p1 = &A[i * N + l]; }
            else if (i > l) { p0 = &A[k * N + i]; p1 = &A[l * N + i]; }
            if (p0) {
                const double a0 = *p0, b0 = *p1;
                *p0 = a0 * c - b0 * s;
                *p1 = a0 * s + b0 * c;
            }
            const double v0 = V[k * N + i], v1 = V[l * N + i];
            V[k * N + i] = v0 * c - v1 * s;
            V[l * N + i] = v0 * s + v1 * c;
        }
        __syncwarp();
        if (lane < 4) {  // indR[k], indC[k], indR[l], indC[l]
            const int idx = lane < 2 ? k : l;
            if ((lane & 1) == 0) {
                if (idx < N - 1) {
                    double m2 = fabs(A[idx * N + idx + 1]);
                    int m = idx + 1;
                    for (i = idx + 2; i < N; i++) {
                        const double val = fabs(A[idx * N + i]);
                        if (m2 < val) m2 = val, m = i;
                    }
                    indR[idx] = m;
                }
            } else if (idx > 0) {
                double m2 = fabs(A[idx]);
                int m = 0;
                for (i = 1; i < idx; i++) {
                    const double val = fabs(A[i * N + idx]);
                    if (m2 < val) m2 = val, m = i;
                }
                indC[idx] = m;
            }
        }
        __syncwarp();
    }
    __syncwarp();
    // eigenvalues descending, rows of V alongside
    for (int k = 0; k < N - 1; k++) {
        int m = k;
        for (int i = k + 1; i < N; i++)
            if (W[m] < W[i]) m = i;
        __syncwarp();
        if (k != m) {
            if (lane == 0) {
                const double tmp = W[m];
                W[m] = W[k];
                W[k] = tmp;
            }
            if (lane < N) {
                const double tmp = V[m * N + lane];
                V[m * N + lane] = V[k * N + lane];
                V[k * N + lane] = tmp;
            }
        }
        __syncwarp();
    }
}

// ---- register-resident form: the one the kernels call --------------------------------------------------------------------
// Measured on a B200 (tools/microbench_jacobi.cu) a rotation of the form above costs ~3000 cycles, of which the fp64
// division / square-root chain that cannot be shortened is ~650.  Spreading the two index searches over the lanes
// (tools/microbench_jacobi.cu keeps that intermediate form, 2000 cycles) still leaves two round trips through shared
// memory on the dependent chain: the index arrays are refreshed (load the rotated rows, butterfly, store, barrier) and
// only then the pivot is searched (load indices, load values, reduce).  Here lane i keeps everything that concerns index i in registers — W[i], the row
// candidate (indR[i], A[i][indR[i]]) and the column candidate (indC[i], A[indC[i]][i]) — and the observation that
// removes the refresh from the chain is that the elements a rotation changes are exactly the pairs the lanes have
// just computed: lane i holds the new (i,k) and (i,l) elements, i.e. together the lanes hold the whole of rows and
// columns k and l.  So the next pivot is the maximum over
//   * the own candidates of the lanes other than k and l (value replaced by the fresh pair element when the stored index
//     is k or l, as the serial code would read it), and
//   * the fresh pair elements themselves, which stand for the four candidates being refreshed (row k, column k, row l,
//     column l: a refreshed candidate's value is the maximum of its group, so the overall maximum is unchanged),
// ordered as OpenCV visits them (rows 0..N-2, then columns 1..N-1; inside a refreshed row or column the lowest index),
// found with three warp reductions.  The refreshed indices themselves (needed only from the pivot search after next)
// are computed one rotation later, in the shadow of that rotation's division / square-root chain.
// Same IEEE operations on every element: bit-identical to jacobi_eig<N> for finite matrices (others go to
// jacobi_eig_warp).  indR / indC: shared-memory scratch of the fall-back only.
template <int N>
__device__ __noinline__ void jacobi_eig_warp3(double* A, double* W, double* V, int* indR, int* indC) {
    static_assert(N >= 2 && N <= 9, "lane layout");
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int BIG = 0x7fffffff;
    const int lane = threadIdx.x & 31;
    {
        bool finite = true;
        for (int e = lane; e < N * N; e += 32) finite = finite && fabs(A[e]) <= DBL_MAX;
        if (!__all_sync(FULL, finite)) {
            jacobi_eig_warp<N>(A, W, V, indR, indC);
            return;
        }
    }
    for (int e = lane; e < N * N; e += 32) V[e] = (e / N == e % N) ? 1. : 0.;
    const bool alane = lane < N;
    const int i = alane ? lane : 0;
    const bool vlane = lane >= 16 && lane < 16 + N;
    const int vi = lane - 16;
    double Wreg = A[i * N + i];
    // own candidates: first maximal |A[i][j]|, j > i, and first maximal |A[j][i]|, j < i
    int ind_r = 0, ind_c = 0;
    double val_r = 0, val_c = 0;
    bool ok_r = alane && i < N - 1, ok_c = alane && i > 0;
    if (ok_r) {
        double mv = fabs(A[i * N + i + 1]);
        ind_r = i + 1;
        for (int j = i + 2; j < N; j++) {
            const double v = fabs(A[i * N + j]);
            if (mv < v) mv = v, ind_r = j;
        }
        val_r = A[i * N + ind_r];
    }
    if (ok_c) {
        double mv = fabs(A[i]);
        ind_c = 0;
        for (int j = 1; j < i; j++) {
            const double v = fabs(A[j * N + i]);
            if (mv < v) mv = v, ind_c = j;
        }
        val_c = A[ind_c * N + i];
    }
    __syncwarp();
    bool have_grp = false;          // the previous rotation's pair elements stand for its four refreshed candidates
    int gk = 0, gl = 0;             // ... its pivot
    double na0 = 0, nb0 = 0;        // ... this lane's new (i, gk) and (i, gl) elements

    for (int it = 0; it < N * N * 30; it++) {
        // ---- account for the last rotation: its pivot lanes are represented by the pair elements, the others re-read ----
        if (have_grp) {
            if (i == gk || i == gl) {
                ok_r = false;
                ok_c = false;
            } else {
                if (ok_r) val_r = ind_r == gk ? na0 : (ind_r == gl ? nb0 : val_r);
                if (ok_c) val_c = ind_c == gk ? na0 : (ind_c == gl ? nb0 : val_c);
            }
        }
        // ---- pivot: best of this lane's entries, then best of the warp ----
        // four entries per lane at most; keys = bit pattern of |value| (-1: no entry), ties to the lower visiting order
        const bool g_a = have_grp && alane && i != gk, g_b = have_grp && alane && i != gl;
        const long long k1 = ok_r ? __double_as_longlong(fabs(val_r)) : -1ll, k2 = ok_c ? __double_as_longlong(fabs(val_c)) : -1ll;
        const long long k3 = g_a ? __double_as_longlong(fabs(na0)) : -1ll, k4 = g_b ? __double_as_longlong(fabs(nb0)) : -1ll;
        const int o1 = i * 16, o2 = (N - 2 + i) * 16;
        const int o3 = (i > gk ? gk : N - 2 + gk) * 16 + i, o4 = (i > gl ? gl : N - 2 + gl) * 16 + i;
        const int kl1 = i | (ind_r << 8), kl2 = ind_c | (i << 8);
        const int kl3 = i > gk ? (gk | (i << 8)) : (i | (gk << 8)), kl4 = i > gl ? (gl | (i << 8)) : (i | (gl << 8));
        const bool s12 = k2 > k1 || (k2 == k1 && o2 < o1);      // entry 2 beats entry 1
        const bool s34 = k4 > k3 || (k4 == k3 && o4 < o3);
        const long long ka = s12 ? k2 : k1, kb = s34 ? k4 : k3;
        const int oa = s12 ? o2 : o1, ob = s34 ? o4 : o3;
        const double va = s12 ? val_c : val_r, vb = s34 ? nb0 : na0;
        const int kla = s12 ? kl2 : kl1, klb = s34 ? kl4 : kl3;
        const bool sab = kb > ka || (kb == ka && ob < oa);
        const long long kbest = sab ? kb : ka;
        const double bv = sab ? vb : va;
        const int bo = kbest < 0 ? BIG : (sab ? ob : oa), bkl = sab ? klb : kla;
        const bool any = bo != BIG;
        const double abv = fabs(bv);
        const unsigned hi = any ? (unsigned)__double2hiint(abv) : 0u, lo = any ? (unsigned)__double2loint(abv) : 0u;
        const unsigned mh = __reduce_max_sync(FULL, hi);
        const unsigned ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
        const bool holder = any && hi == mh && lo == ml;
        const int mo = __reduce_min_sync(FULL, holder ? bo : BIG);
        const int win = __ffs(__ballot_sync(FULL, holder && bo == mo)) - 1;
        const int kl = __shfl_sync(FULL, bkl, win);
        const double p = __shfl_sync(FULL, bv, win);
        const int k = kl & 255, l = kl >> 8;
        if (fabs(p) <= DBL_EPSILON) break;
        const double Wk = __shfl_sync(FULL, Wreg, k), Wl = __shfl_sync(FULL, Wreg, l);
        // ---- operands of the rotation (shared memory is current: barrier at the end of the previous rotation) ----
        const bool rot_a = alane && i != k && i != l;
        int e0 = k * N + i, e1 = l * N + i;
        if (alane) {
            if (i < k) e0 = i * N + k;
            if (i < l) e1 = i * N + l;
        } else if (vlane) {
            e0 = k * N + vi;
            e1 = l * N + vi;
        }
        double a0 = 0, b0 = 0;
        if (rot_a) { a0 = A[e0]; b0 = A[e1]; }
        if (vlane) { a0 = V[e0]; b0 = V[e1]; }
        // ---- rotation parameters (jacobi_rotation with the two quotients by hypot(p, t) on different lanes), and, in the
        // issue slots its three divisions and two square roots leave idle, the indices refreshed by the PREVIOUS rotation:
        // four groups — row gk (i > gk) and column gk (i < gk) over the (i, gk) elements, row gl and column gl over the
        // (i, gl) ones; first maximum = lowest lane.  The stages are placed between the long operations by hand: the
        // compiler does not move instructions across the slow-path branches of a division.
        const bool in_rk = alane && i > gk, in_ck = alane && i < gk, in_rl = alane && i > gl, in_cl = alane && i < gl;
        const double aa = fabs(na0), ab = fabs(nb0);
        const unsigned ha = (unsigned)__double2hiint(aa), la = (unsigned)__double2loint(aa);
        const unsigned hb = (unsigned)__double2hiint(ab), lb = (unsigned)__double2loint(ab);
        const double y = (Wl - Wk) * 0.5;
        const double ap = fabs(p), ay = fabs(y);
        const bool p_big = ap > ay;
        const double big1 = p_big ? ap : ay, small1 = p_big ? ay : ap;
        // stage 1: high words
        const unsigned h_rk = __reduce_max_sync(FULL, in_rk ? ha : 0u), h_ck = __reduce_max_sync(FULL, in_ck ? ha : 0u);
        const unsigned h_rl = __reduce_max_sync(FULL, in_rl ? hb : 0u), h_cl = __reduce_max_sync(FULL, in_cl ? hb : 0u);
        const double q1 = small1 / big1;                           // cv_hypot(p, y), written out
        // stage 2: low words among the holders of the maximal high word
        const unsigned l_rk = __reduce_max_sync(FULL, in_rk && ha == h_rk ? la : 0u), l_ck = __reduce_max_sync(FULL, in_ck && ha == h_ck ? la : 0u);
        const unsigned l_rl = __reduce_max_sync(FULL, in_rl && hb == h_rl ? lb : 0u), l_cl = __reduce_max_sync(FULL, in_cl && hb == h_cl ? lb : 0u);
        const double r1 = big1 * sqrt(1 + q1 * q1);
        const double hyp1 = (p_big || ay > 0) ? r1 : 0.;
        const double tt = ay + hyp1;
        // stage 3: lowest holder of each group
        const int w_rk = __ffs(__ballot_sync(FULL, in_rk && ha == h_rk && la == l_rk)) - 1;
        const int w_ck = __ffs(__ballot_sync(FULL, in_ck && ha == h_ck && la == l_ck)) - 1;
        const int w_rl = __ffs(__ballot_sync(FULL, in_rl && hb == h_rl && lb == l_rl)) - 1;
        const int w_cl = __ffs(__ballot_sync(FULL, in_cl && hb == h_cl && lb == l_cl)) - 1;
        const double q = ap / tt;                                  // |p| <= tt: hypot(p, tt) takes this case
        // stage 4: the winners' values
        const double x_rk = __shfl_sync(FULL, na0, w_rk < 0 ? 0 : w_rk), x_ck = __shfl_sync(FULL, na0, w_ck < 0 ? 0 : w_ck);
        const double x_rl = __shfl_sync(FULL, nb0, w_rl < 0 ? 0 : w_rl), x_cl = __shfl_sync(FULL, nb0, w_cl < 0 ? 0 : w_cl);
        const double sh = tt * sqrt(1 + q * q);
        if (have_grp) {   // lanes gk, gl get their refreshed candidates back (they sat out the pivot search above)
            if (alane && i == gk) { ind_r = w_rk; val_r = x_rk; ok_r = w_rk >= 0; ind_c = w_ck; val_c = x_ck; ok_c = w_ck >= 0; }
            if (alane && i == gl) { ind_r = w_rl; val_r = x_rl; ok_r = w_rl >= 0; ind_c = w_cl; val_c = x_cl; ok_c = w_cl >= 0; }
        }
        const double quo = ((lane & 1) ? p : tt) / sh;
        const double c = __shfl_sync(FULL, quo, 0);
        double sn = __shfl_sync(FULL, quo, 1);
        double t = copysign(q, p) * p;
        if (y < 0) sn = -sn, t = -t;
        // ---- rotate ----
        const double n0 = a0 * c - b0 * sn, n1 = a0 * sn + b0 * c;
        if (rot_a) { A[e0] = n0; A[e1] = n1; }
        if (vlane) { V[e0] = n0; V[e1] = n1; }
        na0 = n0;
        nb0 = n1;
        if (alane && i == k) { Wreg = Wreg - t; nb0 = 0; A[k * N + l] = 0; }   // A[k][l] = 0 belongs to column l's group
        if (alane && i == l) { Wreg = Wreg + t; na0 = 0; }                      // ... and to row k's
        gk = k;
        gl = l;
        have_grp = true;
        __syncwarp();
    }
    if (alane) W[i] = Wreg;
    __syncwarp();
    // eigenvalues descending, rows of V alongside
    for (int k = 0; k < N - 1; k++) {
        int m = k;
        for (int j = k + 1; j < N; j++)
            if (W[m] < W[j]) m = j;
        __syncwarp();
        if (k != m) {
            if (lane == 0) {
                const double tmp = W[m];
                W[m] = W[k];
                W[k] = tmp;
            }
            if (lane < N) {
                const double tmp = V[m * N + lane];
                V[m * N + lane] = V[k * N + lane];
                V[k * N + lane] = tmp;
            }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void mat3_mul(const double* a, const double* b, double* o) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += a[i * 3 + k] * b[k * 3 + j];
            o[i * 3 + j] = s;
        }
}

// Normalisation statistics of a point set: centroids and inverse mean absolute deviations.
struct HNorm {
    double cMx, cMy, cmx, cmy;  // centroids of src (M) and dst (m)
    double sMx, sMy, smx, smy;  // count / sum |coord - centroid|
};

// de-normalise the 9-vector of the smallest eigenvalue and scale by 1/H[2][2]
__device__ __forceinline__ void h_from_eigvec(const double* vec, const HNorm& nm, double* H) {
    const double invHnorm[9] = {1. / nm.smx, 0, nm.cmx, 0, 1. / nm.smy, nm.cmy, 0, 0, 1};
    const double Hnorm2[9] = {nm.sMx, 0, -nm.cMx * nm.sMx, 0, nm.sMy, -nm.cMy * nm.sMy, 0, 0, 1};
    double Ht[9], H0[9];
    mat3_mul(invHnorm, vec, Ht);
    mat3_mul(Ht, Hnorm2, H0);
    const double sc = 1. / H0[8];
    for (int i = 0; i < 9; i++) H[i] = H0[i] * sc;
}

// L^T L (full 9x9, upper triangle accumulated then mirrored by the caller) -> H.  Returns false when
// the normalisation is degenerate.  LtL is destroyed.
__device__ __forceinline__ void h_from_LtL(double* LtL, const HNorm& nm, double* H) {
    double W[9], V[81];
    for (int j = 0; j < 9; j++)
        for (int k = 0; k < j; k++) LtL[j * 9 + k] = LtL[k * 9 + j];
    jacobi_eig<9>(LtL, W, V);
    h_from_eigvec(V + 72, nm, H);
}

__device__ __forceinline__ void h_accumulate_LtL(double* LtL, const HNorm& nm, double Mx, double My, double mx,
                                                 double my) {
    const double x = (mx - nm.cmx) * nm.smx, y = (my - nm.cmy) * nm.smy;
    const double X = (Mx - nm.cMx) * nm.sMx, Y = (My - nm.cMy) * nm.sMy;
    const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
    const double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
    for (int j = 0; j < 9; j++)
        for (int k = j; k < 9; k++) LtL[j * 9 + k] += Lx[j] * Lx[k] + Ly[j] * Ly[k];
}

// 4-point solve.  M = src (x,y) x4, m = dst (x,y) x4 as fp32.  Returns 1 model or 0.
static __device__ int h_solve4(const float* M, const float* m, double* H) {
    HNorm nm = {0, 0, 0, 0, 0, 0, 0, 0};
    const int count = 4;
    for (int i = 0; i < count; i++) {
        nm.cmx += m[2 * i];
        nm.cmy += m[2 * i + 1];
        nm.cMx += M[2 * i];
        nm.cMy += M[2 * i + 1];
    }
    nm.cmx /= count;
    nm.cmy /= count;
    nm.cMx /= count;
    nm.cMy /= count;
    for (int i = 0; i < count; i++) {
        nm.smx += fabs(m[2 * i] - nm.cmx);
        nm.smy += fabs(m[2 * i + 1] - nm.cmy);
        nm.sMx += fabs(M[2 * i] - nm.cMx);
        nm.sMy += fabs(M[2 * i + 1] - nm.cMy);
    }
    if (fabs(nm.smx) < DBL_EPSILON || fabs(nm.smy) < DBL_EPSILON || fabs(nm.sMx) < DBL_EPSILON ||
        fabs(nm.sMy) < DBL_EPSILON)
        return 0;
    nm.smx = count / nm.smx;
    nm.smy = count / nm.smy;
    nm.sMx = count / nm.sMx;
    nm.sMy = count / nm.sMy;
    double LtL[81];
    for (int i = 0; i < 81; i++) LtL[i] = 0;
    for (int i = 0; i < count; i++) h_accumulate_LtL(LtL, nm, M[2 * i], M[2 * i + 1], m[2 * i], m[2 * i + 1]);
    h_from_LtL(LtL, nm, H);
    return 1;
}

// 4-point solve on strided (shared-memory) storage: ws = this thread's column of a [H4_WS_DOUBLES][STRIDE] double
// workspace (packed upper triangle of L^T L 45 | V 81).  Bit-identical to h_solve4.
constexpr int H4_WS_DOUBLES = 45 + 81;
template <int STRIDE>
__device__ __forceinline__ int h_solve4_strided(double* ws, const float* M, const float* m, double* H) {
    HNorm nm = {0, 0, 0, 0, 0, 0, 0, 0};
    const int count = 4;
    for (int i = 0; i < count; i++) {
        nm.cmx += m[2 * i];
        nm.cmy += m[2 * i + 1];
        nm.cMx += M[2 * i];
        nm.cMy += M[2 * i + 1];
    }
    nm.cmx /= count; nm.cmy /= count; nm.cMx /= count; nm.cMy /= count;
    for (int i = 0; i < count; i++) {
        nm.smx += fabs(m[2 * i] - nm.cmx);
        nm.smy += fabs(m[2 * i + 1] - nm.cmy);
        nm.sMx += fabs(M[2 * i] - nm.cMx);
        nm.sMy += fabs(M[2 * i + 1] - nm.cMy);
    }
    if (fabs(nm.smx) < DBL_EPSILON || fabs(nm.smy) < DBL_EPSILON || fabs(nm.sMx) < DBL_EPSILON || fabs(nm.sMy) < DBL_EPSILON)
        return 0;
    nm.smx = count / nm.smx; nm.smy = count / nm.smy; nm.sMx = count / nm.sMx; nm.sMy = count / nm.sMy;
    double* U = ws;
    double* V = ws + 45 * STRIDE;
    for (int e = 0; e < 45; e++) U[e * STRIDE] = 0;
    for (int i = 0; i < count; i++) {
        const double x = ((double)m[2 * i] - nm.cmx) * nm.smx, y = ((double)m[2 * i + 1] - nm.cmy) * nm.smy;
        const double X = ((double)M[2 * i] - nm.cMx) * nm.sMx, Y = ((double)M[2 * i + 1] - nm.cMy) * nm.sMy;
        const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
        const double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
#pragma unroll
        for (int j = 0; j < 9; j++)
#pragma unroll
            for (int k = j; k < 9; k++) U[(((j * (17 - j)) >> 1) + k) * STRIDE] += Lx[j] * Lx[k] + Ly[j] * Ly[k];
    }
    jacobi_eig_packed<9, STRIDE>(U, V);
    double vec[9];
    for (int i = 0; i < 9; ++i) vec[i] = V[(72 + i) * STRIDE];
    h_from_eigvec(vec, nm, H);
    return 1;
}

// Warp-cooperative 4-point / k-point solve.  M, m: the fp32 points (count of them, every lane passes the same pointers);
// jw: the warp's shared-memory workspace.  Every lane returns the same flag and H.  Bit-identical to h_solve4.
__device__ __forceinline__ void h_from_LtL_warp(JacobiWarp9& jw, const HNorm& nm, double* H) {
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < 81; e += 32) {  // mirror the upper triangle down
        const int j = e / 9, k = e % 9;
        if (k < j) jw.A[e] = jw.A[k * 9 + j];
    }
    __syncwarp();
    jacobi_eig_warp3<9>(jw.A, jw.W, jw.V, jw.indR, jw.indC);
    double vec[9];
    for (int i = 0; i < 9; ++i) vec[i] = jw.V[72 + i];
    h_from_eigvec(vec, nm, H);
    __syncwarp();
}

__device__ __forceinline__ int h_solve4_warp(JacobiWarp9& jw, const float* M, const float* m, double* H) {
    const int lane = threadIdx.x & 31;
    HNorm nm = {0, 0, 0, 0, 0, 0, 0, 0};
    const int count = 4;
    for (int i = 0; i < count; i++) {
        nm.cmx += m[2 * i];
        nm.cmy += m[2 * i + 1];
        nm.cMx += M[2 * i];
        nm.cMy += M[2 * i + 1];
    }
    nm.cmx /= count; nm.cmy /= count; nm.cMx /= count; nm.cMy /= count;
    for (int i = 0; i < count; i++) {
        nm.smx += fabs(m[2 * i] - nm.cmx);
        nm.smy += fabs(m[2 * i + 1] - nm.cmy);
        nm.sMx += fabs(M[2 * i] - nm.cMx);
        nm.sMy += fabs(M[2 * i + 1] - nm.cMy);
    }
    if (fabs(nm.smx) < DBL_EPSILON || fabs(nm.smy) < DBL_EPSILON || fabs(nm.sMx) < DBL_EPSILON || fabs(nm.sMy) < DBL_EPSILON)
        return 0;
    nm.smx = count / nm.smx; nm.smy = count / nm.smy; nm.sMx = count / nm.sMx; nm.sMy = count / nm.sMy;
    // entry (j,k), j <= k, of L^T L: the same per-point sums, in point order, as h_accumulate_LtL
    for (int e = lane; e < 81; e += 32) {
        const int j = e / 9, k = e % 9;
        if (j > k) continue;
        double acc = 0;
        for (int i = 0; i < count; i++) {
            const double x = ((double)m[2 * i] - nm.cmx) * nm.smx, y = ((double)m[2 * i + 1] - nm.cmy) * nm.smy;
            const double X = ((double)M[2 * i] - nm.cMx) * nm.sMx, Y = ((double)M[2 * i + 1] - nm.cMy) * nm.sMy;
            const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
            const double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
            acc += Lx[j] * Lx[k] + Ly[j] * Ly[k];
        }
        jw.A[e] = acc;
    }
    __syncwarp();
    h_from_LtL_warp(jw, nm, H);
    return 1;
}

// Fast 4-point solve (throughput mode, not bit-compatible with the Jacobi path): closed form through the
// projective basis of each quadruple, H ~ Q * diag(b_i / a_i) * adj(P), P = [p1 p2 p3] (columns, homogeneous
// source points), a = adj(P) p4, and the same (Q, b) for the destination points; ~150 fp64 operations, all in
// registers.  Agrees with h_solve4 to ~1e-12 relative on well-conditioned samples.
__device__ __forceinline__ int h_solve4_fast(const float* M, const float* m, double* H) {
    double P[4][2], Qd[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        P[i][0] = M[2 * i]; P[i][1] = M[2 * i + 1];
        Qd[i][0] = m[2 * i]; Qd[i][1] = m[2 * i + 1];
    }
    // rows of adj(P): p2 x p3, p3 x p1, p1 x p2 with p = (x, y, 1)
    double adj[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int i = (r + 1) % 3, j = (r + 2) % 3;
        adj[r][0] = P[i][1] - P[j][1];
        adj[r][1] = P[j][0] - P[i][0];
        adj[r][2] = P[i][0] * P[j][1] - P[j][0] * P[i][1];
    }
    double s[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int i = (r + 1) % 3, j = (r + 2) % 3;
        const double a = adj[r][0] * P[3][0] + adj[r][1] * P[3][1] + adj[r][2];
        const double b = (Qd[i][1] - Qd[j][1]) * Qd[3][0] + (Qd[j][0] - Qd[i][0]) * Qd[3][1] +
                         (Qd[i][0] * Qd[j][1] - Qd[j][0] * Qd[i][1]);
        s[r] = b / a;
    }
    double Hp[9];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double t0 = s[0] * adj[0][c], t1 = s[1] * adj[1][c], t2 = s[2] * adj[2][c];
        Hp[0 * 3 + c] = Qd[0][0] * t0 + Qd[1][0] * t1 + Qd[2][0] * t2;
        Hp[1 * 3 + c] = Qd[0][1] * t0 + Qd[1][1] * t1 + Qd[2][1] * t2;
        Hp[2 * 3 + c] = t0 + t1 + t2;
    }
    const double sc = 1. / Hp[8];
    bool finite = true;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        H[i] = Hp[i] * sc;
        finite &= (fabs(H[i]) <= DBL_MAX);
    }
    return finite ? 1 : 0;
}

// -- degeneracy test of a 4-point sample (checkSubset) -------------------------------
__device__ __forceinline__ bool have_collinear4(const float* p) {
    const int i = 3;
    for (int j = 0; j < i; j++) {
        const double dx1 = __fsub_rn(p[2 * j], p[2 * i]);
        const double dy1 = __fsub_rn(p[2 * j + 1], p[2 * i + 1]);
        for (int k = 0; k < j; k++) {
            const double dx2 = __fsub_rn(p[2 * k], p[2 * i]);
            const double dy2 = __fsub_rn(p[2 * k + 1], p[2 * i + 1]);
            if (fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2)))
                return true;
        }
    }
    return false;
}

__device__ __forceinline__ double det3_pts(const float* p, int a, int b, int c) {
    const double a00 = p[2 * a], a01 = p[2 * a + 1], a10 = p[2 * b], a11 = p[2 * b + 1], a20 = p[2 * c],
                 a21 = p[2 * c + 1];
    return a00 * (a11 * 1. - a21 * 1.) - a01 * (a10 * 1. - a20 * 1.) + 1. * (a10 * a21 - a20 * a11);
}

__device__ __forceinline__ bool h_check_subset4(const float* ms1, const float* ms2) {
    if (have_collinear4(ms1) || have_collinear4(ms2)) return false;
    int negative = 0;
    negative += det3_pts(ms1, 0, 1, 2) * det3_pts(ms2, 0, 1, 2) < 0;
    negative += det3_pts(ms1, 1, 2, 3) * det3_pts(ms2, 1, 2, 3) < 0;
    negative += det3_pts(ms1, 0, 2, 3) * det3_pts(ms2, 0, 2, 3) < 0;
    negative += det3_pts(ms1, 0, 1, 3) * det3_pts(ms2, 0, 1, 3) < 0;
    return negative == 0 || negative == 4;
}

}  // namespace b2r

// Register-file access probe for the operand patterns of the pair-mapped scoring loop (score_h.cuh: k3_score_h_pairs):
// warp-instructions per clock per SMSP of FFMA2 streams whose operands come from
//   A: a 64-bit pair shared by G consecutive instructions (.reuse candidate) or distinct pairs
//   B: a 32-bit broadcast scalar (.F32), distinct per instruction, or a shared 64-bit pair, or distinct pairs
//   C: a 32-bit broadcast scalar, an immediate, or a distinct 64-bit accumulator
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I code-reproduction-ransac_b200/csrc tools/regprobe2.cu -o tools/regprobe2
#include <cstdio>
#include <cstdlib>
#include "f32x2.cuh"
using namespace b2r;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
constexpr int G = 12;
enum { SECOND_LEVEL, FIRST_LEVEL, FIRST_LEVEL_IMM, A_TYPE, Q_TYPE, T_TYPE, ALL_DISTINCT, ACC_ONLY, A_TYPE_SCALAR, FULL_EVAL, NKIND };
static const char* names[] = {"acc = P.shared * s_i.F32 + acc_i       (second level)", "acc = P.shared * s_i.F32 + t_i.F32      (first level)",
    "acc = P.shared * s_i.F32 + 1.0          (first level, immediate)", "acc = w_i * Nu.shared + sx_i            (a, b)",
    "acc = b_i * b_i + t_i                   (q, m)", "acc = w_i * w_i                         (t)", "acc = x_i * y_i + z_i  all distinct pairs",
    "acc = acc_i * c.inv + d.inv             (accumulator only)", "acc = w_i * nu.F32 + sx_i               (a, b, point-major mapping)",
    "the 11-operation evaluation, 4 hypotheses x 1 point pair"};
template <int KIND>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float a, float b) {
    f2_t acc[G], w[G], sx[G];
    float s[G], t[G];
    f2_t P = f2_pack(a, b), Nu = f2_pack(b, a);
#pragma unroll
    for (int i = 0; i < G; ++i) {
        acc[i] = f2_pack(a + i, b - i); w[i] = f2_pack(b + 0.5f * i, a + threadIdx.x); sx[i] = f2_pack(a * i + threadIdx.x * 1e-4f, b + i - threadIdx.x * 1e-4f);
        s[i] = a + 0.25f * i + threadIdx.x * 1e-3f; t[i] = b - 0.125f * i + threadIdx.x * 2e-3f;
    }
    float px = a;
    for (int it = 0; it < iters; ++it) {
        px = px * 1.0000001f;               // one scalar op per G packed ops keeps P loop-variant
        P = f2_pack(px, b); Nu = f2_pack(b, px);
        if (KIND == FULL_EVAL) {
            const f2_t X = P, Y = Nu, nu = f2_pack(px, px), nv = f2_pack(b, px), one = f2_dup(1.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const f2_t ww = f2_fma(X, f2_dup(s[j]), f2_fma(Y, f2_dup(t[j]), one));
                const f2_t x1 = f2_fma(X, f2_dup(s[4 + j]), f2_fma(Y, f2_dup(t[4 + j]), f2_dup(s[8 + j])));
                const f2_t y1 = f2_fma(X, f2_dup(t[8 + j]), f2_fma(Y, f2_dup(s[j] + 1.f), f2_dup(t[j] + 1.f)));
                const f2_t aa = f2_fma(ww, nu, x1), bb = f2_fma(ww, nv, y1), tt = f2_mul(ww, ww);
                float t0, t1; f2_unpack(tt, t0, t1);
                acc[j] = f2_fma(aa, aa, f2_fma(bb, bb, f2_fma(acc[j], f2_dup(1e-9f), f2_pack(-t0, -t1))));   // 12 packed ops
            }
            continue;
        }
#pragma unroll
        for (int i = 0; i < G; ++i) {
            if (KIND == SECOND_LEVEL) acc[i] = f2_fma(P, f2_dup(s[i]), acc[i]);
            if (KIND == FIRST_LEVEL) acc[i] = f2_add(acc[i], f2_fma(P, f2_dup(s[i]), f2_dup(t[i])));   // 2 ops: the fma + an add to keep it
            if (KIND == FIRST_LEVEL_IMM) acc[i] = f2_add(acc[i], f2_fma(P, f2_dup(s[i]), f2_dup(1.0f)));
            if (KIND == A_TYPE) acc[i] = f2_fma(acc[i], Nu, sx[i]);
            if (KIND == Q_TYPE) acc[i] = f2_fma(acc[i], acc[i], sx[i]);
            if (KIND == T_TYPE) acc[i] = f2_mul(acc[i], acc[i]);
            if (KIND == ALL_DISTINCT) acc[i] = f2_fma(w[i], sx[(i + 1) % G], acc[i]);
            if (KIND == ACC_ONLY) acc[i] = f2_fma(acc[i], P, Nu);
            if (KIND == A_TYPE_SCALAR) acc[i] = f2_fma(acc[i], f2_dup(px), sx[i]);
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < G; ++i) { float lo, hi; f2_unpack(acc[i], lo, hi); r += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int KIND>
static void run(int nsm) {
    const int ctas = nsm * 4, threads = 256, iters = 16384;
    float* out; CK(cudaMalloc(&out, sizeof(float) * ctas * threads));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int wu = 0; wu < 3; ++wu) probe<KIND><<<ctas, threads>>>(out, iters, 1.0001f, 0.5f);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); probe<KIND><<<ctas, threads>>>(out, iters, 1.0001f, 0.5f); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double ops_per_iter = KIND == FULL_EVAL ? 48.0 : ((KIND == FIRST_LEVEL || KIND == FIRST_LEVEL_IMM) ? 2.0 * G : (double)G);
    const double instr = (double)iters * ops_per_iter * ctas * threads / 32.0;   // packed warp-instructions
    printf("{\"pattern\": \"%s\", \"ms\": %.4f, \"packed_instr_per_clk_per_smsp\": %.3f}\n", names[KIND], best,
           instr / (best * 1e-3) / (nsm * 4) / 1.965e9);
    fflush(stdout);
    CK(cudaFree(out));
}
template <int K> struct Loop { static void go(int nsm) { run<K>(nsm); Loop<K + 1>::go(nsm); } };
template <> struct Loop<NKIND> { static void go(int) {} };
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    Loop<0>::go(prop.multiProcessorCount);
    return 0;
}

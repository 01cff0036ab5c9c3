// Register-file bandwidth probe: FFMA / FFMA2 with 3 distinct register operands vs reusable/uniform operands.
#include <cstdio>
#include <cstdlib>
#include "f32x2.cuh"
using namespace b2r;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
constexpr int ILP = 8;
__constant__ float2 cpts[64];
enum { S3, S2R, P3, P2R, P_SHARED_B, P_UR, NK };
static const char* names[] = {"ffma 3 distinct regs", "ffma 2 regs + invariant", "ffma2 3 distinct regs", "ffma2 2 regs + invariant",
                              "ffma2 b shared by 8 consecutive", "ffma2 b from constant bank (UR)"};
template <int KIND>
__global__ void __launch_bounds__(256) probe(const float* in, float* out, int iters) {
    float x[ILP], y[ILP], z[ILP]; f2_t X[ILP], Y[ILP], Z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        x[i] = in[threadIdx.x + i]; y[i] = in[threadIdx.x + 32 + i]; z[i] = in[threadIdx.x + 64 + i];
        X[i] = f2_pack(x[i], y[i]); Y[i] = f2_pack(y[i], z[i]); Z[i] = f2_pack(z[i], x[i]);
    }
    f2_t inv = Y[0];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            f2_t sh = Y[u];
            f2_t cu = f2_pack(cpts[(it + u) & 63].x, cpts[(it + u) & 63].y);
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (KIND == S3) x[i] = __fmaf_rn(y[i], z[i], x[i]);
                if (KIND == S2R) x[i] = __fmaf_rn(y[i], z[0], x[i]);
                if (KIND == P3) X[i] = f2_fma(Y[i], Z[i], X[i]);
                if (KIND == P2R) X[i] = f2_fma(Y[i], inv, X[i]);
                if (KIND == P_SHARED_B) X[i] = f2_fma(Z[i], sh, X[i]);
                if (KIND == P_UR) X[i] = f2_fma(Z[i], cu, X[i]);
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float lo, hi; f2_unpack(X[i], lo, hi); acc += lo + hi + x[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int KIND> static void run(int nsm, const float* in) {
    const int ctas = nsm * 8, threads = 256, iters = 8192;
    float* out; CK(cudaMalloc(&out, sizeof(float) * ctas * threads));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 5; ++w) probe<KIND><<<ctas, threads>>>(in, out, iters);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); probe<KIND><<<ctas, threads>>>(in, out, iters); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double steps = (double)iters * 4 * ILP * ctas * threads;
    printf("{\"probe\": \"%s\", \"ms\": %.4f, \"warp_instr_per_clk_per_smsp_at_1965\": %.3f}\n", names[KIND], best,
           steps / (best * 1e-3) / 32.0 / (nsm * 4) / 1.965e9);
    fflush(stdout); CK(cudaFree(out));
}
template <int K> struct Loop { static void go(int nsm, const float* in) { run<K>(nsm, in); Loop<K + 1>::go(nsm, in); } };
template <> struct Loop<NK> { static void go(int, const float*) {} };
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    float h[512]; for (int i = 0; i < 512; ++i) h[i] = 1.0f + 1e-6f * i;
    float* in; CK(cudaMalloc(&in, sizeof(h))); CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
    float2 c[64]; for (int i = 0; i < 64; ++i) c[i] = make_float2(1.f + 1e-6f * i, 1.f - 1e-6f * i);
    CK(cudaMemcpyToSymbol(cpts, c, sizeof(c)));
    Loop<0>::go(prop.multiProcessorCount, in);
    return 0;
}

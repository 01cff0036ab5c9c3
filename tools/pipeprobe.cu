// Pipe-throughput probes for sm_100a FP32 datapaths (scalar vs packed f32x2, MUFU, ALU co-issue).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I code-reproduction-ransac_b200/csrc tools/pipeprobe.cu -o tools/pipeprobe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "f32x2.cuh"
using namespace b2r;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
constexpr int ILP = 8;
enum { FFMA, FADD, FMUL, FFMA2, FADD2, FMUL2, MUFU, MIX_FMUL2_FADD2, MIX_FFMA2_FADD2, MIX_FFMA2_FMUL2, MIX_FFMA2_FFMA,
       MIX_FFMA2_FSETP, MIX_FFMA2_MUFU, MIX_FADD2_FADD, MIX_FFMA2_IADD, NKIND };
static const char* names[] = {"ffma", "fadd", "fmul", "ffma2", "fadd2", "fmul2", "mufu_rcp", "fmul2+fadd2", "ffma2+fadd2",
    "ffma2+fmul2", "ffma2+ffma", "ffma2+fsetp_iadd", "4ffma2+mufu", "fadd2+fadd", "ffma2+iadd"};
// lane-level fp32 operations (or mufu ops / alu ops) per inner step per thread, for reporting
template <int KIND>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float a, float b) {
    float x[ILP]; f2_t y[ILP]; int c[ILP];
    f2_t a2 = f2_pack(a, a * 1.0001f), b2 = f2_pack(b, b * 0.999f);
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = a + (float)(threadIdx.x + i); y[i] = f2_pack(a + (float)(threadIdx.x + i), b + (float)i); c[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (KIND == FFMA) x[i] = __fmaf_rn(x[i], a, b);
                if (KIND == FADD) x[i] = __fadd_rn(x[i], b);
                if (KIND == FMUL) x[i] = __fmul_rn(x[i], a);
                if (KIND == FFMA2) y[i] = f2_fma(y[i], a2, b2);
                if (KIND == FADD2) y[i] = f2_add(y[i], b2);
                if (KIND == FMUL2) y[i] = f2_mul(y[i], a2);
                if (KIND == MUFU) x[i] = rcp_approx(x[i]) + 0.0f;  // FADD keeps the chain from being folded
                if (KIND == MIX_FMUL2_FADD2) y[i] = (u & 1) ? f2_mul(y[i], a2) : f2_add(y[i], b2);
                if (KIND == MIX_FFMA2_FADD2) y[i] = (u & 1) ? f2_fma(y[i], a2, b2) : f2_add(y[i], b2);
                if (KIND == MIX_FFMA2_FMUL2) y[i] = (u & 1) ? f2_fma(y[i], a2, b2) : f2_mul(y[i], a2);
                if (KIND == MIX_FFMA2_FFMA) { y[i] = f2_fma(y[i], a2, b2); x[i] = __fmaf_rn(x[i], a, b); }
                if (KIND == MIX_FFMA2_FSETP) { y[i] = f2_fma(y[i], a2, b2); float lo, hi; f2_unpack(y[i], lo, hi); c[i] += (lo <= b) ? 1 : 0; }
                if (KIND == MIX_FFMA2_MUFU) { y[i] = f2_fma(y[i], a2, b2); if (u == 0) x[i] = rcp_approx(x[i]) + 0.0f; }
                if (KIND == MIX_FADD2_FADD) { y[i] = f2_add(y[i], b2); x[i] = __fadd_rn(x[i], b); }
                if (KIND == MIX_FFMA2_IADD) { y[i] = f2_fma(y[i], a2, b2); c[i] = c[i] * 3 + i; }
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float lo, hi; f2_unpack(y[i], lo, hi); acc += lo + hi + x[i] + (float)c[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int KIND>
static void run(int nsm) {
    const int ctas = nsm * 8, threads = 256, iters = 8192;
    float* out; CK(cudaMalloc(&out, sizeof(float) * ctas * threads));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 5; ++w) probe<KIND><<<ctas, threads>>>(out, iters, 1.0001f, 0.5f);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); probe<KIND><<<ctas, threads>>>(out, iters, 1.0001f, 0.5f); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double steps = (double)iters * 4 * ILP * ctas * threads;  // inner steps executed, all threads
    printf("{\"probe\": \"%s\", \"ms\": %.4f, \"thread_steps_per_s\": %.4e, \"warp_steps_per_clk_per_smsp_at_1965\": %.3f}\n",
           names[KIND], best, steps / (best * 1e-3), steps / (best * 1e-3) / 32.0 / (nsm * 4) / 1.965e9);
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

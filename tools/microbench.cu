// Stand-alone probe: (1) FP32 pipe peaks on this GPU (FFMA, FFMA2, FMUL+FADD, FMUL2+FADD2, MUFU.RCP),
// (2) K3 scoring kernel variants timed on a synthetic H x N problem and checked against a host loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -ffp-contract=off \
//        -I code-reproduction-ransac_b200/csrc tools/microbench.cu -o tools/microbench
// Prints one JSON object per line.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cstring>
#include "score_h.cuh"

using namespace b2r;

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

constexpr int ILP = 8;

// kind: 0 FFMA, 1 FFMA2, 2 FMUL+FADD, 3 FMUL2+FADD2, 4 MUFU.RCP, 5 FFMA2 + MUFU (10:2 mix per pair)
template <int KIND>
__global__ void __launch_bounds__(256) pipe_probe(float* out, int iters, float a, float b, long long* cyc) {
    long long t0 = clock64();
    float acc = 0.f;
    if (KIND == 0 || KIND == 2 || KIND == 4) {
        float x[ILP];
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = a + (float)(threadIdx.x + i);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int i = 0; i < ILP; ++i) {
                    if (KIND == 0) x[i] = __fmaf_rn(x[i], a, b);
                    if (KIND == 2) x[i] = (u & 1) ? __fmul_rn(x[i], a) : __fadd_rn(x[i], b);
                    if (KIND == 4) x[i] = rcp_approx(x[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc += x[i];
    } else {
        f2_t x[ILP];
        f2_t a2 = f2_pack(a, a * 1.0001f), b2 = f2_pack(b, b * 0.999f);
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = f2_pack(a + (float)(threadIdx.x + i), b + (float)i);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int i = 0; i < ILP; ++i) {
                    if (KIND == 1) x[i] = f2_fma(x[i], a2, b2);
                    if (KIND == 3) x[i] = (u & 1) ? f2_mul(x[i], a2) : f2_add(x[i], b2);
                    if (KIND == 5) {
                        if (u < 6 || (i & 3)) {  // 6*8 + 2*6 = 60 FFMA2 : 4 pairs of MUFU -> 10 : 2*(2/3)...
                            x[i] = f2_fma(x[i], a2, b2);
                        } else {
                            float lo, hi;
                            f2_unpack(x[i], lo, hi);
                            x[i] = f2_pack(rcp_approx(lo), rcp_approx(hi));
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            float lo, hi;
            f2_unpack(x[i], lo, hi);
            acc += lo + hi;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND>
static void run_probe(const char* name, double lane_ops_per_instr, int nsm) {
    const int ctas = nsm * 8, threads = 256, iters = 4096;
    float* out;
    long long* cyc;
    CK(cudaMalloc(&out, sizeof(float) * ctas * threads));
    CK(cudaMalloc(&cyc, sizeof(long long) * ctas));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) pipe_probe<KIND><<<ctas, threads>>>(out, iters, 1.0001f, 0.5f, cyc);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        pipe_probe<KIND><<<ctas, threads>>>(out, iters, 1.0001f, 0.5f, cyc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    std::vector<long long> hc(ctas);
    CK(cudaMemcpy(hc.data(), cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost));
    double mean_cyc = 0;
    for (auto c : hc) mean_cyc += (double)c;
    mean_cyc /= ctas;
    // warp instructions issued per SM: 8 CTAs/SM * 8 warps * iters * 8 * ILP
    double instr_per_thread = (double)iters * 8 * ILP;
    double total_ops = instr_per_thread * lane_ops_per_instr * (double)ctas * threads;  // scalar fp32 operations
    double ops_per_s = total_ops / (best * 1e-3);
    printf("{\"probe\": \"%s\", \"ms\": %.4f, \"ops_per_s\": %.4e, \"ops_per_clk_per_sm_at_1965\": %.2f, "
           "\"mean_cta_cycles\": %.0f}\n",
           name, best, ops_per_s, ops_per_s / nsm / 1.965e9, mean_cyc);
    fflush(stdout);
    CK(cudaFree(out));
    CK(cudaFree(cyc));
}

// ---------------------------------------------------------------------------------------------
static void host_counts(const std::vector<float>& models, const std::vector<float>& pts4, int N, float thr,
                        const std::vector<int>& which, std::vector<int>& out) {
    out.resize(which.size());
    for (size_t k = 0; k < which.size(); ++k) {
        const float* h = &models[(size_t)which[k] * 8];
        int c = 0;
        for (int i = 0; i < N; ++i) {
            volatile float X = pts4[4 * i], Y = pts4[4 * i + 1], u = pts4[4 * i + 2], v = pts4[4 * i + 3];
            float ww = 1.f / ((h[6] * X + h[7] * Y) + 1.f);
            float dx = ((h[0] * X + h[1] * Y) + h[2]) * ww - u;
            float dy = ((h[3] * X + h[4] * Y) + h[5]) * ww - v;
            float e = dx * dx + dy * dy;
            c += (e <= thr);
        }
        out[k] = c;
    }
}

static int g_extra_smem = 0;
template <int NPAIR, bool EXACT>
static void run_k3(const char* name, const float4* d_models, int H, const PointH* d_pts, int N, float thr,
                   int* d_counts, int tile, int reps, std::vector<int>* result) {
    size_t smem = 128 + (size_t)tile * 16 + (size_t)g_extra_smem;   // g_extra_smem: pad to force fewer resident CTAs per SM (occupancy probe)
    CK(cudaFuncSetAttribute(k3_score_h<NPAIR, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((H + K3_THREADS * 2 * NPAIR - 1) / (K3_THREADS * 2 * NPAIR), (N + tile - 1) / tile);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f, sum = 0;
    for (int r = 0; r < reps + 2; ++r) {
        CK(cudaMemsetAsync(d_counts, 0, sizeof(int) * H));
        CK(cudaEventRecord(e0));
        k3_score_h<NPAIR, EXACT><<<grid, K3_THREADS, smem>>>(d_models, H, H, d_pts, N, thr, d_counts, tile);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2) { best = fminf(best, ms); sum += ms; }
    }
    double evals = (double)H * N;
    printf("{\"k3\": \"%s\", \"npair\": %d, \"exact\": %d, \"tile\": %d, \"H\": %d, \"N\": %d, \"grid\": [%d,%d], "
           "\"ms_best\": %.4f, \"ms_mean\": %.4f, \"evals_per_s_best\": %.4e, \"evals_per_s_mean\": %.4e}\n",
           name, NPAIR, (int)EXACT, tile, H, N, grid.x, grid.y, best, sum / reps, evals / (best * 1e-3),
           evals / (sum / reps * 1e-3));
    fflush(stdout);
    if (result) {
        result->resize(H);
        CK(cudaMemcpy(result->data(), d_counts, sizeof(int) * H, cudaMemcpyDeviceToHost));
    }
}

int main(int argc, char** argv) {
    g_extra_smem = argc > 3 ? atoi(argv[3]) : 0;
    int H = argc > 1 ? atoi(argv[1]) : 100000;
    int N = argc > 2 ? atoi(argv[2]) : 100000;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int nsm = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"smem_optin\": %zu, \"regs_per_sm\": %d, \"l2\": %d}\n",
           prop.name, nsm, prop.clockRate, prop.sharedMemPerBlockOptin, prop.regsPerMultiprocessor, prop.l2CacheSize);

    run_probe<0>("ffma_scalar", 1.0, nsm);
    run_probe<1>("ffma2_packed", 2.0, nsm);
    run_probe<2>("fmul_fadd_scalar", 1.0, nsm);
    run_probe<3>("fmul2_fadd2_packed", 2.0, nsm);
    run_probe<4>("mufu_rcp", 1.0, nsm);
    run_probe<5>("ffma2_mufu_mix", 2.0, nsm);

    // synthetic problem: ground-truth homography, 50% outliers, hypotheses = perturbed truth
    srand(1898);
    auto frand = []() { return (float)rand() / (float)RAND_MAX; };
    const float Ht[8] = {1500.f, 80.f, 600.f, -40.f, -700.f, 100.f, 0.05f, -0.02f};
    std::vector<float> pts4((size_t)N * 4);
    std::vector<PointH> ptsd(N);
    for (int i = 0; i < N; ++i) {
        float X = 0.05f + 0.35f * frand(), Y = -2.6f + 1.7f * frand();
        float w = Ht[6] * X + Ht[7] * Y + 1.f;
        float u = (Ht[0] * X + Ht[1] * Y + Ht[2]) / w + (frand() - 0.5f) * 4.f;
        float v = (Ht[3] * X + Ht[4] * Y + Ht[5]) / w + (frand() - 0.5f) * 4.f;
        if (i & 1) { u = 2142.f * frand(); v = 1620.f * frand(); }
        pts4[4 * i] = X; pts4[4 * i + 1] = Y; pts4[4 * i + 2] = u; pts4[4 * i + 3] = v;
        ptsd[i] = PointH{X, Y, -u, -v};
    }
    std::vector<float> models((size_t)H * 8);
    for (int k = 0; k < H; ++k)
        for (int j = 0; j < 8; ++j) models[(size_t)k * 8 + j] = Ht[j] * (1.f + 0.02f * (frand() - 0.5f) * (float)(k % 7));
    const float thr = 75.f * 75.f;

    float4* d_models;
    PointH* d_pts;
    int* d_counts;
    CK(cudaMalloc(&d_models, sizeof(float) * 8 * H));
    CK(cudaMalloc(&d_pts, sizeof(PointH) * N));
    CK(cudaMalloc(&d_counts, sizeof(int) * H));
    CK(cudaMemcpy(d_models, models.data(), sizeof(float) * 8 * H, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_pts, ptsd.data(), sizeof(PointH) * N, cudaMemcpyHostToDevice));

    std::vector<int> which;
    for (int k = 0; k < 64; ++k) which.push_back((int)(((long long)k * 7919 * 13) % H));
    std::vector<int> ref;
    host_counts(models, pts4, N, thr, which, ref);

    std::vector<int> got;
    auto check = [&](const char* name, bool must_be_exact) {
        int bad = 0, maxd = 0;
        for (size_t k = 0; k < which.size(); ++k) {
            int d = abs(got[which[k]] - ref[k]);
            if (d) ++bad;
            if (d > maxd) maxd = d;
        }
        printf("{\"check\": \"%s\", \"mismatching_hyps\": %d, \"of\": %zu, \"max_count_diff\": %d, \"required_exact\": %d}\n",
               name, bad, which.size(), maxd, (int)must_be_exact);
        fflush(stdout);
    };

    const int reps = 5;
#ifdef K3_SWEEP_FAST
    (void)check;
    run_k3<1, false>("fast_p1_t1024", d_models, H, d_pts, N, thr, d_counts, 1024, reps, &got); check("fast_p1_t1024", false);
    run_k3<2, false>("fast_p2_t512", d_models, H, d_pts, N, thr, d_counts, 512, reps, &got); check("fast_p2_t512", false);
    run_k3<2, false>("fast_p2_t1024", d_models, H, d_pts, N, thr, d_counts, 1024, reps, nullptr);
    run_k3<2, false>("fast_p2_t2048", d_models, H, d_pts, N, thr, d_counts, 2048, reps, nullptr);
    run_k3<3, false>("fast_p3_t512", d_models, H, d_pts, N, thr, d_counts, 512, reps, &got); check("fast_p3_t512", false);
    run_k3<3, false>("fast_p3_t1024", d_models, H, d_pts, N, thr, d_counts, 1024, reps, nullptr);
    run_k3<3, false>("fast_p3_t2048", d_models, H, d_pts, N, thr, d_counts, 2048, reps, nullptr);
#if K3_MIN_CTAS < 3
    run_k3<4, false>("fast_p4_t512", d_models, H, d_pts, N, thr, d_counts, 512, reps, nullptr);
    run_k3<4, false>("fast_p4_t1024", d_models, H, d_pts, N, thr, d_counts, 1024, reps, nullptr);
#endif
#else
    run_k3<4, true>("exact_p4_t1024", d_models, H, d_pts, N, thr, d_counts, 1024, reps, &got); check("exact_p4_t1024", true);
    run_k3<4, false>("fast_p4_t1024", d_models, H, d_pts, N, thr, d_counts, 1024, reps, &got); check("fast_p4_t1024", false);
    run_k3<2, true>("exact_p2_t1024", d_models, H, d_pts, N, thr, d_counts, 1024, reps, &got); check("exact_p2_t1024", true);
    run_k3<2, false>("fast_p2_t1024", d_models, H, d_pts, N, thr, d_counts, 1024, reps, &got); check("fast_p2_t1024", false);
    run_k3<3, true>("exact_p3_t1024", d_models, H, d_pts, N, thr, d_counts, 1024, reps, nullptr);
    run_k3<3, false>("fast_p3_t1024", d_models, H, d_pts, N, thr, d_counts, 1024, reps, nullptr);
    run_k3<4, false>("fast_p4_t2048", d_models, H, d_pts, N, thr, d_counts, 2048, reps, nullptr);
    run_k3<4, false>("fast_p4_t512", d_models, H, d_pts, N, thr, d_counts, 512, reps, nullptr);
    run_k3<4, true>("exact_p4_t2048", d_models, H, d_pts, N, thr, d_counts, 2048, reps, nullptr);
    run_k3<6, false>("fast_p6_t1024", d_models, H, d_pts, N, thr, d_counts, 1024, reps, nullptr);
    run_k3<6, true>("exact_p6_t1024", d_models, H, d_pts, N, thr, d_counts, 1024, reps, nullptr);
#endif
    return 0;
}

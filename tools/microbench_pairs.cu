// Stand-alone probe of a PAIR-MAPPED form of the FAST scoring kernel — the two packed lanes are two points, the hypothesis
// coefficient is the broadcast scalar, so the operand shared by consecutive FFMA2 is the 64-bit one — against the product's
// point-major mapping (score_h.cuh: k3_score_h): what keeps the FMA pipe below its packed peak?  Variants of the same loop: hypotheses per thread, unroll, resident CTAs, counting form (2 LEA.HI / one
// LOP3 / none), points from shared memory or held in registers.  One JSON line per variant.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -I code-reproduction-ransac_b200/csrc \
//        tools/microbench_pairs.cu -o tools/microbench_pairs
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "score_h.cuh"
using namespace b2r;
// thr' = the float above thr; s = thr'^-1/2 (the margin scale of the FAST mode, score_h.cuh)
__device__ __forceinline__ float k3_margin_scale(float thr) {
    const float thr_up = thr < __int_as_float(0x7f800000) ? __uint_as_float(__float_as_uint(thr) + 1u) : thr;
    return rsqrtf(fmaxf(thr_up, 1e-30f));
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// COUNT: 0 = 2 LEA.HI per pair evaluation (product), 1 = one LOP3 (xor of the two margins), 2 = margins summed by FADD2-free
//        integer add of one lane only (1 ALU op), 3 = no per-evaluation ALU work (margins folded with FFMA2 into an accumulator)
// SRC:   0 = shared-memory tile (product), 1 = the tile's first pair held in registers (no LDS in the loop)
template <int NH, int UNROLL, int MINCTAS, int COUNT, int SRC>
__global__ void __launch_bounds__(K3_THREADS, MINCTAS)
k3v(const float4* __restrict__ models, int H, const float4* __restrict__ pairs, int npairs, float thr, int* __restrict__ counts,
    int tile_pairs) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    const float4* tile = reinterpret_cast<const float4*>(smem_raw + 128);
    const int p_begin = blockIdx.y * tile_pairs;
    const int np = min(tile_pairs, npairs - p_begin);
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(bar, (uint32_t)np * 32u); tma_load_1d(smem_raw + 128, pairs + 2 * (size_t)p_begin, (uint32_t)np * 32u, bar); }
    const float s = k3_margin_scale(thr);
    const int h_base = blockIdx.x * (K3_THREADS * NH) + threadIdx.x;
    float hs[NH][8];
#pragma unroll
    for (int j = 0; j < NH; ++j) {
        const int hh = h_base + j * K3_THREADS;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        if (hh < H) { a0 = __ldg(models + 2 * hh); a1 = __ldg(models + 2 * hh + 1); }
        hs[j][0] = a0.x * s; hs[j][1] = a0.y * s; hs[j][2] = a0.z * s; hs[j][3] = a0.w * s;
        hs[j][4] = a1.x * s; hs[j][5] = a1.y * s; hs[j][6] = a1.z; hs[j][7] = a1.w;
    }
    int cnt[NH];
    f2_t facc[NH];
#pragma unroll
    for (int j = 0; j < NH; ++j) { cnt[j] = 0; facc[j] = f2_dup(0.f); }
    const f2_t one = f2_dup(1.0f);
    mbar_wait(bar, 0);
    float4 A0 = tile[0], B0 = tile[1];
#pragma unroll UNROLL
    for (int p = 0; p < np; ++p) {
        float4 A, B;
        if (SRC == 0) { A = tile[2 * p]; B = tile[2 * p + 1]; }
        else { A = A0; B = B0; A0.x += 1e-9f; }
        const f2_t X = f2_pack(A.x, A.y), Y = f2_pack(A.z, A.w), nu = f2_pack(B.x, B.y), nv = f2_pack(B.z, B.w);
#pragma unroll
        for (int j = 0; j < NH; ++j) {
            const f2_t w = f2_fma(X, f2_dup(hs[j][6]), f2_fma(Y, f2_dup(hs[j][7]), one));
            const f2_t sx = f2_fma(X, f2_dup(hs[j][0]), f2_fma(Y, f2_dup(hs[j][1]), f2_dup(hs[j][2])));
            const f2_t sy = f2_fma(X, f2_dup(hs[j][3]), f2_fma(Y, f2_dup(hs[j][4]), f2_dup(hs[j][5])));
            const f2_t a = f2_fma(w, nu, sx);
            const f2_t b = f2_fma(w, nv, sy);
            const f2_t t = f2_mul(w, w);
            float t0, t1, e0, e1;
            f2_unpack(t, t0, t1);
            if (COUNT == 3) {
                facc[j] = f2_fma(a, a, f2_fma(b, b, facc[j]));   // 10 FFMA2 + 1 FMUL2 (t dead -> the compiler may drop it: 10 ops)
                continue;
            }
            f2_unpack(f2_fma(a, a, f2_fma(b, b, f2_pack(-t0, -t1))), e0, e1);
            if (COUNT == 0) { cnt[j] += (int)(__float_as_uint(e0) >> 31); cnt[j] += (int)(__float_as_uint(e1) >> 31); }
            if (COUNT == 1) cnt[j] ^= (int)(__float_as_uint(e0) ^ __float_as_uint(e1));
            if (COUNT == 2) cnt[j] += (int)__float_as_uint(e0) + (int)__float_as_uint(e1);   // one IADD3
        }
    }
#pragma unroll
    for (int j = 0; j < NH; ++j) {
        const int hh = h_base + j * K3_THREADS;
        float f0, f1;
        f2_unpack(facc[j], f0, f1);
        if (hh < H) atomicAdd(counts + hh, cnt[j] + (COUNT == 3 ? (int)(f0 + f1) : 0));
    }
}

template <int NH, int UNROLL, int MINCTAS, int COUNT, int SRC>
static void run(const float4* d_models, int H, const float4* d_pairs, int npairs, float thr, int* d_counts, int tile_pairs) {
    size_t smem = 128 + (size_t)tile_pairs * 32;
    CK(cudaFuncSetAttribute(k3v<NH, UNROLL, MINCTAS, COUNT, SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((H + K3_THREADS * NH - 1) / (K3_THREADS * NH), (npairs + tile_pairs - 1) / tile_pairs);
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k3v<NH, UNROLL, MINCTAS, COUNT, SRC>));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3v<NH, UNROLL, MINCTAS, COUNT, SRC>, K3_THREADS, smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 7; ++r) {
        CK(cudaMemsetAsync(d_counts, 0, sizeof(int) * H));
        CK(cudaEventRecord(e0));
        k3v<NH, UNROLL, MINCTAS, COUNT, SRC><<<grid, K3_THREADS, smem>>>(d_models, H, d_pairs, npairs, thr, d_counts, tile_pairs);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2) best = fminf(best, ms);
    }
    const double evals = 2.0 * npairs * (double)H;
    printf("{\"k3v\": {\"nh\": %d, \"unroll\": %d, \"minctas\": %d, \"count\": %d, \"src\": %d, \"tile_pairs\": %d}, \"regs\": %d, \"ctas_per_sm\": %d, "
           "\"grid\": [%d, %d], \"ms\": %.4f, \"evals_per_s\": %.4e, \"frac_of_11op_pipe_peak\": %.3f}\n",
           NH, UNROLL, MINCTAS, COUNT, SRC, tile_pairs, fa.numRegs, occ, grid.x, grid.y, best, evals / (best * 1e-3),
           evals / (best * 1e-3) / (64.0 / 11.0 * 0.5 * 4 * 148 * 1.965e9));
    fflush(stdout);
}

int main(int argc, char** argv) {
    const int H = argc > 1 ? atoi(argv[1]) : 100000, N = argc > 2 ? atoi(argv[2]) : 100000;
    const int npairs = (N + 1) / 2;
    srand(1898);
    auto frand = []() { return (float)rand() / (float)RAND_MAX; };
    const float Ht[8] = {1500.f, 80.f, 600.f, -40.f, -700.f, 100.f, 0.05f, -0.02f};
    const float thr = 9.f, s = 1.f / sqrtf(thr);
    std::vector<float> pr((size_t)npairs * 8);
    for (int i = 0; i < npairs; ++i) {
        float P[2][4];
        for (int k = 0; k < 2; ++k) {
            float X = 0.05f + 0.35f * frand(), Y = -2.6f + 1.7f * frand(), w = Ht[6] * X + Ht[7] * Y + 1.f;
            float u = (Ht[0] * X + Ht[1] * Y + Ht[2]) / w + (frand() - 0.5f) * 4.f, v = (Ht[3] * X + Ht[4] * Y + Ht[5]) / w + (frand() - 0.5f) * 4.f;
            if ((2 * i + k) & 1) { u = 2142.f * frand(); v = 1620.f * frand(); }
            P[k][0] = X; P[k][1] = Y; P[k][2] = -u * s; P[k][3] = -v * s;
        }
        float* o = &pr[(size_t)i * 8];
        o[0] = P[0][0]; o[1] = P[1][0]; o[2] = P[0][1]; o[3] = P[1][1]; o[4] = P[0][2]; o[5] = P[1][2]; o[6] = P[0][3]; o[7] = P[1][3];
    }
    std::vector<float> models((size_t)H * 8);
    for (int k = 0; k < H; ++k)
        for (int j = 0; j < 8; ++j) models[(size_t)k * 8 + j] = Ht[j] * (1.f + 0.02f * (frand() - 0.5f) * (float)(k % 7));
    float4 *d_models, *d_pairs; int* d_counts;
    CK(cudaMalloc(&d_models, sizeof(float) * 8 * H)); CK(cudaMalloc(&d_pairs, sizeof(float) * 8 * npairs)); CK(cudaMalloc(&d_counts, sizeof(int) * H));
    CK(cudaMemcpy(d_models, models.data(), sizeof(float) * 8 * H, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_pairs, pr.data(), sizeof(float) * 8 * npairs, cudaMemcpyHostToDevice));
#define RUN(NH, UN, MC, CNT, SRC, TP) run<NH, UN, MC, CNT, SRC>(d_models, H, d_pairs, npairs, thr, d_counts, TP)
    RUN(4, 2, 3, 0, 0, 512);    // product
    RUN(4, 2, 3, 1, 0, 512);    // one ALU op per pair evaluation instead of two
    RUN(4, 2, 3, 2, 0, 512);
    RUN(4, 2, 3, 3, 0, 512);    // no ALU work per evaluation
    RUN(4, 2, 3, 0, 1, 512);    // no LDS in the loop
    RUN(4, 2, 3, 3, 1, 512);    // neither
    RUN(4, 1, 3, 0, 0, 512);
    RUN(4, 4, 3, 0, 0, 512);
    RUN(4, 2, 2, 0, 0, 512);
    RUN(4, 2, 4, 0, 0, 512);
    RUN(2, 2, 4, 0, 0, 512);
    RUN(2, 4, 6, 0, 0, 512);
    RUN(3, 2, 3, 0, 0, 512);
    RUN(6, 2, 2, 0, 0, 512);
    RUN(8, 1, 2, 0, 0, 512);
    RUN(8, 2, 1, 0, 0, 512);
    RUN(4, 2, 3, 0, 0, 256);
    RUN(4, 2, 3, 0, 0, 1024);
    RUN(4, 2, 3, 0, 0, 2048);
    return 0;
}

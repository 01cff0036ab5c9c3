// Mapping B of the north star's scoring kernel ("build both, keep the faster"): lanes own POINTS, the
// hypothesis is warp-uniform, inlier flags are counted with a warp ballot + popc — against the product's mapping A (a thread
// owns hypotheses in packed registers, points are broadcast; score_h.cuh).  Same division-free margin, same data.
//   mapB<P>: a warp keeps P x 32 points in registers (lane = point), streams the hypothesis tile from shared memory (two
//            broadcast LDS.128 per hypothesis), 11 scalar FFMA per evaluation, one VOTE + POPC per 32 evaluations, one
//            shared-memory RED per hypothesis and P x 32 points.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -I code-reproduction-ransac_b200/csrc \
//        tools/microbench_mapb.cu -o tools/microbench_mapb          Usage: microbench_mapb [H] [N]   one JSON line per kernel
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "score_h.cuh"
using namespace b2r;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int MB_THREADS = 256, MB_HTILE = 512;   // hypotheses staged per CTA: 512 x 32 B = 16 KB

// grid: x = ceil(H / MB_HTILE), y = ceil(N / (8 warps * 32 * P)).  counts must be zero.
template <int P>
__global__ void __launch_bounds__(MB_THREADS)
mapB(const float4* __restrict__ models, int H, const PointH* __restrict__ pts, int N, float thr, int* __restrict__ counts) {
    __shared__ float4 hs[MB_HTILE * 2];
    __shared__ int cs[MB_HTILE];
    const float thr_up = __uint_as_float(__float_as_uint(thr) + 1u), s = rsqrtf(thr_up);
    const int h0 = blockIdx.x * MB_HTILE, nh = min(MB_HTILE, H - h0);
    for (int i = threadIdx.x; i < nh; i += MB_THREADS) {
        float4 a = __ldg(models + 2 * (h0 + i)), b = __ldg(models + 2 * (h0 + i) + 1);
        a.x *= s; a.y *= s; a.z *= s; a.w *= s; b.x *= s; b.y *= s;      // h0..h5 scaled by thr^-1/2, as the product does
        hs[2 * i] = a; hs[2 * i + 1] = b;
        cs[i] = 0;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p_base = (blockIdx.y * (MB_THREADS / 32) + warp) * (32 * P);
    float X[P], Y[P], nu[P], nv[P];
    bool live[P];
#pragma unroll
    for (int k = 0; k < P; ++k) {
        const int p = p_base + k * 32 + lane;
        live[k] = p < N;
        const PointH pt = live[k] ? pts[p] : PointH{0.f, 0.f, 0.f, 0.f};
        X[k] = pt.X; Y[k] = pt.Y; nu[k] = pt.nu * s; nv[k] = pt.nv * s;
    }
    __syncthreads();
#pragma unroll 2
    for (int i = 0; i < nh; ++i) {
        const float4 a = hs[2 * i], b = hs[2 * i + 1];   // broadcast LDS.128
        int c = 0;
#pragma unroll
        for (int k = 0; k < P; ++k) {
            const float w = __fmaf_rn(b.z, X[k], __fmaf_rn(b.w, Y[k], 1.0f));
            const float sx = __fmaf_rn(a.x, X[k], __fmaf_rn(a.y, Y[k], a.z));
            const float sy = __fmaf_rn(a.w, X[k], __fmaf_rn(b.x, Y[k], b.y));
            const float aa = __fmaf_rn(w, nu[k], sx), bb = __fmaf_rn(w, nv[k], sy);
            const float m = __fmaf_rn(aa, aa, __fmaf_rn(bb, bb, -(w * w)));
            c += __popc(__ballot_sync(0xffffffffu, live[k] && m < 0.f));
        }
        if (lane == 0 && c) atomicAdd(&cs[i], c);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nh; i += MB_THREADS)
        if (cs[i]) atomicAdd(counts + h0 + i, cs[i]);
}

template <typename F>
static float time_it(F launch, int* d_counts, int H) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 6; ++r) {
        CK(cudaMemsetAsync(d_counts, 0, sizeof(int) * H));
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2) best = fminf(best, ms);
    }
    return best;
}

template <int P>
static void run_b(const float4* d_models, int H, const PointH* d_pts, int N, float thr, int* d_counts, const std::vector<int>& ref) {
    dim3 grid((H + MB_HTILE - 1) / MB_HTILE, (N + 8 * 32 * P - 1) / (8 * 32 * P));
    const float ms = time_it([&] { mapB<P><<<grid, MB_THREADS>>>(d_models, H, d_pts, N, thr, d_counts); }, d_counts, H);
    std::vector<int> got(H);
    CK(cudaMemcpy(got.data(), d_counts, sizeof(int) * H, cudaMemcpyDeviceToHost));
    long long maxd = 0;
    for (int k = 0; k < H; ++k) maxd = std::max<long long>(maxd, llabs((long long)got[k] - ref[k]));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, mapB<P>));
    printf("{\"kernel\": \"mapping B (lanes own points, ballot + popc)\", \"points_per_lane\": %d, \"regs\": %d, \"grid\": [%d, %d], \"ms\": %.4f, "
           "\"evals_per_s\": %.4e, \"max_count_diff_vs_mapping_A\": %lld}\n", P, fa.numRegs, grid.x, grid.y, ms, (double)H * N / (ms * 1e-3), maxd);
}

int main(int argc, char** argv) {
    const int H = argc > 1 ? atoi(argv[1]) : 100000, N = argc > 2 ? atoi(argv[2]) : 100000;
    const float thr = 9.0f;
    srand(1898);
    auto frand = []() { return (float)rand() / (float)RAND_MAX; };
    const float Ht[8] = {1500.f, 80.f, 600.f, -40.f, -700.f, 100.f, 0.05f, -0.02f};
    std::vector<PointH> pts(N);
    for (int i = 0; i < N; ++i) {
        float X = 0.05f + 0.35f * frand(), Y = -2.6f + 1.7f * frand(), w = Ht[6] * X + Ht[7] * Y + 1.f;
        float u = (Ht[0] * X + Ht[1] * Y + Ht[2]) / w + (frand() - 0.5f) * 4.f, v = (Ht[3] * X + Ht[4] * Y + Ht[5]) / w + (frand() - 0.5f) * 4.f;
        if (i & 1) { u = 2142.f * frand(); v = 1620.f * frand(); }
        pts[i] = PointH{X, Y, -u, -v};
    }
    std::vector<float> models((size_t)H * 8);
    for (int k = 0; k < H; ++k)
        for (int j = 0; j < 8; ++j) models[(size_t)k * 8 + j] = Ht[j] * (1.f + 0.02f * (frand() - 0.5f) * (float)(k % 7));
    float4* d_models; PointH* d_pts; int* d_counts;
    CK(cudaMalloc(&d_models, sizeof(float) * 8 * H)); CK(cudaMalloc(&d_pts, sizeof(PointH) * N)); CK(cudaMalloc(&d_counts, sizeof(int) * H));
    CK(cudaMemcpy(d_models, models.data(), sizeof(float) * 8 * H, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_pts, pts.data(), sizeof(PointH) * N, cudaMemcpyHostToDevice));
    const int tile = 1024;
    dim3 grid((H + K3_THREADS * 4 - 1) / (K3_THREADS * 4), (N + tile - 1) / tile);
    const size_t smem = 128 + (size_t)tile * 16;
    float ms = time_it([&] { k3_score_h<2, false><<<grid, K3_THREADS, smem>>>(d_models, H, H, d_pts, N, thr, d_counts, tile); }, d_counts, H);
    std::vector<int> ref(H);
    CK(cudaMemcpy(ref.data(), d_counts, sizeof(int) * H, cudaMemcpyDeviceToHost));
    printf("{\"kernel\": \"mapping A = k3_score_h<2,fast> (threads own hypotheses, packed FFMA2)\", \"H\": %d, \"N\": %d, \"ms\": %.4f, \"evals_per_s\": %.4e}\n",
           H, N, ms, (double)H * N / (ms * 1e-3));
    run_b<1>(d_models, H, d_pts, N, thr, d_counts, ref);
    run_b<2>(d_models, H, d_pts, N, thr, d_counts, ref);
    run_b<4>(d_models, H, d_pts, N, thr, d_counts, ref);
    run_b<8>(d_models, H, d_pts, N, thr, d_counts, ref);
    return 0;
}

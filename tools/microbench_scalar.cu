// Probe: K3 fast arithmetic with SCALAR FFMA (NH hypotheses per thread, point operands reused across them) against the packed FFMA2 kernel.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -I code-reproduction-ransac_b200/csrc tools/microbench_scalar.cu -o tools/microbench_scalar
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "score_h.cuh"
using namespace b2r;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int NH, int UNROLL>
__global__ void __launch_bounds__(K3_THREADS, 2)
k3s_score_h(const float4* __restrict__ models, int H, const PointH* __restrict__ pts, int N, float thr, int* __restrict__ counts, int tile_pts) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    const float4* tile = reinterpret_cast<const float4*>(smem_raw + 128);
    const int p_begin = blockIdx.y * tile_pts;
    const int np = min(tile_pts, N - p_begin);
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(bar, (uint32_t)np * 16u); tma_load_1d(smem_raw + 128, pts + p_begin, (uint32_t)np * 16u, bar); }
    const int h_base = blockIdx.x * (K3_THREADS * NH) + threadIdx.x;
    float h[NH][8];
#pragma unroll
    for (int j = 0; j < NH; ++j) {
        const int hh = h_base + j * K3_THREADS;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        if (hh < H) { a0 = __ldg(models + 2 * hh); a1 = __ldg(models + 2 * hh + 1); }
        h[j][0] = a0.x; h[j][1] = a0.y; h[j][2] = a0.z; h[j][3] = a0.w; h[j][4] = a1.x; h[j][5] = a1.y; h[j][6] = a1.z; h[j][7] = a1.w;
    }
    int cnt[NH];
#pragma unroll
    for (int j = 0; j < NH; ++j) cnt[j] = 0;
    mbar_wait(bar, 0);
#pragma unroll UNROLL
    for (int p = 0; p < np; ++p) {
        const float4 pt = tile[p];
#pragma unroll
        for (int j = 0; j < NH; ++j) {
            const float w = __fmaf_rn(h[j][6], pt.x, __fmaf_rn(h[j][7], pt.y, 1.0f));
            const float ww = rcp_approx(w);
            const float sx = __fmaf_rn(h[j][0], pt.x, __fmaf_rn(h[j][1], pt.y, h[j][2]));
            const float sy = __fmaf_rn(h[j][3], pt.x, __fmaf_rn(h[j][4], pt.y, h[j][5]));
            const float dx = __fmaf_rn(sx, ww, pt.z), dy = __fmaf_rn(sy, ww, pt.w);
            const float e = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
            cnt[j] += (e <= thr) ? 1 : 0;
        }
    }
#pragma unroll
    for (int j = 0; j < NH; ++j) {
        const int hh = h_base + j * K3_THREADS;
        if (hh < H) atomicAdd(counts + hh, cnt[j]);
    }
}

template <int NH, int UNROLL>
static void run(const float4* d_models, int H, const PointH* d_pts, int N, float thr, int* d_counts, int tile, std::vector<int>& ref) {
    size_t smem = 128 + (size_t)tile * 16;
    dim3 grid((H + K3_THREADS * NH - 1) / (K3_THREADS * NH), (N + tile - 1) / tile);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); float best = 1e30f;
    for (int r = 0; r < 6; ++r) {
        CK(cudaMemsetAsync(d_counts, 0, sizeof(int) * H)); CK(cudaEventRecord(e0));
        k3s_score_h<NH, UNROLL><<<grid, K3_THREADS, smem>>>(d_models, H, d_pts, N, thr, d_counts, tile);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r >= 2) best = fminf(best, ms);
    }
    std::vector<int> got(H); CK(cudaMemcpy(got.data(), d_counts, sizeof(int) * H, cudaMemcpyDeviceToHost));
    int bad = 0; for (int k = 0; k < H; ++k) bad += got[k] != ref[k];
    printf("{\"k3s\": \"scalar FFMA\", \"nh\": %d, \"unroll\": %d, \"tile\": %d, \"ms\": %.4f, \"evals_per_s\": %.4e, \"count_mismatch_vs_packed\": %d}\n", NH, UNROLL, tile, best, (double)H * N / (best * 1e-3), bad);
    fflush(stdout);
}

int main(int argc, char** argv) {
    int H = argc > 1 ? atoi(argv[1]) : 100000, N = argc > 2 ? atoi(argv[2]) : 100000;
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
    for (int k = 0; k < H; ++k) for (int j = 0; j < 8; ++j) models[(size_t)k * 8 + j] = Ht[j] * (1.f + 0.02f * (frand() - 0.5f) * (float)(k % 7));
    const float thr = 9.f;
    float4* d_models; PointH* d_pts; int* d_counts;
    CK(cudaMalloc(&d_models, sizeof(float) * 8 * H)); CK(cudaMalloc(&d_pts, sizeof(PointH) * N)); CK(cudaMalloc(&d_counts, sizeof(int) * H));
    CK(cudaMemcpy(d_models, models.data(), sizeof(float) * 8 * H, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_pts, pts.data(), sizeof(PointH) * N, cudaMemcpyHostToDevice));
    std::vector<int> ref(H);
    {
        const int tile = 512; size_t smem = 128 + tile * 16;
        dim3 grid((H + K3_THREADS * 8 - 1) / (K3_THREADS * 8), (N + tile - 1) / tile);
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); float best = 1e30f;
        for (int r = 0; r < 6; ++r) {
            CK(cudaMemsetAsync(d_counts, 0, sizeof(int) * H)); CK(cudaEventRecord(e0));
            k3_score_h<4, false><<<grid, K3_THREADS, smem>>>(d_models, H, H, d_pts, N, thr, d_counts, tile);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r >= 2) best = fminf(best, ms);
        }
        CK(cudaMemcpy(ref.data(), d_counts, sizeof(int) * H, cudaMemcpyDeviceToHost));
        printf("{\"k3\": \"packed npair4 tile512\", \"ms\": %.4f, \"evals_per_s\": %.4e}\n", best, (double)H * N / (best * 1e-3));
    }
    run<4, 2>(d_models, H, d_pts, N, thr, d_counts, 512, ref);
    run<8, 2>(d_models, H, d_pts, N, thr, d_counts, 512, ref);
    run<8, 4>(d_models, H, d_pts, N, thr, d_counts, 512, ref);
    run<8, 1>(d_models, H, d_pts, N, thr, d_counts, 1024, ref);
    run<6, 2>(d_models, H, d_pts, N, thr, d_counts, 512, ref);
    run<12, 2>(d_models, H, d_pts, N, thr, d_counts, 512, ref);
    return 0;
}

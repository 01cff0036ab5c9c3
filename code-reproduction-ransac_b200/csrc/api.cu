// C ABI of libransac_b200.so (declared in include/ransac_b200.h).  Host-side orchestration only: every
// arithmetic step of the hot path runs in the kernels of pipeline_h.cuh / score_h.cuh.  There is no CPU
// fallback: each entry point fails with B2R_ERR_CUDA when no device is usable.
#include "host_common.h"
#include "sweep.cuh"
#include "score_h_filt.cuh"

using namespace b2r;

thread_local std::string g_err;

int fail(int code, const char* fmt, const char* a, const char* b) {
    char buf[512];
    snprintf(buf, sizeof(buf), fmt, a, b);
    g_err = buf;
    return code;
}


struct b2r_h_problem {
    int Q = 0, n = 0;
    DevBuf pts;       // [Q][n] PointH
    DevBuf samples;   // [Q][H][4] int32
    DevBuf models;    // [Q][H][8] fp32
    DevBuf counts;    // [Q][H] int32
    DevBuf state;     // [Q] RansacState (replay path) + one int32 "problems not done" counter behind it
    DevBuf keys;      // [Q] u64
    DevBuf sel;       // [Q] HSelect
    DevBuf H;         // [Q][9] fp64
    DevBuf mask;      // [Q][n] u8
    DevBuf rmask;     // [Q][n] u8
    DevBuf info;      // [Q][12] int32
    DevBuf sweep;     // camera sweep: pos3d, pixels, cams, pos2 [Q][n][2], scores [Q][2], M [Q][9], best
    int H_last = 0;
    float stage_ms[5] = {0, 0, 0, 0, 0};
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    void release() {
        pts.release(); samples.release(); models.release(); counts.release(); state.release(); keys.release();
        sel.release(); H.release(); mask.release(); rmask.release(); info.release(); sweep.release();
        for (auto& e : ev)
            if (e) cudaEventDestroy(e), e = nullptr;
    }
};

extern "C" {

int b2r_version(void) { return 100; }
const char* b2r_last_error(void) { return g_err.c_str(); }

int b2r_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

b2r_ctx* b2r_ctx_create(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        fail(B2R_ERR_CUDA, "no usable CUDA device (%s); ransac_b200 has no CPU fallback%s",
             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return nullptr;
    }
    if (device < 0 || device >= n) {
        fail(B2R_ERR_ARG, "device index out of range%s%s");
        return nullptr;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        fail(B2R_ERR_CUDA, "cudaSetDevice: %s%s", cudaGetErrorString(e));
        return nullptr;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        fail(B2R_ERR_CUDA, "cudaGetDeviceProperties: %s%s", cudaGetErrorString(e));
        return nullptr;
    }
    if (prop.major != 10) {
        char cc[32];
        snprintf(cc, sizeof(cc), "%d.%d", prop.major, prop.minor);
        fail(B2R_ERR_CUDA, "device %s has compute capability %s; this library is built for sm_100a only", prop.name, cc);
        return nullptr;
    }
    b2r_ctx* c = new b2r_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        fail(B2R_ERR_CUDA, "cudaStreamCreate: %s%s", cudaGetErrorString(e));
        delete c;
        return nullptr;
    }
    return c;
}

void b2r_ctx_destroy(b2r_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->cached) {
        c->cached->release();
        delete c->cached;
    }
    b2r_p_problem_destroy(c->cached_p);
    c->in_a.release(); c->in_b.release(); c->scratch0.release(); c->scratch1.release(); c->scratch2.release();
    c->scratch3.release(); c->gscratch.release(); c->pin_in.release(); c->pin_out.release();
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

void* b2r_ctx_stream(b2r_ctx* c) { return c ? (void*)c->stream : nullptr; }

int b2r_ctx_synchronize(b2r_ctx* c) {
    if (!c) return fail(B2R_ERR_ARG, "null context%s%s");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return B2R_OK;
}

int b2r_ctx_launch_count(b2r_ctx* c) { return c ? c->launches : 0; }

void b2r_default_h_params(b2r_h_params* p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->thr = 3.0;
    p->max_iters = 2000;
    p->confidence = 0.995;
    p->sampler = B2R_SAMPLER_CV_REPLAY;
    p->seed = 0;
    p->arith = B2R_ARITH_EXACT;
    p->mask_semantics = B2R_MASK_CV413;
    p->refine = 1;
    p->hyp_begin = 0;
    p->solver = B2R_SOLVER_EXACT;
    p->reserved = 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
static int check_params(const b2r_h_params* p) {
    if (!p) return fail(B2R_ERR_ARG, "null params%s%s");
    if (!(p->thr >= 0)) return fail(B2R_ERR_ARG, "thr must be >= 0%s%s");
    if (p->sampler != B2R_SAMPLER_CV_REPLAY && p->sampler != B2R_SAMPLER_PHILOX) return fail(B2R_ERR_ARG, "bad sampler%s%s");
    if (p->arith != B2R_ARITH_EXACT && p->arith != B2R_ARITH_FAST && p->arith != B2R_ARITH_EXACT_UNFILTERED)
        return fail(B2R_ERR_ARG, "bad arith%s%s");
    if (p->solver != B2R_SOLVER_EXACT && p->solver != B2R_SOLVER_FAST) return fail(B2R_ERR_ARG, "bad solver%s%s");
    if (p->max_iters > (1 << 30)) return fail(B2R_ERR_ARG, "max_iters too large%s%s");
    if (p->refine < B2R_REFINE_NONE || p->refine > B2R_REFINE_PARALLEL) return fail(B2R_ERR_ARG, "bad refine%s%s");
    // the arg-max keys carry the global hypothesis id in 32 bits (count << 32 | ~id): ids beyond 2^32 would alias
    if (p->sampler == B2R_SAMPLER_PHILOX &&
        (p->hyp_begin < 0 || p->hyp_begin + (long long)(p->max_iters > 1 ? p->max_iters : 1) > (1LL << 32)))
        return fail(B2R_ERR_ARG, "PHILOX hypothesis ids must lie in [0, 2^32): hyp_begin >= 0 and hyp_begin + max_iters <= 2^32%s%s");
    return B2R_OK;
}

// k_solve_h4_smem / k_jacobi_packed9 use K2S_SMEM = 31.5 KB of dynamic shared memory: below the 48 KB every kernel may
// ask for without cudaFuncSetAttribute
static_assert(K2S_SMEM <= 48 * 1024, "k_solve_h4_smem would need the dynamic shared memory opt-in");
static int k2s_prepare(b2r_ctx*) { return B2R_OK; }

template <int NPAIR>
static int launch_k3(b2r_ctx* c, const float4* models, int H, int H_stride, const PointH* pts, int n, float thr_sq, int* counts,
                     int Q, int arith, int max_tile) {
    // tile: as many points per CTA as keeps >= ~4 CTAs per SM slot in flight
    int tile = max_tile;
    const long long hyp_blocks = (H + K3_THREADS * 2 * NPAIR - 1) / (K3_THREADS * 2 * NPAIR);
    while (tile > 128 && hyp_blocks * ((n + tile - 1) / tile) * Q < 8LL * c->sm_count) tile >>= 1;
    if (tile > n) tile = ((n + 7) / 8) * 8;
    const size_t smem = 128 + (size_t)tile * 16;
    dim3 grid((unsigned)hyp_blocks, (unsigned)((n + tile - 1) / tile), (unsigned)Q);
    // a handful of points (the reference's 12) cannot repay the per-hypothesis set-up of the filter: the un-fused kernel then
    if (arith == B2R_ARITH_EXACT && n < 64) arith = B2R_ARITH_EXACT_UNFILTERED;
    if (arith == B2R_ARITH_EXACT) {   // the same counts as the un-fused sequence, through the filtered predicate (score_h_filt.cuh)
        static bool optin[64] = {false};   // > 48 KB of dynamic shared memory: once per device
        if (c->device < 64 && !optin[c->device]) {
            CU(cudaFuncSetAttribute(k3_score_h_filt<NPAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k3_filt_smem(1024, NPAIR)));
            optin[c->device] = true;
        }
        LAUNCH(c, (k3_score_h_filt<NPAIR>), grid, K3_THREADS, k3_filt_smem(tile, NPAIR), models, H, H_stride, pts, n, thr_sq, counts,
               tile);
    } else if (arith == B2R_ARITH_EXACT_UNFILTERED)
        LAUNCH(c, (k3_score_h<NPAIR, true>), grid, K3_THREADS, smem, models, H, H_stride, pts, n, thr_sq, counts, tile);
    else
        LAUNCH(c, (k3_score_h<NPAIR, false>), grid, K3_THREADS, smem, models, H, H_stride, pts, n, thr_sq, counts, tile);
    CU(cudaGetLastError());
    return B2R_OK;
}

// Scores hypotheses [begin, begin+H) of every problem; models/counts are [Q][H_stride] arrays.
static int score_models(b2r_ctx* c, const float4* models, int H, const PointH* pts, int n, float thr_sq, int* counts,
                        int Q, int arith, int H_stride = 0, int begin = 0) {
    if (H_stride == 0) H_stride = H;
    if (begin == 0 && H == H_stride) {
        CU(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)Q * H, c->stream));
    } else {
        CU(cudaMemset2DAsync(counts + begin, sizeof(int) * (size_t)H_stride, 0, sizeof(int) * (size_t)H, (size_t)Q, c->stream));
    }
    // measured on B200 at 100k x 100k (profiles/r01g_microbench_k3_forms.jsonl): both arithmetic modes are fastest with 2
    // hypothesis pairs per thread and 1024-point tiles (fast: 2.83e12/s with the division-free margin; exact: 1.54e12/s)
    return launch_k3<2>(c, models + 2 * (size_t)begin, H, H_stride, pts, n, thr_sq, counts + begin, Q, arith, 1024);
}

static int problem_reserve(b2r_h_problem* pr, int Q, int n, int H) {
    CU(pr->pts.reserve(sizeof(PointH) * (size_t)Q * n));
    CU(pr->samples.reserve(sizeof(int) * 4 * (size_t)Q * H));
    CU(pr->models.reserve(sizeof(float) * 8 * (size_t)Q * H));
    CU(pr->counts.reserve(sizeof(int) * (size_t)Q * H));
    CU(pr->state.reserve(sizeof(RansacState) * (size_t)Q + 64));
    CU(pr->keys.reserve(sizeof(unsigned long long) * (size_t)Q));
    CU(pr->sel.reserve(sizeof(HSelect) * (size_t)Q));
    CU(pr->H.reserve(sizeof(double) * 9 * (size_t)Q));
    CU(pr->mask.reserve((size_t)Q * n));
    CU(pr->rmask.reserve((size_t)Q * n));
    CU(pr->info.reserve(sizeof(int) * 12 * (size_t)Q));
    for (auto& e : pr->ev)
        if (!e) CU(cudaEventCreate(&e));
    return B2R_OK;
}

static int upload_points(b2r_ctx* c, b2r_h_problem* pr, const double* src, const double* dst, int dst_shared, int Q, int n) {
    const size_t sb = sizeof(double) * 2 * (size_t)Q * n, db = sizeof(double) * 2 * (size_t)(dst_shared ? 1 : Q) * n;
    CU(c->in_a.reserve(sb));
    CU(c->in_b.reserve(db));
    CU(cudaMemcpyAsync(c->in_a.p, src, sb, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->in_b.p, dst, db, cudaMemcpyHostToDevice, c->stream));
    const size_t total = (size_t)Q * n;
    LAUNCH(c, k_pack_points_h, (unsigned)((total + 255) / 256), 256, 0, c->in_a.as<double>(), c->in_b.as<double>(),
           dst_shared, Q, n, pr->pts.as<PointH>());
    CU(cudaGetLastError());
    pr->Q = Q;
    pr->n = n;
    return B2R_OK;
}

// stage 1: sample + solve + score (+ per-problem argmax key for PHILOX).  The replay path runs OpenCV's sequential loop
// in growing chunks of iterations and stops as soon as every problem of the batch has reached its iteration bound.
static int run_score(b2r_ctx* c, b2r_h_problem* pr, const b2r_h_params* p) {
    const int Q = pr->Q, n = pr->n, H = p->max_iters > 1 ? p->max_iters : 1;
    int rc = problem_reserve(pr, Q, n, H);
    if (rc) return rc;
    pr->H_last = H;
    const float thr_sq = (float)(p->thr * p->thr);
    CU(cudaEventRecord(pr->ev[0], c->stream));
    if (n > 4 && p->sampler == B2R_SAMPLER_PHILOX) {
        dim3 grid((unsigned)((H + 127) / 128), (unsigned)Q);
        if (p->solver == B2R_SOLVER_EXACT) {
            // sample only, then the shared-memory Jacobi kernel (7x the throughput of solving in the sampling thread)
            LAUNCH(c, k_philox_sample_solve_h<false>, grid, 128, 0, pr->pts.as<PointH>(), n, H, (long long)p->hyp_begin, p->seed,
                   pr->samples.as<int>(), (float4*)nullptr);
            if ((rc = k2s_prepare(c))) return rc;
            dim3 g2((unsigned)((H + K2S_THREADS - 1) / K2S_THREADS), (unsigned)Q);
            LAUNCH(c, k_solve_h4_smem, g2, K2S_THREADS, K2S_SMEM, pr->pts.as<PointH>(), n, pr->samples.as<int>(), H, 0, H,
                   (const RansacState*)nullptr, pr->models.as<float4>(), (double*)nullptr, (uint8_t*)nullptr, (uint8_t*)nullptr);
        } else {
            LAUNCH(c, k_philox_sample_solve_h<true>, grid, 128, 0, pr->pts.as<PointH>(), n, H, (long long)p->hyp_begin, p->seed,
                   pr->samples.as<int>(), pr->models.as<float4>());
        }
        CU(cudaGetLastError());
        CU(cudaEventRecord(pr->ev[1], c->stream));
        rc = score_models(c, pr->models.as<float4>(), H, pr->pts.as<PointH>(), n, thr_sq, pr->counts.as<int>(), Q, p->arith);
        if (rc) return rc;
        CU(cudaEventRecord(pr->ev[2], c->stream));
        CU(cudaMemsetAsync(pr->keys.p, 0, sizeof(unsigned long long) * Q, c->stream));
        int gx = (H + 255) / 256;
        if (gx > 4 * c->sm_count) gx = 4 * c->sm_count;
        LAUNCH(c, k_argmax_key, dim3((unsigned)gx, (unsigned)Q), 256, 0, pr->counts.as<int>(), H,
               (unsigned long long)p->hyp_begin, pr->keys.as<unsigned long long>());
        CU(cudaGetLastError());
        return B2R_OK;
    }
    if (n > 4) {
        RansacState* st = pr->state.as<RansacState>();
        int* not_done = reinterpret_cast<int*>(st + Q);
        LAUNCH(c, k_state_init, (unsigned)((Q + 127) / 128), 128, 0, st, p->max_iters, Q);
        // chunk lengths 64, 128, 256, ... (boundaries 64, 192, 448, 960, ...): every chunk costs one solver latency, so the lengths
        // grow geometrically; on the reference's data OpenCV stops after 13-173 iterations (median 24): two chunks at most
        int active = Q;   // problems still iterating (from the "not done" counter of the previous chunk)
        for (int begin = 0, len = 64; begin < H; begin += len, len *= 2) {
            if (len > H - begin) len = H - begin;
            CU(cudaMemsetAsync(not_done, 0, sizeof(int), c->stream));
            LAUNCH(c, k_cv_sample_h, (unsigned)Q, 32, 0, pr->pts.as<PointH>(), n, H, begin, len, pr->samples.as<int>(), st, Q);
            if (p->solver == B2R_SOLVER_EXACT && (long long)active * len <= 20LL * c->sm_count) {   // crossover measured at ~3200 solves (tools/microbench_jacobi.cu)
                // few solves (a single problem, or the stragglers of a batch: finished problems' warps exit at once):
                // one warp each — 4x less latency than a thread each, and the redundant fp64 work does not matter
                dim3 grid((unsigned)((len + 7) / 8), (unsigned)Q);
                LAUNCH(c, k_solve_h4_warp, grid, 256, 0, pr->pts.as<PointH>(), n, pr->samples.as<int>(), H, begin, len,
                       (const RansacState*)st, pr->models.as<float4>(), (double*)nullptr, (uint8_t*)nullptr, (uint8_t*)nullptr);
            } else if (p->solver == B2R_SOLVER_EXACT) {
                if ((rc = k2s_prepare(c))) return rc;
                dim3 grid((unsigned)((len + K2S_THREADS - 1) / K2S_THREADS), (unsigned)Q);
                LAUNCH(c, k_solve_h4_smem, grid, K2S_THREADS, K2S_SMEM, pr->pts.as<PointH>(), n, pr->samples.as<int>(), H, begin, len,
                       (const RansacState*)st, pr->models.as<float4>(), (double*)nullptr, (uint8_t*)nullptr, (uint8_t*)nullptr);
            } else {
                dim3 grid((unsigned)((len + 127) / 128), (unsigned)Q);
                LAUNCH(c, k_solve_h4, grid, 128, 0, pr->pts.as<PointH>(), n, pr->samples.as<int>(), H, begin, len,
                       (const RansacState*)st, pr->models.as<float4>(), (double*)nullptr, (uint8_t*)nullptr, (uint8_t*)nullptr, p->solver);
            }
            CU(cudaGetLastError());
            rc = score_models(c, pr->models.as<float4>(), len, pr->pts.as<PointH>(), n, thr_sq, pr->counts.as<int>(), Q, p->arith, H,
                              begin);
            if (rc) return rc;
            LAUNCH(c, k_select_cv_chunk, (unsigned)((Q + 127) / 128), 128, 0, pr->counts.as<int>(), H, begin, len, n, p->confidence, 4,
                   st, pr->sel.as<HSelect>(), Q, not_done);
            CU(cudaGetLastError());
            if (begin + len >= H) break;
            int nd = 0;
            CU(cudaMemcpyAsync(&nd, not_done, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            if (nd == 0) break;
            active = nd;
        }
    }
    CU(cudaEventRecord(pr->ev[1], c->stream));  // replay path: the whole loop is accounted to stage 0
    CU(cudaEventRecord(pr->ev[2], c->stream));
    return B2R_OK;
}

__global__ void k_select_n4(HSelect* sel, int* samples, int Q) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    HSelect s;
    s.best = 0; s.best_count = 4; s.iters_run = 0; s.pad = 0;
    sel[q] = s;
    reinterpret_cast<int4*>(samples)[q] = make_int4(0, 1, 2, 3);
}

// The winner may live on another rank: rebuild its sample from the global hypothesis id into slot 0 of `samples`.
__global__ void k_resample_winner(const PointH* __restrict__ pts, int n, const unsigned long long* __restrict__ keys,
                                  uint64_t seed, int Hs, int* __restrict__ samples, HSelect* __restrict__ sel, int H_total,
                                  int Q) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const unsigned long long key = keys[q];
    const int count = (int)(key >> 32);
    const unsigned long long gid = 0xFFFFFFFFull - (key & 0xFFFFFFFFull);
    HSelect s;
    s.best = -1; s.best_count = 0; s.iters_run = H_total; s.pad = 0;
    if (count > 3) {
        const PointH* P = pts + (size_t)q * n;
        int idx[4];
        float ms1[8], ms2[8];
        bool found = false;
        for (int attempt = 0; attempt < PHILOX_MAX_ATTEMPTS && !found; ++attempt) {
            const Philox4 r = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)attempt, (uint32_t)q,
                                            (uint32_t)seed, (uint32_t)(seed >> 32));
            distinct4(r, (uint32_t)n, idx);
            gather4(P, idx, ms1, ms2);
            found = h_check_subset4(ms1, ms2);
        }
        if (found) {
            reinterpret_cast<int4*>(samples)[(size_t)q * Hs] = make_int4(idx[0], idx[1], idx[2], idx[3]);
            s.best = 0;
            s.best_count = count;
            s.pad = (int)(gid & 0x7fffffff);
        }
    }
    sel[q] = s;
}

// K4 launch: one thread-block cluster per problem, one CTA for small problems; a single problem of 4096 points or more
// runs on a cooperative grid instead (reductions through global memory + grid barriers).
static int launch_finalize(b2r_ctx* c, const PointH* pts, int n, const int* samples, int Hs, const HSelect* sel, float thr_sq,
                           int mask_semantics, int refine, int solver, double* H_out, uint8_t* mask_out, uint8_t* rmask_out,
                           int* info, const uint8_t* ext_mask, const double* ext_H, int Q, const float4* models = nullptr) {
    // problems of the reference's size with the exact solver: sums in OpenCV's order, eigen-solves only (kernel comment)
    int seq = (solver == B2R_SOLVER_EXACT && refine == B2R_REFINE_CV && n > 4 && n <= 128) ? 1 : 0;
    if (Q == 1 && n >= 4096) {
        constexpr int GT = 512;
        static thread_local int coop_ok[16] = {0};
        int& ok = coop_ok[c->device & 15];
        if (ok == 0) {
            int per_sm = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_finalize_h<GT, true>, GT, 0));
            ok = per_sm >= 1 ? 1 : -1;
        }
        // the passes over the points are bound by the SM's fp64 rate (~150 operations per point and LM evaluation), the ~10
        // reductions of a call by the grid barrier and the gather of the partials, which grow with the CTAs that join:
        // measured (tools/perf_finalize_grid.py, profiles/r02s_finalize_grid_tuning.jsonl) one point per thread is best up
        // to ~20 000 points (20 000 points: 0.075 ms against 0.114 ms for a cluster of four CTAs), two above (100 000
        // points: 0.093 ms; 148 CTAs 0.105); below 4096 points a single cluster is faster
        const int ppt = n >= 32768 ? 2 : 1;
        const int ctas = ok > 0 ? std::min(c->sm_count, std::max(2, (n + GT * ppt - 1) / (GT * ppt))) : -1;
        if (ctas > 0) {
            CU(c->gscratch.reserve(sizeof(double) * 2 * (size_t)ctas * RED_MAX));
            double* gs = c->gscratch.as<double>();
            void* args[] = {(void*)&pts, (void*)&n, (void*)&samples, (void*)&Hs, (void*)&sel, (void*)&thr_sq, (void*)&mask_semantics,
                            (void*)&refine, (void*)&solver, (void*)&H_out, (void*)&mask_out, (void*)&rmask_out, (void*)&info,
                            (void*)&ext_mask, (void*)&ext_H, (void*)&gs, (void*)&models, (void*)&seq};
            CU(cudaLaunchCooperativeKernel((const void*)k_finalize_h<GT, true>, dim3((unsigned)ctas), dim3(GT), args, 0, c->stream));
            c->launches++;
            return B2R_OK;
        }
    }
    // a big batch of mid-size problems is bound by the sequential part of each problem (one thread's small linear algebra):
    // 64-thread CTAs keep eight problems per SM in flight instead of four
    // measured (tools/perf_cfg1.py): one 128-thread CTA finalizes 2000 points in 0.20 ms, a 1024-thread CTA in 0.25 ms — the LM's
    // sequential linear algebra and the depth of the reductions, not the passes over the points, bound a mid-size problem
    // ... and 512-thread CTAs (127 registers: the 1024-thread instantiation is capped at 64 and spilled 1.8 KB per thread) take
    // over at 512 points with the closed-form solver, at 4096 with the exact one (tools/perf_cfg1.py: 5000 points 0.28 -> 0.17 ms,
    // 20 000 points 0.34 -> 0.19 ms, 2000 points 0.21 -> 0.17 ms)
    const bool many = Q >= 4 * c->sm_count && n >= 256 && n < 4096;
    const int big_n = solver == B2R_SOLVER_EXACT ? 4096 : 512;
    const int threads = many ? 64 : (n >= big_n ? 512 : 128);
    const int csize = n >= 32768 ? 8 : (n >= 8192 ? 4 : (n >= 4096 ? 2 : 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(Q * csize));
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = seq ? sizeof(double) * 20 * (size_t)n : 0;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    double* no_scratch = nullptr;
    if (threads == 512)
        CU(cudaLaunchKernelEx(&cfg, k_finalize_h<512, false>, pts, n, samples, Hs, sel, thr_sq, mask_semantics, refine, solver, H_out,
                              mask_out, rmask_out, info, ext_mask, ext_H, no_scratch, models, seq));
    else if (threads == 64)
        CU(cudaLaunchKernelEx(&cfg, k_finalize_h<64, false>, pts, n, samples, Hs, sel, thr_sq, mask_semantics, refine, solver, H_out,
                              mask_out, rmask_out, info, ext_mask, ext_H, no_scratch, models, seq));
    else
        CU(cudaLaunchKernelEx(&cfg, k_finalize_h<128, false>, pts, n, samples, Hs, sel, thr_sq, mask_semantics, refine, solver, H_out,
                              mask_out, rmask_out, info, ext_mask, ext_H, no_scratch, models, seq));
    c->launches++;
    return B2R_OK;
}

// stage 2: select + finalize.  keys_host != nullptr: use these (globally reduced) keys instead of the local ones.
static int run_finish(b2r_ctx* c, b2r_h_problem* pr, const b2r_h_params* p, const uint64_t* keys_host,
                      const uint64_t* keys_dev = nullptr) {
    const int Q = pr->Q, n = pr->n, H = pr->H_last;
    const float thr_sq = (float)(p->thr * p->thr);
    if (n == 4) {
        LAUNCH(c, k_select_n4, (unsigned)((Q + 127) / 128), 128, 0, pr->sel.as<HSelect>(), pr->samples.as<int>(), Q);
    } else if (p->sampler == B2R_SAMPLER_PHILOX) {
        if (keys_host || keys_dev) {
            if (keys_host) CU(cudaMemcpyAsync(pr->keys.p, keys_host, sizeof(uint64_t) * Q, cudaMemcpyHostToDevice, c->stream));
            else CU(cudaMemcpyAsync(pr->keys.p, keys_dev, sizeof(uint64_t) * Q, cudaMemcpyDeviceToDevice, c->stream));
            LAUNCH(c, k_resample_winner, (unsigned)((Q + 127) / 128), 128, 0, pr->pts.as<PointH>(), n,
                   pr->keys.as<unsigned long long>(), p->seed, H, pr->samples.as<int>(), pr->sel.as<HSelect>(), H, Q);
        } else {
            LAUNCH(c, k_select_from_keys, (unsigned)((Q + 127) / 128), 128, 0, pr->keys.as<unsigned long long>(),
                   (unsigned long long)p->hyp_begin, H, 4, pr->sel.as<HSelect>(), Q);
        }
    }   // CV_REPLAY: run_score's last k_select_cv_chunk already left the selection in pr->sel
    CU(cudaGetLastError());
    CU(cudaEventRecord(pr->ev[3], c->stream));
    const int Hs = n == 4 ? 1 : H;
    // the winner's fp32 model as K3 scored it is still in pr->models unless it came from another rank's shard
    const float4* stored = (n > 4 && !keys_host && !keys_dev) ? pr->models.as<float4>() : nullptr;
    int rc = launch_finalize(c, pr->pts.as<PointH>(), n, pr->samples.as<int>(), Hs, pr->sel.as<HSelect>(), thr_sq,
                             p->mask_semantics, p->refine, p->solver, pr->H.as<double>(), pr->mask.as<uint8_t>(),
                             pr->rmask.as<uint8_t>(), pr->info.as<int>(), nullptr, nullptr, Q, stored);
    if (rc) return rc;
    if (n > 4 && (keys_host || keys_dev))
        LAUNCH(c, k_patch_best_iter, (unsigned)((Q + 127) / 128), 128, 0, pr->info.as<int>(), pr->keys.as<unsigned long long>(),
               (long long)p->hyp_begin, H, Q);
    CU(cudaGetLastError());
    CU(cudaEventRecord(pr->ev[4], c->stream));
    return B2R_OK;
}

static int fetch(b2r_ctx* c, b2r_h_problem* pr, double* H_out, uint8_t* mask_out, b2r_h_info* info_out) {
    const int Q = pr->Q, n = pr->n;
    const size_t hb = sizeof(double) * 9 * (size_t)Q, mb = (size_t)Q * n, ib = sizeof(int) * 12 * (size_t)Q;
    CU(c->pin_out.reserve(hb + mb + ib + 64));
    char* base = (char*)c->pin_out.p;
    CU(cudaMemcpyAsync(base, pr->H.p, hb, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(base + hb, pr->info.p, ib, cudaMemcpyDeviceToHost, c->stream));
    if (mask_out) CU(cudaMemcpyAsync(base + hb + ib, pr->mask.p, mb, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (H_out) memcpy(H_out, base, hb);
    if (info_out) memcpy(info_out, base + hb, ib);
    if (mask_out) memcpy(mask_out, base + hb + ib, mb);
    for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&pr->stage_ms[i], pr->ev[i], pr->ev[i + 1]);
    cudaEventElapsedTime(&pr->stage_ms[4], pr->ev[0], pr->ev[4]);
    cudaGetLastError();
    return B2R_OK;
}

static_assert(sizeof(b2r_h_info) == 12 * sizeof(int32_t), "b2r_h_info layout");

extern "C" {

b2r_h_problem* b2r_h_problem_upload(b2r_ctx* c, const double* src, const double* dst, int32_t dst_shared, int32_t Q,
                                    int32_t n) {
    if (!c || !src || !dst || Q < 1 || Q > 65535 || n < 4) {
        fail(B2R_ERR_ARG, "b2r_h_problem_upload: need 1 <= Q <= 65535 and n >= 4 points (cv2.findHomography raises below 4)%s%s");
        return nullptr;
    }
    if (cudaSetDevice(c->device) != cudaSuccess) {
        fail(B2R_ERR_CUDA, "cudaSetDevice failed%s%s");
        return nullptr;
    }
    b2r_h_problem* pr = new b2r_h_problem();
    if (pr->pts.reserve(sizeof(PointH) * (size_t)Q * n) != cudaSuccess || upload_points(c, pr, src, dst, dst_shared, Q, n) ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
        if (g_err.empty()) fail(B2R_ERR_CUDA, "upload failed%s%s");
        pr->release();
        delete pr;
        return nullptr;
    }
    return pr;
}

int b2r_h_problem_reupload(b2r_ctx* c, b2r_h_problem* pr, const double* src, const double* dst, int32_t dst_shared,
                           int32_t Q, int32_t n) {
    if (!c || !pr || !src || !dst || Q < 1 || Q > 65535 || n < 4) return fail(B2R_ERR_ARG, "bad argument (need 1 <= Q <= 65535, n >= 4)%s%s");
    CU(cudaSetDevice(c->device));
    CU(pr->pts.reserve(sizeof(PointH) * (size_t)Q * n));
    return upload_points(c, pr, src, dst, dst_shared, Q, n);
}

void b2r_h_problem_free(b2r_ctx* c, b2r_h_problem* pr) {
    if (!pr) return;
    if (c) cudaSetDevice(c->device);
    pr->release();
    delete pr;
}

int b2r_h_problem_run(b2r_ctx* c, b2r_h_problem* pr, const b2r_h_params* p) {
    if (!c || !pr) return fail(B2R_ERR_ARG, "null argument%s%s");
    int rc = check_params(p);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    if ((rc = run_score(c, pr, p))) return rc;
    return run_finish(c, pr, p, nullptr);
}

int b2r_h_problem_score_shard(b2r_ctx* c, b2r_h_problem* pr, const b2r_h_params* p, uint64_t* keys_out) {
    if (!c || !pr || !keys_out) return fail(B2R_ERR_ARG, "null argument%s%s");
    int rc = check_params(p);
    if (rc) return rc;
    if (p->sampler != B2R_SAMPLER_PHILOX) return fail(B2R_ERR_ARG, "hypothesis sharding needs the PHILOX sampler%s%s");
    if (pr->n <= 4) return fail(B2R_ERR_ARG, "sharding needs n > 4%s%s");
    CU(cudaSetDevice(c->device));
    if ((rc = run_score(c, pr, p))) return rc;
    CU(cudaMemcpyAsync(keys_out, pr->keys.p, sizeof(uint64_t) * pr->Q, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2R_OK;
}

int b2r_h_problem_finish(b2r_ctx* c, b2r_h_problem* pr, const b2r_h_params* p, const uint64_t* keys) {
    if (!c || !pr || !keys) return fail(B2R_ERR_ARG, "null argument%s%s");
    int rc = check_params(p);
    if (rc) return rc;
    if (p->sampler != B2R_SAMPLER_PHILOX) return fail(B2R_ERR_ARG, "finishing from reduced keys needs the PHILOX sampler%s%s");
    CU(cudaSetDevice(c->device));
    return run_finish(c, pr, p, keys);
}

int b2r_h_problem_score_shard_dev(b2r_ctx* c, b2r_h_problem* pr, const b2r_h_params* p, uint64_t* keys_dev_out) {
    if (!c || !pr || !keys_dev_out) return fail(B2R_ERR_ARG, "null argument%s%s");
    int rc = check_params(p);
    if (rc) return rc;
    if (p->sampler != B2R_SAMPLER_PHILOX) return fail(B2R_ERR_ARG, "hypothesis sharding needs the PHILOX sampler%s%s");
    if (pr->n <= 4) return fail(B2R_ERR_ARG, "sharding needs n > 4%s%s");
    CU(cudaSetDevice(c->device));
    if ((rc = run_score(c, pr, p))) return rc;
    CU(cudaMemcpyAsync(keys_dev_out, pr->keys.p, sizeof(uint64_t) * pr->Q, cudaMemcpyDeviceToDevice, c->stream));
    return B2R_OK;
}

int b2r_h_problem_finish_dev(b2r_ctx* c, b2r_h_problem* pr, const b2r_h_params* p, const uint64_t* keys_dev) {
    if (!c || !pr || !keys_dev) return fail(B2R_ERR_ARG, "null argument%s%s");
    int rc = check_params(p);
    if (rc) return rc;
    if (p->sampler != B2R_SAMPLER_PHILOX) return fail(B2R_ERR_ARG, "finishing from reduced keys needs the PHILOX sampler%s%s");
    CU(cudaSetDevice(c->device));
    return run_finish(c, pr, p, nullptr, keys_dev);
}

int b2r_h_problem_fetch(b2r_ctx* c, b2r_h_problem* pr, double* H_out, uint8_t* mask_out, b2r_h_info* info_out) {
    if (!c || !pr) return fail(B2R_ERR_ARG, "null argument%s%s");
    CU(cudaSetDevice(c->device));
    return fetch(c, pr, H_out, mask_out, info_out);
}

int b2r_h_problem_peek_hyps(b2r_ctx* c, b2r_h_problem* pr, int32_t q, int32_t first, int32_t count, int32_t* samples_out,
                            float* models8_out, int32_t* counts_out) {
    if (!c || !pr || q < 0 || q >= pr->Q || first < 0 || count < 1 || (long long)first + count > pr->H_last)
        return fail(B2R_ERR_ARG, "peek_hyps: slots [first, first + count) must lie inside the last run's hypotheses%s%s");
    CU(cudaSetDevice(c->device));
    const size_t slot = (size_t)q * pr->H_last + first;
    if (samples_out)
        CU(cudaMemcpyAsync(samples_out, pr->samples.as<int>() + 4 * slot, sizeof(int) * 4 * (size_t)count, cudaMemcpyDeviceToHost, c->stream));
    if (models8_out)
        CU(cudaMemcpyAsync(models8_out, pr->models.as<float>() + 8 * slot, sizeof(float) * 8 * (size_t)count, cudaMemcpyDeviceToHost, c->stream));
    if (counts_out)
        CU(cudaMemcpyAsync(counts_out, pr->counts.as<int>() + slot, sizeof(int) * (size_t)count, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2R_OK;
}

int b2r_h_problem_stage_ms(b2r_ctx* c, b2r_h_problem* pr, float ms_out[5]) {
    if (!c || !pr || !ms_out) return fail(B2R_ERR_ARG, "null argument%s%s");
    memcpy(ms_out, pr->stage_ms, sizeof(float) * 5);
    return B2R_OK;
}

int b2r_find_homography_batch(b2r_ctx* c, const double* src, const double* dst, int32_t dst_shared, int32_t Q, int32_t n,
                              const b2r_h_params* p, double* H_out, uint8_t* mask_out, b2r_h_info* info_out) {
    if (!c || !src || !dst || !H_out) return fail(B2R_ERR_ARG, "null argument%s%s");
    if (Q < 1 || Q > 65535) return fail(B2R_ERR_ARG, "Q must be in 1..65535 (one grid dimension per problem): split larger batches%s%s");
    if (n < 4) return fail(B2R_ERR_ARG, "findHomography needs at least 4 point pairs (cv2 raises cv2.error)%s%s");
    int rc = check_params(p);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    if (!c->cached) c->cached = new b2r_h_problem();
    b2r_h_problem* pr = c->cached;
    CU(pr->pts.reserve(sizeof(PointH) * (size_t)Q * n));
    if ((rc = upload_points(c, pr, src, dst, dst_shared, Q, n))) return rc;
    if ((rc = run_score(c, pr, p))) return rc;
    if ((rc = run_finish(c, pr, p, nullptr))) return rc;
    return fetch(c, pr, H_out, mask_out, info_out);
}

int b2r_find_homography(b2r_ctx* c, const double* src, const double* dst, int32_t n, const b2r_h_params* p, double* H_out,
                        uint8_t* mask_out, b2r_h_info* info_out) {
    b2r_h_info info;
    int rc = b2r_find_homography_batch(c, src, dst, 1, 1, n, p, H_out, mask_out, &info);
    if (rc) return rc;
    if (info_out) *info_out = info;
    return info.status;
}


// ---- the camera-location sweep (find_homographies + arg-min), fused --------------------------------------------------
int b2r_camera_sweep(b2r_ctx* c, const double* pos3d, const double* pixels, int32_t n, const double* cams, int32_t Q,
                     const b2r_h_params* p, double* scores_out, double* M_out, double* H_out, uint8_t* mask_out,
                     b2r_h_info* info_out, int32_t* best_out) {
    if (!c || !pos3d || !pixels || !cams || !scores_out) return fail(B2R_ERR_ARG, "null argument%s%s");
    if (Q < 1 || Q > 65535) return fail(B2R_ERR_ARG, "Q must be in 1..65535 (one grid dimension per problem): split larger batches%s%s");
    if (n < 4) return fail(B2R_ERR_ARG, "findHomography needs at least 4 point pairs (cv2 raises cv2.error)%s%s");
    int rc = check_params(p);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    if (!c->cached) c->cached = new b2r_h_problem();
    b2r_h_problem* pr = c->cached;
    const size_t total = (size_t)Q * n;
    CU(pr->pts.reserve(sizeof(PointH) * total));
    // layout of the sweep buffer (doubles): pos3d 3n | pixels 2n | cams 3Q | pos2 2Qn | scores 2Q | M 9Q | best (int)
    const size_t o_pix = 3 * (size_t)n, o_cam = o_pix + 2 * (size_t)n, o_pos2 = o_cam + 3 * (size_t)Q, o_sc = o_pos2 + 2 * total,
                 o_M = o_sc + 2 * (size_t)Q, o_best = o_M + 9 * (size_t)Q;
    CU(pr->sweep.reserve(sizeof(double) * (o_best + 1)));
    double* sw = pr->sweep.as<double>();
    CU(cudaMemcpyAsync(sw, pos3d, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(sw + o_pix, pixels, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(sw + o_cam, cams, sizeof(double) * 3 * Q, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_sweep_prologue, (unsigned)((total + 255) / 256), 256, 0, sw, sw + o_pix, sw + o_cam, Q, n, sw + o_pos2,
           pr->pts.as<PointH>());
    CU(cudaGetLastError());
    pr->Q = Q;
    pr->n = n;
    if ((rc = run_score(c, pr, p))) return rc;
    if ((rc = run_finish(c, pr, p, nullptr))) return rc;
    LAUNCH(c, k_sweep_epilogue, (unsigned)Q, 128, 0, pr->H.as<double>(), pr->mask.as<uint8_t>(), pr->info.as<int>(), sw + o_pos2,
           sw + o_pix, n, p->thr, sw + o_sc, sw + o_M);
    LAUNCH(c, k_sweep_argmin, 1, 256, 0, sw + o_sc, Q, reinterpret_cast<int*>(sw + o_best));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(scores_out, sw + o_sc, sizeof(double) * 2 * Q, cudaMemcpyDeviceToHost, c->stream));
    if (M_out) CU(cudaMemcpyAsync(M_out, sw + o_M, sizeof(double) * 9 * Q, cudaMemcpyDeviceToHost, c->stream));
    int best = 0;
    CU(cudaMemcpyAsync(&best, sw + o_best, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    rc = fetch(c, pr, H_out, mask_out, info_out);   // synchronises the stream
    if (rc) return rc;
    if (best_out) *best_out = best;
    return B2R_OK;
}

// ---- building blocks ------------------------------------------------------------------------------------------------
static int upload_f32_points(b2r_ctx* c, const float* src, const float* dst, int n, DevBuf& out) {
    CU(c->in_a.reserve(sizeof(float) * 2 * (size_t)n));
    CU(c->in_b.reserve(sizeof(float) * 2 * (size_t)n));
    CU(out.reserve(sizeof(PointH) * (size_t)n));
    CU(cudaMemcpyAsync(c->in_a.p, src, sizeof(float) * 2 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->in_b.p, dst, sizeof(float) * 2 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_pack_points_h_f32, (unsigned)((n + 255) / 256), 256, 0, c->in_a.as<float>(), c->in_b.as<float>(), n,
           out.as<PointH>());
    CU(cudaGetLastError());
    return B2R_OK;
}

int b2r_score_h(b2r_ctx* c, const float* models, int32_t n_models, const float* src, const float* dst, int32_t n,
                float thr_sq, int32_t arith, int32_t* counts_out) {
    if (!c || !models || !src || !dst || !counts_out || n_models < 1 || n < 1) return fail(B2R_ERR_ARG, "bad argument%s%s");
    CU(cudaSetDevice(c->device));
    int rc = upload_f32_points(c, src, dst, n, c->scratch0);
    if (rc) return rc;
    CU(c->scratch1.reserve(sizeof(float) * 8 * (size_t)n_models));
    CU(c->scratch2.reserve(sizeof(int) * (size_t)n_models));
    CU(cudaMemcpyAsync(c->scratch1.p, models, sizeof(float) * 8 * (size_t)n_models, cudaMemcpyHostToDevice, c->stream));
    if ((rc = score_models(c, c->scratch1.as<float4>(), n_models, c->scratch0.as<PointH>(), n, thr_sq, c->scratch2.as<int>(),
                           1, arith)))
        return rc;
    CU(cudaMemcpyAsync(counts_out, c->scratch2.p, sizeof(int) * (size_t)n_models, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2R_OK;
}

int b2r_solve_h4(b2r_ctx* c, const float* src, const float* dst, int32_t n, const int32_t* idx, int32_t n_samples,
                 int32_t solver, double* H_out, uint8_t* ok_out, uint8_t* subset_ok_out) {
    if (!c || !src || !dst || !idx || !H_out || n < 4 || n_samples < 1) return fail(B2R_ERR_ARG, "bad argument%s%s");
    for (size_t i = 0; i < (size_t)n_samples * 4; ++i)
        if (idx[i] < 0 || idx[i] >= n) return fail(B2R_ERR_ARG, "sample index out of range%s%s");
    CU(cudaSetDevice(c->device));
    int rc = upload_f32_points(c, src, dst, n, c->scratch0);
    if (rc) return rc;
    CU(c->scratch1.reserve(sizeof(int) * 4 * (size_t)n_samples));
    CU(c->scratch2.reserve(sizeof(double) * 9 * (size_t)n_samples));
    CU(c->scratch3.reserve(2 * (size_t)n_samples));
    CU(cudaMemcpyAsync(c->scratch1.p, idx, sizeof(int) * 4 * (size_t)n_samples, cudaMemcpyHostToDevice, c->stream));
    uint8_t* ok_d = c->scratch3.as<uint8_t>();
    uint8_t* sub_d = ok_d + n_samples;
    if (solver == B2R_SOLVER_EXACT_WARP) {
        LAUNCH(c, k_solve_h4_warp, dim3((unsigned)((n_samples + 7) / 8), 1), 256, 0, c->scratch0.as<PointH>(), n,
               c->scratch1.as<int>(), n_samples, 0, n_samples, (const RansacState*)nullptr, (float4*)nullptr, c->scratch2.as<double>(),
               ok_d, sub_d);
    } else if (solver == B2R_SOLVER_EXACT) {
        if ((rc = k2s_prepare(c))) return rc;
        LAUNCH(c, k_solve_h4_smem, dim3((unsigned)((n_samples + K2S_THREADS - 1) / K2S_THREADS), 1), K2S_THREADS, K2S_SMEM,
               c->scratch0.as<PointH>(), n, c->scratch1.as<int>(), n_samples, 0, n_samples, (const RansacState*)nullptr,
               (float4*)nullptr, c->scratch2.as<double>(), ok_d, sub_d);
    } else
        LAUNCH(c, k_solve_h4, dim3((unsigned)((n_samples + 127) / 128), 1), 128, 0, c->scratch0.as<PointH>(), n,
               c->scratch1.as<int>(), n_samples, 0, n_samples, (const RansacState*)nullptr, (float4*)nullptr, c->scratch2.as<double>(),
               ok_d, sub_d, solver);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(H_out, c->scratch2.p, sizeof(double) * 9 * (size_t)n_samples, cudaMemcpyDeviceToHost, c->stream));
    if (ok_out) CU(cudaMemcpyAsync(ok_out, ok_d, (size_t)n_samples, cudaMemcpyDeviceToHost, c->stream));
    if (subset_ok_out) CU(cudaMemcpyAsync(subset_ok_out, sub_d, (size_t)n_samples, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2R_OK;
}

int b2r_sample_cv(b2r_ctx* c, const float* src, const float* dst, int32_t n, int32_t n_iters, int32_t* idx_out,
                  int32_t* n_generated_out) {
    if (!c || !src || !dst || !idx_out || n < 5 || n_iters < 1) return fail(B2R_ERR_ARG, "bad argument (need n > 4)%s%s");
    CU(cudaSetDevice(c->device));
    int rc = upload_f32_points(c, src, dst, n, c->scratch0);
    if (rc) return rc;
    CU(c->scratch1.reserve(sizeof(int) * 4 * (size_t)n_iters));
    CU(c->scratch2.reserve(sizeof(RansacState)));
    LAUNCH(c, k_state_init, 1, 32, 0, c->scratch2.as<RansacState>(), (int)n_iters, 1);
    LAUNCH(c, k_cv_sample_h, 1, 32, 0, c->scratch0.as<PointH>(), n, n_iters, 0, n_iters, c->scratch1.as<int>(),
           c->scratch2.as<RansacState>(), 1);
    CU(cudaGetLastError());
    RansacState st_h;
    CU(cudaMemcpyAsync(&st_h, c->scratch2.p, sizeof(RansacState), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const int gen = st_h.gen;
    CU(cudaMemcpyAsync(idx_out, c->scratch1.p, sizeof(int) * 4 * (size_t)gen, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (n_generated_out) *n_generated_out = gen;
    return B2R_OK;
}

int b2r_sample_philox(b2r_ctx* c, const float* src, const float* dst, int32_t n, uint64_t seed, int32_t q, int64_t hyp_begin,
                      int32_t n_hyp, int32_t* idx_out) {
    if (!c || !src || !dst || !idx_out || n < 5 || n_hyp < 1 || q != 0)
        return fail(B2R_ERR_ARG, "bad argument (need n > 4; only problem index 0 is exposed here)%s%s");
    CU(cudaSetDevice(c->device));
    int rc = upload_f32_points(c, src, dst, n, c->scratch0);
    if (rc) return rc;
    CU(c->scratch1.reserve(sizeof(int) * 4 * (size_t)n_hyp));
    LAUNCH(c, k_philox_sample_solve_h<false>, dim3((unsigned)((n_hyp + 127) / 128), 1), 128, 0, c->scratch0.as<PointH>(), n, n_hyp,
           (long long)hyp_begin, seed, c->scratch1.as<int>(), (float4*)nullptr);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(idx_out, c->scratch1.p, sizeof(int) * 4 * (size_t)n_hyp, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2R_OK;
}

}  // extern "C"

__global__ void k_select_dummy(HSelect* sel) {
    HSelect s;
    s.best = 0; s.best_count = 0; s.iters_run = 0; s.pad = 0;
    sel[0] = s;
}

// ---- building block: the three forms of the eigen-solver on caller-supplied matrices -----------------------------------
template <int N>
__global__ void __launch_bounds__(64) k_jacobi_thread(const double* __restrict__ mats, int n_mat, double* __restrict__ Wout,
                                                     double* __restrict__ Vout) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_mat) return;
    double A[N * N], W[N], V[N * N];
    for (int e = 0; e < N * N; ++e) A[e] = mats[(size_t)g * N * N + e];
    jacobi_eig<N>(A, W, V);
    for (int e = 0; e < N; ++e) Wout[(size_t)g * N + e] = W[e];
    for (int e = 0; e < N * N; ++e) Vout[(size_t)g * N * N + e] = V[e];
}

template <int N>
__global__ void __launch_bounds__(256) k_jacobi_warp(const double* __restrict__ mats, int n_mat, double* __restrict__ Wout,
                                                    double* __restrict__ Vout) {
    __shared__ JacobiWarp9 jw[8];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = blockIdx.x * 8 + w;
    if (g >= n_mat) return;
    for (int e = lane; e < N * N; e += 32) jw[w].A[e] = mats[(size_t)g * N * N + e];
    __syncwarp();
    jacobi_eig_warp3<N>(jw[w].A, jw[w].W, jw[w].V, jw[w].indR, jw[w].indC);
    __syncwarp();
    if (lane < N) Wout[(size_t)g * N + lane] = jw[w].W[lane];
    for (int e = lane; e < N * N; e += 32) Vout[(size_t)g * N * N + e] = jw[w].V[e];
}

__global__ void __launch_bounds__(K2S_THREADS) k_jacobi_packed9(const double* __restrict__ mats, int n_mat,
                                                                double* __restrict__ Wout, double* __restrict__ Vout) {
    extern __shared__ __align__(16) double k2s_ws[];
    const int g = blockIdx.x * K2S_THREADS + threadIdx.x;
    if (g >= n_mat) return;
    double* U = k2s_ws + threadIdx.x;
    double* V = U + 45 * K2S_THREADS;
    for (int r = 0, e = 0; r < 9; ++r)
        for (int c = r; c < 9; ++c, ++e) U[e * K2S_THREADS] = mats[(size_t)g * 81 + r * 9 + c];
    jacobi_eig_packed<9, K2S_THREADS>(U, V);
    for (int r = 0; r < 9; ++r) Wout[(size_t)g * 9 + r] = U[(((r * (17 - r)) >> 1) + r) * K2S_THREADS];
    for (int e = 0; e < 81; ++e) Vout[(size_t)g * 81 + e] = V[e * K2S_THREADS];
}

template <int N>
static int jacobi_launch(b2r_ctx* c, int form, const double* mats, int n_mat, double* W, double* V) {
    if (form == 0) LAUNCH(c, k_jacobi_thread<N>, (unsigned)((n_mat + 63) / 64), 64, 0, mats, n_mat, W, V);
    else LAUNCH(c, k_jacobi_warp<N>, (unsigned)((n_mat + 7) / 8), 256, 0, mats, n_mat, W, V);
    return B2R_OK;
}

extern "C" {

int b2r_jacobi_eig(b2r_ctx* c, const double* A, int32_t n_mat, int32_t n, int32_t form, double* W_out, double* V_out) {
    if (!c || !A || !W_out || !V_out || n_mat < 1) return fail(B2R_ERR_ARG, "null argument or no matrices%s%s");
    if (n < 2 || n > 9 || form < 0 || form > 2 || (form == 2 && n != 9))
        return fail(B2R_ERR_ARG, "b2r_jacobi_eig: 2 <= n <= 9, form 0..2 (form 2: n = 9 only)%s%s");
    CU(cudaSetDevice(c->device));
    const size_t ab = sizeof(double) * (size_t)n_mat * n * n, wb = sizeof(double) * (size_t)n_mat * n;
    CU(c->scratch0.reserve(ab));
    CU(c->scratch1.reserve(ab));
    CU(c->scratch2.reserve(wb));
    double *dA = c->scratch0.as<double>(), *dV = c->scratch1.as<double>(), *dW = c->scratch2.as<double>();
    CU(cudaMemcpyAsync(dA, A, ab, cudaMemcpyHostToDevice, c->stream));
    int rc = B2R_OK;
    if (form == 2) {
        if ((rc = k2s_prepare(c))) return rc;
        LAUNCH(c, k_jacobi_packed9, (unsigned)((n_mat + K2S_THREADS - 1) / K2S_THREADS), K2S_THREADS, K2S_SMEM, dA, n_mat, dW, dV);
    } else {
        switch (n) {
            case 2: rc = jacobi_launch<2>(c, form, dA, n_mat, dW, dV); break;
            case 3: rc = jacobi_launch<3>(c, form, dA, n_mat, dW, dV); break;
            case 4: rc = jacobi_launch<4>(c, form, dA, n_mat, dW, dV); break;
            case 5: rc = jacobi_launch<5>(c, form, dA, n_mat, dW, dV); break;
            case 6: rc = jacobi_launch<6>(c, form, dA, n_mat, dW, dV); break;
            case 7: rc = jacobi_launch<7>(c, form, dA, n_mat, dW, dV); break;
            case 8: rc = jacobi_launch<8>(c, form, dA, n_mat, dW, dV); break;
            default: rc = jacobi_launch<9>(c, form, dA, n_mat, dW, dV); break;
        }
    }
    if (rc) return rc;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(W_out, dW, wb, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(V_out, dV, ab, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2R_OK;
}

// refine-only entry: a finalize launch whose winning sample is replaced by a caller-supplied model and mask
int b2r_refine_h(b2r_ctx* c, const float* src, const float* dst, int32_t n, const uint8_t* mask, double* H_io,
                 int32_t* lm_iters_out) {
    if (!c || !src || !dst || !mask || !H_io || n < 5) return fail(B2R_ERR_ARG, "bad argument (need n > 4)%s%s");
    int k = 0;
    for (int i = 0; i < n; ++i) k += mask[i] != 0;
    if (k < 4) return fail(B2R_ERR_ARG, "refinement needs at least 4 masked points%s%s");
    CU(cudaSetDevice(c->device));
    int rc = upload_f32_points(c, src, dst, n, c->scratch0);
    if (rc) return rc;
    // scratch1: [mask n][rmask n][mask_out n] ; scratch2: [H_in 9][H_out 9] doubles + sel + info
    CU(c->scratch1.reserve(3 * (size_t)n));
    CU(c->scratch2.reserve(sizeof(double) * 18 + sizeof(HSelect) + sizeof(int) * 12 + 64));
    uint8_t* m_in = c->scratch1.as<uint8_t>();
    double* H_in = c->scratch2.as<double>();
    double* H_out = H_in + 9;
    HSelect* sel = reinterpret_cast<HSelect*>(H_out + 9);
    int* info = reinterpret_cast<int*>(sel + 1);
    CU(cudaMemcpyAsync(m_in, mask, (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(H_in, H_io, sizeof(double) * 9, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_select_dummy, 1, 1, 0, sel);
    if ((rc = launch_finalize(c, c->scratch0.as<PointH>(), n, nullptr, 1, sel, 0.f, B2R_MASK_LEGACY, 1, B2R_SOLVER_EXACT, H_out,
                              m_in + 2 * (size_t)n, m_in + (size_t)n, info, m_in, H_in, 1)))
        return rc;
    int info_h[12];
    CU(cudaMemcpyAsync(H_io, H_out, sizeof(double) * 9, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(info_h, info, sizeof(info_h), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (lm_iters_out) *lm_iters_out = info_h[9];
    return B2R_OK;
}

int b2r_selftest_rcp(b2r_ctx* c, uint64_t* mismatches_out, uint64_t* tested_out) {
    if (!c || !mismatches_out || !tested_out) return fail(B2R_ERR_ARG, "null argument%s%s");
    CU(cudaSetDevice(c->device));
    CU(c->scratch0.reserve(16));
    CU(cudaMemsetAsync(c->scratch0.p, 0, 16, c->stream));
    LAUNCH(c, k_selftest_rcp, (unsigned)(c->sm_count * 8), 256, 0, c->scratch0.as<unsigned long long>(),
           c->scratch0.as<unsigned long long>() + 1);
    CU(cudaGetLastError());
    unsigned long long h[2];
    CU(cudaMemcpyAsync(h, c->scratch0.p, 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *mismatches_out = h[0];
    *tested_out = h[1];
    return B2R_OK;
}

int b2r_probe_fp32_peak(b2r_ctx* c, double* fma_per_s_out, double* ffma2_per_s_out) {
    if (!c || !fma_per_s_out) return fail(B2R_ERR_ARG, "null argument%s%s");
    CU(cudaSetDevice(c->device));
    const int ctas = c->sm_count * 8, threads = 256, iters = 8192;
    CU(c->scratch0.reserve(sizeof(float) * (size_t)ctas * threads));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    for (int packed = 0; packed < 2; ++packed) {
        float best = 1e30f;
        for (int r = 0; r < 8; ++r) {
            CU(cudaEventRecord(e0, c->stream));
            if (packed)
                LAUNCH(c, k_probe_fma<1>, ctas, threads, 0, c->scratch0.as<float>(), iters, 1.0001f, 0.5f);
            else
                LAUNCH(c, k_probe_fma<0>, ctas, threads, 0, c->scratch0.as<float>(), iters, 1.0001f, 0.5f);
            CU(cudaEventRecord(e1, c->stream));
            CU(cudaEventSynchronize(e1));
            float ms;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 3 && ms < best) best = ms;
        }
        const double lane_fma = (double)iters * 4 * 8 * (packed ? 2 : 1) * (double)ctas * threads;
        if (packed) {
            if (ffma2_per_s_out) *ffma2_per_s_out = lane_fma / (best * 1e-3);
        } else {
            *fma_per_s_out = lane_fma / (best * 1e-3);
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return B2R_OK;
}

}  // extern "C"

#ifdef B2R_FIN_PROFILE
// development probe (see FINCLK in pipeline_h.cuh): read (and optionally clear) the per-section cycle counters
extern "C" __attribute__((visibility("default"))) int b2r_debug_fin_clocks(unsigned long long* out, int reset) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (out && cudaMemcpyFromSymbol(out, b2r::g_fin_clk, sizeof(unsigned long long) * 32) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long z[32] = {0};
        if (cudaMemcpyToSymbol(b2r::g_fin_clk, z, sizeof(z)) != cudaSuccess) return -1;
    }
    return 0;
}
#endif

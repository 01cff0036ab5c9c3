// C ABI of libransac_b200.so, PnP path: cv2.solvePnPRansac (main_v1.py:497-502; testpro-K.py:72-75 as a batch over
// intrinsics) and cv2.solvePnPRefineLM (main_v1.py:508-509).  Host-side orchestration only; every arithmetic step
// runs in the kernels of pipeline_p.cuh / score_p.cuh / pnp_solver.cuh.  No CPU fallback.
#include "host_common.h"
#include "pipeline_p.cuh"
#include "score_p_filt.cuh"

using namespace b2r;

struct b2r_p_problem {
    int Q = 0, n = 0, P = 0;  // Q problems (intrinsics), n points, P point sets (1 = shared by all problems, else Q)
    DevBuf raw_obj, raw_img;  // [P][n][3], [P][n][2] fp64 as the caller passed them
    DevBuf px, pf, centre;    // [P][n] PointPX, PointPF; [P][3] fp64
    DevBuf win;               // [Q][6] fp64: the winner's minimal model (k_winner_model_p)
    DevBuf Kq;                // [Q][4] fx, fy, cx, cy
    DevBuf samples;           // [Q][H][5] int32
    DevBuf mx, mf;            // [Q][H][12] fp64 / fp32 models
    DevBuf rt;                // [Q][H][6] rvec | tvec of every hypothesis (replay path: the finalize kernel reads the winner's)
    bool rt_valid = false;
    DevBuf counts, state, keys, sel;   // state: [Q] RansacState + one int32 "not done" counter
    DevBuf rmask, pose, info_i, info_d, inliers, ninl;
    int H_last = 0;
    float stage_ms[5] = {0, 0, 0, 0, 0};
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    void release() {
        raw_obj.release(); raw_img.release(); px.release(); pf.release(); centre.release(); Kq.release(); win.release();
        samples.release(); mx.release(); mf.release(); rt.release(); counts.release(); state.release(); keys.release(); sel.release();
        rmask.release(); pose.release(); info_i.release(); info_d.release(); inliers.release(); ninl.release();
        for (auto& e : ev)
            if (e) cudaEventDestroy(e), e = nullptr;
    }
    size_t pts_stride() const { return P == 1 ? 0 : (size_t)n; }
};

void b2r_p_problem_destroy(b2r_p_problem* pr) {
    if (!pr) return;
    pr->release();
    delete pr;
}

static int check_p_params(const b2r_p_params* p) {
    if (!p) return fail(B2R_ERR_ARG, "null params%s%s");
    if (!(p->thr >= 0)) return fail(B2R_ERR_ARG, "reprojectionError must be >= 0%s%s");
    if (p->sampler != B2R_SAMPLER_CV_REPLAY && p->sampler != B2R_SAMPLER_PHILOX) return fail(B2R_ERR_ARG, "bad sampler%s%s");
    if (p->arith != B2R_ARITH_EXACT && p->arith != B2R_ARITH_FAST && p->arith != B2R_ARITH_EXACT_UNFILTERED)
        return fail(B2R_ERR_ARG, "bad arith%s%s");
    if (p->solver != B2R_SOLVER_EXACT && p->solver != B2R_SOLVER_FAST) return fail(B2R_ERR_ARG, "bad solver%s%s");
    if (p->max_iters > (1 << 30)) return fail(B2R_ERR_ARG, "iterationsCount too large%s%s");
    // the arg-max keys carry the global hypothesis id in 32 bits (count << 32 | ~id): ids beyond 2^32 would alias
    if (p->sampler == B2R_SAMPLER_PHILOX &&
        (p->hyp_begin < 0 || p->hyp_begin + (long long)(p->max_iters > 1 ? p->max_iters : 1) > (1LL << 32)))
        return fail(B2R_ERR_ARG, "PHILOX hypothesis ids must lie in [0, 2^32): hyp_begin >= 0 and hyp_begin + iterationsCount <= 2^32%s%s");
    return B2R_OK;
}

static int k4_from_K(const double* K, int Q, std::vector<double>& out) {
    out.resize((size_t)Q * 4);
    for (int q = 0; q < Q; ++q) {
        const double* k = K + 9 * (size_t)q;
        out[4 * q] = k[0]; out[4 * q + 1] = k[4]; out[4 * q + 2] = k[2]; out[4 * q + 3] = k[5];
        if (!(k[0] != 0) || !(k[4] != 0)) return fail(B2R_ERR_ARG, "camera matrix needs non-zero focal lengths%s%s");
    }
    return B2R_OK;
}

static int p_upload(b2r_ctx* c, b2r_p_problem* pr, const double* obj, const double* img, int pts_shared, int Q, int n,
                    const double* K) {
    const int P = pts_shared ? 1 : Q;
    std::vector<double> k4;
    int rc = k4_from_K(K, Q, k4);
    if (rc) return rc;
    const size_t ob = sizeof(double) * 3 * (size_t)P * n, ib = sizeof(double) * 2 * (size_t)P * n;
    CU(pr->raw_obj.reserve(ob));
    CU(pr->raw_img.reserve(ib));
    CU(pr->px.reserve(sizeof(PointPX) * (size_t)P * n));
    CU(pr->pf.reserve(sizeof(PointPF) * (size_t)P * n));
    CU(pr->centre.reserve(sizeof(double) * 3 * (size_t)P));
    CU(pr->Kq.reserve(sizeof(double) * 4 * (size_t)Q));
    CU(cudaMemcpyAsync(pr->raw_obj.p, obj, ob, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(pr->raw_img.p, img, ib, cudaMemcpyHostToDevice, c->stream));
    CU(c->pin_in.reserve(sizeof(double) * 4 * (size_t)Q));
    CU(cudaStreamSynchronize(c->stream));  // pin_in may still feed an earlier copy
    memcpy(c->pin_in.p, k4.data(), sizeof(double) * 4 * (size_t)Q);
    CU(cudaMemcpyAsync(pr->Kq.p, c->pin_in.p, sizeof(double) * 4 * (size_t)Q, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_centre_p, (unsigned)P, 256, 0, pr->raw_obj.as<double>(), n, pr->centre.as<double>());
    const size_t total = (size_t)P * n;
    LAUNCH(c, k_pack_points_p, (unsigned)((total + 255) / 256), 256, 0, pr->raw_obj.as<double>(), pr->raw_img.as<double>(), P, n,
           pr->centre.as<double>(), pr->px.as<PointPX>(), pr->pf.as<PointPF>());
    CU(cudaGetLastError());
    pr->Q = Q; pr->n = n; pr->P = P;
    return B2R_OK;
}

static int p_reserve(b2r_p_problem* pr, int Q, int n, int H, bool exact) {
    CU(pr->samples.reserve(sizeof(int) * PNP_MP * (size_t)Q * H));
    if (exact) CU(pr->mx.reserve(sizeof(double) * 12 * (size_t)Q * H));
    else CU(pr->mf.reserve(sizeof(float) * 12 * (size_t)Q * H));
    CU(pr->counts.reserve(sizeof(int) * (size_t)Q * H));
    CU(pr->state.reserve(sizeof(RansacState) * (size_t)Q + 64));
    CU(pr->keys.reserve(sizeof(unsigned long long) * (size_t)Q));
    CU(pr->sel.reserve(sizeof(HSelect) * (size_t)Q));
    CU(pr->rmask.reserve((size_t)Q * n));
    CU(pr->pose.reserve(sizeof(double) * 6 * (size_t)Q));
    CU(pr->info_i.reserve(sizeof(int) * 12 * (size_t)Q));
    CU(pr->info_d.reserve(sizeof(double) * 8 * (size_t)Q));
    CU(pr->inliers.reserve(sizeof(int) * (size_t)Q * n));
    CU(pr->ninl.reserve(sizeof(int) * (size_t)Q));
    for (auto& e : pr->ev)
        if (!e) CU(cudaEventCreate(&e));
    return B2R_OK;
}

static int pick_tile(b2r_ctx* c, long long hyp_blocks, int n, int Q, int max_tile) {
    int tile = max_tile;
    while (tile > 128 && hyp_blocks * ((n + tile - 1) / tile) * Q < 8LL * c->sm_count) tile >>= 1;
    if (tile > n) tile = ((n + 7) / 8) * 8;
    return tile;
}

static int zero_counts(b2r_ctx* c, int* counts, int Q, int H, int H_stride, int begin) {
    if (begin == 0 && H == H_stride) CU(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)Q * H, c->stream));
    else CU(cudaMemset2DAsync(counts + begin, sizeof(int) * (size_t)H_stride, 0, sizeof(int) * (size_t)H, (size_t)Q, c->stream));
    return B2R_OK;
}

// score hypotheses [begin, begin+H) of every problem; mx/counts are [Q][H_stride] arrays
// pf / centre != nullptr: the filtered predicate (score_p_filt.cuh: the same counts); nullptr: OpenCV's sequence for every evaluation
static int score_p_exact(b2r_ctx* c, const double* mx, int H, const PointPX* px, size_t stride, int n, const double* Kq,
                         float thr_sq, int* counts, int Q, int H_stride = 0, int begin = 0, const PointPF* pf = nullptr,
                         const double* centre = nullptr, size_t centre_stride = 0) {
    if (H_stride == 0) H_stride = H;
    int rc = zero_counts(c, counts, Q, H, H_stride, begin);
    if (rc) return rc;
    if (pf && centre && n >= 64) {   // fewer points cannot repay the per-hypothesis set-up of the filter
        constexpr int NP = 2;
        const long long hb = (H + K3_THREADS * 2 * NP - 1) / (K3_THREADS * 2 * NP);
        const int tile = pick_tile(c, hb, n, Q, 1024);
        static bool optin[64] = {false};   // > 48 KB of dynamic shared memory: once per device
        if (c->device < 64 && !optin[c->device]) {
            CU(cudaFuncSetAttribute(k3_score_p_filt<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k3p_filt_smem(1024, NP)));
            optin[c->device] = true;
        }
        dim3 grid((unsigned)hb, (unsigned)((n + tile - 1) / tile), (unsigned)Q);
        LAUNCH(c, (k3_score_p_filt<NP>), grid, K3_THREADS, k3p_filt_smem(tile, NP), mx + 12 * (size_t)begin, H, H_stride, px, pf, stride, n, Kq,
               centre, centre_stride, thr_sq, counts + begin, tile);
        CU(cudaGetLastError());
        return B2R_OK;
    }
    const bool two = (long long)H * Q > 1024LL * c->sm_count;
    const int per_cta = K3P_THREADS * (two ? 2 : 1);
    const long long hb = (H + per_cta - 1) / per_cta;
    const int tile = pick_tile(c, hb, n, Q, 1024);
    dim3 grid((unsigned)hb, (unsigned)((n + tile - 1) / tile), (unsigned)Q);
    const size_t smem = 128 + (size_t)tile * 32;
    const double* m0 = mx + 12 * (size_t)begin;
    if (two) LAUNCH(c, (k3_score_p_exact<2>), grid, K3P_THREADS, smem, m0, H, H_stride, px, stride, n, Kq, thr_sq, counts + begin, tile);
    else LAUNCH(c, (k3_score_p_exact<1>), grid, K3P_THREADS, smem, m0, H, H_stride, px, stride, n, Kq, thr_sq, counts + begin, tile);
    CU(cudaGetLastError());
    return B2R_OK;
}

static int score_p_fast(b2r_ctx* c, const float4* mf, int H, const PointPF* pf, size_t stride, int n, float thr_sq, int* counts,
                        int Q, int H_stride = 0, int begin = 0) {
    if (H_stride == 0) H_stride = H;
    int rc = zero_counts(c, counts, Q, H, H_stride, begin);
    if (rc) return rc;
    const bool big = (long long)H * Q >= 32768;   // 2 hypothesis pairs per thread once there are enough hypothesis blocks
    const int per_cta = K3_THREADS * 2 * (big ? 2 : 1);
    const long long hb = (H + per_cta - 1) / per_cta;
    const int tile = pick_tile(c, hb, n, Q, 1024);
    dim3 grid((unsigned)hb, (unsigned)((n + tile - 1) / tile), (unsigned)Q);
    const size_t smem = 128 + (size_t)tile * 32;
    const float4* m0 = mf + 3 * (size_t)begin;
    if (big) LAUNCH(c, (k3_score_p_fast<2>), grid, K3_THREADS, smem, m0, H, H_stride, pf, stride, n, thr_sq, counts + begin, tile);
    else LAUNCH(c, (k3_score_p_fast<1>), grid, K3_THREADS, smem, m0, H, H_stride, pf, stride, n, thr_sq, counts + begin, tile);
    CU(cudaGetLastError());
    return B2R_OK;
}

// stage 1: sample + solve + score (+ per-problem argmax key for PHILOX).  The replay path runs OpenCV's sequential loop in
// growing chunks of iterations (256, 512, ...) and stops once every problem of the batch has reached its iteration bound.
static int p_run_score(b2r_ctx* c, b2r_p_problem* pr, const b2r_p_params* p) {
    const int Q = pr->Q, n = pr->n;
    int H = p->max_iters > 1 ? p->max_iters : 1;
    if (n == PNP_MP) H = 1;  // OpenCV solves the only possible subset once
    const bool exact = p->arith != B2R_ARITH_FAST;
    const bool filt = p->arith == B2R_ARITH_EXACT;
    int rc = p_reserve(pr, Q, n, H, exact);
    if (rc) return rc;
    pr->H_last = H;
    const float thr_sq = (float)(p->thr * p->thr);
    const bool philox = p->sampler == B2R_SAMPLER_PHILOX;
    const size_t cstride = pr->P == 1 ? 0 : 3;
    double* mx = exact ? pr->mx.as<double>() : nullptr;
    float4* mf = exact ? nullptr : pr->mf.as<float4>();
    CU(cudaEventRecord(pr->ev[0], c->stream));
    pr->rt_valid = false;
    if (philox) {
        dim3 grid((unsigned)((H + 63) / 64), (unsigned)Q);
        if (p->solver)
            LAUNCH(c, k_epnp_solve_p<true>, grid, 64, 0, pr->px.as<PointPX>(), pr->pts_stride(), n, H, 0, H, (const RansacState*)nullptr,
                   pr->Kq.as<double>(), pr->centre.as<double>(), cstride, 1, (long long)p->hyp_begin, p->seed, pr->samples.as<int>(), mx,
                   mf, (double*)nullptr, (uint8_t*)nullptr);
        else
            LAUNCH(c, k_epnp_solve_p<false>, grid, 64, 0, pr->px.as<PointPX>(), pr->pts_stride(), n, H, 0, H, (const RansacState*)nullptr,
                   pr->Kq.as<double>(), pr->centre.as<double>(), cstride, 1, (long long)p->hyp_begin, p->seed, pr->samples.as<int>(), mx,
                   mf, (double*)nullptr, (uint8_t*)nullptr);
        CU(cudaGetLastError());
        CU(cudaEventRecord(pr->ev[1], c->stream));
        if (exact) rc = score_p_exact(c, mx, H, pr->px.as<PointPX>(), pr->pts_stride(), n, pr->Kq.as<double>(), thr_sq, pr->counts.as<int>(), Q, 0, 0,
                                      filt ? pr->pf.as<PointPF>() : nullptr, pr->centre.as<double>(), cstride);
        else rc = score_p_fast(c, mf, H, pr->pf.as<PointPF>(), pr->pts_stride(), n, thr_sq, pr->counts.as<int>(), Q);
        if (rc) return rc;
        CU(cudaEventRecord(pr->ev[2], c->stream));
        CU(cudaMemsetAsync(pr->keys.p, 0, sizeof(unsigned long long) * Q, c->stream));
        int gx = (H + 255) / 256;
        if (gx > 4 * c->sm_count) gx = 4 * c->sm_count;
        LAUNCH(c, k_argmax_key, dim3((unsigned)gx, (unsigned)Q), 256, 0, pr->counts.as<int>(), H, (unsigned long long)p->hyp_begin,
               pr->keys.as<unsigned long long>());
        CU(cudaGetLastError());
        return B2R_OK;
    }
    RansacState* st = pr->state.as<RansacState>();
    int* not_done = reinterpret_cast<int*>(st + Q);
    CU(pr->rt.reserve(sizeof(double) * 6 * (size_t)Q * H));
    pr->rt_valid = true;
    LAUNCH(c, k_state_init, (unsigned)((Q + 127) / 128), 128, 0, st, n == PNP_MP ? 1 : (int)p->max_iters, Q);
    for (int begin = 0, len = 256; begin < H; begin += len, len *= 2) {
        if (len > H - begin) len = H - begin;
        CU(cudaMemsetAsync(not_done, 0, sizeof(int), c->stream));
        LAUNCH(c, k_cv_sample_p, (unsigned)Q, 32, 0, n, H, begin, len, pr->samples.as<int>(), st, Q);
        dim3 grid((unsigned)((len + 63) / 64), (unsigned)Q);
        if (p->solver)
            LAUNCH(c, k_epnp_solve_p<true>, grid, 64, 0, pr->px.as<PointPX>(), pr->pts_stride(), n, H, begin, len, (const RansacState*)st,
                   pr->Kq.as<double>(), pr->centre.as<double>(), cstride, 0, 0LL, (uint64_t)0, pr->samples.as<int>(), mx, mf,
                   pr->rt.as<double>(), (uint8_t*)nullptr);
        else
            LAUNCH(c, k_epnp_solve_p<false>, grid, 64, 0, pr->px.as<PointPX>(), pr->pts_stride(), n, H, begin, len, (const RansacState*)st,
                   pr->Kq.as<double>(), pr->centre.as<double>(), cstride, 0, 0LL, (uint64_t)0, pr->samples.as<int>(), mx, mf,
                   pr->rt.as<double>(), (uint8_t*)nullptr);
        CU(cudaGetLastError());
        if (exact) rc = score_p_exact(c, mx, len, pr->px.as<PointPX>(), pr->pts_stride(), n, pr->Kq.as<double>(), thr_sq, pr->counts.as<int>(), Q, H, begin,
                                      filt ? pr->pf.as<PointPF>() : nullptr, pr->centre.as<double>(), cstride);
        else rc = score_p_fast(c, mf, len, pr->pf.as<PointPF>(), pr->pts_stride(), n, thr_sq, pr->counts.as<int>(), Q, H, begin);
        if (rc) return rc;
        LAUNCH(c, k_select_cv_chunk, (unsigned)((Q + 127) / 128), 128, 0, pr->counts.as<int>(), H, begin, len, n, p->confidence, PNP_MP,
               st, pr->sel.as<HSelect>(), Q, not_done);
        CU(cudaGetLastError());
        if (begin + len >= H) break;
        int nd = 0;
        CU(cudaMemcpyAsync(&nd, not_done, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (nd == 0) break;
    }
    CU(cudaEventRecord(pr->ev[1], c->stream));  // replay path: the whole loop is accounted to stage 0
    CU(cudaEventRecord(pr->ev[2], c->stream));
    return B2R_OK;
}

// n == modelPoints: the one possible subset is the model (OpenCV skips RANSAC altogether)
__global__ void k_select_single(HSelect* sel, int Q) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    HSelect s;
    s.best = 0; s.best_count = PNP_MP; s.iters_run = 1; s.pad = 0;
    sel[q] = s;
}

template <typename Kern, typename... Args>
static int launch_cluster(b2r_ctx* c, Kern kern, int n_clusters, int csize, int threads, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_clusters * csize));
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CU(cudaLaunchKernelEx(&cfg, kern, args...));
    c->launches++;
    return B2R_OK;
}

static int p_run_finish(b2r_ctx* c, b2r_p_problem* pr, const b2r_p_params* p, const uint64_t* keys_host,
                        const uint64_t* keys_dev = nullptr) {
    const int Q = pr->Q, n = pr->n, H = pr->H_last;
    const float thr_sq = (float)(p->thr * p->thr);
    int Hs = H;
    if (n == PNP_MP) {
        LAUNCH(c, k_select_single, (unsigned)((Q + 127) / 128), 128, 0, pr->sel.as<HSelect>(), Q);
    } else if (p->sampler == B2R_SAMPLER_PHILOX) {
        if (keys_host || keys_dev) {
            if (keys_host) CU(cudaMemcpyAsync(pr->keys.p, keys_host, sizeof(uint64_t) * Q, cudaMemcpyHostToDevice, c->stream));
            else CU(cudaMemcpyAsync(pr->keys.p, keys_dev, sizeof(uint64_t) * Q, cudaMemcpyDeviceToDevice, c->stream));
            LAUNCH(c, k_resample_winner_p, (unsigned)((Q + 127) / 128), 128, 0, n, pr->keys.as<unsigned long long>(), p->seed, H,
                   pr->samples.as<int>(), pr->sel.as<HSelect>(), H, Q);
        } else {
            LAUNCH(c, k_select_from_keys, (unsigned)((Q + 127) / 128), 128, 0, pr->keys.as<unsigned long long>(),
                   (unsigned long long)p->hyp_begin, H, PNP_MP, pr->sel.as<HSelect>(), Q);
        }
    }   // CV_REPLAY: p_run_score's last k_select_cv_chunk already left the selection in pr->sel
    CU(cudaGetLastError());
    CU(cudaEventRecord(pr->ev[3], c->stream));
    const int csize = n >= 32768 ? 8 : (n >= 8192 ? 2 : 1);
    int rc;
    CU(pr->win.reserve(sizeof(double) * 6 * (size_t)Q));
    const double* rt_kept = pr->rt_valid && !keys_host && !keys_dev ? pr->rt.as<double>() : nullptr;
    if (p->solver)
        LAUNCH(c, k_winner_model_p<true>, (unsigned)((Q + 31) / 32), 32, 0, (const PointPX*)pr->px.as<PointPX>(), pr->pts_stride(),
               (const int*)pr->samples.as<int>(), Hs, (const HSelect*)pr->sel.as<HSelect>(), (const double*)pr->Kq.as<double>(), rt_kept,
               pr->win.as<double>(), Q);
    else
        LAUNCH(c, k_winner_model_p<false>, (unsigned)((Q + 31) / 32), 32, 0, (const PointPX*)pr->px.as<PointPX>(), pr->pts_stride(),
               (const int*)pr->samples.as<int>(), Hs, (const HSelect*)pr->sel.as<HSelect>(), (const double*)pr->Kq.as<double>(), rt_kept,
               pr->win.as<double>(), Q);
    CU(cudaGetLastError());
    if (n >= 1024)
        rc = launch_cluster(c, k_finalize_p<512>, Q, csize, 512, (const PointPX*)pr->px.as<PointPX>(), pr->pts_stride(),
                            (const double*)pr->raw_obj.as<double>(), (const double*)pr->raw_img.as<double>(), pr->pts_stride(), n,
                            (const int*)pr->samples.as<int>(), Hs, (const HSelect*)pr->sel.as<HSelect>(),
                            (const double*)pr->Kq.as<double>(), thr_sq, (int)p->refine, (int)(n == PNP_MP),
                            (const double*)pr->win.as<double>(), pr->rmask.as<uint8_t>(),
                            pr->pose.as<double>(), pr->info_i.as<int>(), pr->info_d.as<double>());
    else
        rc = launch_cluster(c, k_finalize_p<128>, Q, csize, 128, (const PointPX*)pr->px.as<PointPX>(), pr->pts_stride(),
                            (const double*)pr->raw_obj.as<double>(), (const double*)pr->raw_img.as<double>(), pr->pts_stride(), n,
                            (const int*)pr->samples.as<int>(), Hs, (const HSelect*)pr->sel.as<HSelect>(),
                            (const double*)pr->Kq.as<double>(), thr_sq, (int)p->refine, (int)(n == PNP_MP),
                            (const double*)pr->win.as<double>(), pr->rmask.as<uint8_t>(),
                            pr->pose.as<double>(), pr->info_i.as<int>(), pr->info_d.as<double>());
    if (rc) return rc;
    if (n > PNP_MP && (keys_host || keys_dev))
        LAUNCH(c, k_patch_best_iter, (unsigned)((Q + 127) / 128), 128, 0, pr->info_i.as<int>(), pr->keys.as<unsigned long long>(),
               (long long)p->hyp_begin, H, Q);
    LAUNCH(c, k_compact_inliers, (unsigned)Q, 1024, 0, pr->rmask.as<uint8_t>(), n, pr->inliers.as<int>(), pr->ninl.as<int>());
    CU(cudaGetLastError());
    CU(cudaEventRecord(pr->ev[4], c->stream));
    return B2R_OK;
}

static_assert(sizeof(b2r_p_info) == 12 * sizeof(int32_t) + 8 * sizeof(double), "b2r_p_info layout");

static int p_fetch(b2r_ctx* c, b2r_p_problem* pr, double* rvec, double* tvec, int32_t* inliers, int32_t* n_inl, b2r_p_info* info) {
    const int Q = pr->Q, n = pr->n;
    const size_t pb = sizeof(double) * 6 * (size_t)Q, ib = sizeof(int) * 12 * (size_t)Q, db = sizeof(double) * 8 * (size_t)Q,
                 nb = sizeof(int) * (size_t)Q, lb = inliers ? sizeof(int) * (size_t)Q * n : 0;
    CU(c->pin_out.reserve(pb + ib + db + nb + lb + 64));
    char* base = (char*)c->pin_out.p;
    CU(cudaMemcpyAsync(base, pr->pose.p, pb, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(base + pb, pr->info_d.p, db, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(base + pb + db, pr->info_i.p, ib, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(base + pb + db + ib, pr->ninl.p, nb, cudaMemcpyDeviceToHost, c->stream));
    if (inliers) CU(cudaMemcpyAsync(base + pb + db + ib + nb, pr->inliers.p, lb, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const double* pose = (const double*)base;
    const double* idd = (const double*)(base + pb);
    const int* ii = (const int*)(base + pb + db);
    const int* ni = (const int*)(base + pb + db + ib);
    for (int q = 0; q < Q; ++q) {
        if (rvec) memcpy(rvec + 3 * q, pose + 6 * q, sizeof(double) * 3);
        if (tvec) memcpy(tvec + 3 * q, pose + 6 * q + 3, sizeof(double) * 3);
        if (n_inl) n_inl[q] = ii[12 * q] == 0 ? ni[q] : 0;
        if (info) {
            memcpy(&info[q], ii + 12 * q, sizeof(int) * 12);
            memcpy(reinterpret_cast<char*>(&info[q]) + sizeof(int) * 12, idd + 8 * q, sizeof(double) * 8);
        }
    }
    if (inliers) memcpy(inliers, base + pb + db + ib + nb, lb);
    for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&pr->stage_ms[i], pr->ev[i], pr->ev[i + 1]);
    cudaEventElapsedTime(&pr->stage_ms[4], pr->ev[0], pr->ev[4]);
    cudaGetLastError();
    return B2R_OK;
}

static int check_pnp_shape(int Q, int n) {
    if (Q < 1 || Q > 65535) return fail(B2R_ERR_ARG, "Q must be in 1..65535 (one grid dimension per problem): split larger batches%s%s");
    if (n < 4) return fail(B2R_ERR_ARG, "solvePnPRansac needs at least 4 correspondences (cv2 raises cv2.error)%s%s");
    if (n == 4)
        return fail(B2R_ERR_ARG, "n == 4 takes OpenCV's P3P branch, which is not on the reference's path (12 points) and not "
                                 "implemented%s%s");
    return B2R_OK;
}

// fp32 rows of the fast model from (R | t) fp64 models (building block for the K3 parity tests)
__global__ void k_fast_models_from_rt(const double* __restrict__ mx, int H, const double* __restrict__ K4, const double* __restrict__ centre,
                                      float4* __restrict__ mf) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= H) return;
    double R[9], t[3];
    bool ok = true;
    for (int i = 0; i < 9; ++i) { R[i] = mx[(size_t)g * 12 + i]; ok &= R[i] == R[i]; }
    for (int i = 0; i < 3; ++i) { t[i] = mx[(size_t)g * 12 + 9 + i]; ok &= t[i] == t[i]; }
    store_fast_model(mf, g, R, t, K4, centre, ok);
}

extern "C" {

void b2r_default_p_params(b2r_p_params* p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->thr = 8.0;          // cv2 defaults: iterationsCount=100, reprojectionError=8.0, confidence=0.99
    p->max_iters = 100;
    p->confidence = 0.99;
    p->sampler = B2R_SAMPLER_CV_REPLAY;
    p->arith = B2R_ARITH_EXACT;
    p->refine = 1;
}

b2r_p_problem* b2r_p_problem_upload(b2r_ctx* c, const double* obj, const double* img, int32_t pts_shared, int32_t Q, int32_t n,
                                    const double* K) {
    if (!c || !obj || !img || !K || check_pnp_shape(Q, n)) {
        if (!c || !obj || !img || !K) fail(B2R_ERR_ARG, "null argument%s%s");
        return nullptr;
    }
    if (cudaSetDevice(c->device) != cudaSuccess) {
        fail(B2R_ERR_CUDA, "cudaSetDevice failed%s%s");
        return nullptr;
    }
    b2r_p_problem* pr = new b2r_p_problem();
    if (p_upload(c, pr, obj, img, pts_shared, Q, n, K) || cudaStreamSynchronize(c->stream) != cudaSuccess) {
        if (g_err.empty()) fail(B2R_ERR_CUDA, "upload failed%s%s");
        b2r_p_problem_destroy(pr);
        return nullptr;
    }
    return pr;
}

int b2r_p_problem_reupload(b2r_ctx* c, b2r_p_problem* pr, const double* obj, const double* img, int32_t pts_shared, int32_t Q,
                           int32_t n, const double* K) {
    if (!c || !pr || !obj || !img || !K) return fail(B2R_ERR_ARG, "null argument%s%s");
    int rc = check_pnp_shape(Q, n);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    return p_upload(c, pr, obj, img, pts_shared, Q, n, K);
}

void b2r_p_problem_free(b2r_ctx* c, b2r_p_problem* pr) {
    if (!pr) return;
    if (c) cudaSetDevice(c->device);
    b2r_p_problem_destroy(pr);
}

int b2r_p_problem_run(b2r_ctx* c, b2r_p_problem* pr, const b2r_p_params* p) {
    if (!c || !pr) return fail(B2R_ERR_ARG, "null argument%s%s");
    int rc = check_p_params(p);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    if ((rc = p_run_score(c, pr, p))) return rc;
    return p_run_finish(c, pr, p, nullptr);
}

int b2r_p_problem_score_shard(b2r_ctx* c, b2r_p_problem* pr, const b2r_p_params* p, uint64_t* keys_out) {
    if (!c || !pr || !keys_out) return fail(B2R_ERR_ARG, "null argument%s%s");
    int rc = check_p_params(p);
    if (rc) return rc;
    if (p->sampler != B2R_SAMPLER_PHILOX) return fail(B2R_ERR_ARG, "hypothesis sharding needs the PHILOX sampler%s%s");
    if (pr->n <= PNP_MP) return fail(B2R_ERR_ARG, "sharding needs n > 5%s%s");
    CU(cudaSetDevice(c->device));
    if ((rc = p_run_score(c, pr, p))) return rc;
    CU(cudaMemcpyAsync(keys_out, pr->keys.p, sizeof(uint64_t) * pr->Q, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2R_OK;
}

int b2r_p_problem_finish(b2r_ctx* c, b2r_p_problem* pr, const b2r_p_params* p, const uint64_t* keys) {
    if (!c || !pr || !keys) return fail(B2R_ERR_ARG, "null argument%s%s");
    int rc = check_p_params(p);
    if (rc) return rc;
    if (p->sampler != B2R_SAMPLER_PHILOX) return fail(B2R_ERR_ARG, "finishing from reduced keys needs the PHILOX sampler%s%s");
    CU(cudaSetDevice(c->device));
    return p_run_finish(c, pr, p, keys);
}

int b2r_p_problem_score_shard_dev(b2r_ctx* c, b2r_p_problem* pr, const b2r_p_params* p, uint64_t* keys_dev_out) {
    if (!c || !pr || !keys_dev_out) return fail(B2R_ERR_ARG, "null argument%s%s");
    int rc = check_p_params(p);
    if (rc) return rc;
    if (p->sampler != B2R_SAMPLER_PHILOX) return fail(B2R_ERR_ARG, "hypothesis sharding needs the PHILOX sampler%s%s");
    if (pr->n <= PNP_MP) return fail(B2R_ERR_ARG, "sharding needs n > 5%s%s");
    CU(cudaSetDevice(c->device));
    if ((rc = p_run_score(c, pr, p))) return rc;
    CU(cudaMemcpyAsync(keys_dev_out, pr->keys.p, sizeof(uint64_t) * pr->Q, cudaMemcpyDeviceToDevice, c->stream));
    return B2R_OK;
}

int b2r_p_problem_finish_dev(b2r_ctx* c, b2r_p_problem* pr, const b2r_p_params* p, const uint64_t* keys_dev) {
    if (!c || !pr || !keys_dev) return fail(B2R_ERR_ARG, "null argument%s%s");
    int rc = check_p_params(p);
    if (rc) return rc;
    if (p->sampler != B2R_SAMPLER_PHILOX) return fail(B2R_ERR_ARG, "finishing from reduced keys needs the PHILOX sampler%s%s");
    CU(cudaSetDevice(c->device));
    return p_run_finish(c, pr, p, nullptr, keys_dev);
}

int b2r_p_problem_fetch(b2r_ctx* c, b2r_p_problem* pr, double* rvec_out, double* tvec_out, int32_t* inliers_out,
                        int32_t* n_inliers_out, b2r_p_info* info_out) {
    if (!c || !pr) return fail(B2R_ERR_ARG, "null argument%s%s");
    CU(cudaSetDevice(c->device));
    return p_fetch(c, pr, rvec_out, tvec_out, inliers_out, n_inliers_out, info_out);
}

int b2r_p_problem_stage_ms(b2r_ctx* c, b2r_p_problem* pr, float ms_out[5]) {
    if (!c || !pr || !ms_out) return fail(B2R_ERR_ARG, "null argument%s%s");
    memcpy(ms_out, pr->stage_ms, sizeof(float) * 5);
    return B2R_OK;
}

int b2r_solve_pnp_ransac_batch(b2r_ctx* c, const double* obj, const double* img, int32_t pts_shared, int32_t Q, int32_t n,
                               const double* K, const b2r_p_params* p, double* rvec_out, double* tvec_out, int32_t* inliers_out,
                               int32_t* n_inliers_out, b2r_p_info* info_out) {
    if (!c || !obj || !img || !K || !rvec_out || !tvec_out) return fail(B2R_ERR_ARG, "null argument%s%s");
    int rc = check_pnp_shape(Q, n);
    if (rc) return rc;
    if ((rc = check_p_params(p))) return rc;
    CU(cudaSetDevice(c->device));
    if (!c->cached_p) c->cached_p = new b2r_p_problem();
    b2r_p_problem* pr = c->cached_p;
    if ((rc = p_upload(c, pr, obj, img, pts_shared, Q, n, K))) return rc;
    if ((rc = p_run_score(c, pr, p))) return rc;
    if ((rc = p_run_finish(c, pr, p, nullptr))) return rc;
    return p_fetch(c, pr, rvec_out, tvec_out, inliers_out, n_inliers_out, info_out);
}

int b2r_solve_pnp_ransac(b2r_ctx* c, const double* obj, const double* img, int32_t n, const double* K, const b2r_p_params* p,
                         double* rvec_out, double* tvec_out, int32_t* inliers_out, int32_t* n_inliers_out, b2r_p_info* info_out) {
    b2r_p_info info;
    int32_t ni = 0;
    int rc = b2r_solve_pnp_ransac_batch(c, obj, img, 1, 1, n, K, p, rvec_out, tvec_out, inliers_out, &ni, &info);
    if (rc) return rc;
    if (n_inliers_out) *n_inliers_out = ni;
    if (info_out) *info_out = info;
    return info.status;
}

int b2r_solve_pnp_refine_lm(b2r_ctx* c, const double* obj, const double* img, int32_t k, const double* K, double* rvec_io,
                            double* tvec_io, int32_t max_iters, int32_t* iters_out) {
    if (!c || !obj || !img || !K || !rvec_io || !tvec_io || k < 3) return fail(B2R_ERR_ARG, "bad argument (need >= 3 points)%s%s");
    CU(cudaSetDevice(c->device));
    std::vector<double> k4;
    int rc = k4_from_K(K, 1, k4);
    if (rc) return rc;
    const size_t ob = sizeof(double) * 3 * (size_t)k, ib = sizeof(double) * 2 * (size_t)k;
    CU(c->in_a.reserve(ob));
    CU(c->in_b.reserve(ib));
    CU(c->scratch0.reserve(sizeof(double) * 10 + sizeof(int) * 2));
    double small[10];
    memcpy(small, rvec_io, sizeof(double) * 3);
    memcpy(small + 3, tvec_io, sizeof(double) * 3);
    memcpy(small + 6, k4.data(), sizeof(double) * 4);
    CU(cudaMemcpyAsync(c->in_a.p, obj, ob, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->in_b.p, img, ib, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->scratch0.p, small, sizeof(small), cudaMemcpyHostToDevice, c->stream));
    double* pose = c->scratch0.as<double>();
    int* it_d = reinterpret_cast<int*>(pose + 10);
    const int csize = k >= 32768 ? 8 : (k >= 8192 ? 2 : 1);
    if (k >= 1024)
        rc = launch_cluster(c, k_refine_lm_p<512>, 1, csize, 512, (const double*)c->in_a.as<double>(), (const double*)c->in_b.as<double>(),
                            (int)k, (const double*)(pose + 6), (int)(max_iters > 0 ? max_iters : 20), pose, it_d);
    else
        rc = launch_cluster(c, k_refine_lm_p<128>, 1, csize, 128, (const double*)c->in_a.as<double>(), (const double*)c->in_b.as<double>(),
                            (int)k, (const double*)(pose + 6), (int)(max_iters > 0 ? max_iters : 20), pose, it_d);
    if (rc) return rc;
    int it_h = 0;
    CU(cudaMemcpyAsync(small, pose, sizeof(double) * 6, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(&it_h, it_d, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(rvec_io, small, sizeof(double) * 3);
    memcpy(tvec_io, small + 3, sizeof(double) * 3);
    if (iters_out) *iters_out = it_h;
    return B2R_OK;
}

// ---- building blocks ------------------------------------------------------------------------------------------------
// pack n points into scratch: returns device pointers through a temporary problem object owned by the context
static int bb_points(b2r_ctx* c, const double* obj, const double* img, int n, const double* K, b2r_p_problem** out) {
    if (!c->cached_p) c->cached_p = new b2r_p_problem();
    int rc = p_upload(c, c->cached_p, obj, img, 1, 1, n, K);
    *out = c->cached_p;
    return rc;
}

int b2r_score_p(b2r_ctx* c, const double* models_Rt, int32_t n_models, const double* obj, const double* img, int32_t n,
                const double* K, float thr_sq, int32_t arith, int32_t* counts_out) {
    if (!c || !models_Rt || !obj || !img || !K || !counts_out || n_models < 1 || n < 1) return fail(B2R_ERR_ARG, "bad argument%s%s");
    CU(cudaSetDevice(c->device));
    b2r_p_problem* pr;
    int rc = bb_points(c, obj, img, n, K, &pr);
    if (rc) return rc;
    CU(c->scratch1.reserve(sizeof(double) * 12 * (size_t)n_models));
    CU(c->scratch2.reserve(sizeof(int) * (size_t)n_models));
    CU(cudaMemcpyAsync(c->scratch1.p, models_Rt, sizeof(double) * 12 * (size_t)n_models, cudaMemcpyHostToDevice, c->stream));
    if (arith != B2R_ARITH_FAST) {
        rc = score_p_exact(c, c->scratch1.as<double>(), n_models, pr->px.as<PointPX>(), 0, n, pr->Kq.as<double>(), thr_sq,
                           c->scratch2.as<int>(), 1, 0, 0, arith == B2R_ARITH_EXACT ? pr->pf.as<PointPF>() : nullptr, pr->centre.as<double>(), 0);
    } else {
        CU(c->scratch3.reserve(sizeof(float) * 12 * (size_t)n_models));
        LAUNCH(c, k_fast_models_from_rt, (unsigned)((n_models + 127) / 128), 128, 0, c->scratch1.as<double>(), n_models,
               pr->Kq.as<double>(), pr->centre.as<double>(), c->scratch3.as<float4>());
        CU(cudaGetLastError());
        rc = score_p_fast(c, c->scratch3.as<float4>(), n_models, pr->pf.as<PointPF>(), 0, n, thr_sq, c->scratch2.as<int>(), 1);
    }
    if (rc) return rc;
    CU(cudaMemcpyAsync(counts_out, c->scratch2.p, sizeof(int) * (size_t)n_models, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2R_OK;
}

int b2r_pnp_minimal_models(b2r_ctx* c, const double* obj, const double* img, int32_t n, const double* K, const int32_t* idx,
                           int32_t n_samples, int32_t solver, double* rvec_out, double* tvec_out, double* R_out, uint8_t* ok_out) {
    if (!c || !obj || !img || !K || !idx || n < 5 || n_samples < 1) return fail(B2R_ERR_ARG, "bad argument%s%s");
    for (size_t i = 0; i < (size_t)n_samples * PNP_MP; ++i)
        if (idx[i] < 0 || idx[i] >= n) return fail(B2R_ERR_ARG, "sample index out of range%s%s");
    CU(cudaSetDevice(c->device));
    b2r_p_problem* pr;
    int rc = bb_points(c, obj, img, n, K, &pr);
    if (rc) return rc;
    CU(c->scratch1.reserve(sizeof(int) * PNP_MP * (size_t)n_samples));
    CU(c->scratch2.reserve(sizeof(double) * 18 * (size_t)n_samples));
    CU(c->scratch3.reserve((size_t)n_samples));
    CU(cudaMemcpyAsync(c->scratch1.p, idx, sizeof(int) * PNP_MP * (size_t)n_samples, cudaMemcpyHostToDevice, c->stream));
    double* mx = c->scratch2.as<double>();
    double* rt = mx + 12 * (size_t)n_samples;
    if (solver == B2R_SOLVER_FAST)
        LAUNCH(c, k_epnp_solve_p<true>, dim3((unsigned)((n_samples + 63) / 64), 1), 64, 0, pr->px.as<PointPX>(), (size_t)0, n, n_samples, 0,
               n_samples, (const RansacState*)nullptr, pr->Kq.as<double>(), pr->centre.as<double>(), (size_t)0, 0, 0LL, (uint64_t)0,
               c->scratch1.as<int>(), mx, (float4*)nullptr, rt, c->scratch3.as<uint8_t>());
    else
        LAUNCH(c, k_epnp_solve_p<false>, dim3((unsigned)((n_samples + 63) / 64), 1), 64, 0, pr->px.as<PointPX>(), (size_t)0, n, n_samples, 0,
               n_samples, (const RansacState*)nullptr, pr->Kq.as<double>(), pr->centre.as<double>(), (size_t)0, 0, 0LL, (uint64_t)0,
               c->scratch1.as<int>(), mx, (float4*)nullptr, rt, c->scratch3.as<uint8_t>());
    CU(cudaGetLastError());
    std::vector<double> h(18 * (size_t)n_samples);
    CU(cudaMemcpyAsync(h.data(), mx, sizeof(double) * 18 * (size_t)n_samples, cudaMemcpyDeviceToHost, c->stream));
    if (ok_out) CU(cudaMemcpyAsync(ok_out, c->scratch3.p, (size_t)n_samples, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int s = 0; s < n_samples; ++s) {
        if (R_out) memcpy(R_out + 9 * (size_t)s, h.data() + 12 * (size_t)s, sizeof(double) * 9);
        if (rvec_out) memcpy(rvec_out + 3 * (size_t)s, h.data() + 12 * (size_t)n_samples + 6 * (size_t)s, sizeof(double) * 3);
        if (tvec_out) memcpy(tvec_out + 3 * (size_t)s, h.data() + 12 * (size_t)n_samples + 6 * (size_t)s + 3, sizeof(double) * 3);
    }
    return B2R_OK;
}

int b2r_sample_cv_p(b2r_ctx* c, int32_t n, int32_t n_iters, int32_t* idx_out) {
    if (!c || !idx_out || n < 5 || n_iters < 1) return fail(B2R_ERR_ARG, "bad argument (need n >= 5)%s%s");
    CU(cudaSetDevice(c->device));
    CU(c->scratch1.reserve(sizeof(int) * PNP_MP * (size_t)n_iters));
    CU(c->scratch2.reserve(sizeof(RansacState)));
    LAUNCH(c, k_state_init, 1, 32, 0, c->scratch2.as<RansacState>(), (int)n_iters, 1);
    LAUNCH(c, k_cv_sample_p, 1, 32, 0, (int)n, (int)n_iters, 0, (int)n_iters, c->scratch1.as<int>(), c->scratch2.as<RansacState>(), 1);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(idx_out, c->scratch1.p, sizeof(int) * PNP_MP * (size_t)n_iters, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2R_OK;
}

}  // extern "C"

// Host-side plumbing shared by the translation units of libransac_b200.so (api.cu: homography path,
// api_pnp.cu: PnP path): error reporting, growable device / pinned buffers, the context.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

// the library is built with -fvisibility=hidden; exactly the symbols of the public header are exported
#pragma GCC visibility push(default)
#include "../../include/ransac_b200.h"
#pragma GCC visibility pop

extern thread_local std::string g_err;
int fail(int code, const char* fmt, const char* a = "", const char* b = "");

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) return fail(B2R_ERR_CUDA, "CUDA error: %s  [%s]", cudaGetErrorString(e_), #call); \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct b2r_h_problem;
struct b2r_p_problem;
void b2r_p_problem_destroy(b2r_p_problem* pr);  // api_pnp.cu

struct b2r_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    DevBuf in_a, in_b, scratch0, scratch1, scratch2, scratch3;  // staging for the host-pointer entry points
    DevBuf gscratch;  // partial sums of the grid-wide finalize reductions (2 x CTAs x RED_MAX doubles)
    PinnedBuf pin_in, pin_out;
    b2r_h_problem* cached = nullptr;  // reusable problem storage of b2r_find_homography[_batch]
    b2r_p_problem* cached_p = nullptr;  // same for b2r_solve_pnp_ransac[_batch]
    int launches = 0;
};

#define LAUNCH(ctx, kernel, grid, block, smem, ...)                \
    do {                                                           \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); \
        (ctx)->launches++;                                         \
    } while (0)


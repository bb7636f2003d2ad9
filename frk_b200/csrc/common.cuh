// Shared internals of libfrkb200: context, error reporting, launch accounting, scans.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/frkb200.h"

struct frk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    int64_t launches = 0;
    // growable scratch (device) + small pinned host staging
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    double *h_pinned = nullptr;   // 4096 doubles
    double *d_small = nullptr;    // 4096 doubles
};

void frk_set_error(const char *fmt, ...);

#define FRK_CHECK_CUDA(expr)                                                            \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            frk_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

#define FRK_REQUIRE(cond, ...)                                                          \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            frk_set_error(__VA_ARGS__);                                                 \
            return 2;                                                                   \
        }                                                                               \
    } while (0)

// every kernel launch goes through this so that frk_launch_count() is exact
#define FRK_LAUNCH(ctx, kernel, grid, block, smem, ...)                                 \
    do {                                                                                \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                \
        (ctx)->launches++;                                                              \
        FRK_CHECK_CUDA(cudaGetLastError());                                             \
    } while (0)

int frk_scratch(frk_ctx *ctx, size_t bytes, void **out);   // grow-only scratch buffer

// exclusive scan of int32 counts[n] -> out[n+1] (out[n] = total); total also returned to host
int frk_exclusive_scan_i32(frk_ctx *ctx, const int *d_in, int64_t n, int *d_out, int64_t *h_total);

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ---------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

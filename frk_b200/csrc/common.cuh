// Shared internals of libfrkb200: context, error reporting, launch accounting, scans.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <array>
#include <map>
#include <string>
#include <vector>
#include "../../include/frkb200.h"

// per-kernel timing (bench.py roofline): CUDA events around every launch while profiling is on
struct frk_prof_rec { const char *name; cudaEvent_t e0, e1; double flops, bytes; };
struct frk_prof_stat { int64_t launches = 0; double ms = 0, flops = 0, bytes = 0; };

struct frk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;             // lane contexts (frk_ctx_create_lane) own their main stream
    cudaStream_t stream2 = nullptr;      // side stream: the bulk of potrf's inner update runs beside the next diagonal factorisation
    cudaStream_t stream3 = nullptr;      // second side stream: the rest of the trailing SYRK overlaps the whole next panel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_panel = nullptr, ev_next = nullptr, ev_rest[2] = {nullptr, nullptr};
    int sm_count = 148;
    int quad_ws_min = 8;                 // tiles per column from which the quadratic form takes the warp-specialised kernel (bucket.cu); a context
                                         // whose work runs BESIDE latency-bound chains sets 24: that kernel takes every SM's shared memory
    int64_t launches = 0;
    // growable scratch (device) + small pinned host staging
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    double *h_pinned = nullptr;   // 4096 doubles
    double *d_small = nullptr;    // 4096 doubles
    double *d_winv = nullptr;     // 8 x 16 x 16: inverses of the 16 x 16 diagonal blocks of the current panel's factor (potrf: diag128 -> panel solve)
    // profiling
    bool profiling = false;
    double next_flops = 0, next_bytes = 0;          // algorithmic work of the next launch (FRK_WORK)
    const char *next_tag = nullptr;                 // profile name of the next launch (default: the kernel's name)
    std::vector<frk_prof_rec> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    std::map<std::string, frk_prof_stat> prof_stats;
    // CUDA-graph cache for the launch-latency-bound factorisations: key = (op, r, pointers)
    bool use_graphs = true;
    bool deterministic = false;          // frk_use_deterministic: fixed-order reductions (the Gram flush) instead of fp64 atomics
    struct GraphEntry { cudaGraphExec_t exec; int64_t nodes; };
    std::map<std::array<uintptr_t, 5>, GraphEntry> graphs;
    // lane contexts (frk_ctx_lane): created on first use, kept for the life of this context so that models come and go cheaply
    std::vector<frk_ctx *> lanes_hi, lanes_lo;
    // in-library NCCL communicator (comm.cu): row-partitioned runs without a host-language reduction hook
    void *comm = nullptr;                // ncclComm_t
    int comm_rank = 0, comm_world = 1;
    double *comm_h = nullptr, *comm_d = nullptr;   // pinned / device staging for the small host-side reductions
    cudaStream_t comm_stream = nullptr;  // ... which run on their own stream: they carry host data and must not queue behind a factorisation
};
static constexpr int FRK_COMM_STAGE = 4096;
int frk_pack_upper(frk_ctx *ctx, int r, const double *d_G, double *d_P);      // column j, rows 0..j -> P[j(j+1)/2 + i]
int frk_unpack_upper(frk_ctx *ctx, int r, const double *d_P, double *d_G);

int frk_prof_begin(frk_ctx *ctx, const char *name, cudaEvent_t *e1);   // records e0; returns 0
int frk_prof_harvest(frk_ctx *ctx);

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
        cudaEvent_t _e1 = nullptr;                                                      \
        if ((ctx)->profiling) { if (frk_prof_begin((ctx), (ctx)->next_tag ? (ctx)->next_tag : #kernel, &_e1)) return 1; }   \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                \
        if (_e1) cudaEventRecord(_e1, (ctx)->stream);                                   \
        (ctx)->launches++;                                                              \
        (ctx)->next_flops = (ctx)->next_bytes = 0; (ctx)->next_tag = nullptr;               \
        FRK_CHECK_CUDA(cudaGetLastError());                                             \
    } while (0)

// declare the ALGORITHMIC work (flops, bytes) of the next FRK_LAUNCH ("roofline accounting")
#define FRK_TAG(ctx, tag) do { (ctx)->next_tag = (tag); } while (0)
#define FRK_WORK(ctx, fl, by) do { (ctx)->next_flops = (double)(fl); (ctx)->next_bytes = (double)(by); } while (0)

int frk_scratch(frk_ctx *ctx, size_t bytes, void **out);   // grow-only scratch buffer

// a second context on the parent's device with its own streams, events, staging and graph cache: independent work (e.g. the
// speculative Nelder-Mead evaluations of one M-step) is enqueued on lanes and runs concurrently with the parent's stream
int frk_ctx_create_lane(const frk_ctx *parent, bool high_priority, frk_ctx **out);
// enqueue-only halves of frk_potrf / frk_dot (dense.cu): results are staged in the context's pinned buffer and read after a stream
// synchronisation, so that one host thread can keep several contexts busy at once
int frk_potrf_async(frk_ctx *ctx, int r, double *d_A);
int frk_block_scatter_scaled(frk_ctx *ctx, int r, double *d_A, int ni, const int *d_idx, const double *d_in, double alpha);   // dense.cu
int frk_potrf_finish(frk_ctx *ctx, double *h_logdet);
int frk_potrf_status(frk_ctx *ctx, double *h_logdet);   // non-failing: order of the first bad leading minor, 0 if positive definite
int frk_dot_async(frk_ctx *ctx, int64_t n, const double *d_A, const double *d_B);
static constexpr int FRK_PINNED_DOT = 4080;          // slot of frk_dot_async's result in frk_ctx::h_pinned
// lane `idx` of `parent` (cached in the parent, destroyed with it)
int frk_ctx_lane(frk_ctx *parent, int idx, bool high_priority, frk_ctx **out);

// exclusive scan of int32 counts[n] -> out[n+1] (out[n] = total); total also returned to host
int frk_exclusive_scan_i32(frk_ctx *ctx, const int *d_in, int64_t n, int *d_out, int64_t *h_total);

// basis.cu: rows of eval_basis as a dense row-major block, evaluated on demand (implicit matrices of bucket.cu)
int frk_eval_basis_dense_rows(frk_ctx *ctx, const frk_basis *b, int64_t n, const double *d_s, int64_t ld_s, int normalise, double *d_out);

// dense.cu
int frk_dsyrk_like_lower(frk_ctx *ctx, int n, int K, const double *d_A, int lda, const double *d_B, int ldb, double *d_C, int ldc);
int frk_mirror_lower(frk_ctx *ctx, int r, double *d_A);      // A[j,i] = A[i,j] for i > j
int frk_dtrmm_lower(frk_ctx *ctx, int n, int N, const double *d_X, int ldx, const double *d_B, int ldb, double *d_C, int ldc);
// bucket.cu
int frk_rowbuckets_is_dense(const frk_rowbuckets *bk);
int frk_quadform_dense_factor(frk_ctx *ctx, int64_t n, int r, const double *d_val, const frk_rowbuckets *bk, const double *d_Linv,
                              const double *d_mu, double *d_q, double *d_t);

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ---------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

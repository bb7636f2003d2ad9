// Context, error string, memory helpers and the int32 exclusive scan used by the CSR builders.
#include "common.cuh"
#include <cstdarg>

static thread_local char g_err[1024] = "";

void frk_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *frk_last_error(void) { return g_err; }
extern "C" int frk_abi_version(void) { return FRKB200_ABI_VERSION; }

extern "C" int frk_ctx_create(int device, void *stream, frk_ctx **out) {
    FRK_REQUIRE(out != nullptr, "frk_ctx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        frk_set_error("frk_ctx_create: no CUDA device available (%s); libfrkb200 has no CPU fallback",
                      cudaGetErrorString(e));
        return 1;
    }
    FRK_REQUIRE(device >= 0 && device < ndev, "frk_ctx_create: device %d out of range (0..%d)", device, ndev - 1);
    FRK_CHECK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    FRK_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    FRK_REQUIRE(prop.major >= 10, "frk_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);
    frk_ctx *c = new frk_ctx();
    c->device = device;
    c->stream = (cudaStream_t)stream;
    c->sm_count = prop.multiProcessorCount;
    {   // keep stream-ordered allocations cached in the pool across synchronisations (default: trimmed at every sync)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    int prio_lo = 0, prio_hi = 0;                                      // critical-path branch high, bulk trailing update low
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    FRK_CHECK_CUDA(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_hi));
    FRK_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    FRK_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    FRK_CHECK_CUDA(cudaStreamCreateWithPriority(&c->stream3, cudaStreamNonBlocking, prio_lo));
    FRK_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_panel, cudaEventDisableTiming));
    FRK_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_next, cudaEventDisableTiming));
    FRK_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_rest[0], cudaEventDisableTiming));
    FRK_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_rest[1], cudaEventDisableTiming));
    FRK_CHECK_CUDA(cudaMallocHost((void **)&c->h_pinned, 4096 * sizeof(double)));
    FRK_CHECK_CUDA(cudaMalloc((void **)&c->d_small, 4096 * sizeof(double)));
    FRK_CHECK_CUDA(cudaMalloc((void **)&c->d_winv, 2 * 64 * 64 * sizeof(double)));
    *out = c;
    return 0;
}

int frk_ctx_create_lane(const frk_ctx *parent, bool high_priority, frk_ctx **out) {
    FRK_REQUIRE(parent && out, "frk_ctx_create_lane: NULL argument");
    FRK_CHECK_CUDA(cudaSetDevice(parent->device));
    // a lane that factorises carries a critical path (potf2 / trsm chain) on its main stream: highest priority, like stream2;
    // a lane for bulk work beside a latency-bound phase gets the lowest
    cudaStream_t s = nullptr;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    FRK_CHECK_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high_priority ? prio_hi : prio_lo));
    frk_ctx *c = nullptr;
    if (int rc = frk_ctx_create(parent->device, (void *)s, &c)) { cudaStreamDestroy(s); return rc; }
    c->own_stream = true;
    c->use_graphs = parent->use_graphs;
    *out = c;
    return 0;
}

int frk_ctx_lane(frk_ctx *parent, int idx, bool high_priority, frk_ctx **out) {
    FRK_REQUIRE(parent && out && idx >= 0 && idx < 1024, "frk_ctx_lane: bad argument");
    std::vector<frk_ctx *> &v = high_priority ? parent->lanes_hi : parent->lanes_lo;
    if ((int)v.size() <= idx) v.resize(idx + 1, nullptr);
    if (!v[idx]) { if (int rc = frk_ctx_create_lane(parent, high_priority, &v[idx])) return rc; }
    v[idx]->use_graphs = parent->use_graphs;
    v[idx]->deterministic = parent->deterministic;
    *out = v[idx];
    return 0;
}

extern "C" int frk_ctx_destroy(frk_ctx *ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    frk_comm_destroy(ctx);
    for (frk_ctx *l : ctx->lanes_hi) frk_ctx_destroy(l);
    for (frk_ctx *l : ctx->lanes_lo) frk_ctx_destroy(l);
    frk_prof_harvest(ctx);
    for (auto &kv : ctx->graphs) cudaGraphExecDestroy(kv.second.exec);
    for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->stream2) { cudaStreamSynchronize(ctx->stream2); cudaStreamDestroy(ctx->stream2); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->stream3) { cudaStreamSynchronize(ctx->stream3); cudaStreamDestroy(ctx->stream3); }
    if (ctx->ev_panel) cudaEventDestroy(ctx->ev_panel);
    if (ctx->ev_next) cudaEventDestroy(ctx->ev_next);
    if (ctx->ev_rest[0]) cudaEventDestroy(ctx->ev_rest[0]);
    if (ctx->ev_rest[1]) cudaEventDestroy(ctx->ev_rest[1]);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->d_small) cudaFree(ctx->d_small);
    if (ctx->d_winv) cudaFree(ctx->d_winv);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

extern "C" int frk_sync(frk_ctx *ctx) {
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int frk_malloc(frk_ctx *ctx, size_t bytes, void **d_out) {
    FRK_CHECK_CUDA(cudaSetDevice(ctx->device));
    FRK_CHECK_CUDA(cudaMalloc(d_out, bytes ? bytes : 8));
    return 0;
}

extern "C" int frk_free(frk_ctx *ctx, void *d_ptr) {
    FRK_CHECK_CUDA(cudaSetDevice(ctx->device));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    FRK_CHECK_CUDA(cudaFree(d_ptr));
    return 0;
}

extern "C" int frk_h2d(frk_ctx *ctx, void *d_dst, const void *h_src, size_t bytes) {
    FRK_CHECK_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

extern "C" int frk_d2h(frk_ctx *ctx, void *h_dst, const void *d_src, size_t bytes) {
    FRK_CHECK_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int frk_d2d(frk_ctx *ctx, void *d_dst, const void *d_src, size_t bytes) {
    FRK_CHECK_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

extern "C" int frk_memset0(frk_ctx *ctx, void *d_dst, size_t bytes) {
    FRK_CHECK_CUDA(cudaMemsetAsync(d_dst, 0, bytes, ctx->stream));
    return 0;
}

// ---------------------------------------------------------------------------------
// per-kernel profiling
// ---------------------------------------------------------------------------------
int frk_prof_begin(frk_ctx *ctx, const char *name, cudaEvent_t *e1) {
    if (ctx->prof_pending.size() >= 8192) { if (int rc = frk_prof_harvest(ctx)) return rc; }
    cudaEvent_t ev[2];
    for (int k = 0; k < 2; k++) {
        if (!ctx->prof_pool.empty()) { ev[k] = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
        else FRK_CHECK_CUDA(cudaEventCreate(&ev[k]));
    }
    FRK_CHECK_CUDA(cudaEventRecord(ev[0], ctx->stream));
    ctx->prof_pending.push_back({name, ev[0], ev[1], ctx->next_flops, ctx->next_bytes});
    *e1 = ev[1];
    return 0;
}

int frk_prof_harvest(frk_ctx *ctx) {
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->prof_pending) {
        float ms = 0.f;
        FRK_CHECK_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
        frk_prof_stat &st = ctx->prof_stats[r.name];
        st.launches++; st.ms += ms; st.flops += r.flops; st.bytes += r.bytes;
        ctx->prof_pool.push_back(r.e0);
        ctx->prof_pool.push_back(r.e1);
    }
    ctx->prof_pending.clear();
    return 0;
}

extern "C" int frk_profile(frk_ctx *ctx, int enable) {
    FRK_REQUIRE(ctx, "frk_profile: NULL ctx");
    if (int rc = frk_prof_harvest(ctx)) return rc;
    if (enable && !ctx->profiling) ctx->prof_stats.clear();
    ctx->profiling = enable != 0;
    return 0;
}

extern "C" int frk_profile_report(frk_ctx *ctx, char *buf, size_t cap) {
    FRK_REQUIRE(ctx && buf && cap > 2, "frk_profile_report: bad argument");
    if (int rc = frk_prof_harvest(ctx)) return rc;
    std::string out = "[";
    bool first = true;
    for (auto &kv : ctx->prof_stats) {
        char line[768];
        std::string nm = kv.first;
        for (auto &ch : nm) if (ch == '"' || ch == '\\') ch = ' ';
        snprintf(line, sizeof(line), "%s{\"kernel\": \"%s\", \"launches\": %lld, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}",
                 first ? "" : ", ", nm.c_str(), (long long)kv.second.launches, kv.second.ms, kv.second.flops, kv.second.bytes);
        out += line;
        first = false;
    }
    out += "]";
    FRK_REQUIRE(out.size() + 1 <= cap, "frk_profile_report: buffer too small (%zu needed)", out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}

extern "C" int frk_use_graphs(frk_ctx *ctx, int enable) {
    FRK_REQUIRE(ctx, "frk_use_graphs: NULL ctx");
    ctx->use_graphs = enable != 0;
    for (frk_ctx *l : ctx->lanes_hi) if (l) l->use_graphs = ctx->use_graphs;
    for (frk_ctx *l : ctx->lanes_lo) if (l) l->use_graphs = ctx->use_graphs;
    return 0;
}

extern "C" int frk_use_deterministic(frk_ctx *ctx, int enable) {
    FRK_REQUIRE(ctx, "frk_use_deterministic: NULL ctx");
    ctx->deterministic = enable != 0;
    for (frk_ctx *l : ctx->lanes_hi) if (l) l->deterministic = ctx->deterministic;
    for (frk_ctx *l : ctx->lanes_lo) if (l) l->deterministic = ctx->deterministic;
    return 0;
}

extern "C" int64_t frk_launch_count(frk_ctx *ctx) { return ctx->launches; }

int frk_scratch(frk_ctx *ctx, size_t bytes, void **out) {
    if (bytes > ctx->scratch_bytes) {
        FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch) FRK_CHECK_CUDA(cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        size_t want = bytes + bytes / 4 + 4096;
        FRK_CHECK_CUDA(cudaMalloc(&ctx->scratch, want));
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return 0;
}

// ---------------------------------------------------------------------------------
// exclusive scan: three passes (block sums -> scan of block sums -> rescan + offset).
// 1024 threads x 4 items per block.  Block sums are int64 so that an overflowing total
// (nnz >= 2^31, which R's dgCMatrix cannot hold either) is detected, not wrapped.
// ---------------------------------------------------------------------------------
constexpr int SCAN_T = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_T * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int *smem_warp /*32*/, int &block_total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) smem_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = smem_warp[lane];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        smem_warp[lane] = wi - w;            // exclusive warp offsets
        if (lane == 31) smem_warp[32] = wi;  // block total
    }
    __syncthreads();
    block_total = smem_warp[32];
    int res = incl - v + smem_warp[wid];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_T) scan_block_sums(const int *__restrict__ in, int64_t n, int64_t *__restrict__ bsum) {
    __shared__ int sw[33];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) s += in[base + k];
    int tot;
    block_exclusive_scan(s, sw, tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) scan_of_block_sums(int64_t *bsum, int64_t nb, int64_t *total) {
    // single block; sequential over chunks of 1024
    __shared__ int64_t sh[1024];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < nb; c0 += 1024) {
        int64_t i = c0 + threadIdx.x;
        int64_t v = i < nb ? bsum[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int64_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        int64_t incl = sh[threadIdx.x];
        if (i < nb) bsum[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SCAN_T) scan_apply(const int *__restrict__ in, int64_t n, const int64_t *__restrict__ boff,
                                                     const int64_t *__restrict__ total, int *__restrict__ out) {
    __shared__ int sw[33];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int tot;
    int ex = block_exclusive_scan(s, sw, tot);
    int64_t off = boff[blockIdx.x] + ex;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = (int)off;
        off += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = (int)(*total);
}

int frk_exclusive_scan_i32(frk_ctx *ctx, const int *d_in, int64_t n, int *d_out, int64_t *h_total) {
    int64_t nb = cdiv(n > 0 ? n : 1, SCAN_TILE);
    // scratch use is local to this call: block sums + total live in d_small if they fit, else scratch tail
    int64_t *bsum = nullptr;
    void *tmp = nullptr;
    if ((nb + 1) * sizeof(int64_t) <= 4096 * sizeof(double)) {
        bsum = (int64_t *)ctx->d_small;
    } else {
        FRK_CHECK_CUDA(cudaMallocAsync(&tmp, (nb + 1) * sizeof(int64_t), ctx->stream));
        bsum = (int64_t *)tmp;
    }
    int64_t *total = bsum + nb;
    FRK_LAUNCH(ctx, scan_block_sums, (unsigned)nb, SCAN_T, 0, d_in, n, bsum);
    FRK_LAUNCH(ctx, scan_of_block_sums, 1, 1024, 0, bsum, nb, total);
    FRK_LAUNCH(ctx, scan_apply, (unsigned)nb, SCAN_T, 0, d_in, n, bsum, total, d_out);
    int64_t tot = 0;
    FRK_CHECK_CUDA(cudaMemcpyAsync(&tot, total, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (tmp) FRK_CHECK_CUDA(cudaFreeAsync(tmp, ctx->stream));
    FRK_REQUIRE(tot < (int64_t)2147483647, "scan total %lld exceeds the int32 CSR limit (2^31-1); shard the rows",
                (long long)tot);
    if (h_total) *h_total = tot;
    return 0;
}

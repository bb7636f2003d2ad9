// In-library NCCL communicator for row-partitioned (multi-GPU) runs.
//
// the context holds the communicator, so that an R process (one per GPU, no Python, no torch) reaches
// the multi-GPU path through the same C ABI: `frk_comm_unique_id` on rank 0, the 128 bytes handed to the other ranks by
// whatever the host has (a file, MPI, a socket), `frk_comm_init` on every rank.  Afterwards every frk_model created on the
// context with a NULL reduction hook reduces over this communicator:
//   * ONE fp64 sum all-reduce per EM iteration of [packed upper triangle of S'D^-1 S | S'D^-1 rho | rho'D^-1 rho | sum log d]
//     (r(r+1)/2 + r + 2 doubles; R/SREfit.R:194,252,254 -- what crossprod() sums over rows),
//   * O(p^2)-double all-reduces of the M-step moments (R/SREfit.R:335-356),
//   * a real ncclBroadcast of the winning K_i^-1 when the Nelder-Mead candidates are distributed over the ranks.
// libnccl is opened with dlopen at first use (no link-time dependency: the library loads in processes without NCCL, and
// inside a torch process the already-loaded libnccl.so.2 is reused).  FRKB200_NCCL_LIB overrides the library name.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.handle) return 0;
    const char *names[] = {getenv("FRKB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (h) break;
    }
    FRK_REQUIRE(h, "frk_comm: cannot open libnccl.so.2 (%s); set FRKB200_NCCL_LIB", dlerror());
#define FRK_SYM(field, name)                                                                     \
    do {                                                                                         \
        *(void **)(&g_nccl.field) = dlsym(h, name);                                              \
        if (!g_nccl.field) { frk_set_error("frk_comm: libnccl has no symbol %s", name); dlclose(h); return 1; } \
    } while (0)
    FRK_SYM(GetUniqueId, "ncclGetUniqueId");
    FRK_SYM(CommInitRank, "ncclCommInitRank");
    FRK_SYM(CommDestroy, "ncclCommDestroy");
    FRK_SYM(AllReduce, "ncclAllReduce");
    FRK_SYM(Broadcast, "ncclBroadcast");
    FRK_SYM(GetErrorString, "ncclGetErrorString");
#undef FRK_SYM
    g_nccl.handle = h;
    return 0;
}

#define FRK_CHECK_NCCL(expr)                                                                     \
    do {                                                                                         \
        ncclResult_t _r = (expr);                                                                \
        if (_r != ncclSuccess) {                                                                 \
            frk_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

}  // namespace

extern "C" int frk_comm_unique_id(void *out128) {
    FRK_REQUIRE(out128, "frk_comm_unique_id: NULL buffer");
    static_assert(sizeof(ncclUniqueId) == FRK_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    if (int rc = nccl_load()) return rc;
    ncclUniqueId id;
    FRK_CHECK_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return 0;
}

extern "C" int frk_comm_init(frk_ctx *ctx, int rank, int world, const void *id128) {
    FRK_REQUIRE(ctx && id128 && world >= 1 && rank >= 0 && rank < world, "frk_comm_init: bad argument");
    FRK_REQUIRE(!ctx->comm, "frk_comm_init: this context already has a communicator");
    if (int rc = nccl_load()) return rc;
    FRK_CHECK_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    FRK_CHECK_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
    ctx->comm = (void *)comm;
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    if (!ctx->comm_h) FRK_CHECK_CUDA(cudaMallocHost((void **)&ctx->comm_h, FRK_COMM_STAGE * sizeof(double)));
    if (!ctx->comm_d) FRK_CHECK_CUDA(cudaMalloc((void **)&ctx->comm_d, FRK_COMM_STAGE * sizeof(double)));
    if (!ctx->comm_stream) FRK_CHECK_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    return 0;
}

extern "C" int frk_comm_destroy(frk_ctx *ctx) {
    if (!ctx || !ctx->comm) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
    g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
    ctx->comm_rank = 0;
    ctx->comm_world = 1;
    if (ctx->comm_h) { cudaFreeHost(ctx->comm_h); ctx->comm_h = nullptr; }
    if (ctx->comm_d) { cudaFree(ctx->comm_d); ctx->comm_d = nullptr; }
    if (ctx->comm_stream) { cudaStreamDestroy(ctx->comm_stream); ctx->comm_stream = nullptr; }
    return 0;
}

extern "C" int frk_comm_rank(const frk_ctx *ctx) { return ctx ? ctx->comm_rank : 0; }
extern "C" int frk_comm_world(const frk_ctx *ctx) { return ctx ? ctx->comm_world : 1; }

// in place, on the context's stream; op: 0 = sum, 1 = max
extern "C" int frk_comm_allreduce(frk_ctx *ctx, double *d_buf, int64_t n, int op) {
    FRK_REQUIRE(ctx && d_buf && n >= 0, "frk_comm_allreduce: bad argument");
    if (!ctx->comm || n == 0) return 0;
    FRK_CHECK_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)n, ncclFloat64, op == 1 ? ncclMax : ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    return 0;
}

extern "C" int frk_comm_broadcast(frk_ctx *ctx, double *d_buf, int64_t n, int root) {
    FRK_REQUIRE(ctx && d_buf && n >= 0 && root >= 0 && root < ctx->comm_world, "frk_comm_broadcast: bad argument");
    if (!ctx->comm || n == 0) return 0;
    FRK_CHECK_NCCL(g_nccl.Broadcast(d_buf, d_buf, (size_t)n, ncclFloat64, root, (ncclComm_t)ctx->comm, ctx->stream));
    return 0;
}

// a few host doubles (M-step moments, Nelder-Mead values): staged through the context's pinned / device staging pair
extern "C" int frk_comm_allreduce_host(frk_ctx *ctx, double *h_buf, int64_t n, int op) {
    FRK_REQUIRE(ctx && h_buf && n >= 0, "frk_comm_allreduce_host: bad argument");
    if (!ctx->comm || n == 0) return 0;
    for (int64_t o = 0; o < n; o += FRK_COMM_STAGE) {
        const int64_t k = std::min<int64_t>(FRK_COMM_STAGE, n - o);
        cudaStream_t st = ctx->comm_stream;
        memcpy(ctx->comm_h, h_buf + o, (size_t)k * sizeof(double));
        FRK_CHECK_CUDA(cudaMemcpyAsync(ctx->comm_d, ctx->comm_h, (size_t)k * sizeof(double), cudaMemcpyHostToDevice, st));
        FRK_CHECK_NCCL(g_nccl.AllReduce(ctx->comm_d, ctx->comm_d, (size_t)k, ncclFloat64, op == 1 ? ncclMax : ncclSum, (ncclComm_t)ctx->comm, st));
        FRK_CHECK_CUDA(cudaMemcpyAsync(ctx->comm_h, ctx->comm_d, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, st));
        FRK_CHECK_CUDA(cudaStreamSynchronize(st));
        memcpy(h_buf + o, ctx->comm_h, (size_t)k * sizeof(double));
    }
    return 0;
}

// ---------------------------------------------------------------------------------
// packed upper triangle <-> full column-major storage (what crosses the ranks is r(r+1)/2 doubles, not r^2)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_upper_kernel(int r, const double *__restrict__ G, double *__restrict__ P, int unpack_into_G,
                                                         double *__restrict__ Gout) {
    // column j holds rows 0..j at P[j(j+1)/2 ...]; one CTA per column (grid-stride), coalesced along the column
    for (int j = blockIdx.x; j < r; j += gridDim.x) {
        const size_t po = (size_t)j * (j + 1) / 2, go = (size_t)j * r;
        if (!unpack_into_G) for (int i = threadIdx.x; i <= j; i += blockDim.x) P[po + i] = G[go + i];
        else                for (int i = threadIdx.x; i <= j; i += blockDim.x) Gout[go + i] = P[po + i];
    }
}

int frk_pack_upper(frk_ctx *ctx, int r, const double *d_G, double *d_P) {
    FRK_LAUNCH(ctx, pack_upper_kernel, (unsigned)std::min(r, ctx->sm_count * 8), 256, 0, r, d_G, d_P, 0, (double *)nullptr);
    return 0;
}

int frk_unpack_upper(frk_ctx *ctx, int r, const double *d_P, double *d_G) {
    FRK_LAUNCH(ctx, pack_upper_kernel, (unsigned)std::min(r, ctx->sm_count * 8), 256, 0, r, (const double *)nullptr, (double *)d_P, 1, d_G);
    return 0;
}

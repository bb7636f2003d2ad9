// Host-buffer entry points: what the R `.Call` shim (r/src/frkb200_R.c) binds.
// Each one uploads its host arguments, runs the device path, and downloads the result; nothing
// here computes on the CPU.
#include "common.cuh"

namespace {
struct DevBuf {
    void *p = nullptr;
    cudaStream_t s;
    explicit DevBuf(cudaStream_t st) : s(st) {}
    ~DevBuf() { if (p) cudaFreeAsync(p, s); }
    int alloc(size_t bytes) {
        FRK_CHECK_CUDA(cudaMallocAsync(&p, bytes ? bytes : 8, s));
        return 0;
    }
    template <typename T> T *as() { return (T *)p; }
};
}  // namespace

extern "C" int frk_eval_basis_host(frk_ctx *ctx, const frk_basis *b, int64_t n, const double *h_s, int normalise,
                                   int *h_rowptr, int *h_col, double *h_val, int64_t *h_nnz) {
    FRK_REQUIRE(ctx && b && h_rowptr && h_nnz && (n == 0 || h_s), "frk_eval_basis_host: NULL argument");
    const int d = frk_basis_dim(b);
    DevBuf s(ctx->stream), rp(ctx->stream), col(ctx->stream), val(ctx->stream);
    if (int rc = s.alloc((size_t)n * d * sizeof(double))) return rc;
    if (int rc = rp.alloc((size_t)(n + 1) * sizeof(int))) return rc;
    FRK_CHECK_CUDA(cudaMemcpyAsync(s.p, h_s, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    int64_t nnz = 0;
    if (int rc = frk_eval_basis_count(ctx, b, n, s.as<double>(), n, rp.as<int>(), &nnz)) return rc;
    *h_nnz = nnz;
    FRK_CHECK_CUDA(cudaMemcpyAsync(h_rowptr, rp.p, (size_t)(n + 1) * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (h_col == nullptr || h_val == nullptr) {   // size query
        FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    if (int rc = col.alloc((size_t)nnz * sizeof(int))) return rc;
    if (int rc = val.alloc((size_t)nnz * sizeof(double))) return rc;
    if (int rc = frk_eval_basis_fill(ctx, b, n, s.as<double>(), n, rp.as<int>(), col.as<int>(), val.as<double>(), normalise)) return rc;
    FRK_CHECK_CUDA(cudaMemcpyAsync(h_col, col.p, (size_t)nnz * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FRK_CHECK_CUDA(cudaMemcpyAsync(h_val, val.p, (size_t)nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int frk_grid_lookup_host(frk_ctx *ctx, int64_t n, const double *h_xy, const double *h_offset,
                                    const double *h_cellsize, const int *h_dims, const int *h_cell2bau, int64_t ncell,
                                    int *h_bau) {
    FRK_REQUIRE(ctx && h_bau && (n == 0 || h_xy), "frk_grid_lookup_host: NULL argument");
    DevBuf xy(ctx->stream), map(ctx->stream), out(ctx->stream);
    if (int rc = xy.alloc((size_t)n * 2 * sizeof(double))) return rc;
    if (int rc = out.alloc((size_t)n * sizeof(int))) return rc;
    FRK_CHECK_CUDA(cudaMemcpyAsync(xy.p, h_xy, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (h_cell2bau) {
        FRK_REQUIRE(ncell == (int64_t)h_dims[0] * h_dims[1], "frk_grid_lookup_host: cell2bau must have cells.dim[1]*cells.dim[2] entries");
        if (int rc = map.alloc((size_t)ncell * sizeof(int))) return rc;
        FRK_CHECK_CUDA(cudaMemcpyAsync(map.p, h_cell2bau, (size_t)ncell * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (int rc = frk_grid_lookup(ctx, n, xy.as<double>(), n, h_offset, h_cellsize, h_dims, h_cell2bau ? map.as<int>() : nullptr,
                                 out.as<int>())) return rc;
    FRK_CHECK_CUDA(cudaMemcpyAsync(h_bau, out.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// over(SpatialPoints, SpatialPolygons BAUs) from host buffers (hexagon / irregular BAUs): the polygon index is built, used and dropped
extern "C" int frk_polygon_lookup_host(frk_ctx *ctx, int npoly, const int *h_ptr, const double *h_vx, const double *h_vy, int64_t n,
                                       const double *h_xy, int *h_bau) {
    FRK_REQUIRE(ctx && h_ptr && h_vx && h_vy && h_bau && npoly > 0 && (n == 0 || h_xy), "frk_polygon_lookup_host: NULL argument");
    if (n == 0) return 0;
    frk_polygons *P = nullptr;
    if (int rc = frk_polygons_create(ctx, npoly, h_ptr, h_vx, h_vy, &P)) return rc;
    DevBuf xy(ctx->stream), out(ctx->stream);
    auto run = [&]() -> int {
        if (int rc = xy.alloc((size_t)n * 2 * sizeof(double))) return rc;
        if (int rc = out.alloc((size_t)n * sizeof(int))) return rc;
        FRK_CHECK_CUDA(cudaMemcpyAsync(xy.p, h_xy, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if (int rc = frk_polygon_lookup(ctx, P, n, xy.as<double>(), n, 0, npoly, out.as<int>())) return rc;
        FRK_CHECK_CUDA(cudaMemcpyAsync(h_bau, out.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        return 0;
    };
    const int rc = run();
    frk_polygons_destroy(P);
    return rc;
}

extern "C" int frk_quadform_host(frk_ctx *ctx, int64_t n, int r, const int *h_rowptr, const int *h_col, const double *h_val,
                                 const double *h_M, const double *h_mu, double *h_q, double *h_t) {
    FRK_REQUIRE(ctx && h_rowptr && (h_q || h_t), "frk_quadform_host: NULL argument");
    const int64_t nnz = h_rowptr[n];
    DevBuf rp(ctx->stream), col(ctx->stream), val(ctx->stream), M(ctx->stream), mu(ctx->stream), q(ctx->stream), t(ctx->stream);
    if (int rc = rp.alloc((size_t)(n + 1) * sizeof(int))) return rc;
    if (int rc = col.alloc((size_t)nnz * sizeof(int))) return rc;
    if (int rc = val.alloc((size_t)nnz * sizeof(double))) return rc;
    FRK_CHECK_CUDA(cudaMemcpyAsync(rp.p, h_rowptr, (size_t)(n + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    FRK_CHECK_CUDA(cudaMemcpyAsync(col.p, h_col, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    FRK_CHECK_CUDA(cudaMemcpyAsync(val.p, h_val, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (h_M && h_q) {
        if (int rc = M.alloc((size_t)r * r * sizeof(double))) return rc;
        if (int rc = q.alloc((size_t)n * sizeof(double))) return rc;
        FRK_CHECK_CUDA(cudaMemcpyAsync(M.p, h_M, (size_t)r * r * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (h_mu && h_t) {
        if (int rc = mu.alloc((size_t)r * sizeof(double))) return rc;
        if (int rc = t.alloc((size_t)n * sizeof(double))) return rc;
        FRK_CHECK_CUDA(cudaMemcpyAsync(mu.p, h_mu, (size_t)r * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (int rc = frk_quadform_spmv(ctx, n, r, rp.as<int>(), col.as<int>(), val.as<double>(), M.as<double>(), mu.as<double>(),
                                   q.as<double>(), t.as<double>())) return rc;
    if (q.p) FRK_CHECK_CUDA(cudaMemcpyAsync(h_q, q.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (t.p) FRK_CHECK_CUDA(cudaMemcpyAsync(h_t, t.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int frk_chol2inv_host(frk_ctx *ctx, int r, const double *h_A, double *h_Ainv, double *h_logdet) {
    FRK_REQUIRE(ctx && h_A && h_Ainv && r > 0, "frk_chol2inv_host: NULL argument");
    DevBuf A(ctx->stream), Ai(ctx->stream), W(ctx->stream);
    const size_t bytes = (size_t)r * r * sizeof(double);
    if (int rc = A.alloc(bytes)) return rc;
    if (int rc = Ai.alloc(bytes)) return rc;
    if (int rc = W.alloc(bytes)) return rc;
    FRK_CHECK_CUDA(cudaMemcpyAsync(A.p, h_A, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = frk_potrf(ctx, r, A.as<double>(), h_logdet)) return rc;
    if (int rc = frk_potri(ctx, r, A.as<double>(), Ai.as<double>(), W.as<double>())) return rc;
    FRK_CHECK_CUDA(cudaMemcpyAsync(h_Ainv, Ai.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

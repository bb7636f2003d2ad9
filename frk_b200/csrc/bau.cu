// Point -> BAU incidence (C_Z), binning of observations into BAUs, and the S = C_Z S0 row gather.
//
// Reference: map_data_to_BAUs(SpatialPoints) R/geometryfns.R:1222-1320 (sp::over on a pixel grid,
// then group_by(BAU) %>% summarise_all(mean)), BuildC R/geometryfns.R:1557-1568, S <- Cmat %*% S0
// R/SRE.R:502.  sp::over on SpatialPixels is sp's getGridIndex (sp is not vendored in the reference;
// restated from its published algorithm): idx_d = round((x_d - offset_d)/cellsize_d), rows counted
// from the top, outside -> NA.  R's round() is half-to-even == rint().
#include "common.cuh"

__global__ void __launch_bounds__(256) grid_lookup_kernel(int64_t n, const double *__restrict__ xy, int64_t ld, double ox, double oy,
                                                          double cx, double cy, int nx, int ny,
                                                          const int *__restrict__ cell2bau, int expand_grid_order, int lo,
                                                          int hi, int *__restrict__ bau) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double fx = rint(__ddiv_rn(__dsub_rn(xy[i], ox), cx));
        double fy = rint(__ddiv_rn(__dsub_rn(xy[i + ld], oy), cy));
        fy = (double)ny - (fy + 1.0);                       // this.idx = cells.dim[2] - (this.idx + 1)
        int out = -1;
        if (fx >= 0.0 && fx < (double)nx && fy >= 0.0 && fy < (double)ny) {   // NaN fails every test -> NA
            int64_t cell = (int64_t)fy * nx + (int64_t)fx;
            if (cell2bau) out = __ldg(cell2bau + cell);
            else if (expand_grid_order) out = (int)(((int64_t)ny - 1 - (int64_t)fy) * nx + (int64_t)fx);   // BAU rows bottom-up
            else out = (int)cell;
            if (out >= 0) out = (out >= lo && out < hi) ? out - lo : -1;
        }
        bau[i] = out;
    }
}

extern "C" int frk_grid_lookup(frk_ctx *ctx, int64_t n, const double *d_xy, int64_t ld_xy, const double *h_offset,
                               const double *h_cellsize, const int *h_dims, const int *d_cell2bau, int *d_bau) {
    return frk_grid_lookup_range(ctx, n, d_xy, ld_xy, h_offset, h_cellsize, h_dims, d_cell2bau, 0, 0, 2147483647, d_bau);
}

extern "C" int frk_grid_lookup_range(frk_ctx *ctx, int64_t n, const double *d_xy, int64_t ld_xy, const double *h_offset,
                                     const double *h_cellsize, const int *h_dims, const int *d_cell2bau,
                                     int expand_grid_order, int bau_lo, int bau_hi, int *d_bau) {
    FRK_REQUIRE(ctx && h_offset && h_cellsize && h_dims && d_bau, "frk_grid_lookup: NULL argument");
    FRK_REQUIRE(h_cellsize[0] > 0 && h_cellsize[1] > 0 && h_dims[0] > 0 && h_dims[1] > 0, "frk_grid_lookup: bad grid");
    FRK_REQUIRE((int64_t)h_dims[0] * h_dims[1] < 2147483647LL, "frk_grid_lookup: grid has too many cells for int32");
    if (n == 0) return 0;
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, grid_lookup_kernel, grid, 256, 0, n, d_xy, ld_xy, h_offset[0], h_offset[1], h_cellsize[0], h_cellsize[1],
               h_dims[0], h_dims[1], d_cell2bau, expand_grid_order, bau_lo, bau_hi, d_bau);
    return 0;
}

// ---------------------------------------------------------------------------------
// binning: deterministic counting sort by BAU, then per-BAU mean in ascending point order
// ---------------------------------------------------------------------------------
__global__ void bin_count_kernel(int64_t n, const int *__restrict__ bau, int *__restrict__ cnt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int b = bau[i];
        if (b >= 0) atomicAdd(cnt + b, 1);
    }
}

__global__ void bin_flag_kernel(int64_t nbau, const int *__restrict__ cnt, int *__restrict__ flag) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbau; i += (int64_t)gridDim.x * blockDim.x)
        flag[i] = cnt[i] > 0;
}

// slot[b] (exclusive scan of occupancy flags) is the output row of BAU b; start[b] (exclusive scan of
// counts) is where its members go in the member list.
__global__ void bin_scatter_kernel(int64_t n, const int *__restrict__ bau, const int *__restrict__ start, int *__restrict__ fill,
                                   int *__restrict__ members) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int b = bau[i];
        if (b >= 0) {
            int k = atomicAdd(fill + b, 1);
            members[start[b] + k] = (int)i;
        }
    }
}

__global__ void bin_reduce_kernel(int64_t nbau, int64_t n, int ncols, const int *__restrict__ cnt, const int *__restrict__ slot,
                                  const int *__restrict__ start, int *__restrict__ members, const double *__restrict__ vals,
                                  int *__restrict__ obs_bau, double *__restrict__ means, int *__restrict__ counts) {
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nbau; b += (int64_t)gridDim.x * blockDim.x) {
        int c = cnt[b];
        if (c == 0) continue;
        int *mem = members + start[b];
        for (int a = 1; a < c; a++) {   // restore ascending point order (atomics scrambled it)
            int v = mem[a], k = a - 1;
            while (k >= 0 && mem[k] > v) { mem[k + 1] = mem[k]; k--; }
            mem[k + 1] = v;
        }
        int row = slot[b];
        obs_bau[row] = (int)b;
        counts[row] = c;
        for (int col = 0; col < ncols; col++) {
            const double *v = vals + (int64_t)col * n;
            // R's mean(): s = sum(x)/n, then s + sum(x - s)/n
            double s = 0.0;
            for (int a = 0; a < c; a++) s += v[mem[a]];
            s /= (double)c;
            double t = 0.0;
            for (int a = 0; a < c; a++) t += v[mem[a]] - s;
            means[(int64_t)col * n + row] = s + t / (double)c;
        }
    }
}

extern "C" int frk_bin_mean(frk_ctx *ctx, int64_t n, const int *d_bau, int64_t nbau, int ncols, const double *d_vals,
                            int *d_obs_bau, double *d_means, int *d_counts, int64_t *h_m) {
    FRK_REQUIRE(ctx && d_bau && d_obs_bau && d_means && d_counts && h_m, "frk_bin_mean: NULL argument");
    FRK_REQUIRE(nbau > 0 && nbau < 2147483647LL && n < 2147483647LL, "frk_bin_mean: sizes exceed int32");
    *h_m = 0;
    if (n == 0) return 0;
    // scratch: cnt[nbau] | slot[nbau+1] | start[nbau+1] | fill[nbau] | flag[nbau] | members[n]
    void *raw = nullptr;
    size_t ints = (size_t)(5 * nbau + 2 + n);
    if (int rc = frk_scratch(ctx, ints * sizeof(int), &raw)) return rc;
    int *cnt = (int *)raw, *slot = cnt + nbau, *start = slot + nbau + 1, *fill = start + nbau + 1, *flag = fill + nbau,
        *members = flag + nbau;
    FRK_CHECK_CUDA(cudaMemsetAsync(cnt, 0, nbau * sizeof(int), ctx->stream));
    FRK_CHECK_CUDA(cudaMemsetAsync(fill, 0, nbau * sizeof(int), ctx->stream));
    unsigned gp = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16);
    unsigned gb = (unsigned)std::min<int64_t>(cdiv(nbau, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, bin_count_kernel, gp, 256, 0, n, d_bau, cnt);
    FRK_LAUNCH(ctx, bin_flag_kernel, gb, 256, 0, nbau, cnt, flag);
    int64_t m = 0, kept = 0;
    if (int rc = frk_exclusive_scan_i32(ctx, flag, nbau, slot, &m)) return rc;
    if (int rc = frk_exclusive_scan_i32(ctx, cnt, nbau, start, &kept)) return rc;
    FRK_LAUNCH(ctx, bin_scatter_kernel, gp, 256, 0, n, d_bau, start, fill, members);
    FRK_LAUNCH(ctx, bin_reduce_kernel, gb, 256, 0, nbau, n, ncols, cnt, slot, start, members, d_vals, d_obs_bau, d_means, d_counts);
    *h_m = m;
    return 0;
}

// ---------------------------------------------------------------------------------
// CSR row gather  S = C_Z S0  (one 1 per row of C_Z)
// ---------------------------------------------------------------------------------
__global__ void gather_count_kernel(int64_t m, const int *__restrict__ rows, const int *__restrict__ rp, int *__restrict__ cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) { int r = rows[i]; cnt[i] = rp[r + 1] - rp[r]; }
}

__global__ void __launch_bounds__(256) gather_fill_kernel(int64_t m, const int *__restrict__ rows, const int *__restrict__ rps,
                                                          const int *__restrict__ cs, const double *__restrict__ vs,
                                                          const int *__restrict__ rpd, int *__restrict__ cd, double *__restrict__ vd) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= m) return;
    int src = rps[rows[row]], dst = rpd[row], k = rpd[row + 1] - dst;
    for (int t = lane; t < k; t += 32) { cd[dst + t] = cs[src + t]; vd[dst + t] = vs[src + t]; }
}

extern "C" int frk_csr_gather_count(frk_ctx *ctx, int64_t m, const int *d_rows, const int *d_rowptr_src, int *d_rowptr_dst,
                                    int64_t *h_nnz) {
    if (m == 0) { FRK_CHECK_CUDA(cudaMemsetAsync(d_rowptr_dst, 0, sizeof(int), ctx->stream)); if (h_nnz) *h_nnz = 0; return 0; }
    int *counts = nullptr;
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&counts, m * sizeof(int), ctx->stream));
    FRK_LAUNCH(ctx, gather_count_kernel, (unsigned)cdiv(m, 256), 256, 0, m, d_rows, d_rowptr_src, counts);
    int rc = frk_exclusive_scan_i32(ctx, counts, m, d_rowptr_dst, h_nnz);
    cudaFreeAsync(counts, ctx->stream);
    return rc;
}

extern "C" int frk_csr_gather_fill(frk_ctx *ctx, int64_t m, const int *d_rows, const int *d_rowptr_src, const int *d_col_src,
                                   const double *d_val_src, const int *d_rowptr_dst, int *d_col_dst, double *d_val_dst) {
    if (m == 0) return 0;
    FRK_LAUNCH(ctx, gather_fill_kernel, (unsigned)cdiv(m * 32, 256), 256, 0, m, d_rows, d_rowptr_src, d_col_src, d_val_src,
               d_rowptr_dst, d_col_dst, d_val_dst);
    return 0;
}

// ---------------------------------------------------------------------------------
// BAU centroids of a pixel grid in expand.grid order (x fastest), rows [lo, hi):
// coordinates(BAUs) as SRE() passes them to eval_basis (R/SRE.R:412, R/geometryfns.R:629-637).
// out: (hi-lo) x 2 column-major.
// ---------------------------------------------------------------------------------
__global__ void grid_centroids_kernel(int64_t lo, int64_t n, int nx, const double *__restrict__ xg, const double *__restrict__ yg,
                                      double *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = lo + i;
        out[i] = __ldg(xg + (b % nx));
        out[i + n] = __ldg(yg + (b / nx));
    }
}

extern "C" int frk_grid_centroids(frk_ctx *ctx, int64_t lo, int64_t hi, int nx, const double *d_xgrid, const double *d_ygrid,
                                  double *d_out) {
    FRK_REQUIRE(ctx && d_xgrid && d_ygrid && d_out && nx > 0 && hi >= lo, "frk_grid_centroids: bad argument");
    const int64_t n = hi - lo;
    if (n == 0) return 0;
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, grid_centroids_kernel, grid, 256, 0, lo, n, nx, d_xgrid, d_ygrid, d_out);
    return 0;
}

// Point -> BAU incidence (C_Z), binning of observations into BAUs, and the S = C_Z S0 row gather.
//
// Reference: map_data_to_BAUs(SpatialPoints) R/geometryfns.R:1222-1320 (sp::over on a pixel grid,
// then group_by(BAU) %>% summarise_all(mean)), BuildC R/geometryfns.R:1557-1568, S <- Cmat %*% S0
// R/SRE.R:502.  sp::over on SpatialPixels is sp's getGridIndex (sp is not vendored in the reference;
// restated from its published algorithm): idx_d = round((x_d - offset_d)/cellsize_d), rows counted
// from the top, outside -> NA.  R's round() is half-to-even == rint().
#include "common.cuh"
#include <cub/cub.cuh>

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
// map_data_to_BAUs(ST) R/geometryfns.R:1419-1547: time slice i holds the data with t1_i <= time < t1_{i+1}; the last slice takes
// everything from its start on (:1462-1468); each slice is then binned in space.  With the ST BAUs ordered space-fastest
// (auto_BAUs, :979-980) that is one combined index  bau = spatial + ns * slice.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) st_bau_index_kernel(int64_t n, const int *__restrict__ sp, const double *__restrict__ t, int nt,
                                                           const double *__restrict__ brk, int64_t ns, int64_t lo, int64_t hi,
                                                           int *__restrict__ bau) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double ti = t[i];
        const int s = sp[i];
        int out = -1;
        if (s >= 0 && ti >= __ldg(brk)) {            // NaN and times before the first interval -> NA
            int a = 0, b = nt;                       // largest a with brk[a] <= ti
            while (b - a > 1) {
                int mid = (a + b) >> 1;
                if (__ldg(brk + mid) <= ti) a = mid; else b = mid;
            }
            const int64_t g = (int64_t)s + ns * a;
            if (g >= lo && g < hi) out = (int)(g - lo);
        }
        bau[i] = out;
    }
}

extern "C" int frk_st_bau_index(frk_ctx *ctx, int64_t n, const int *d_spatial, const double *d_time, int nt, const double *h_breaks,
                                int64_t ns, int64_t bau_lo, int64_t bau_hi, int *d_bau) {
    FRK_REQUIRE(ctx && h_breaks && d_bau && (n == 0 || (d_spatial && d_time)), "frk_st_bau_index: NULL argument");
    FRK_REQUIRE(nt > 0 && ns > 0, "frk_st_bau_index: need at least one time interval and one spatial BAU");
    FRK_REQUIRE(ns * (int64_t)nt < 2147483647LL, "frk_st_bau_index: too many space-time BAUs for int32");
    for (int k = 1; k < nt; k++) FRK_REQUIRE(h_breaks[k] > h_breaks[k - 1], "frk_st_bau_index: interval starts must be increasing");
    if (n == 0) return 0;
    double *d_brk = nullptr;
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&d_brk, nt * sizeof(double), ctx->stream));
    // pageable source: the copy is staged before the call returns, so h_breaks may be freed by the caller
    FRK_CHECK_CUDA(cudaMemcpyAsync(d_brk, h_breaks, nt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, st_bau_index_kernel, grid, 256, 0, n, d_spatial, d_time, nt, d_brk, ns, bau_lo, bau_hi, d_bau);
    cudaFreeAsync(d_brk, ctx->stream);
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
                                  unsigned sum_mask, int *__restrict__ obs_bau, double *__restrict__ means, int *__restrict__ counts) {
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
            if (col < 32 && ((sum_mask >> col) & 1u)) { means[(int64_t)col * n + row] = s; continue; }   // sum_variables: .safe_sum, :1294-1299
            s /= (double)c;
            double t = 0.0;
            for (int a = 0; a < c; a++) t += v[mem[a]] - s;
            means[(int64_t)col * n + row] = s + t / (double)c;
        }
    }
}

extern "C" int frk_bin_mean(frk_ctx *ctx, int64_t n, const int *d_bau, int64_t nbau, int ncols, const double *d_vals,
                            int *d_obs_bau, double *d_means, int *d_counts, int64_t *h_m) {
    return frk_bin_mean_sum(ctx, n, d_bau, nbau, ncols, d_vals, 0u, d_obs_bau, d_means, d_counts, h_m);
}

extern "C" int frk_bin_mean_sum(frk_ctx *ctx, int64_t n, const int *d_bau, int64_t nbau, int ncols, const double *d_vals, unsigned sum_mask,
                                int *d_obs_bau, double *d_means, int *d_counts, int64_t *h_m) {
    FRK_REQUIRE(ctx && d_bau && d_obs_bau && d_means && d_counts && h_m, "frk_bin_mean: NULL argument");
    FRK_REQUIRE(sum_mask == 0u || ncols <= 32, "frk_bin_mean_sum: sum_variables among more than 32 columns");
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
    FRK_LAUNCH(ctx, bin_reduce_kernel, gb, 256, 0, nbau, n, ncols, cnt, slot, start, members, d_vals, sum_mask, d_obs_bau, d_means, d_counts);
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

// ---------------------------------------------------------------------------------
// map_data_to_BAUs for polygon BAUs (ISEA3H hexagons): over(SpatialPoints, SpatialPolygons), R/geometryfns.R:1266,1680.
// sp is not vendored; its point.in.polygon (pip.c InPoly) is O'Rourke's crossing test, restated: shift the ring so that the
// point is the origin, count edges crossing the positive / negative x axis with x = (x_i y_{i-1} - x_{i-1} y_i) / (y_{i-1} - y_i).
// Interior, edge and vertex all count as "in"; the first polygon in index order wins.  A uniform cell grid over the polygons'
// bounding boxes (built on the host at frk_polygons_create) gives each point its ascending candidate list.
// ---------------------------------------------------------------------------------
struct frk_polygons {
    frk_ctx *ctx = nullptr;
    int npoly = 0;
    double x0 = 0, y0 = 0, hx = 1, hy = 1;
    int nx = 1, ny = 1;
    int *ptr = nullptr, *cell_start = nullptr, *cand = nullptr;
    double *vx = nullptr, *vy = nullptr, *bbox = nullptr;     // bbox: xmin, xmax, ymin, ymax per polygon
    std::vector<void *> allocs;
};

__device__ __forceinline__ int in_poly_dev(double px, double py, const double *__restrict__ vx, const double *__restrict__ vy, int n) {
    int rcross = 0, lcross = 0;
    double x1 = __dsub_rn(__ldg(vx + n - 1), px), y1 = __dsub_rn(__ldg(vy + n - 1), py);     // vertex i-1
    for (int i = 0; i < n; i++) {
        const double x = __dsub_rn(__ldg(vx + i), px), y = __dsub_rn(__ldg(vy + i), py);
        if (x == 0.0 && y == 0.0) return 3;
        const bool rstrad = (y > 0) != (y1 > 0), lstrad = (y < 0) != (y1 < 0);
        if (rstrad || lstrad) {
            const double xc = __ddiv_rn(__dsub_rn(__dmul_rn(x, y1), __dmul_rn(x1, y)), __dsub_rn(y1, y));
            if (rstrad && xc > 0) rcross++;
            if (lstrad && xc < 0) lcross++;
        }
        x1 = x; y1 = y;
    }
    if ((rcross & 1) != (lcross & 1)) return 2;
    return (rcross & 1) ? 1 : 0;
}

__global__ void __launch_bounds__(256) polygon_lookup_kernel(int64_t n, const double *__restrict__ xy, int64_t ld, double x0, double y0,
                                                             double hx, double hy, int nx, int ny, const int *__restrict__ cell_start,
                                                             const int *__restrict__ cand, const int *__restrict__ ptr,
                                                             const double *__restrict__ vx, const double *__restrict__ vy,
                                                             const double *__restrict__ bbox, int lo, int hi, int *__restrict__ bau,
                                                             int *__restrict__ multi) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double px = xy[i], py = xy[i + ld];
        int out = -1;
        double fx = floor((px - x0) / hx), fy = floor((py - y0) / hy);
        if (fx == (double)nx && px == x0 + hx * nx) fx = nx - 1;       // closed upper edges of the covered rectangle
        if (fy == (double)ny && py == y0 + hy * ny) fy = ny - 1;
        if (fx >= 0 && fx < nx && fy >= 0 && fy < ny) {                // NaN fails every test -> NA
            const int64_t c = (int64_t)fy * nx + (int64_t)fx;
            for (int k = cell_start[c]; k < cell_start[c + 1]; k++) {
                const int pid = __ldg(cand + k);
                const double *bb = bbox + 4 * (size_t)pid;
                if (px < __ldg(bb) || px > __ldg(bb + 1) || py < __ldg(bb + 2) || py > __ldg(bb + 3)) continue;
                const int b = ptr[pid];
                if (in_poly_dev(px, py, vx + b, vy + b, ptr[pid + 1] - b) != 0) {
                    if (out < 0) { out = pid; if (!multi) break; }
                    else { atomicAdd(multi, 1); break; }                 // a second polygon holds this point too
                }
            }
        }
        if (out >= 0) out = (out >= lo && out < hi) ? out - lo : -1;
        bau[i] = out;
    }
}

extern "C" int frk_polygons_destroy(frk_polygons *P) {
    if (!P) return 0;
    cudaSetDevice(P->ctx->device);
    cudaStreamSynchronize(P->ctx->stream);
    for (void *p : P->allocs) cudaFree(p);
    delete P;
    return 0;
}

extern "C" int frk_polygons_create(frk_ctx *ctx, int npoly, const int *h_ptr, const double *h_vx, const double *h_vy, frk_polygons **out) {
    FRK_REQUIRE(ctx && out && npoly > 0 && h_ptr && h_vx && h_vy, "frk_polygons_create: bad argument");
    frk_polygons *P = new frk_polygons();
    P->ctx = ctx; P->npoly = npoly;
    const int nv = h_ptr[npoly];
    std::vector<double> bbox(4 * (size_t)npoly);
    double X0 = 1e300, X1 = -1e300, Y0 = 1e300, Y1 = -1e300, wsum = 0, hsum = 0;
    for (int k = 0; k < npoly; k++) {
        FRK_REQUIRE(h_ptr[k + 1] - h_ptr[k] >= 3, "frk_polygons_create: polygon %d has fewer than 3 vertices", k + 1);
        double a = 1e300, b = -1e300, c = 1e300, d = -1e300;
        for (int v = h_ptr[k]; v < h_ptr[k + 1]; v++) { a = std::min(a, h_vx[v]); b = std::max(b, h_vx[v]); c = std::min(c, h_vy[v]); d = std::max(d, h_vy[v]); }
        bbox[4 * k] = a; bbox[4 * k + 1] = b; bbox[4 * k + 2] = c; bbox[4 * k + 3] = d;
        X0 = std::min(X0, a); X1 = std::max(X1, b); Y0 = std::min(Y0, c); Y1 = std::max(Y1, d);
        wsum += b - a; hsum += d - c;
    }
    // cells about the size of an average polygon, at most 2048 x 2048
    const double W = std::max(X1 - X0, 1e-300), H = std::max(Y1 - Y0, 1e-300);
    P->nx = std::max(1, std::min(2048, (int)std::floor(W / std::max(wsum / npoly, W / 2048))));
    P->ny = std::max(1, std::min(2048, (int)std::floor(H / std::max(hsum / npoly, H / 2048))));
    P->x0 = X0; P->y0 = Y0; P->hx = W / P->nx; P->hy = H / P->ny;
    const int64_t ncell = (int64_t)P->nx * P->ny;
    std::vector<int> start(ncell + 1, 0), cnt(ncell, 0), cand;
    auto range = [&](double lo, double hi, double o, double h, int n, int &i0, int &i1) {
        i0 = std::max(0, (int)std::floor((lo - o) / h) - 1);
        i1 = std::min(n - 1, (int)std::floor((hi - o) / h) + 1);
    };
    for (int pass = 0; pass < 2; pass++) {
        for (int k = 0; k < npoly; k++) {                              // ascending k -> ascending candidate lists
            int ix0, ix1, iy0, iy1;
            range(bbox[4 * k], bbox[4 * k + 1], P->x0, P->hx, P->nx, ix0, ix1);
            range(bbox[4 * k + 2], bbox[4 * k + 3], P->y0, P->hy, P->ny, iy0, iy1);
            for (int iy = iy0; iy <= iy1; iy++)
                for (int ix = ix0; ix <= ix1; ix++) {
                    const int64_t c = (int64_t)iy * P->nx + ix;
                    if (pass == 0) start[c + 1]++;
                    else cand[start[c] + cnt[c]++] = k;
                }
        }
        if (pass == 0) {
            for (int64_t c = 0; c < ncell; c++) start[c + 1] += start[c];
            cand.resize(start[ncell]);
        }
    }
    auto up = [&](const void *h, size_t bytes, void **d) -> int {
        FRK_CHECK_CUDA(cudaMalloc(d, std::max<size_t>(bytes, 8)));
        P->allocs.push_back(*d);
        FRK_CHECK_CUDA(cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice));
        return 0;
    };
    int rc = 0;
    rc = rc || up(h_ptr, (size_t)(npoly + 1) * sizeof(int), (void **)&P->ptr) || up(h_vx, (size_t)nv * sizeof(double), (void **)&P->vx) ||
         up(h_vy, (size_t)nv * sizeof(double), (void **)&P->vy) || up(bbox.data(), bbox.size() * sizeof(double), (void **)&P->bbox) ||
         up(start.data(), start.size() * sizeof(int), (void **)&P->cell_start) || up(cand.data(), cand.size() * sizeof(int), (void **)&P->cand);
    if (rc) { frk_polygons_destroy(P); return 1; }
    *out = P;
    return 0;
}

// d_bau[i] = 0-based polygon (BAU) row of point i relative to bau_lo, -1 for NA (outside every polygon, or outside [bau_lo, bau_hi))
extern "C" int frk_polygon_lookup(frk_ctx *ctx, const frk_polygons *P, int64_t n, const double *d_xy, int64_t ld_xy, int bau_lo, int bau_hi,
                                  int *d_bau) {
    FRK_REQUIRE(ctx && P && d_bau && (n == 0 || d_xy), "frk_polygon_lookup: NULL argument");
    if (n == 0) return 0;
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, polygon_lookup_kernel, grid, 256, 0, n, d_xy, ld_xy, P->x0, P->y0, P->hx, P->hy, P->nx, P->ny, P->cell_start, P->cand,
               P->ptr, P->vx, P->vy, P->bbox, bau_lo, bau_hi, d_bau, (int *)nullptr);
    return 0;
}

// as frk_polygon_lookup over all polygons, and *h_multi = number of points that lie in MORE than one polygon
extern "C" int frk_polygon_lookup_count(frk_ctx *ctx, const frk_polygons *P, int64_t n, const double *d_xy, int64_t ld_xy, int *d_first,
                                        int64_t *h_multi) {
    FRK_REQUIRE(ctx && P && d_first && h_multi && (n == 0 || d_xy), "frk_polygon_lookup_count: NULL argument");
    *h_multi = 0;
    if (n == 0) return 0;
    int *cnt = (int *)(ctx->d_small + 3000);
    FRK_CHECK_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int), ctx->stream));
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, polygon_lookup_kernel, grid, 256, 0, n, d_xy, ld_xy, P->x0, P->y0, P->hx, P->hy, P->nx, P->ny, P->cell_start, P->cand,
               P->ptr, P->vx, P->vy, P->bbox, 0, 2147483647, d_first, cnt);
    int h = 0;
    FRK_CHECK_CUDA(cudaMemcpyAsync(&h, cnt, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    *h_multi = h;
    return 0;
}

// ---------------------------------------------------------------------------------
// General incidence of points in (possibly overlapping) polygons: BuildC(SpatialPolygons) for prediction polygons, C_P
// (R/geometryfns.R:1572-1616, R/SREpredict.R:119-128): every polygon gets ALL the points (BAU centroids) it contains.
// Pass 1 counts the containing polygons per point, pass 2 emits (polygon, point) pairs, a stable radix sort by polygon turns
// them into CSR rows with ascending point index; the weights are BAUs$wts / row sum (normalise_wts).
// ---------------------------------------------------------------------------------
template <bool EMIT>
__global__ void __launch_bounds__(256) polygon_incidence_kernel(int64_t n, const double *__restrict__ xy, int64_t ld, double x0, double y0,
                                                                double hx, double hy, int nx, int ny, const int *__restrict__ cell_start,
                                                                const int *__restrict__ cand, const int *__restrict__ ptr,
                                                                const double *__restrict__ vx, const double *__restrict__ vy,
                                                                const double *__restrict__ bbox, int *__restrict__ cnt_or_off,
                                                                int *__restrict__ pid_out, int *__restrict__ pt_out, int *__restrict__ pcnt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double px = xy[i], py = xy[i + ld];
        int k = 0;
        const int o = EMIT ? cnt_or_off[i] : 0;
        double fx = floor((px - x0) / hx), fy = floor((py - y0) / hy);
        if (fx == (double)nx && px == x0 + hx * nx) fx = nx - 1;
        if (fy == (double)ny && py == y0 + hy * ny) fy = ny - 1;
        if (fx >= 0 && fx < nx && fy >= 0 && fy < ny) {
            const int64_t c = (int64_t)fy * nx + (int64_t)fx;
            for (int e = cell_start[c]; e < cell_start[c + 1]; e++) {
                const int pid = __ldg(cand + e);
                const double *bb = bbox + 4 * (size_t)pid;
                if (px < __ldg(bb) || px > __ldg(bb + 1) || py < __ldg(bb + 2) || py > __ldg(bb + 3)) continue;
                const int b = ptr[pid];
                if (in_poly_dev(px, py, vx + b, vy + b, ptr[pid + 1] - b) != 0) {
                    if (EMIT) { pid_out[o + k] = pid; pt_out[o + k] = (int)i; atomicAdd(pcnt + pid, 1); }
                    k++;
                }
            }
        }
        if (!EMIT) cnt_or_off[i] = k;
    }
}

__global__ void __launch_bounds__(256) incidence_weights_kernel(int64_t m, const int *__restrict__ rowptr, const int *__restrict__ members,
                                                                const double *__restrict__ wts, double *__restrict__ weights,
                                                                int *__restrict__ empty) {
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= m) return;
    const int a = rowptr[j], b = rowptr[j + 1];
    if (a == b) { if (lane == 0) atomicAdd(empty, 1); return; }
    double s = 0.0;
    for (int k = a + lane; k < b; k += 32) s += wts ? wts[members[k]] : 1.0;
    s = warp_sum(s);
    for (int k = a + lane; k < b; k += 32) weights[k] = (wts ? wts[members[k]] : 1.0) / s;
}

extern "C" int frk_polygon_incidence_count(frk_ctx *ctx, const frk_polygons *P, int64_t n, const double *d_xy, int64_t ld_xy,
                                           int *d_ptoff, int64_t *h_npairs) {
    FRK_REQUIRE(ctx && P && d_xy && d_ptoff && h_npairs && n > 0 && n < 2147483647LL, "frk_polygon_incidence_count: bad argument");
    int *cnt = nullptr;
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&cnt, n * sizeof(int), ctx->stream));
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, polygon_incidence_kernel<false>, grid, 256, 0, n, d_xy, ld_xy, P->x0, P->y0, P->hx, P->hy, P->nx, P->ny, P->cell_start,
               P->cand, P->ptr, P->vx, P->vy, P->bbox, cnt, (int *)nullptr, (int *)nullptr, (int *)nullptr);
    const int rc = frk_exclusive_scan_i32(ctx, cnt, n, d_ptoff, h_npairs);
    cudaFreeAsync(cnt, ctx->stream);
    return rc;
}

extern "C" int frk_polygon_incidence_fill(frk_ctx *ctx, const frk_polygons *P, int64_t n, const double *d_xy, int64_t ld_xy,
                                          const int *d_ptoff, int64_t npairs, const double *d_wts, int *d_rowptr, int *d_members,
                                          double *d_weights, int64_t *h_empty) {
    FRK_REQUIRE(ctx && P && d_xy && d_ptoff && d_rowptr && d_members && d_weights && h_empty && n > 0 && npairs >= 0 && npairs < 2147483647LL,
                "frk_polygon_incidence_fill: bad argument");
    const int m = P->npoly;
    const int64_t np1 = std::max<int64_t>(npairs, 1);
    int *buf = nullptr;
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&buf, (size_t)(3 * np1 + m + 1) * sizeof(int), ctx->stream));
    int *pid = buf, *pid2 = pid + np1, *pt = pid2 + np1, *pcnt = pt + np1;
    FRK_CHECK_CUDA(cudaMemsetAsync(pcnt, 0, (size_t)(m + 1) * sizeof(int), ctx->stream));
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, polygon_incidence_kernel<true>, grid, 256, 0, n, d_xy, ld_xy, P->x0, P->y0, P->hx, P->hy, P->nx, P->ny, P->cell_start,
               P->cand, P->ptr, P->vx, P->vy, P->bbox, (int *)d_ptoff, pid, pt, pcnt);
    int rc = 0;
    void *tmp = nullptr;
    if (npairs > 0) {      // pairs were emitted in ascending point order: a stable sort by polygon keeps it inside every row
        int end_bit = 1;
        while ((1LL << end_bit) < m) end_bit++;
        size_t tmp_bytes = 0;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, pid, pid2, pt, d_members, (int)npairs, 0, end_bit, ctx->stream);
        if (e == cudaSuccess) e = cudaMallocAsync(&tmp, tmp_bytes, ctx->stream);
        if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, pid, pid2, pt, d_members, (int)npairs, 0, end_bit, ctx->stream);
        if (e != cudaSuccess) { frk_set_error("frk_polygon_incidence_fill: sort failed (%s)", cudaGetErrorString(e)); rc = 1; }
        ctx->launches += 1;
    }
    int64_t tot = 0;
    if (!rc) rc = frk_exclusive_scan_i32(ctx, pcnt, m, d_rowptr, &tot);
    if (!rc) {
        cudaMemsetAsync(pcnt + m, 0, sizeof(int), ctx->stream);
        incidence_weights_kernel<<<(unsigned)cdiv((int64_t)m * 32, 256), 256, 0, ctx->stream>>>(m, d_rowptr, d_members, d_wts, d_weights, pcnt + m);
        ctx->launches += 1;
        int e = 0;
        cudaMemcpyAsync(&e, pcnt + m, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { frk_set_error("frk_polygon_incidence_fill: weights kernel failed"); rc = 1; }
        *h_empty = e;
    }
    if (tmp) cudaFreeAsync(tmp, ctx->stream);
    cudaFreeAsync(buf, ctx->stream);
    return rc;
}

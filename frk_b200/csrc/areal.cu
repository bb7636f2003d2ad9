// Areal (polygon-support) data: BuildC(SpatialPolygons) + normalise_wts, and the aggregations  S = C_Z S0,  X = C_Z X_BAU,
// Vfs = diag(C_Z diag(fs) C_Z')  that SRE() forms from it.
//
// Reference: BuildC(data = "SpatialPolygons") R/geometryfns.R:1572-1616 -- observation i gets the BAUs whose centroid lies in
// its polygon (over(BAU_as_points, this_poly)), weight BAUs$wts; SRE() R/SRE.R:472-502 -- Cmat = sparseMatrix(i, j, x),
// rows divided by their sums (normalise_wts = TRUE, :483-490), S = Cmat %*% S0, Vfs = tcrossprod(Cmat %*% Diagonal(sqrt(fs))),
// X from the BAU-averaged covariates (map_data_to_BAUs(SpatialPolygons), R/geometryfns.R:1328-1380: .safe_mean over the BAUs).
// Device path: footprints that share no BAU (Vfs stays diagonal); the point-in-polygon test itself is frk_polygon_lookup (bau.cu).
#include "common.cuh"
#include <algorithm>
#include <cub/cub.cuh>

__global__ void areal_keys_kernel(int64_t N, const int *__restrict__ group, int m, int *__restrict__ key, int *__restrict__ idx,
                                  int *__restrict__ cnt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        int g = group[i];
        if (g < 0 || g >= m) g = m;                 // "in no footprint" sorts last
        else atomicAdd(cnt + g, 1);
        key[i] = g;
        idx[i] = (int)i;
    }
}

// one warp per footprint: normalise the weights (rows of Cmat divided by their sums); count the empty footprints
__global__ void __launch_bounds__(256) areal_weights_kernel(int64_t m, const int *__restrict__ rowptr, const int *__restrict__ members,
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

extern "C" int frk_areal_incidence(frk_ctx *ctx, int64_t N, const int *d_group, int64_t m, const double *d_wts, int *d_rowptr,
                                   int *d_members, double *d_weights, int64_t *h_nmem, int64_t *h_empty) {
    FRK_REQUIRE(ctx && d_group && d_rowptr && d_members && d_weights && h_nmem && h_empty, "frk_areal_incidence: NULL argument");
    FRK_REQUIRE(N > 0 && m > 0 && N < 2147483647LL && m < 2147483647LL, "frk_areal_incidence: sizes out of range");
    int *key = nullptr, *idx = nullptr, *key2 = nullptr, *cnt = nullptr;
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&key, (3 * N + m + 1) * sizeof(int), ctx->stream));
    idx = key + N; key2 = idx + N; cnt = key2 + N;
    FRK_CHECK_CUDA(cudaMemsetAsync(cnt, 0, (m + 1) * sizeof(int), ctx->stream));
    const unsigned g = (unsigned)std::min<int64_t>(cdiv(N, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, areal_keys_kernel, g, 256, 0, N, d_group, (int)m, key, idx, cnt);
    // stable radix sort by footprint: the BAU rows of a footprint come out in ascending order (j_idx of BuildC)
    size_t tmp_bytes = 0;
    int end_bit = 1;
    while ((1LL << end_bit) <= m) end_bit++;
    void *tmp = nullptr;
    int rc = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key, key2, idx, d_members, (int)N, 0, end_bit, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&tmp, tmp_bytes, ctx->stream);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key, key2, idx, d_members, (int)N, 0, end_bit, ctx->stream);
    if (e != cudaSuccess) { frk_set_error("frk_areal_incidence: sort failed (%s)", cudaGetErrorString(e)); rc = 1; }
    ctx->launches += 1;
    if (!rc) rc = frk_exclusive_scan_i32(ctx, cnt, m, d_rowptr, h_nmem);
    if (!rc) {
        cudaMemsetAsync(cnt + m, 0, sizeof(int), ctx->stream);
        areal_weights_kernel<<<(unsigned)cdiv(m * 32, 256), 256, 0, ctx->stream>>>(m, d_rowptr, d_members, d_wts, d_weights, cnt + m);
        ctx->launches += 1;
        int ne = 0;
        cudaMemcpyAsync(&ne, cnt + m, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { frk_set_error("frk_areal_incidence: weights kernel failed"); rc = 1; }
        *h_empty = ne;
    }
    if (tmp) cudaFreeAsync(tmp, ctx->stream);
    cudaFreeAsync(key, ctx->stream);
    return rc;
}

// ---------------------------------------------------------------------------------
// S = C_Z S0: one CTA per footprint.  The union of the member rows' columns is marked in a shared-memory bitmap; values are
// accumulated member by member (a member row has distinct columns, so a pass is conflict-free and the summation order is
// the member order: deterministic), then compacted in ascending column order.
// ---------------------------------------------------------------------------------
constexpr int AG_T = 256;

__global__ void __launch_bounds__(AG_T) aggregate_count_kernel(int64_t m, int r, const int *__restrict__ crp, const int *__restrict__ mem,
                                                               const int *__restrict__ rp0, const int *__restrict__ col0,
                                                               int *__restrict__ cnt) {
    extern __shared__ unsigned ag_bits[];
    const int nw = (r + 31) / 32;
    __shared__ int total;
    for (int64_t j = blockIdx.x; j < m; j += gridDim.x) {
        for (int w = threadIdx.x; w < nw; w += AG_T) ag_bits[w] = 0u;
        if (threadIdx.x == 0) total = 0;
        __syncthreads();
        for (int k = crp[j]; k < crp[j + 1]; k++) {
            const int row = mem[k];
            for (int e = rp0[row] + threadIdx.x; e < rp0[row + 1]; e += AG_T) {
                const int c = col0[e];
                atomicOr(&ag_bits[c >> 5], 1u << (c & 31));
            }
        }
        __syncthreads();
        int s = 0;
        for (int w = threadIdx.x; w < nw; w += AG_T) s += __popc(ag_bits[w]);
        atomicAdd(&total, s);
        __syncthreads();
        if (threadIdx.x == 0) cnt[j] = total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(AG_T) aggregate_fill_kernel(int64_t m, int r, const int *__restrict__ crp, const int *__restrict__ mem,
                                                              const double *__restrict__ wgt, const int *__restrict__ rp0,
                                                              const int *__restrict__ col0, const double *__restrict__ val0,
                                                              const int *__restrict__ rp, int *__restrict__ col, double *__restrict__ val) {
    extern __shared__ __align__(16) unsigned char ag_raw[];
    double *acc = reinterpret_cast<double *>(ag_raw);
    const int nw = (r + 31) / 32;
    unsigned *bits = reinterpret_cast<unsigned *>(acc + r);
    int *woff = reinterpret_cast<int *>(bits + nw);
    for (int64_t j = blockIdx.x; j < m; j += gridDim.x) {
        for (int c = threadIdx.x; c < r; c += AG_T) acc[c] = 0.0;
        for (int w = threadIdx.x; w < nw; w += AG_T) bits[w] = 0u;
        __syncthreads();
        for (int k = crp[j]; k < crp[j + 1]; k++) {
            const int row = mem[k];
            const double w = wgt[k];
            for (int e = rp0[row] + threadIdx.x; e < rp0[row + 1]; e += AG_T) {
                const int c = col0[e];
                acc[c] += w * val0[e];
                atomicOr(&bits[c >> 5], 1u << (c & 31));
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {                      // offsets of the bitmap words (r/32 <= ~800 words: a serial scan is cheap here)
            int o = 0;
            for (int w = 0; w < nw; w++) { woff[w] = o; o += __popc(bits[w]); }
        }
        __syncthreads();
        const int base = rp[j];
        for (int w = threadIdx.x; w < nw; w += AG_T) {
            unsigned b = bits[w];
            int o = base + woff[w];
            while (b) {
                const int bit = __ffs(b) - 1;
                const int c = w * 32 + bit;
                col[o] = c; val[o] = acc[c]; o++;
                b &= b - 1;
            }
        }
        __syncthreads();
    }
}

extern "C" int frk_csr_aggregate_count(frk_ctx *ctx, int64_t m, int r, const int *d_cz_rowptr, const int *d_cz_members,
                                       const int *d_rowptr0, const int *d_col0, int *d_rowptr, int64_t *h_nnz) {
    FRK_REQUIRE(ctx && d_cz_rowptr && d_cz_members && d_rowptr0 && d_col0 && d_rowptr && m > 0 && r > 0, "frk_csr_aggregate_count: bad argument");
    int *cnt = nullptr;
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&cnt, m * sizeof(int), ctx->stream));
    const size_t smem = (size_t)((r + 31) / 32) * sizeof(unsigned);
    const unsigned grid = (unsigned)std::min<int64_t>(m, (int64_t)ctx->sm_count * 8);
    FRK_LAUNCH(ctx, aggregate_count_kernel, grid, AG_T, smem, m, r, d_cz_rowptr, d_cz_members, d_rowptr0, d_col0, cnt);
    const int rc = frk_exclusive_scan_i32(ctx, cnt, m, d_rowptr, h_nnz);
    cudaFreeAsync(cnt, ctx->stream);
    return rc;
}

extern "C" int frk_csr_aggregate_fill(frk_ctx *ctx, int64_t m, int r, const int *d_cz_rowptr, const int *d_cz_members,
                                      const double *d_cz_weights, const int *d_rowptr0, const int *d_col0, const double *d_val0,
                                      const int *d_rowptr, int *d_col, double *d_val) {
    FRK_REQUIRE(ctx && d_cz_rowptr && d_cz_members && d_cz_weights && d_rowptr0 && d_rowptr && m > 0 && r > 0, "frk_csr_aggregate_fill: bad argument");
    const int nw = (r + 31) / 32;
    const size_t smem = (size_t)r * sizeof(double) + (size_t)nw * (sizeof(unsigned) + sizeof(int));
    FRK_REQUIRE(smem <= 200 * 1024, "frk_csr_aggregate_fill: r = %d basis functions do not fit the shared-memory accumulator (max ~25000)", r);
    static bool attr = false;
    if (!attr) {
        FRK_CHECK_CUDA(cudaFuncSetAttribute(aggregate_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    const unsigned grid = (unsigned)std::min<int64_t>(m, (int64_t)ctx->sm_count * 4);
    FRK_LAUNCH(ctx, aggregate_fill_kernel, grid, AG_T, smem, m, r, d_cz_rowptr, d_cz_members, d_cz_weights, d_rowptr0, d_col0, d_val0,
               d_rowptr, d_col, d_val);
    return 0;
}

// out[j + m*c] = sum_k w_jk^power in[member_k + ld_in*c]   (power 1: X = C_Z X_BAU; power 2 with in = fs: diag(C_Z diag(fs) C_Z'))
__global__ void __launch_bounds__(256) areal_dense_kernel(int64_t m, const int *__restrict__ crp, const int *__restrict__ mem,
                                                          const double *__restrict__ wgt, int power, int ncols,
                                                          const double *__restrict__ in, int64_t ld_in, double *__restrict__ out) {
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= m) return;
    for (int c = 0; c < ncols; c++) {
        double s = 0.0;
        for (int k = crp[j] + lane; k < crp[j + 1]; k += 32) {
            const double w = wgt[k];
            s += (power == 2 ? w * w : w) * in[mem[k] + ld_in * c];
        }
        s = warp_sum(s);
        if (lane == 0) out[j + m * c] = s;
    }
}

extern "C" int frk_areal_aggregate(frk_ctx *ctx, int64_t m, const int *d_cz_rowptr, const int *d_cz_members, const double *d_cz_weights,
                                   int power, int ncols, const double *d_in, int64_t ld_in, double *d_out) {
    FRK_REQUIRE(ctx && d_cz_rowptr && d_cz_members && d_cz_weights && d_in && d_out && m > 0 && ncols > 0 && (power == 1 || power == 2),
                "frk_areal_aggregate: bad argument");
    FRK_LAUNCH(ctx, areal_dense_kernel, (unsigned)cdiv(m * 32, 256), 256, 0, m, d_cz_rowptr, d_cz_members, d_cz_weights, power, ncols,
               d_in, ld_in, d_out);
    return 0;
}

// Bucketed row-streaming kernels: the E-step Gram / right-hand side and the sparse quadratic form, with the
// r x r operand staged per bucket in shared memory.
//
// Reference: S'D^-1 S and S'D^-1 (Z - X alpha) of .loglik.ind / .SRE.Estep.ind (R/SREfit.R:194,252,254);
// rowSums((S0 Cov) * S0) of .batch_compute_var (R/SREutils.R:286-300); Omega_diag1 (R/SREfit.R:332-334).
//
// Rows of a basis matrix whose supports overlap the same few basis functions are grouped into a BUCKET
// (key = the row's last column index, i.e. the highest finest-resolution function in its support).  The
// union U of the columns touched by a bucket is small (<= 64 for FRK's local bases: ~31 non-zeros per row,
// union ~42).  frk_csr_bucket_rows() computes, once per matrix, the row permutation, each bucket's column
// list, and for every stored entry its LOCAL column index within the bucket (one byte).  Then one CTA per
// bucket
//   * gram:      accumulates the dense |U| x |U| tile  sum_j w_j s_j s_j'  in registers (rows densified into
//                shared memory, 4x4 register tiles, no atomics) and flushes it once to G with fp64 atomics --
//                ~100x fewer global atomics than one atomic per (row, pair);
//   * quadform:  gathers M[U,U] once into shared memory and evaluates q_i = s_i' M s_i per row from there --
//                the r x r matrix is read |U|^2 times per bucket from L2 instead of k^2 times per row.
// The per-iteration kernels stream val (8 B) + local index (1 B) per non-zero; the 4-byte global column index
// is not read at all.  A bucket whose union exceeds 64 columns (or, for the quadratic form, a row with more
// than 40 non-zeros: tensor-product bases) takes the direct kernels' arithmetic inside the same launch.
#include "common.cuh"
#include "coupled.cuh"
#include <cub/cub.cuh>

constexpr int UMAX = 64;          // columns per bucket handled through shared memory
constexpr int KMAX = 40;          // non-zeros per row handled by the thread-per-row quadratic form
constexpr int HSIZE = 1024;       // open-addressing table for column -> local index (> UMAX + threads: never fills)
constexpr int MAXP_B = 4;
struct AlphaB { double a[MAXP_B]; };
constexpr int QD_LDX = 36;                   // X is stored TRANSPOSED, Xt[col][row]: lane = row scatters are conflict-free, and with a
                                             // stride of 36 both the A fragments and the epilogue reads hit every bank pair at most twice

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------
// chunk metadata: CH rows of a bucket at a time; entries of the chunk are then walked as one flat range so that
// every thread has independent, mostly coalesced loads in flight (rows themselves are scattered by the permutation)
// ---------------------------------------------------------------------------------
template <int CH>
struct ChunkMeta {
    int beg[CH];      // first entry of the row in the CSR arrays
    int off[CH + 1];  // exclusive prefix sum of the row lengths within the chunk
    int row[CH];
};

// all threads; blockDim.x >= CH is NOT required
template <int CH>
__device__ __forceinline__ void chunk_load(ChunkMeta<CH> &C, const int *__restrict__ perm, int c0, int nr, const int *__restrict__ rowptr) {
    for (int rl = threadIdx.x; rl < CH; rl += blockDim.x) {
        int b = 0, k = 0, row = -1;
        if (rl < nr) { row = perm[c0 + rl]; b = rowptr[row]; k = rowptr[row + 1] - b; }
        C.beg[rl] = b; C.row[rl] = row; C.off[rl + 1] = k;
    }
    if (threadIdx.x == 0) C.off[0] = 0;
    __syncthreads();
    if (threadIdx.x < 32) {          // warp 0: inclusive scan over CH lengths (CH <= 256)
        int carry = 0;
        for (int base = 0; base < CH; base += 32) {
            int v = C.off[base + threadIdx.x + 1], inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if ((int)threadIdx.x >= o) inc += t; }
            C.off[base + threadIdx.x + 1] = carry + inc;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncthreads();
}

// row (within the chunk) that owns flat entry g
template <int CH>
__device__ __forceinline__ int chunk_row_of(const ChunkMeta<CH> &C, int g) {
    int lo = 0, hi = CH;             // off[lo] <= g < off[hi]
#pragma unroll
    for (int s = CH / 2; s >= 1; s >>= 1) { const int mid = lo + s; if (mid < hi && C.off[mid] <= g) lo = mid; }
    return lo;
}

// ---------------------------------------------------------------------------------
// bucketing (one-off per matrix): counting sort of rows by key, per-bucket column union, per-entry local index
// ---------------------------------------------------------------------------------
__global__ void bucket_count_kernel(int64_t n, const int *__restrict__ rowptr, const int *__restrict__ col, int *__restrict__ key,
                                    int *__restrict__ cnt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = rowptr[i + 1];
        const int k = e > rowptr[i] ? col[e - 1] : 0;
        key[i] = k;
        atomicAdd(cnt + k, 1);
    }
}

__global__ void iota_kernel(int64_t n, int *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (int)i;
}

constexpr int BU_T = 256, BU_CH = 256;
struct UnionSmem {
    int keys[HSIZE];
    int ids[HSIZE];
    int ucol[UMAX];
    int rank[UMAX], sorted[UMAX];
    int nU;
    ChunkMeta<BU_CH> C;
};

__device__ __forceinline__ unsigned hslot(int c) { return ((unsigned)c * 2654435761u) >> 22; }

__global__ void __launch_bounds__(BU_T) bucket_union_kernel(int nbuckets, const int *__restrict__ bptr, const int *__restrict__ perm,
                                                            const int *__restrict__ rowptr, const int *__restrict__ col,
                                                            int *__restrict__ ucol_out, int *__restrict__ nU_out,
                                                            unsigned char *__restrict__ lid) {
    __shared__ UnionSmem S;
    for (int b = blockIdx.x; b < nbuckets; b += gridDim.x) {
        const int b0 = bptr[b], b1 = bptr[b + 1];
        if (b1 <= b0) { if (threadIdx.x == 0) nU_out[b] = 0; continue; }
        __syncthreads();
        for (int s = threadIdx.x; s < HSIZE; s += BU_T) S.keys[s] = -1;
        if (threadIdx.x == 0) S.nU = 0;
        __syncthreads();
        for (int pass = 0; pass < 2; pass++) {       // pass 0: union; pass 1: local index of every entry
            if (pass == 1 && S.nU > UMAX) break;
            if (pass == 1) {
                // The ids handed out in pass 0 follow the order in which threads won their atomics.  Re-number the union in
                // ascending column order so that the local layout -- hence every summation order downstream -- is the same in every run.
                const int nU0 = S.nU;
                if (threadIdx.x < nU0) {
                    const int c = S.ucol[threadIdx.x];
                    int rk = 0;
                    for (int u = 0; u < nU0; u++) rk += S.ucol[u] < c;
                    S.rank[threadIdx.x] = rk;
                    S.sorted[rk] = c;
                }
                __syncthreads();
                for (int s2 = threadIdx.x; s2 < HSIZE; s2 += BU_T)
                    if (S.keys[s2] != -1) S.ids[s2] = S.rank[S.ids[s2]];
                if (threadIdx.x < nU0) S.ucol[threadIdx.x] = S.sorted[threadIdx.x];
                __syncthreads();
            }
            for (int c0 = b0; c0 < b1; c0 += BU_CH) {
                const int nr = min(BU_CH, b1 - c0);
                __syncthreads();
                chunk_load<BU_CH>(S.C, perm, c0, nr, rowptr);
                const int total = S.C.off[BU_CH];
                for (int g = threadIdx.x; g < total; g += BU_T) {
                    const int rl = chunk_row_of<BU_CH>(S.C, g);
                    const int e = S.C.beg[rl] + g - S.C.off[rl];
                    const int c = col[e];
                    unsigned s = hslot(c);
                    if (pass == 0) {
                        while (true) {
                            if (*(volatile int *)&S.nU > UMAX) break;                 // overflow: direct path for this bucket
                            int cur = *(volatile int *)&S.keys[s];
                            if (cur == c) break;                                      // common case: no atomic
                            if (cur == -1) {
                                cur = atomicCAS(&S.keys[s], -1, c);
                                if (cur == -1) {
                                    const int id = atomicAdd(&S.nU, 1);
                                    if (id < UMAX) { S.ucol[id] = c; S.ids[s] = id; }
                                    break;
                                }
                                if (cur == c) break;
                            }
                            s = (s + 1) & (HSIZE - 1);
                        }
                    } else {
                        while (S.keys[s] != c) s = (s + 1) & (HSIZE - 1);
                        lid[e] = (unsigned char)S.ids[s];
                    }
                }
            }
            __syncthreads();
        }
        const int nU = S.nU;
        if (threadIdx.x == 0) nU_out[b] = nU;
        if (nU <= UMAX && threadIdx.x < nU) ucol_out[(size_t)b * UMAX + threadIdx.x] = S.ucol[threadIdx.x];
    }
}

extern "C" int frk_csr_bucket_rows(frk_ctx *ctx, int64_t n, const int *d_rowptr, const int *d_col, int ncols, int *d_perm,
                                   int *d_bptr, int *d_ucol, int *d_nU, unsigned char *d_lid) {
    FRK_REQUIRE(ctx && d_rowptr && d_perm && d_bptr && d_ucol && d_nU && d_lid && ncols > 0 && n >= 0 && n < 2147483647LL,
                "frk_csr_bucket_rows: bad argument");
    if (n == 0) {
        FRK_CHECK_CUDA(cudaMemsetAsync(d_bptr, 0, (size_t)(ncols + 1) * sizeof(int), ctx->stream));
        FRK_CHECK_CUDA(cudaMemsetAsync(d_nU, 0, (size_t)ncols * sizeof(int), ctx->stream));
        return 0;
    }
    int *key = nullptr;
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&key, (size_t)(3 * n + (size_t)ncols) * sizeof(int), ctx->stream));
    int *key2 = key + n, *idx = key2 + n, *cnt = idx + n;
    FRK_CHECK_CUDA(cudaMemsetAsync(cnt, 0, (size_t)ncols * sizeof(int), ctx->stream));
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, bucket_count_kernel, grid, 256, 0, n, d_rowptr, d_col, key, cnt);
    int64_t tot = 0;
    int rc = frk_exclusive_scan_i32(ctx, cnt, ncols, d_bptr, &tot);
    if (!rc) {
        // stable radix sort of the rows by key: within a bucket the rows stay in ascending order, so every per-bucket summation
        // downstream runs in the same order in every run (an atomic scatter would hand out the slots in race order)
        FRK_LAUNCH(ctx, iota_kernel, grid, 256, 0, n, idx);
        int end_bit = 1;
        while ((1LL << end_bit) < ncols) end_bit++;
        size_t tmp_bytes = 0;
        void *tmp = nullptr;
        cudaError_t es = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key, key2, idx, d_perm, (int)n, 0, end_bit, ctx->stream);
        if (es == cudaSuccess) es = cudaMallocAsync(&tmp, tmp_bytes, ctx->stream);
        if (es == cudaSuccess) es = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key, key2, idx, d_perm, (int)n, 0, end_bit, ctx->stream);
        if (tmp) cudaFreeAsync(tmp, ctx->stream);
        if (es != cudaSuccess) {
            cudaFreeAsync(key, ctx->stream);
            frk_set_error("frk_csr_bucket_rows: sort failed (%s)", cudaGetErrorString(es));
            return 1;
        }
        ctx->launches += 1;
        FRK_LAUNCH(ctx, bucket_union_kernel, (unsigned)std::min<int64_t>(ncols, (int64_t)ctx->sm_count * 8), BU_T, 0, ncols, d_bptr,
                   d_perm, d_rowptr, d_col, d_ucol, d_nU, d_lid);
    }
    cudaFreeAsync(key, ctx->stream);
    return rc;
}

// ---------------------------------------------------------------------------------
// frk_rowbuckets: the bucket arrays plus a bucket-ordered, transposed copy of the matrix ("ELL tiles").
// A tile is 32 consecutive rows of a bucket (in perm order) stored entry-major in lane-interleaved vectors:
// val as double2 [K/2][32] (entries 2j, 2j+1 of a row adjacent), lid as uchar4 [K/4][32]; K = the longest row of the
// tile rounded up to a multiple of 4.  Shorter rows are padded with val = 0 and the lid of their LAST real entry: a
// consumer that scatters from the last entry to the first needs no test for padding.  A warp reads a tile with fully
// coalesced 16-byte / 4-byte loads -- no per-row pointer chasing in the per-iteration kernels.
// ---------------------------------------------------------------------------------
struct frk_rowbuckets {
    frk_ctx *ctx = nullptr;
    int64_t n = 0;
    int ncols = 0;
    int *perm = nullptr, *bptr = nullptr, *ucol = nullptr, *nU = nullptr;
    unsigned char *lid = nullptr;
    // ELL tiles
    // implicit: dense rows that are never stored (frk_rowbuckets_create_implicit): row i = eval_basis(ibasis, icoords[i, ]), evaluated
    // chunk by chunk inside the Gram / quadratic-form calls
    bool implicit = false;
    const frk_basis *ibasis = nullptr;
    const double *icoords = nullptr;
    int64_t ild = 0;
    int inorm = 0;
    bool dense = false;            // every row holds all ncols columns (non-compact basis functions): val IS the dense matrix, row-major;
                                   // the Gram and the quadratic form are then plain GEMMs and no buckets are built
    int T = 0;
    int64_t nnz = 0;
    int max_nU = 0;                // widest column set of a non-empty bucket
    int nblk_hist[10] = {0};       // non-empty buckets by (nU + 8) / 8 = 8-column blocks incl. the Gram kernel's virtual column
    bool all_ell = false;          // every non-empty bucket has <= UMAX columns and every tile an ELL copy (no direct-path rows)
    int *btile = nullptr;          // [ncols+1] first tile of every bucket
    int *tile_K = nullptr;         // [T] entries per row in the tile; -1: a row is longer than KMAX (direct path)
    int *tile_off = nullptr;       // [T+1] offset of the tile in ell_val / ell_lid (entries)
    double *ell_val = nullptr;
    unsigned char *ell_lid = nullptr;
    // deterministic Gram (frk_use_deterministic): per-bucket tile staging and the inverted index column -> (bucket, local id),
    // built on first use
    mutable double *stage = nullptr, *stage_rhs = nullptr, *stage_scal = nullptr;
    mutable int *inv_ptr = nullptr, *inv_bucket = nullptr, *inv_lid = nullptr;
    mutable std::vector<void *> allocs;
};

__global__ void ell_tiles_per_bucket_kernel(int ncols, const int *__restrict__ bptr, const int *__restrict__ nU, int *__restrict__ cnt) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < ncols) cnt[b] = nU[b] > UMAX ? 0 : (bptr[b + 1] - bptr[b] + 31) / 32;
}

// one warp per bucket-tile: K = longest row
__global__ void __launch_bounds__(256) ell_tile_K_kernel(int ncols, const int *__restrict__ bptr, const int *__restrict__ btile,
                                                         const int *__restrict__ perm, const int *__restrict__ rowptr,
                                                         int *__restrict__ tile_K, int *__restrict__ tile_len) {
    const int lane = threadIdx.x & 31;
    for (int b = blockIdx.x; b < ncols; b += gridDim.x) {
        const int t0 = btile[b], t1 = btile[b + 1], r0 = bptr[b], r1 = bptr[b + 1];
        for (int t = t0 + (threadIdx.x >> 5); t < t1; t += blockDim.x >> 5) {
            const int p = r0 + (t - t0) * 32 + lane;
            int k = 0;
            if (p < r1) { const int row = perm[p]; k = rowptr[row + 1] - rowptr[row]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) k = max(k, __shfl_xor_sync(0xffffffffu, k, o));
            k = (k + 3) & ~3;
            if (lane == 0) { tile_K[t] = k > KMAX ? -1 : k; tile_len[t] = k > KMAX ? 0 : k * 32; }
        }
    }
}

// number of buckets / tiles that need the direct path (bucket wider than UMAX columns, row longer than KMAX)
__global__ void ell_direct_count_kernel(int ncols, const int *__restrict__ bptr, const int *__restrict__ nU, int T, const int *__restrict__ tile_K,
                                        int *__restrict__ out) {
    int bad = 0, mx = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ncols; i += gridDim.x * blockDim.x)
        if (bptr[i + 1] > bptr[i]) { bad += nU[i] > UMAX; mx = max(mx, nU[i]); atomicAdd(out + 2 + min((nU[i] + 8) >> 3, 9), 1); }
    if (mx) atomicMax(out + 1, mx);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x) bad += tile_K[i] < 0;
    if (bad) atomicAdd(out, bad);
}

__global__ void __launch_bounds__(256) ell_fill_kernel(int ncols, const int *__restrict__ bptr, const int *__restrict__ btile,
                                                       const int *__restrict__ perm, const int *__restrict__ rowptr,
                                                       const double *__restrict__ val, const unsigned char *__restrict__ lid,
                                                       const int *__restrict__ tile_K, const int *__restrict__ tile_off,
                                                       double *__restrict__ ell_val, unsigned char *__restrict__ ell_lid) {
    const int lane = threadIdx.x & 31;
    for (int b = blockIdx.x; b < ncols; b += gridDim.x) {
        const int t0 = btile[b], t1 = btile[b + 1], r0 = bptr[b], r1 = bptr[b + 1];
        for (int t = t0 + (threadIdx.x >> 5); t < t1; t += blockDim.x >> 5) {
            const int K = tile_K[t];
            if (K <= 0) continue;
            const int p = r0 + (t - t0) * 32 + lane;
            int beg = 0, k = 0;
            if (p < r1) { const int row = perm[p]; beg = rowptr[row]; k = rowptr[row + 1] - beg; }
            const size_t off = (size_t)tile_off[t];
            const unsigned char padl = k > 0 ? lid[beg + k - 1] : (unsigned char)0;
            for (int e = 0; e < K; e++) {
                ell_val[off + (size_t)(e >> 1) * 64 + lane * 2 + (e & 1)] = e < k ? val[beg + e] : 0.0;
                ell_lid[off + (size_t)(e >> 2) * 128 + lane * 4 + (e & 3)] = e < k ? lid[beg + e] : padl;
            }
        }
    }
}

template <typename T>
static int rb_alloc(frk_rowbuckets *B, T **out, size_t n) {
    void *p = nullptr;
    FRK_CHECK_CUDA(cudaMallocAsync(&p, std::max<size_t>(n * sizeof(T), 16), B->ctx->stream));   // pooled
    B->allocs.push_back(p);
    *out = (T *)p;
    return 0;
}

extern "C" int frk_rowbuckets_destroy(frk_rowbuckets *B) {
    if (!B) return 0;
    cudaSetDevice(B->ctx->device);
    cudaStreamSynchronize(B->ctx->stream);
    for (void *p : B->allocs) cudaFreeAsync(p, B->ctx->stream);
    delete B;
    return 0;
}

extern "C" int frk_rowbuckets_create(frk_ctx *ctx, int64_t n, int ncols, const int *d_rowptr, const int *d_col, const double *d_val,
                                     frk_rowbuckets **out) {
    FRK_REQUIRE(ctx && out && d_rowptr && ncols > 0 && n >= 0, "frk_rowbuckets_create: bad argument");
    frk_rowbuckets *B = new frk_rowbuckets();
    B->ctx = ctx; B->n = n; B->ncols = ncols;
    auto fail = [&](int rc) { frk_rowbuckets_destroy(B); return rc; };
    int nnz = 0;
    if (int rc = frk_d2h(ctx, &nnz, d_rowptr + n, sizeof(int))) return fail(rc);
    B->nnz = nnz;
    if (n > 0 && ncols > UMAX && (int64_t)nnz == n * (int64_t)ncols) {   // rows have ascending unique columns < ncols: all of them present
        B->dense = true;
        *out = B;
        return 0;
    }
    if (rb_alloc(B, &B->perm, (size_t)std::max<int64_t>(n, 1)) || rb_alloc(B, &B->bptr, (size_t)ncols + 1) ||
        rb_alloc(B, &B->ucol, (size_t)ncols * UMAX) || rb_alloc(B, &B->nU, (size_t)ncols) || rb_alloc(B, &B->lid, (size_t)std::max(nnz, 1)) ||
        rb_alloc(B, &B->btile, (size_t)ncols + 1))
        return fail(1);
    if (int rc = frk_csr_bucket_rows(ctx, n, d_rowptr, d_col, ncols, B->perm, B->bptr, B->ucol, B->nU, B->lid)) return fail(rc);
    if (n == 0) { cudaMemsetAsync(B->btile, 0, (size_t)(ncols + 1) * sizeof(int), ctx->stream); *out = B; return 0; }
    // tiles per bucket -> btile
    int *cnt = nullptr;
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&cnt, (size_t)ncols * sizeof(int), ctx->stream));
    FRK_LAUNCH(ctx, ell_tiles_per_bucket_kernel, (unsigned)cdiv(ncols, 256), 256, 0, ncols, B->bptr, B->nU, cnt);
    int64_t T = 0;
    int rc = frk_exclusive_scan_i32(ctx, cnt, ncols, B->btile, &T);
    cudaFreeAsync(cnt, ctx->stream);
    if (rc) return fail(rc);
    B->T = (int)T;
    if (T == 0) { *out = B; return 0; }
    int *tile_len = nullptr;
    if (rb_alloc(B, &B->tile_K, (size_t)T) || rb_alloc(B, &B->tile_off, (size_t)T + 1)) return fail(1);
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&tile_len, (size_t)T * sizeof(int), ctx->stream));
    const unsigned grid = (unsigned)std::min<int64_t>(ncols, (int64_t)ctx->sm_count * 8);
    FRK_LAUNCH(ctx, ell_tile_K_kernel, grid, 256, 0, ncols, B->bptr, B->btile, B->perm, d_rowptr, B->tile_K, tile_len);
    int64_t total = 0;
    rc = frk_exclusive_scan_i32(ctx, tile_len, T, B->tile_off, &total);
    cudaFreeAsync(tile_len, ctx->stream);
    if (rc) return fail(rc);
    if (rb_alloc(B, &B->ell_val, (size_t)std::max<int64_t>(total, 1)) || rb_alloc(B, &B->ell_lid, (size_t)std::max<int64_t>(total, 1) + 16))
        return fail(1);
    FRK_LAUNCH(ctx, ell_fill_kernel, grid, 256, 0, ncols, B->bptr, B->btile, B->perm, d_rowptr, d_val, B->lid, B->tile_K, B->tile_off,
               B->ell_val, B->ell_lid);
    {
        int *bad = nullptr, h_bad[12] = {1, 0};
        FRK_CHECK_CUDA(cudaMallocAsync((void **)&bad, 12 * sizeof(int), ctx->stream));
        FRK_CHECK_CUDA(cudaMemsetAsync(bad, 0, 12 * sizeof(int), ctx->stream));
        FRK_LAUNCH(ctx, ell_direct_count_kernel, (unsigned)std::min<int64_t>(cdiv(std::max<int64_t>(ncols, T), 256), 1024), 256, 0, ncols, B->bptr,
                   B->nU, (int)T, B->tile_K, bad);
        rc = frk_d2h(ctx, h_bad, bad, 12 * sizeof(int));
        cudaFreeAsync(bad, ctx->stream);
        if (rc) return fail(rc);
        B->all_ell = h_bad[0] == 0;
        B->max_nU = h_bad[1];
        for (int k = 0; k < 10; k++) B->nblk_hist[k] = h_bad[2 + k];
    }
    *out = B;
    return 0;
}

extern "C" int frk_rowbuckets_create_implicit(frk_ctx *ctx, int64_t n, const frk_basis *basis, const double *d_coords, int64_t ld_coords,
                                              int normalise, frk_rowbuckets **out) {
    FRK_REQUIRE(ctx && basis && out && n >= 0 && (n == 0 || d_coords) && ld_coords >= n, "frk_rowbuckets_create_implicit: bad argument");
    frk_rowbuckets *B = new frk_rowbuckets();
    B->ctx = ctx; B->n = n; B->ncols = frk_basis_n(basis); B->nnz = n * (int64_t)B->ncols;
    B->dense = true; B->implicit = true; B->ibasis = basis; B->icoords = d_coords; B->ild = ld_coords; B->inorm = normalise;
    *out = B;
    return 0;
}

// same pattern, new values (coupled.cu re-whitens S whenever sigma2fs changes): only the ELL copy holds values
int frk_rowbuckets_refresh(frk_rowbuckets *B, frk_ctx *ctx, const int *d_rowptr, const double *d_val) {
    FRK_REQUIRE(B && ctx && d_rowptr && d_val, "frk_rowbuckets_refresh: NULL argument");
    if (B->dense || B->T == 0 || B->n == 0) return 0;
    const unsigned grid = (unsigned)std::min<int64_t>(B->ncols, (int64_t)ctx->sm_count * 8);
    FRK_LAUNCH(ctx, ell_fill_kernel, grid, 256, 0, B->ncols, B->bptr, B->btile, B->perm, d_rowptr, d_val, B->lid, B->tile_K, B->tile_off,
               B->ell_val, B->ell_lid);
    return 0;
}

// ---------------------------------------------------------------------------------
// Gram + rhs
// ---------------------------------------------------------------------------------
constexpr int GR_T = 256;         // 16 x 16 threads, 4 x 4 register tile each -> 64 x 64
constexpr int GR_ROWS = 64;       // rows densified per pass
constexpr int GR_LD = UMAX + 1;

struct GramSmem {
    ChunkMeta<GR_ROWS> C;
    int ucol[UMAX];
    double A[GR_ROWS][GR_LD];     // densified rows
    double w[GR_ROWS], wr[GR_ROWS];
    double red[GR_T / 32][2];
};

__global__ void __launch_bounds__(GR_T) gram_bucket_kernel(int nbuckets, const int *__restrict__ bptr, const int *__restrict__ perm,
                                                           const int *__restrict__ ucol_all, const int *__restrict__ nU_all,
                                                           const unsigned char *__restrict__ lid, int64_t m, int r,
                                                           const int *__restrict__ rowptr, const int *__restrict__ col,
                                                           const double *__restrict__ val, const double *__restrict__ z,
                                                           const double *__restrict__ X, int p, AlphaB alpha,
                                                           const double *__restrict__ ve, const double *__restrict__ vfs, double sigma2fs,
                                                           double *__restrict__ acc, double *__restrict__ stage,
                                                           double *__restrict__ stage_rhs, double *__restrict__ stage_scal) {
    __shared__ GramSmem S;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double *G = acc, *rhs = acc + (size_t)r * r;
    double s_quad = 0.0, s_logd = 0.0;

    for (int b = blockIdx.x; b < nbuckets; b += gridDim.x) {
        const int b0 = bptr[b], b1 = bptr[b + 1];
        if (b1 <= b0) continue;
        const int nU = nU_all[b];
        if (nU > UMAX) {
            // direct path: one warp per row, one fp64 atomic per (row, pair) -- also in deterministic mode (frk_use_deterministic
            // fixes the order of the staged buckets only; rows wider than UMAX columns are summed in atomic order)
            for (int pr = b0 + wid; pr < b1; pr += GR_T / 32) {
                const int row = perm[pr], beg = rowptr[row], k = rowptr[row + 1] - beg;
                double rho = z[row];
                for (int q = 0; q < p; q++) rho -= X[row + (int64_t)q * m] * alpha.a[q];
                const double d = sigma2fs * vfs[row] + ve[row], w = 1.0 / d;
                if (lane == 0) { s_quad += rho * rho * w; s_logd += log(d); }
                for (int t = lane; t < k; t += 32) atomicAdd(rhs + col[beg + t], w * rho * val[beg + t]);
                for (int a = 0; a < k; a++) {
                    const int ca = col[beg + a];
                    const double wa = w * val[beg + a];
                    for (int bb = a + lane; bb < k; bb += 32) {     // upper triangle whatever the column order of the caller's CSR
                        const int cb = col[beg + bb];
                        atomicAdd(G + min(ca, cb) + (size_t)r * max(ca, cb), wa * val[beg + bb]);
                    }
                }
            }
            continue;
        }
        __syncthreads();
        if (threadIdx.x < nU) S.ucol[threadIdx.x] = ucol_all[(size_t)b * UMAX + threadIdx.x];
        double t[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) t[i][j] = 0.0;
        double rh = 0.0;                                           // thread u < nU: rhs entry of local column u
        const bool active = (ty <= tx) && (ty * 4 < nU) && (tx * 4 < nU);
        for (int c0 = b0; c0 < b1; c0 += GR_ROWS) {
            const int nr = min(GR_ROWS, b1 - c0);
            __syncthreads();
            for (int idx = threadIdx.x; idx < GR_ROWS * GR_LD; idx += GR_T) (&S.A[0][0])[idx] = 0.0;
            chunk_load<GR_ROWS>(S.C, perm, c0, nr, rowptr);        // (contains the barriers that order the zero fill)
            if (threadIdx.x < nr) {                                // per-row weights
                const int row = S.C.row[threadIdx.x];
                double rho = z[row];
                for (int q = 0; q < p; q++) rho -= X[row + (int64_t)q * m] * alpha.a[q];
                const double d = sigma2fs * vfs[row] + ve[row], w = 1.0 / d;
                S.w[threadIdx.x] = w; S.wr[threadIdx.x] = w * rho;
                s_quad += rho * rho * w; s_logd += log(d);
            }
            const int total = S.C.off[GR_ROWS];
            for (int g0 = threadIdx.x; g0 < total; g0 += 4 * GR_T) {   // densify: flat walk over the chunk's entries
                int rl[4];
                double v[4];
                unsigned char l[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int g = g0 + u * GR_T;
                    rl[u] = -1;
                    if (g < total) {
                        rl[u] = chunk_row_of<GR_ROWS>(S.C, g);
                        const int e = S.C.beg[rl[u]] + g - S.C.off[rl[u]];
                        v[u] = val[e]; l[u] = lid[e];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (rl[u] >= 0) S.A[rl[u]][l[u]] = v[u];
            }
            __syncthreads();
            if (active) {
#pragma unroll 2
                for (int rl = 0; rl < nr; rl++) {
                    const double w = S.w[rl];
                    double a[4], bb[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) a[i] = S.A[rl][ty * 4 + i];
#pragma unroll
                    for (int j = 0; j < 4; j++) bb[j] = S.A[rl][tx * 4 + j] * w;
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) t[i][j] += a[i] * bb[j];
                }
            }
            if (threadIdx.x < nU) {
                for (int rl = 0; rl < nr; rl++) rh += S.A[rl][threadIdx.x] * S.wr[rl];
            }
        }
        // flush: each unordered pair of local columns once, into the upper triangle (row <= col) of G -- or, in deterministic mode,
        // into this bucket's staged tile (local pair (la <= lb)), which gram_ordered_flush_kernel adds in a fixed order
        if (active) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int la = ty * 4 + i, lb = tx * 4 + j;
                    if (la < nU && lb < nU && (ty < tx || la <= lb)) {
                        if (stage) stage[((size_t)b * UMAX + la) * UMAX + lb] = t[i][j];
                        else if (t[i][j] != 0.0) {
                            const int ca = S.ucol[la], cb = S.ucol[lb];
                            atomicAdd(G + min(ca, cb) + (size_t)r * max(ca, cb), t[i][j]);
                        }
                    }
                }
            }
        }
        if (threadIdx.x < nU) {
            if (stage) stage_rhs[(size_t)b * UMAX + threadIdx.x] = rh;
            else atomicAdd(rhs + S.ucol[threadIdx.x], rh);
        }
    }
    s_quad = warp_sum(s_quad);
    s_logd = warp_sum(s_logd);
    __syncthreads();
    if (lane == 0) { S.red[wid][0] = s_quad; S.red[wid][1] = s_logd; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, bsum = 0;
        for (int i = 0; i < GR_T / 32; i++) { a += S.red[i][0]; bsum += S.red[i][1]; }
        if (stage) { stage_scal[2 * blockIdx.x] = a; stage_scal[2 * blockIdx.x + 1] = bsum; }
        else { atomicAdd(rhs + r, a); atomicAdd(rhs + r + 1, bsum); }
    }
}


// ---------------------------------------------------------------------------------
// Gram + rhs, second generation: the ELL tiles on the fp64 tensor cores.
// One CTA per bucket (grid-stride), the bucket's 32-row tiles dealt to its four warps.  For every tile the warp densifies the rows
// into Xt[column][row] (the quadratic form's layout) and accumulates the upper block-triangle of  X' diag(w) X  with DMMA m8n8k4:
// the A fragment of block row mb and the B fragment of block column nb are the SAME shared-memory values (Xt[8 b + g][4 ks + tg]),
// the B side multiplied by the row weight.  A virtual column holding rho (the residual) sits right after the bucket's columns, so
// S'D^-1 rho and sum(rho^2 / d) come out of the same product as one more column.  The local |U| x |U| tile lives in registers for the
// whole bucket; the four warps' copies are summed through shared memory and flushed once (fp64 atomics into the upper triangle of G,
// or the staged tile in deterministic mode).
// ---------------------------------------------------------------------------------
constexpr int GE_WARPS = 4, GE_T = GE_WARPS * 32;
constexpr int GE_MAXBLK = 7;                 // 8-column blocks held in registers per bucket (28 accumulator fragments): column sets of up to 55
constexpr int GE_UP = GE_MAXBLK * 8;

struct GramEllSmem {
    double Xt[GE_WARPS][GE_UP][QD_LDX];
    double w[GE_WARPS][32];
    int ucol[UMAX];
};

template <int NBLK>
__device__ __forceinline__ void gram_ell_bucket(GramEllSmem &S, int wid, int lane, int b, int r0, int r1, int nU, const int *__restrict__ perm,
                                                const int *__restrict__ ucol_all, int t0, int t1, const int *__restrict__ tile_K,
                                                const int *__restrict__ tile_off, const double *__restrict__ ell_val,
                                                const unsigned char *__restrict__ ell_lid, int64_t m, int r, const double *__restrict__ z,
                                                const double *__restrict__ X, int p, const AlphaB &alpha, const double *__restrict__ ve,
                                                const double *__restrict__ vfs, double sigma2fs, double *__restrict__ acc_out,
                                                double *__restrict__ stage, double *__restrict__ stage_rhs, double &s_quad, double &s_logd) {
    constexpr int NACC = NBLK * (NBLK + 1) / 2;
    const int g = lane >> 2, tg = lane & 3;
    double (*Xt)[QD_LDX] = S.Xt[wid];
    double *sw = S.w[wid];
    int *ucol = S.ucol;
    __syncthreads();                                           // the previous bucket's flush is done with Xt and ucol
    if ((int)threadIdx.x < nU) ucol[threadIdx.x] = ucol_all[(size_t)b * UMAX + threadIdx.x];
    double acc[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i][0] = acc[i][1] = 0.0;
    for (int tl = t0 + wid; tl < t1; tl += GE_WARPS) {         // the bucket's tiles are dealt to the CTA's warps
        const int K = tile_K[tl];
        const int p0 = r0 + (tl - t0) * 32, nrows = min(32, r1 - p0);
        const int row = lane < nrows ? perm[p0 + lane] : -1;
        const size_t off = (size_t)tile_off[tl];
        if (tl + GE_WARPS < t1) {                              // pull this warp's next tile towards L2
            const size_t noff = (size_t)tile_off[tl + GE_WARPS];
            const int nK = max(tile_K[tl + GE_WARPS], 0);
            const char *pv = reinterpret_cast<const char *>(ell_val + noff);
            for (int c = lane; c < nK * 2; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pv + (size_t)c * 128));
        }
        const double2 *pv = reinterpret_cast<const double2 *>(ell_val + off) + lane;
        const uchar4 *pl = reinterpret_cast<const uchar4 *>(ell_lid + off) + lane;
        double rho = 0.0, w = 0.0;
        if (row >= 0) {
            rho = z[row];
            for (int q = 0; q < p; q++) rho -= X[row + (int64_t)q * m] * alpha.a[q];
            const double d = sigma2fs * vfs[row] + ve[row];
            w = 1.0 / d;
            s_logd += log(d);
        }
        for (int idx = lane; idx < NBLK * 8 * 16; idx += 32) {
            const int cc = idx >> 4, rr = (idx & 15) * 2;
            *reinterpret_cast<double2 *>(&Xt[cc][rr]) = make_double2(0.0, 0.0);
        }
        __syncwarp();
        double *xs = &Xt[0][lane];
        // entries in rounds of 16 per row (the accumulators leave room for 40 registers of tile data, not for 100); the LAST round first:
        // padding (val 0 at the row's last real column) must be overwritten by the real entry
        for (int j0 = ((K + 15) >> 4) * 4 - 4; j0 >= 0; j0 -= 4) {
            double2 v[8];
            uchar4 l[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (4 * (j0 + j) < K) { v[2 * j] = pv[(2 * (j0 + j)) * 32]; v[2 * j + 1] = pv[(2 * (j0 + j) + 1) * 32]; l[j] = pl[(j0 + j) * 32]; }
#pragma unroll
            for (int j = 3; j >= 0; j--)
                if (4 * (j0 + j) < K) {
                    xs[l[j].w * QD_LDX] = v[2 * j + 1].y;
                    xs[l[j].z * QD_LDX] = v[2 * j + 1].x;
                    xs[l[j].y * QD_LDX] = v[2 * j].y;
                    xs[l[j].x * QD_LDX] = v[2 * j].x;
                }
        }
        xs[nU * QD_LDX] = rho;                                 // the virtual column
        sw[lane] = w;
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
            const double wk = sw[4 * ks + tg];
            double a[NBLK];
#pragma unroll
            for (int mb = 0; mb < NBLK; mb++) a[mb] = Xt[8 * mb + g][4 * ks + tg];
#pragma unroll
            for (int nb = 0; nb < NBLK; nb++) {
                const double bf = a[nb] * wk;
#pragma unroll
                for (int mb = 0; mb <= nb; mb++) dmma884(acc[nb * (nb + 1) / 2 + mb][0], acc[nb * (nb + 1) / 2 + mb][1], a[mb], bf);
            }
        }
        __syncwarp();
    }
    // reduce over the warps: every warp parks its fragments in its own (now idle) Xt area, fragment-major; then the CTA's threads
    // each sum one entry over the warps and flush it.  Entry e of fragment f: local row 8 mb + e / 8, local column 8 nb + e % 8.
    double *park = &S.Xt[wid][0][0];
#pragma unroll
    for (int f = 0; f < NACC; f++) {
        park[f * 64 + g * 8 + 2 * tg] = acc[f][0];
        park[f * 64 + g * 8 + 2 * tg + 1] = acc[f][1];
    }
    __syncthreads();
    double *G = acc_out, *rhs = acc_out + (size_t)r * r;
    for (int idx = threadIdx.x; idx < NACC * 64; idx += GE_T) {
        const int f = idx >> 6, e = idx & 63;
        int nb = 0;
        while ((nb + 1) * (nb + 2) / 2 <= f) nb++;
        const int mb = f - nb * (nb + 1) / 2;
        const int la = 8 * mb + (e >> 3), lb = 8 * nb + (e & 7);
        if (la > lb || lb > nU) continue;
        double val = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < GE_WARPS; w2++) val += (&S.Xt[w2][0][0])[idx];
        if (lb == nU) {                                        // the virtual column: right-hand side / sum(rho^2 w)
            if (la == nU) s_quad += val;
            else if (stage) stage_rhs[(size_t)b * UMAX + la] = val;
            else atomicAdd(rhs + ucol[la], val);
        } else if (stage) stage[((size_t)b * UMAX + la) * UMAX + lb] = val;
        else if (val != 0.0) {
            const int ca = ucol[la], cb = ucol[lb];
            atomicAdd(G + min(ca, cb) + (size_t)r * max(ca, cb), val);
        }
    }
}

// NBLK = blocks held in registers; the launch handles the buckets with LO < blocks <= NBLK (one instantiation per kernel: several
// in one kernel make ptxas spill the accumulators)
template <int NBLK>
__global__ void __launch_bounds__(GE_T, 2) gram_ell_kernel(int lo, int nbuckets, const int *__restrict__ bptr, const int *__restrict__ perm,
                                                           const int *__restrict__ ucol_all, const int *__restrict__ nU_all,
                                                           const int *__restrict__ btile, const int *__restrict__ tile_K,
                                                           const int *__restrict__ tile_off, const double *__restrict__ ell_val,
                                                           const unsigned char *__restrict__ ell_lid, int64_t m, int r,
                                                           const double *__restrict__ z, const double *__restrict__ X, int p, AlphaB alpha,
                                                           const double *__restrict__ ve, const double *__restrict__ vfs, double sigma2fs,
                                                           double *__restrict__ acc, double *__restrict__ stage,
                                                           double *__restrict__ stage_rhs, double *__restrict__ stage_scal) {
    extern __shared__ __align__(16) unsigned char gsm_raw[];
    GramEllSmem &S = *reinterpret_cast<GramEllSmem *>(gsm_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double s_quad = 0.0, s_logd = 0.0;
    for (int b = blockIdx.x; b < nbuckets; b += gridDim.x) {
        const int r0 = bptr[b], r1 = bptr[b + 1];
        if (r1 <= r0) continue;
        const int nU = nU_all[b];
        const int nblk = (nU + 8) >> 3;                        // blocks of 8 columns, the virtual column included
        if (nblk <= lo || nblk > NBLK) continue;
        gram_ell_bucket<NBLK>(S, wid, lane, b, r0, r1, nU, perm, ucol_all, btile[b], btile[b + 1], tile_K, tile_off, ell_val, ell_lid, m, r, z, X, p,
                              alpha, ve, vfs, sigma2fs, acc, stage, stage_rhs, s_quad, s_logd);
    }
    // s_quad: the thread that flushed local entry (nU, nU) added it, one per bucket; s_logd: one term per row
    s_quad = warp_sum(s_quad);
    s_logd = warp_sum(s_logd);
    if (lane == 0) {
        const int gw = blockIdx.x * GE_WARPS + wid;
        if (stage) { stage_scal[2 * gw] += s_quad; stage_scal[2 * gw + 1] += s_logd; }      // (zeroed before the first class is launched)
        else { atomicAdd(acc + (size_t)r * r + r, s_quad); atomicAdd(acc + (size_t)r * r + r + 1, s_logd); }
    }
}

template <int NBLK>
static int launch_gram_ell(frk_ctx *ctx, int lo, unsigned grid, int r, const frk_rowbuckets *bk, int64_t m, const double *d_z, const double *d_X,
                           int p, const AlphaB &al, const double *d_ve, const double *d_vfs, double sigma2fs, double *d_acc, double *st,
                           double *st_rhs, double *st_scal, double &bytes) {
    static bool attr = false;
    if (!attr) {
        FRK_CHECK_CUDA(cudaFuncSetAttribute(gram_ell_kernel<NBLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GramEllSmem)));
        attr = true;
    }
    FRK_TAG(ctx, "gram_ell_kernel");                        // the classes of one pass are one kernel for the profile
    FRK_WORK(ctx, 0, bytes);                                // ... whose algorithmic bytes are booked on the first launch
    bytes = 0;
    FRK_LAUNCH(ctx, gram_ell_kernel<NBLK>, grid, GE_T, sizeof(GramEllSmem), lo, r, bk->bptr, bk->perm, bk->ucol, bk->nU, bk->btile, bk->tile_K,
               bk->tile_off, bk->ell_val, bk->ell_lid, m, r, d_z, d_X, p, al, d_ve, d_vfs, sigma2fs, d_acc, st, st_rhs, st_scal);
    return 0;
}


// ---------------------------------------------------------------------------------
// Gram + rhs, third generation (FRK_GRAM_V2 = 1 selects the kernels above for A/B timing).  Same arithmetic -- ELL tiles densified
// into Xt[column][row], upper block-triangle of X' diag(w) X by DMMA, virtual residual column -- but the OUTPUT is split over the
// four warps of the CTA instead of the tiles: warp w owns the column blocks w and 6 - w of the local tile (8, 8, 8 and 4 fragments
// of 8 x 8), all four warps densify every tile together (a quarter of the entries each) into ONE shared copy and multiply from it.
// 14 accumulator fragments per warp instead of 28 (128 registers: 4 CTAs per SM instead of 2, one instantiation for every bucket
// width, one launch), no cross-warp reduction at the end, and a tile's loads are four times shorter.  Measured at C3: 0.32 ms
// against 0.38 ms (a variant that also issued the next tile's loads before the products measured slower, 0.43 ms).
// ---------------------------------------------------------------------------------
struct Gram3Smem {
    double Xt[GE_UP][QD_LDX];
    double w[32];
    int ucol[UMAX];
};

__global__ void __launch_bounds__(GE_T, 4) gram_ell3_kernel(int nbuckets, const int *__restrict__ bptr, const int *__restrict__ perm,
                                                            const int *__restrict__ ucol_all, const int *__restrict__ nU_all,
                                                            const int *__restrict__ btile, const int *__restrict__ tile_K,
                                                            const int *__restrict__ tile_off, const double *__restrict__ ell_val,
                                                            const unsigned char *__restrict__ ell_lid, int64_t m, int r,
                                                            const double *__restrict__ z, const double *__restrict__ X, int p, AlphaB alpha,
                                                            const double *__restrict__ ve, const double *__restrict__ vfs, double sigma2fs,
                                                            double *__restrict__ acc_out, double *__restrict__ stage,
                                                            double *__restrict__ stage_rhs, double *__restrict__ stage_scal) {
    __shared__ __align__(16) Gram3Smem S;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3;
    const int nbA = wid, nbB = GE_MAXBLK - 1 - wid;               // this warp's column blocks (the same block twice for wid = 3)
    double s_quad = 0.0, s_logd = 0.0;
    double *G = acc_out, *rhs = acc_out + (size_t)r * r;
    for (int b = blockIdx.x; b < nbuckets; b += gridDim.x) {
        const int r0 = bptr[b], r1 = bptr[b + 1];
        if (r1 <= r0) continue;
        const int nU = nU_all[b], t0 = btile[b], t1 = btile[b + 1];
        const int nblk = (nU + 8) >> 3;                            // 8-column blocks, the virtual column included (<= GE_MAXBLK)
        __syncthreads();                                           // the previous bucket is done with ucol
        if ((int)threadIdx.x < nU) S.ucol[threadIdx.x] = ucol_all[(size_t)b * UMAX + threadIdx.x];
        double acc[2][GE_MAXBLK][2];
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int i = 0; i < GE_MAXBLK; i++) acc[q][i][0] = acc[q][i][1] = 0.0;
        for (int tl = t0; tl < t1; tl++) {
            const int K = tile_K[tl];
            const int p0 = r0 + (tl - t0) * 32, nrows = min(32, r1 - p0);
            const size_t off = (size_t)tile_off[tl];
            if (tl + 1 < t1) {                                     // pull the next tile towards L2 (each warp a quarter of its lines)
                const size_t noff = (size_t)tile_off[tl + 1];
                const int nK = max(tile_K[tl + 1], 0);
                const char *pn = reinterpret_cast<const char *>(ell_val + noff);
                for (int c = threadIdx.x; c < nK * 2; c += GE_T) asm volatile("prefetch.global.L2 [%0];" ::"l"(pn + (size_t)c * 128));
            }
            // this warp's quarter of the tile: entry groups j = wid, wid + 4, ... (four entries of every row each); lane = row
            const double2 *pv = reinterpret_cast<const double2 *>(ell_val + off) + lane;
            const uchar4 *pl = reinterpret_cast<const uchar4 *>(ell_lid + off) + lane;
            double2 v[(KMAX / 4 + 3) / 4 * 2];
            uchar4 l[(KMAX / 4 + 3) / 4];
#pragma unroll
            for (int q = 0; q < (KMAX / 4 + 3) / 4; q++) {
                const int j = wid + 4 * q;
                if (4 * j < K) { v[2 * q] = pv[(2 * j) * 32]; v[2 * q + 1] = pv[(2 * j + 1) * 32]; l[q] = pl[j * 32]; }
            }
            double rho = 0.0, w = 0.0;
            if (wid == 0) {                                        // row weights and residuals: one warp, lane = row
                const int row = lane < nrows ? perm[p0 + lane] : -1;
                if (row >= 0) {
                    rho = z[row];
                    for (int q = 0; q < p; q++) rho -= X[row + (int64_t)q * m] * alpha.a[q];
                    const double d = sigma2fs * vfs[row] + ve[row];
                    w = 1.0 / d;
                    s_logd += log(d);
                }
            }
            __syncthreads();                                       // everybody has finished multiplying from the previous tile
            for (int idx = threadIdx.x; idx < nblk * 8 * 16; idx += GE_T) {
                const int cc = idx >> 4, rr = (idx & 15) * 2;
                *reinterpret_cast<double2 *>(&S.Xt[cc][rr]) = make_double2(0.0, 0.0);
            }
            __syncthreads();
            // scatter: padding entries (value 0) are skipped, so the order in which the warps write does not matter
            double *xs = &S.Xt[0][lane];
#pragma unroll
            for (int q = 0; q < (KMAX / 4 + 3) / 4; q++) {
                const int j = wid + 4 * q;
                if (4 * j < K) {
                    if (v[2 * q].x != 0.0) xs[l[q].x * QD_LDX] = v[2 * q].x;
                    if (v[2 * q].y != 0.0) xs[l[q].y * QD_LDX] = v[2 * q].y;
                    if (v[2 * q + 1].x != 0.0) xs[l[q].z * QD_LDX] = v[2 * q + 1].x;
                    if (v[2 * q + 1].y != 0.0) xs[l[q].w * QD_LDX] = v[2 * q + 1].y;
                }
            }
            if (wid == 0) { xs[nU * QD_LDX] = rho; S.w[lane] = w; }
            __syncthreads();
#pragma unroll
            for (int ks = 0; ks < 8; ks++) {
                const double wk = S.w[4 * ks + tg];
                double a[GE_MAXBLK];
#pragma unroll
                for (int mb = 0; mb < GE_MAXBLK; mb++) a[mb] = (mb <= nbB && mb < nblk) ? S.Xt[8 * mb + g][4 * ks + tg] : 0.0;
                if (nbA < nblk) {
                    const double bf = S.Xt[8 * nbA + g][4 * ks + tg] * wk;
#pragma unroll
                    for (int mb = 0; mb < GE_MAXBLK; mb++) if (mb <= nbA) dmma884(acc[0][mb][0], acc[0][mb][1], a[mb], bf);
                }
                if (nbB < nblk && nbB != nbA) {
                    const double bf = S.Xt[8 * nbB + g][4 * ks + tg] * wk;
#pragma unroll
                    for (int mb = 0; mb < GE_MAXBLK; mb++) if (mb <= nbB) dmma884(acc[1][mb][0], acc[1][mb][1], a[mb], bf);
                }
            }
        }
        // flush this warp's fragments: local entries (8 mb + g, 8 nb + 2 tg + e)
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int nb = q ? nbB : nbA;
            if (nb >= nblk || (q && nbB == nbA)) continue;
#pragma unroll
            for (int mb = 0; mb < GE_MAXBLK; mb++) {
                if (mb > nb) continue;
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int la = 8 * mb + g, lb = 8 * nb + 2 * tg + e;
                    const double val = acc[q][mb][e];
                    if (la > lb || lb > nU) continue;
                    if (lb == nU) {                                // the virtual column: right-hand side / sum(rho^2 w)
                        if (la == nU) s_quad += val;
                        else if (stage) stage_rhs[(size_t)b * UMAX + la] = val;
                        else atomicAdd(rhs + S.ucol[la], val);
                    } else if (stage) stage[((size_t)b * UMAX + la) * UMAX + lb] = val;
                    else if (val != 0.0) {
                        const int ca = S.ucol[la], cb = S.ucol[lb];
                        atomicAdd(G + min(ca, cb) + (size_t)r * max(ca, cb), val);
                    }
                }
            }
        }
    }
    s_quad = warp_sum(s_quad);
    s_logd = warp_sum(s_logd);
    if (lane == 0) {
        const int gw = blockIdx.x * GE_WARPS + wid;
        if (stage) { stage_scal[2 * gw] = s_quad; stage_scal[2 * gw + 1] = s_logd; }
        else { atomicAdd(acc_out + (size_t)r * r + r, s_quad); atomicAdd(acc_out + (size_t)r * r + r + 1, s_logd); }
    }
}

// Deterministic flush: one warp owns row a of G.  For every bucket k (ascending) whose column set holds a, with local id la, it adds
// the staged pair values (a, cb), cb >= a, and the right-hand-side entry; block 0 sums the per-CTA scalars in CTA order.
__global__ void __launch_bounds__(256) gram_ordered_flush_kernel(int r, const int *__restrict__ inv_ptr, const int *__restrict__ inv_bucket,
                                                                 const int *__restrict__ inv_lid, const int *__restrict__ ucol_all,
                                                                 const int *__restrict__ nU_all, const double *__restrict__ stage,
                                                                 const double *__restrict__ stage_rhs, const double *__restrict__ stage_scal,
                                                                 int ncta, double *__restrict__ acc) {
    const int lane = threadIdx.x & 31;
    const int a = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    double *G = acc, *rhs = acc + (size_t)r * r;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double q = 0.0, l = 0.0;
        for (int c = 0; c < ncta; c++) { q += stage_scal[2 * c]; l += stage_scal[2 * c + 1]; }
        rhs[r] += q; rhs[r + 1] += l;
    }
    if (a >= r) return;
    double rh = 0.0;
    for (int e = inv_ptr[a]; e < inv_ptr[a + 1]; e++) {
        const int k = inv_bucket[e], la = inv_lid[e], nU = nU_all[k];
        const int *uc = ucol_all + (size_t)k * UMAX;
        const double *T = stage + (size_t)k * UMAX * UMAX;
        for (int lb = lane; lb < nU; lb += 32) {
            const int cb = uc[lb];
            if (cb < a) continue;
            const double v = la <= lb ? T[(size_t)la * UMAX + lb] : T[(size_t)lb * UMAX + la];
            G[a + (size_t)r * cb] += v;                            // this warp is the only writer of row a; buckets in ascending order
        }
        rh += stage_rhs[(size_t)k * UMAX + la];
    }
    if (lane == 0) rhs[a] += rh;
}

// staging buffers and the inverted index column -> (bucket, local id) for the ordered flush; buckets with more than UMAX columns
// (direct path: atomics) are not indexed.  Built once per row-bucket set, on the host from the device's column sets.
static int gram_prepare_ordered(frk_ctx *ctx, const frk_rowbuckets *bk, int r) {
    if (bk->inv_ptr) return 0;
    std::vector<int> ucol((size_t)r * UMAX), nU(r);
    FRK_CHECK_CUDA(cudaMemcpyAsync(ucol.data(), bk->ucol, ucol.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FRK_CHECK_CUDA(cudaMemcpyAsync(nU.data(), bk->nU, nU.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<int> ptr(r + 1, 0);
    for (int k = 0; k < r; k++)
        if (nU[k] > 0 && nU[k] <= UMAX) for (int l = 0; l < nU[k]; l++) ptr[ucol[(size_t)k * UMAX + l] + 1]++;
    for (int a = 0; a < r; a++) ptr[a + 1] += ptr[a];
    std::vector<int> ib(std::max(ptr[r], 1)), il(std::max(ptr[r], 1)), fill(ptr.begin(), ptr.end() - 1);
    for (int k = 0; k < r; k++)                                   // ascending k within every column
        if (nU[k] > 0 && nU[k] <= UMAX) for (int l = 0; l < nU[k]; l++) {
            const int a = ucol[(size_t)k * UMAX + l], o = fill[a]++;
            ib[o] = k; il[o] = l;
        }
    auto alloc = [&](void **p, size_t bytes) -> int {
        FRK_CHECK_CUDA(cudaMallocAsync(p, std::max<size_t>(bytes, 16), ctx->stream));
        bk->allocs.push_back(*p);
        return 0;
    };
    const unsigned grid = (unsigned)std::min<int64_t>(r, (int64_t)ctx->sm_count * 8);
    int *d_ptr = nullptr, *d_ib = nullptr, *d_il = nullptr;
    if (alloc((void **)&d_ptr, ptr.size() * sizeof(int)) || alloc((void **)&d_ib, ib.size() * sizeof(int)) ||
        alloc((void **)&d_il, il.size() * sizeof(int)) || alloc((void **)&bk->stage, (size_t)r * UMAX * UMAX * sizeof(double)) ||
        alloc((void **)&bk->stage_rhs, (size_t)r * UMAX * sizeof(double)) || alloc((void **)&bk->stage_scal, 2 * (4 * (size_t)grid + 8) * sizeof(double)))
        return 1;
    FRK_CHECK_CUDA(cudaMemcpyAsync(d_ptr, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    FRK_CHECK_CUDA(cudaMemcpyAsync(d_ib, ib.data(), ib.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    FRK_CHECK_CUDA(cudaMemcpyAsync(d_il, il.data(), il.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));             // the host vectors go out of scope
    bk->inv_bucket = d_ib; bk->inv_lid = d_il;
    bk->inv_ptr = d_ptr;
    return 0;
}

// ---------------------------------------------------------------------------------
// Dense rows (basis functions without compact support: Gaussian, exponential, Matern): the CSR values are the n x r matrix in
// row-major order, i.e. A = S' (r x n, column-major, ld r).  The Gram is then  A diag(w) A'  and the quadratic forms are the
// column dots of A with M A -- GEMMs on the fp64 tensor path instead of r^2/2 atomics per row.  Columns are processed in chunks
// so that the scaled copy / the product stays below ~512 MB.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dense_weights_kernel(int64_t c0, int64_t nc, int r, int64_t m, const double *__restrict__ A,
                                                            const double *__restrict__ z, const double *__restrict__ X, int p, AlphaB alpha,
                                                            const double *__restrict__ ve, const double *__restrict__ vfs, double sigma2fs,
                                                            double *__restrict__ B, double *__restrict__ wr, double *__restrict__ part) {
    // one warp per column (= row of S) of the chunk: B[:, i] = w_i A[:, i], wr_i = w_i rho_i; per-block partials of rho^2 w and log d
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __shared__ double sq[8], sl[8];
    double q = 0.0, l = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 8 + wid; i < nc; i += (int64_t)gridDim.x * 8) {
        const int64_t row = c0 + i;
        double rho = z[row];
        for (int k = 0; k < p; k++) rho -= X[row + (int64_t)k * m] * alpha.a[k];
        const double d = sigma2fs * vfs[row] + ve[row], w = 1.0 / d;
        const double *a = A + (size_t)i * r;                          // A: the chunk's rows
        double *b = B + (size_t)i * r;
        for (int c = lane; c < r; c += 32) b[c] = w * a[c];
        if (lane == 0) { wr[i] = w * rho; q += rho * rho * w; l += log(d); }
    }
    if (lane == 0) { sq[wid] = q; sl[wid] = l; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < 8; w++) { a += sq[w]; b += sl[w]; }
        part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b;
    }
}

__global__ void dense_scalars_kernel(int nblocks, const double *__restrict__ part, double *__restrict__ out2) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double a = 0, b = 0;
        for (int k = 0; k < nblocks; k++) { a += part[2 * k]; b += part[2 * k + 1]; }     // fixed order
        out2[0] += a; out2[1] += b;
    }
}

static int gram_rhs_dense(frk_ctx *ctx, const frk_rowbuckets *bk, int64_t m, int r, const double *A, const double *d_z, const double *d_X, int p,
                          const AlphaB &al, const double *d_ve, const double *d_vfs, double sigma2fs, double *d_acc) {
    // chunk: the scaled copy B (and, for an implicit matrix, the evaluated rows themselves) stays below ~512 MB each
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(m, (int64_t)(512.0 * 1024 * 1024 / (8.0 * r))));
    const unsigned grid = (unsigned)std::min<int64_t>(cdiv(chunk, 8), (int64_t)ctx->sm_count * 8);
    const bool imp = bk->implicit;
    double *B = nullptr;
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&B, ((size_t)chunk * r * (imp ? 2 : 1) + chunk + 2 * (size_t)grid) * sizeof(double), ctx->stream));
    double *wr = B + (size_t)chunk * r, *part = wr + chunk, *Aeval = part + 2 * (size_t)grid;
    double *G = d_acc, *rhs = d_acc + (size_t)r * r;
    int rc = 0;
    for (int64_t c0 = 0; c0 < m && !rc; c0 += chunk) {
        const int64_t nc = std::min(chunk, m - c0);
        const double *Ac = imp ? nullptr : A + (size_t)c0 * r;
        if (imp) {                                                     // evaluate-and-accumulate: the rows exist only for this chunk
            rc = frk_eval_basis_dense_rows(ctx, bk->ibasis, nc, bk->icoords + c0, bk->ild, bk->inorm, Aeval);
            if (rc) break;
            Ac = Aeval;
        }
        do {
            FRK_LAUNCH(ctx, dense_weights_kernel, grid, 256, 0, c0, nc, r, m, Ac, d_z, d_X, p, al, d_ve, d_vfs, sigma2fs, B, wr, part);
            FRK_LAUNCH(ctx, dense_scalars_kernel, 1, 32, 0, (int)grid, part, rhs + r);
        } while (0);
        // G += A_chunk B_chunk'  (r x r, K = nc; symmetric: only the lower triangle is computed, mirrored below);  rhs += A_chunk wr
        rc = frk_dsyrk_like_lower(ctx, r, (int)nc, Ac, r, B, r, G, r);
        if (!rc) rc = frk_dgemm_nt(ctx, r, 1, (int)nc, 1.0, Ac, r, wr, 1, 1.0, rhs, r);
    }
    if (!rc) rc = frk_mirror_lower(ctx, r, G);                         // the accumulator's consumers read the upper triangle
    cudaFreeAsync(B, ctx->stream);
    return rc;
}

// q_i = a_i . y_i (square = 0: Y = M A) or y_i . y_i (square = 1: Y = L^-1 A, M = L^-T L^-1);  t_i = a_i . mu
__global__ void __launch_bounds__(256) dense_coldot_kernel(int64_t nc, int r, const double *__restrict__ A, const double *__restrict__ Y,
                                                           const double *__restrict__ mu, double *__restrict__ q, double *__restrict__ t,
                                                           int square) {
    const int lane = threadIdx.x & 31;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nc; i += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const double *a = A + (size_t)i * r;
        double sq = 0.0, st = 0.0;
        if (Y) {
            const double *y = Y + (size_t)i * r;
            if (square) for (int c = lane; c < r; c += 32) sq += y[c] * y[c];
            else for (int c = lane; c < r; c += 32) sq += a[c] * y[c];
        }
        if (mu) for (int c = lane; c < r; c += 32) st += a[c] * __ldg(mu + c);
        sq = warp_sum(sq); st = warp_sum(st);
        if (lane == 0) { if (q && Y) q[i] = sq; if (t && mu) t[i] = st; }
    }
}

// d_Linv != NULL: M = L^-T L^-1 with the lower-triangular L^-1 given (what frk_potri leaves in its work buffer): q_i = |L^-1 a_i|^2,
// a triangular product -- half the flops of M A
static int quadform_dense(frk_ctx *ctx, const frk_rowbuckets *bk, int64_t n, int r, const double *A, const double *d_M, const double *d_mu,
                          double *d_q, double *d_t, const double *d_Linv = nullptr) {
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)(512.0 * 1024 * 1024 / (8.0 * r))));
    const bool imp = bk->implicit;
    double *Y = nullptr, *Aeval = nullptr;
    if ((d_M || d_Linv) && d_q) FRK_CHECK_CUDA(cudaMallocAsync((void **)&Y, (size_t)chunk * r * sizeof(double), ctx->stream));
    if (imp) FRK_CHECK_CUDA(cudaMallocAsync((void **)&Aeval, (size_t)chunk * r * sizeof(double), ctx->stream));
    int rc = 0;
    for (int64_t c0 = 0; c0 < n && !rc; c0 += chunk) {
        const int64_t nc = std::min(chunk, n - c0);
        const double *Ac = imp ? Aeval : A + (size_t)c0 * r;
        if (imp) { rc = frk_eval_basis_dense_rows(ctx, bk->ibasis, nc, bk->icoords + c0, bk->ild, bk->inorm, Aeval); if (rc) break; }
        // Y = M A_chunk  (M symmetric r x r; A_chunk is K x N with k contiguous), or the triangular Y = L^-1 A_chunk
        if (Y && d_Linv) rc = frk_dtrmm_lower(ctx, r, (int)nc, d_Linv, r, Ac, r, Y, r);
        else if (Y) rc = frk_dgemm(ctx, 0, 1, r, (int)nc, r, 1.0, d_M, r, Ac, r, 0.0, Y, r);
        if (rc) break;
        const unsigned grid = (unsigned)std::min<int64_t>(cdiv(nc * 32, 256), (int64_t)ctx->sm_count * 16);
        do {
            FRK_LAUNCH(ctx, dense_coldot_kernel, grid, 256, 0, nc, r, Ac, (const double *)Y, d_mu, d_q ? d_q + c0 : nullptr, d_t ? d_t + c0 : nullptr,
                       d_Linv ? 1 : 0);
        } while (0);
    }
    if (Y) cudaFreeAsync(Y, ctx->stream);
    if (Aeval) cudaFreeAsync(Aeval, ctx->stream);
    return rc;
}

extern "C" int frk_gram_rhs_bucketed(frk_ctx *ctx, int64_t m, int r, const int *d_rowptr, const int *d_col, const double *d_val,
                                     const frk_rowbuckets *bk, const double *d_z, const double *d_X, int p, const double *h_alpha,
                                     const double *d_ve, const double *d_vfs, double sigma2fs, double *d_acc) {
    FRK_REQUIRE(ctx && d_acc && bk && (d_rowptr || bk->implicit), "frk_gram_rhs_bucketed: NULL argument");
    FRK_REQUIRE(bk->n == m && bk->ncols == r, "frk_gram_rhs_bucketed: the row buckets belong to a %lld x %d matrix", (long long)bk->n, bk->ncols);
    const int *d_perm = bk->perm, *d_bptr = bk->bptr, *d_ucol = bk->ucol, *d_nU = bk->nU;       // (all NULL for a dense matrix)
    const unsigned char *d_lid = bk->lid;
    FRK_REQUIRE(p >= 0 && p <= MAXP_B, "frk_gram_rhs_bucketed: %d fixed effects; the device path handles at most %d", p, MAXP_B);
    if (m == 0) return 0;
    AlphaB al{};
    for (int q = 0; q < p; q++) al.a[q] = h_alpha[q];
    if (bk->dense) return gram_rhs_dense(ctx, bk, m, r, d_val, d_z, d_X, p, al, d_ve, d_vfs, sigma2fs, d_acc);
    unsigned grid = (unsigned)std::min<int64_t>(r, (int64_t)ctx->sm_count * 8);
    static const bool legacy = getenv("FRK_GRAM_LEGACY") != nullptr;      // A/B timing: the first-generation (DFMA register-tile) kernel
    const bool ell = bk->all_ell && bk->T > 0 && bk->max_nU < GE_MAXBLK * 8 && !legacy;
    if (ell) grid = (unsigned)std::min<int64_t>(r, (int64_t)ctx->sm_count * 2);
    // work of the launch: val 8 B + local id 1 B per stored entry, row weights, the r x r accumulator once
    double *st = nullptr, *st_rhs = nullptr, *st_scal = nullptr;
    if (ctx->deterministic) {
        if (int rc = gram_prepare_ordered(ctx, bk, r)) return rc;
        st = bk->stage; st_rhs = bk->stage_rhs; st_scal = bk->stage_scal;
    }
    static const bool gram_v2 = getenv("FRK_GRAM_V2") != nullptr;         // A/B timing: the second-generation (tile-split) kernels
    if (ell && !gram_v2) {
        grid = (unsigned)std::min<int64_t>(r, (int64_t)ctx->sm_count * 4);
        FRK_WORK(ctx, 0, (double)bk->nnz * 9 + (double)m * (4 + 8 * (3 + p)) + ((double)r * r + r) * 8);
        FRK_LAUNCH(ctx, gram_ell3_kernel, grid, GE_T, 0, r, d_bptr, d_perm, d_ucol, d_nU, bk->btile, bk->tile_K, bk->tile_off, bk->ell_val,
                   bk->ell_lid, m, r, d_z, d_X, p, al, d_ve, d_vfs, sigma2fs, d_acc, st, st_rhs, st_scal);
    } else if (ell) {
        if (st_scal) FRK_CHECK_CUDA(cudaMemsetAsync(st_scal, 0, 2 * (size_t)grid * GE_WARPS * sizeof(double), ctx->stream));
        // one launch per class of column-set sizes present in the matrix: <= 4, 5-6, 7 blocks of 8 columns
        const int *h = bk->nblk_hist;
        double bytes = (double)bk->nnz * 9 + (double)m * (4 + 8 * (3 + p)) + ((double)r * r + r) * 8;
        if (h[0] + h[1] + h[2] + h[3] + h[4] > 0)
            if (int rc = launch_gram_ell<4>(ctx, 0, grid, r, bk, m, d_z, d_X, p, al, d_ve, d_vfs, sigma2fs, d_acc, st, st_rhs, st_scal, bytes)) return rc;
        if (h[5] + h[6] > 0)
            if (int rc = launch_gram_ell<6>(ctx, 4, grid, r, bk, m, d_z, d_X, p, al, d_ve, d_vfs, sigma2fs, d_acc, st, st_rhs, st_scal, bytes)) return rc;
        if (h[7] > 0)
            if (int rc = launch_gram_ell<GE_MAXBLK>(ctx, 6, grid, r, bk, m, d_z, d_X, p, al, d_ve, d_vfs, sigma2fs, d_acc, st, st_rhs, st_scal, bytes)) return rc;
    } else
        FRK_LAUNCH(ctx, gram_bucket_kernel, grid, GR_T, 0, r, d_bptr, d_perm, d_ucol, d_nU, d_lid, m, r, d_rowptr, d_col, d_val, d_z, d_X,
                   p, al, d_ve, d_vfs, sigma2fs, d_acc, st, st_rhs, st_scal);
    if (ctx->deterministic)
        FRK_LAUNCH(ctx, gram_ordered_flush_kernel, (unsigned)cdiv((int64_t)r * 32, 256), 256, 0, r, bk->inv_ptr, bk->inv_bucket, bk->inv_lid,
                   d_ucol, d_nU, bk->stage, bk->stage_rhs, bk->stage_scal, (int)(ell ? grid * GE_WARPS : grid), d_acc);
    return 0;
}

// ---------------------------------------------------------------------------------
// quadratic form + SpMV
// ---------------------------------------------------------------------------------
constexpr int QE_WARPS = 4, QE_T = QE_WARPS * 32;
constexpr int QF_LD = UMAX + 1;

__device__ __forceinline__ void cpa16(void *dst, const void *src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src));
}

// direct evaluation for one row by one warp (rows/buckets that do not fit the shared-memory path)
__device__ __forceinline__ void quad_row_direct(int row, int r, const int *__restrict__ rowptr, const int *__restrict__ col,
                                                const double *__restrict__ val, const double *__restrict__ M,
                                                const double *__restrict__ mu, double *__restrict__ q, double *__restrict__ t, int lane) {
    const int beg = rowptr[row], k = rowptr[row + 1] - beg;
    double tq = 0.0, tt = 0.0;
    if (mu) for (int b = lane; b < k; b += 32) tt += val[beg + b] * __ldg(mu + col[beg + b]);
    if (M) {
        for (int a = 0; a < k; a++) {
            const double *Mc = M + (size_t)r * col[beg + a];
            double part = 0.0;
            for (int b = lane; b < k; b += 32) part += __ldg(Mc + col[beg + b]) * val[beg + b];
            tq += val[beg + a] * part;
        }
    }
    tq = warp_sum(tq);
    tt = warp_sum(tt);
    if (lane == 0) {
        if (q && M) q[row] = tq;
        if (t && mu) t[row] = tt;
    }
}

// One CTA per bucket (grid-stride).  Within a bucket every row lives on the bucket's <= 64 columns U, so for a tile of
// 32 rows the quadratic forms are the row sums of (X M[U,U]) o X with X the 32 x |U| densified tile: a small dense fp64
// product that runs on the tensor cores (DMMA m8n8k4), the one place on this path where the dense panel IS the bottleneck
// (the gather formulation is bound by shared-memory latency, not by HBM).  M is symmetric, so only the block-upper-triangular
// part of M[U,U] (8-column blocks, off-diagonal blocks doubled when staged) is multiplied: 168 instead of 288 DMMAs per tile
// at |U| = 48.  Each warp owns ELL tiles; lane = row while densifying (and for t = s'mu), DMMA fragment layout afterwards.
constexpr int QD_LDM = UMAX + 8;             // 72: M row stride (= 8 mod 16 -> the 32 B-fragment addresses cover the 16 bank pairs twice)

struct QuadDmmaSmem {
    double Xt[QE_WARPS][UMAX][QD_LDX];
    double MU[UMAX][QD_LDM];                 // MU[k][n], k-block < n-block doubled, k-block > n-block unused
    double muU[UMAX];
    int ucol[UMAX];
};


// Y = X MU (block-upper part) for one 32-row tile and q_row = sum_n Y[row][n] X[row][n], NBLK 8-column blocks.
// On return qs[mt] holds the finished sum for row 8 mt + g on all four lanes of group g.
template <int NBLK>
__device__ __forceinline__ void tile_dmma(const double (*Xt)[QD_LDX], const double (*MU)[QD_LDM], int g, int tg, double qs[4]) {
    double acc[4][NBLK][2];
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int nt = 0; nt < NBLK; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    const double *xa = &Xt[tg][g];
    const double *mb = &MU[tg][g];
#pragma unroll
    for (int ks = 0; ks < 2 * NBLK; ks++) {                        // k-step of 4 columns; contributes to n-blocks >= ks/2
        double a[4];
#pragma unroll
        for (int mt = 0; mt < 4; mt++) a[mt] = xa[(4 * ks) * QD_LDX + 8 * mt];
#pragma unroll
        for (int nt = ks / 2; nt < NBLK; nt++) {
            const double bf = mb[(4 * ks) * QD_LDM + 8 * nt];
#pragma unroll
            for (int mt = 0; mt < 4; mt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], bf);
        }
    }
    // this lane holds Y[8mt + g][8nt + 2tg + {0,1}]
    const double *xe = &Xt[2 * tg][g];
#pragma unroll
    for (int mt = 0; mt < 4; mt++) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int nt = 0; nt < NBLK; nt++) {
            s0 += acc[mt][nt][0] * xe[(8 * nt) * QD_LDX + 8 * mt];
            s1 += acc[mt][nt][1] * xe[(8 * nt + 1) * QD_LDX + 8 * mt];
        }
        double v = s0 + s1;
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        qs[mt] = v;
    }
}

__global__ void __launch_bounds__(QE_T, 2) quadform_ell_kernel(int nbuckets, const int *__restrict__ bptr, const int *__restrict__ perm,
                                                               const int *__restrict__ ucol_all, const int *__restrict__ nU_all,
                                                               const int *__restrict__ btile, const int *__restrict__ tile_K,
                                                               const int *__restrict__ tile_off, const double *__restrict__ ell_val,
                                                               const unsigned char *__restrict__ ell_lid, int r,
                                                               const int *__restrict__ rowptr, const int *__restrict__ col,
                                                               const double *__restrict__ val, const double *__restrict__ M,
                                                               const double *__restrict__ mu, double *__restrict__ q,
                                                               double *__restrict__ t) {
    extern __shared__ __align__(16) unsigned char qsm_raw[];
    QuadDmmaSmem &S = *reinterpret_cast<QuadDmmaSmem *>(qsm_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = lane >> 2, tg = lane & 3;
    for (int b = blockIdx.x; b < nbuckets; b += gridDim.x) {
        const int r0 = bptr[b], r1 = bptr[b + 1];
        if (r1 <= r0) continue;
        const int nU = nU_all[b];
        if (nU > UMAX) {
            for (int pr = r0 + wid; pr < r1; pr += QE_WARPS) quad_row_direct(perm[pr], r, rowptr, col, val, M, mu, q, t, lane);
            continue;
        }
        const int nblk = (nU + 7) >> 3, Up = nblk * 8;
        __syncthreads();
        if (threadIdx.x < nU) S.ucol[threadIdx.x] = ucol_all[(size_t)b * UMAX + threadIdx.x];
        __syncthreads();
        if (M) {                                                   // stage the block-upper part of M[U,U]; zero padding
            for (int idx = threadIdx.x; idx < Up * Up; idx += QE_T) {
                const int n = idx % Up, k = idx / Up;
                double v = 0.0;
                if (k < nU && n < nU && (k >> 3) <= (n >> 3)) {
                    v = __ldg(M + S.ucol[n] + (size_t)r * S.ucol[k]);
                    if ((k >> 3) < (n >> 3)) v *= 2.0;
                }
                S.MU[k][n] = v;
            }
        }
        if (threadIdx.x < UMAX) S.muU[threadIdx.x] = (mu && threadIdx.x < nU) ? __ldg(mu + S.ucol[threadIdx.x]) : 0.0;
        __syncthreads();
        double (*Xt)[QD_LDX] = S.Xt[wid];
        const int t0 = btile[b], t1 = btile[b + 1];
        for (int tl = t0 + wid; tl < t1; tl += QE_WARPS) {
            const int K = tile_K[tl];
            const int p0 = r0 + (tl - t0) * 32, nrows = min(32, r1 - p0);
            const int row = lane < nrows ? perm[p0 + lane] : -1;
            if (K < 0) {                                           // a row longer than KMAX: direct, one row at a time
                for (int rl = 0; rl < nrows; rl++) quad_row_direct(__shfl_sync(0xffffffffu, row, rl), r, rowptr, col, val, M, mu, q, t, lane);
                continue;
            }
            // densify: zero the Up x 32 tile, then lane = row scatters its entries (padding entries carry val = 0)
            for (int idx = lane; idx < Up * 16; idx += 32) {
                const int cc = idx >> 4, rr = (idx & 15) * 2;
                *reinterpret_cast<double2 *>(&Xt[cc][rr]) = make_double2(0.0, 0.0);
            }
            const size_t off = (size_t)tile_off[tl];
            if (tl + QE_WARPS < t1) {                              // pull this warp's next tile towards L2 while this one is processed
                const size_t noff = (size_t)tile_off[tl + QE_WARPS];
                const int nK = max(tile_K[tl + QE_WARPS], 0);
                const char *pv = reinterpret_cast<const char *>(ell_val + noff);
                for (int c = lane; c < nK * 2; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pv + (size_t)c * 128));
                if (lane < (nK + 3) / 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(ell_lid + noff + (size_t)lane * 128));
            }
            __syncwarp();
            double tt = 0.0;
            {   // coalesced vector reads of the tile: ALL (<= 40) entries of the row in flight before the first scatter
                const double2 *pv = reinterpret_cast<const double2 *>(ell_val + off) + lane;
                const uchar4 *pl = reinterpret_cast<const uchar4 *>(ell_lid + off) + lane;
                double2 v[KMAX / 2];
                uchar4 l[KMAX / 4];
#pragma unroll
                for (int j = 0; j < KMAX / 4; j++)
                    if (4 * j < K) { v[2 * j] = pv[(2 * j) * 32]; v[2 * j + 1] = pv[(2 * j + 1) * 32]; l[j] = pl[j * 32]; }
                double *xs = &Xt[0][lane];
                double ta = 0.0, tb = 0.0;
#pragma unroll
                for (int j = KMAX / 4 - 1; j >= 0; j--)            // last entry first: padding (val 0 at the last real column) gets overwritten
                    if (4 * j < K) {
                        xs[l[j].w * QD_LDX] = v[2 * j + 1].y; ta += v[2 * j + 1].y * S.muU[l[j].w];
                        xs[l[j].z * QD_LDX] = v[2 * j + 1].x; tb += v[2 * j + 1].x * S.muU[l[j].z];
                        xs[l[j].y * QD_LDX] = v[2 * j].y;     ta += v[2 * j].y * S.muU[l[j].y];
                        xs[l[j].x * QD_LDX] = v[2 * j].x;     tb += v[2 * j].x * S.muU[l[j].x];
                    }
                tt = ta + tb;
            }
            __syncwarp();
            if (M) {
                double qs[4];
                switch (nblk) {                                    // compile-time block counts: no predicates or address math in the loops
                    case 1: tile_dmma<1>(Xt, S.MU, g, tg, qs); break;
                    case 2: tile_dmma<2>(Xt, S.MU, g, tg, qs); break;
                    case 3: tile_dmma<3>(Xt, S.MU, g, tg, qs); break;
                    case 4: tile_dmma<4>(Xt, S.MU, g, tg, qs); break;
                    case 5: tile_dmma<5>(Xt, S.MU, g, tg, qs); break;
                    case 6: tile_dmma<6>(Xt, S.MU, g, tg, qs); break;
                    case 7: tile_dmma<7>(Xt, S.MU, g, tg, qs); break;
                    default: tile_dmma<8>(Xt, S.MU, g, tg, qs); break;
                }
                // row 8mt + g is complete on the 4 lanes of group g: lane tg == mt stores it
                if (q) {
#pragma unroll
                    for (int mt = 0; mt < 4; mt++) {
                        const int rr = __shfl_sync(0xffffffffu, row, 8 * mt + g);
                        if (tg == mt && rr >= 0) q[rr] = qs[mt];
                    }
                }
            }
            if (t && mu && row >= 0) t[row] = tt;
            __syncwarp();
        }
    }
}

// The same product for HALF of a 32-row tile (m-tiles 2 half, 2 half + 1): two consumer warps share one densified tile.
// One warp cannot issue DMMAs faster than every ~32 cycles (measured: profiles/r02x), half the rate of the pipe, so it takes
// two warps per scheduler multiplying at the same time to fill it.
// Shared-memory layout of the warp-specialised kernel.  A 64-bit shared load is served one half-warp (g = 0..3 or 4..7, tg = 0..3)
// per wavefront, so the 16 addresses of a half-warp must fall into 16 different bank pairs:
//   M[U,U] fragments  MU[4ks + tg][8nt + g]:  row stride = 4 mod 16 (68)  ->  pair 4 tg + g
//   A fragments       Xt[4ks + tg][8mt + g]:  stride 36                    ->  pair 4 tg + g
//   epilogue          Xt[8nt + 2tg + e][8mt + g]: 72 tg = 8 tg mod 16 would collide for tg and tg + 2, so rows are stored XOR 4
//                     in the columns with bit 2 set (Xt[c][row ^ (c & 4)]): pair 8 (tg & 1) + 4 e + (g ^ 4 (tg >> 1)), all different;
//                     the A fragments see one value of (c & 4) per k-step and stay conflict-free.
// (profiles/r02z: 610 wavefronts per tile for the consumers' loads against 352 ideal before this layout.)
constexpr int QW_LDM = UMAX + 4;
template <int NBLK>
__device__ __forceinline__ void tile_dmma_half(const double (*Xt)[QD_LDX], const double (*MU)[QW_LDM], int g, int tg, int half, double qs[2],
                                               double ts[2]) {
    double acc[2][NBLK][2];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < NBLK; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    const double *xa0 = &Xt[tg][g + 16 * half], *xa1 = &Xt[tg][(g ^ 4) + 16 * half];
    const double *mb = &MU[tg][g];
    const double *xe = &Xt[2 * tg][(g ^ (4 * (tg >> 1))) + 16 * half];
    double s0[2] = {0.0, 0.0}, s1[2] = {0.0, 0.0};
    // row sums of Y o X for column block nt (this lane: columns 8nt + 2tg + {0,1}); block nt is final after k-step 2nt + 1, so its
    // share of the epilogue is issued one k-step later, between the DMMAs of the blocks still being accumulated
    auto fold = [&](int nt) {
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            s0[mt] += acc[mt][nt][0] * xe[(8 * nt) * QD_LDX + 8 * mt];
            s1[mt] += acc[mt][nt][1] * xe[(8 * nt + 1) * QD_LDX + 8 * mt];
        }
    };
#pragma unroll
    for (int ks = 0; ks < 2 * NBLK; ks++) {
        double a[2];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) a[mt] = ((ks & 1) ? xa1 : xa0)[(4 * ks) * QD_LDX + 8 * mt];
#pragma unroll
        for (int nt = ks / 2; nt < NBLK; nt++) {
            const double bf = mb[(4 * ks) * QW_LDM + 8 * nt];
#pragma unroll
            for (int mt = 0; mt < 2; mt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], bf);
        }
        if (ks >= 2 && !(ks & 1)) fold(ks / 2 - 1);
    }
    fold(NBLK - 1);
    // the LAST column of the last block is never a real column here (the caller keeps one spare): MU holds mu there, X is zero, so
    // Y[row][last] = s_row'mu comes out of the same product; lanes with tg == 3 hold it for rows 8 mt + g
    ts[0] = acc[0][NBLK - 1][1]; ts[1] = acc[1][NBLK - 1][1];
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
        double v = s0[mt] + s1[mt];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        qs[mt] = v;
    }
}

// ---------------------------------------------------------------------------------
// Warp-specialised quadratic form (round 2, third generation).  quadform_ell_kernel above makes every warp alternate between
// a latency-bound phase (load the ELL tile, densify it: ~1000 instructions) and a DMMA-bound one (168 DMMAs at |U| = 48), with
// two warps per scheduler: the DMMA pipe is busy 40 % of the time.  Here ONE CTA of 12 warps owns an SM: warps 0-7 (two per
// scheduler, each taking 16 of a tile's 32 rows) only multiply, warps 8-11 only produce -- producer p+8 feeds consumers p and
// p+4 through a double-buffered densified tile.
//   producer: the ELL tile (K x 32 values + K x 32 local ids, contiguous in HBM) arrives by ONE bulk-asynchronous copy pair
//             (cp.async.bulk ... mbarrier::complete_tx::bytes); the lanes pull their row into registers, the next tile's copy is
//             issued into the same staging area at once (so it flies while this tile is densified and while the producer waits
//             for a free buffer), then zero + scatter into Xt[buf] and t = s'mu on the way;
//   consumer: waits for Xt[buf], Y = X MU (block-upper) and the row sums of Y o X on the fp64 tensor cores, releases the buffer.
// Handshakes are mbarriers in shared memory (full/empty per buffer, one transaction barrier per staging area); the CTA-wide
// barrier is used at bucket boundaries only (M[U,U] is restaged).  Spin loops are bounded: a lost arrival traps instead of hanging.
// ---------------------------------------------------------------------------------
constexpr int QW_PAIRS = 4, QW_T = 3 * QW_PAIRS * 32;
constexpr int QW_STAGE_BYTES = KMAX * 32 * 9;               // values + ids of the longest tile

struct QuadWsSmem {
    double MU[UMAX][QW_LDM];
    double Xt[QW_PAIRS][2][UMAX][QD_LDX];
    __align__(16) unsigned char stage[QW_PAIRS][QW_STAGE_BYTES];
    int ucol[UMAX];
    unsigned long long full_xt[QW_PAIRS][2], empty_xt[QW_PAIRS][2], full_ell[QW_PAIRS];
};
static_assert(sizeof(QuadWsSmem) <= 232448, "QuadWsSmem exceeds the 227 KB a CTA may opt into");

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned a = smem_u32(bar);
    for (unsigned spin = 0;; spin++) {
        unsigned ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
        if (spin > (1u << 22)) __trap();                    // seconds of spinning: a handshake was lost -- fail loudly, never hang the GPU
    }
}
// global -> shared bulk copy (the TMA engine, no tensor map: the tile is one contiguous range), completion counted on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct QuadWsArgs {
    int nbuckets, r;
    const int *bptr, *perm, *ucol_all, *nU_all, *btile, *tile_K, *tile_off;
    const double *ell_val;
    const unsigned char *ell_lid;
    const int *rowptr, *col;
    const double *val, *M, *mu;
    double *q, *t;
};

// Register budgets (setmaxnreg): the kernel is compiled for 168 registers per thread (384 threads, one CTA per SM); the two consumer
// warpgroups give registers back (the 16-row half tile needs 64 for its accumulators), the producer warpgroup takes them (a whole
// ELL tile row -- 40 values and 40 ids -- lives in registers between the staging area and the scatter):
// 8 warps x 32 x 136 + 4 warps x 32 x 232 = 64512 <= 65536.
constexpr int QW_REGS_CONSUMER = 136, QW_REGS_PRODUCER = 232;

// Consumer side of one bucket: this warp's half (16 rows) of every tile of its producer.  Kl = tile_K of the warp's next 32 tiles
// (one per lane), it = running count of tiles taken over (buffer index and mbarrier phase).
template <int NBLK>
__device__ __forceinline__ void quadform_ws_consume(const QuadWsArgs &A, QuadWsSmem &S, int t0, int t1, int r0, int r1, int Kl, unsigned &it) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int pair = wid & (QW_PAIRS - 1), half = (wid >> 2) & 1;
    const int g = lane >> 2, tg = lane & 3;
    const int *__restrict__ perm = A.perm;
    double *__restrict__ q = A.q;
    double *__restrict__ t = (A.t && A.mu) ? A.t : nullptr;
    unsigned live = __ballot_sync(0xffffffffu, Kl > 0);            // tiles that come through the pipeline (K <= 0: the producer's business)
    // the two rows this lane reports per tile: 16 half + 8 mt + g (q by the lane with tg == mt, t by tg == 3); their ids are
    // fetched one tile ahead so that the L2 round trip never sits in front of a tile's products
    int nrow0 = -1, nrow1 = -1;
    { const int p0 = r0 + pair * 32 + 16 * half + g; if (t0 + pair < t1) { nrow0 = p0 < r1 ? perm[p0] : -1; nrow1 = p0 + 8 < r1 ? perm[p0 + 8] : -1; } }
    for (int j = 0, tl = t0 + pair; tl < t1; j++, tl += QW_PAIRS) {
        if ((j & 31) == 0 && j > 0) {
            const int tt = tl + QW_PAIRS * lane;
            live = __ballot_sync(0xffffffffu, tt < t1 && A.tile_K[tt] > 0);
        }
        const int row0 = nrow0, row1 = nrow1;
        if (tl + QW_PAIRS < t1) {
            const int p1 = r0 + (tl + QW_PAIRS - t0) * 32 + 16 * half + g;
            nrow0 = p1 < r1 ? perm[p1] : -1; nrow1 = p1 + 8 < r1 ? perm[p1 + 8] : -1;
        }
        if (!((live >> (j & 31)) & 1u)) continue;
        const unsigned buf = it & 1;
        mbar_wait(&S.full_xt[pair][buf], (it >> 1) & 1);
        double qs[2], ts[2];
        tile_dmma_half<NBLK>(S.Xt[pair][buf], S.MU, g, tg, half, qs, ts);
        mbar_arrive(&S.empty_xt[pair][buf]);                       // qs depends on every shared-memory read of this lane
        it++;
        if (row0 >= 0) { if (q && tg == 0) q[row0] = qs[0]; if (t && tg == 3) t[row0] = ts[0]; }
        if (row1 >= 0) { if (q && tg == 1) q[row1] = qs[1]; if (t && tg == 3) t[row1] = ts[1]; }
    }
}

// Both roles walk the same buckets and meet at the same three CTA-wide barriers per bucket (M[U,U] is restaged by all 384 threads);
// between them a producer and its two consumers only talk through the mbarriers.
template <bool PRODUCER>
__device__ __forceinline__ void quadform_ws_role(const QuadWsArgs &A, QuadWsSmem &S) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int pair = wid & (QW_PAIRS - 1), half = (wid >> 2) & 1;
    const int g = lane >> 2, tg = lane & 3;
    const int r = A.r;
    const int *__restrict__ tile_K = A.tile_K;
    const int *__restrict__ tile_off = A.tile_off;
    const int *__restrict__ perm = A.perm;
    const double *__restrict__ M = A.M;
    const double *__restrict__ mu = A.mu;
    double *__restrict__ q = A.q;
    double *__restrict__ t = A.t;
    unsigned it = 0;                                               // tiles handed from producer to consumers so far (same count on both sides)
    unsigned ell_it = 0;                                           // producer: bulk copies waited for so far
    for (int b = blockIdx.x; b < A.nbuckets; b += gridDim.x) {
        const int r0 = A.bptr[b], r1 = A.bptr[b + 1];
        if (r1 <= r0) continue;
        const int nU = A.nU_all[b];
        if (nU >= UMAX) {
            for (int pr = r0 + wid; pr < r1; pr += 3 * QW_PAIRS) quad_row_direct(perm[pr], r, A.rowptr, A.col, A.val, M, mu, q, t, lane);
            continue;
        }
        const int nblk = (nU + 8) >> 3, Up = nblk * 8;              // at least one spare column: the last one carries mu (t = S mu)
        const int t0 = A.btile[b], t1 = A.btile[b + 1];
        unsigned char *stg = S.stage[pair];
        bool pending = false;                                      // producer: the copy of its next tile is already in flight
        // tile metadata of this warp's next 32 tiles, one per lane (a dependent L2 round trip per tile otherwise)
        int Kl, offl = 0;
        { const int tt = t0 + pair + QW_PAIRS * lane; Kl = tt < t1 ? tile_K[tt] : 0; if (PRODUCER && tt < t1) offl = tile_off[tt]; }
        __syncthreads();                                           // the consumers are done with the previous bucket's M[U,U]
        if (PRODUCER) {                                            // first tile of the bucket: start its copy before M[U,U] is staged
            const int tl = t0 + pair;
            const int K = tl < t1 ? __shfl_sync(0xffffffffu, Kl, 0) : -1;
            if (K > 0) {
                if (lane == 0) {
                    const size_t off = (size_t)offl;
                    mbar_arrive_expect_tx(&S.full_ell[pair], (unsigned)K * 288u);
                    bulk_g2s(stg, A.ell_val + off, (unsigned)K * 256u, &S.full_ell[pair]);
                    bulk_g2s(stg + KMAX * 256, A.ell_lid + off, (unsigned)K * 32u, &S.full_ell[pair]);
                }
                pending = true;
            }
        }
        if (threadIdx.x < nU) S.ucol[threadIdx.x] = A.ucol_all[(size_t)b * UMAX + threadIdx.x];
        __syncthreads();
        for (int idx = threadIdx.x; idx < Up * Up; idx += QW_T) {  // the block-upper part of M[U,U]; zero padding; mu in the last column
            const int n = idx % Up, k = idx / Up;
            double v = 0.0;
            if (k < nU && n < nU && (k >> 3) <= (n >> 3)) {
                v = __ldg(M + S.ucol[n] + (size_t)r * S.ucol[k]);
                if ((k >> 3) < (n >> 3)) v *= 2.0;
            } else if (n == Up - 1 && k < nU && mu) v = __ldg(mu + S.ucol[k]);
            S.MU[k][n] = v;
        }
        __syncthreads();
        if (PRODUCER) {
            for (int j = 0, tl = t0 + pair; tl < t1; j++, tl += QW_PAIRS) {
                if ((j & 31) == 0 && j > 0) {
                    const int tt = tl + QW_PAIRS * lane;
                    Kl = tt < t1 ? tile_K[tt] : 0; offl = tt < t1 ? tile_off[tt] : 0;
                }
                const int K = __shfl_sync(0xffffffffu, Kl, j & 31);
                const int offc = __shfl_sync(0xffffffffu, offl, j & 31);
                if (K <= 0) {                                      // a row longer than KMAX: direct, one row at a time (K == 0: empty rows)
                    const int p0 = r0 + (tl - t0) * 32, nrows = min(32, r1 - p0);
                    const int row = lane < nrows ? perm[p0 + lane] : -1;
                    if (K < 0) for (int rl = 0; rl < nrows; rl++) quad_row_direct(__shfl_sync(0xffffffffu, row, rl), r, A.rowptr, A.col, A.val, M, mu, q, t, lane);
                    else if (row >= 0) { if (q) q[row] = 0.0; if (t && mu) t[row] = 0.0; }
                    continue;
                }
                if (!pending && lane == 0) {
                    const size_t off = (size_t)offc;
                    mbar_arrive_expect_tx(&S.full_ell[pair], (unsigned)K * 288u);
                    bulk_g2s(stg, A.ell_val + off, (unsigned)K * 256u, &S.full_ell[pair]);
                    bulk_g2s(stg + KMAX * 256, A.ell_lid + off, (unsigned)K * 32u, &S.full_ell[pair]);
                }
                mbar_wait(&S.full_ell[pair], ell_it & 1); ell_it++;
                double2 v[KMAX / 2];
                uchar4 l[KMAX / 4];
                {
                    const double2 *pv = reinterpret_cast<const double2 *>(stg) + lane;
                    const uchar4 *pl = reinterpret_cast<const uchar4 *>(stg + KMAX * 256) + lane;
#pragma unroll
                    for (int jj = 0; jj < KMAX / 4; jj++)
                        if (4 * jj < K) { v[2 * jj] = pv[(2 * jj) * 32]; v[2 * jj + 1] = pv[(2 * jj + 1) * 32]; l[jj] = pl[jj * 32]; }
                }
                __syncwarp();
                {   // the staging area is free again: start the copy of this producer's next tile (no global load of this warp is
                    // outstanding here -- the fence below would wait for it)
                    const int nt = tl + QW_PAIRS;
                    int nK, noffi;
                    if (((j + 1) & 31) != 0) {
                        nK = __shfl_sync(0xffffffffu, Kl, (j + 1) & 31); noffi = __shfl_sync(0xffffffffu, offl, (j + 1) & 31);
                        if (nt >= t1) nK = -1;
                    } else {                                       // the next batch of metadata is not loaded yet (once in 32 tiles)
                        nK = nt < t1 ? tile_K[nt] : -1; noffi = nK > 0 ? tile_off[nt] : 0;
                    }
                    pending = nK > 0;
                    if (pending && lane == 0) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // our generic-proxy reads before the async-proxy writes
                        const size_t noff = (size_t)noffi;
                        mbar_arrive_expect_tx(&S.full_ell[pair], (unsigned)nK * 288u);
                        bulk_g2s(stg, A.ell_val + noff, (unsigned)nK * 256u, &S.full_ell[pair]);
                        bulk_g2s(stg + KMAX * 256, A.ell_lid + noff, (unsigned)nK * 32u, &S.full_ell[pair]);
                    }
                }
                const unsigned buf = it & 1;
                mbar_wait(&S.empty_xt[pair][buf], ((it >> 1) & 1) ^ 1);
                double (*Xt)[QD_LDX] = S.Xt[pair][buf];
                for (int idx = lane; idx < Up * 16; idx += 32) {
                    const int cc = idx >> 4, rr = (idx & 15) * 2;
                    *reinterpret_cast<double2 *>(&Xt[cc][rr]) = make_double2(0.0, 0.0);
                }
                __syncwarp();
                double *xs = &Xt[0][0];
#pragma unroll
                for (int jj = KMAX / 4 - 1; jj >= 0; jj--)         // last entry first: padding (val 0 at the last real column) gets overwritten
                    if (4 * jj < K) {                              // Xt[c][row ^ (c & 4)]: see the layout note above tile_dmma_half
                        xs[l[jj].w * QD_LDX + (lane ^ (l[jj].w & 4))] = v[2 * jj + 1].y;
                        xs[l[jj].z * QD_LDX + (lane ^ (l[jj].z & 4))] = v[2 * jj + 1].x;
                        xs[l[jj].y * QD_LDX + (lane ^ (l[jj].y & 4))] = v[2 * jj].y;
                        xs[l[jj].x * QD_LDX + (lane ^ (l[jj].x & 4))] = v[2 * jj].x;
                    }
                mbar_arrive(&S.full_xt[pair][buf]);                // release: this lane's writes are visible to whoever sees the phase flip
                it++;
            }
        } else {
            switch (nblk) {                                        // the block count is a property of the bucket: one specialised tile loop each
                case 1: quadform_ws_consume<1>(A, S, t0, t1, r0, r1, Kl, it); break;
                case 2: quadform_ws_consume<2>(A, S, t0, t1, r0, r1, Kl, it); break;
                case 3: quadform_ws_consume<3>(A, S, t0, t1, r0, r1, Kl, it); break;
                case 4: quadform_ws_consume<4>(A, S, t0, t1, r0, r1, Kl, it); break;
                case 5: quadform_ws_consume<5>(A, S, t0, t1, r0, r1, Kl, it); break;
                case 6: quadform_ws_consume<6>(A, S, t0, t1, r0, r1, Kl, it); break;
                case 7: quadform_ws_consume<7>(A, S, t0, t1, r0, r1, Kl, it); break;
                default: quadform_ws_consume<8>(A, S, t0, t1, r0, r1, Kl, it); break;
            }
        }
    }
}

__global__ void __launch_bounds__(QW_T, 1) quadform_ws_kernel(const __grid_constant__ QuadWsArgs A) {
    extern __shared__ __align__(16) unsigned char qsm_raw[];
    QuadWsSmem &S = *reinterpret_cast<QuadWsSmem *>(qsm_raw);
    if (threadIdx.x < QW_PAIRS) {
        mbar_init(&S.full_xt[threadIdx.x][0], 32); mbar_init(&S.full_xt[threadIdx.x][1], 32);
        mbar_init(&S.empty_xt[threadIdx.x][0], 64); mbar_init(&S.empty_xt[threadIdx.x][1], 64);
        mbar_init(&S.full_ell[threadIdx.x], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= 2 * QW_PAIRS * 32) {                        // warpgroup 2: producers
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(QW_REGS_PRODUCER));
        quadform_ws_role<true>(A, S);
    } else {                                                       // warpgroups 0, 1: consumers
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(QW_REGS_CONSUMER));
        quadform_ws_role<false>(A, S);
    }
}

int frk_rowbuckets_is_dense(const frk_rowbuckets *bk) { return bk && bk->dense; }

// dense / implicit matrices only: the quadratic forms against M = L^-T L^-1 through the triangular factor (model.cu)
int frk_quadform_dense_factor(frk_ctx *ctx, int64_t n, int r, const double *d_val, const frk_rowbuckets *bk, const double *d_Linv,
                              const double *d_mu, double *d_q, double *d_t) {
    FRK_REQUIRE(ctx && bk && bk->dense && d_Linv && bk->n == n && bk->ncols == r, "frk_quadform_dense_factor: bad argument");
    if (n == 0) return 0;
    return quadform_dense(ctx, bk, n, r, d_val, nullptr, d_mu, d_q, d_t, d_Linv);
}

extern "C" int frk_quadform_spmv_bucketed(frk_ctx *ctx, int64_t n, int r, const int *d_rowptr, const int *d_col, const double *d_val,
                                          const frk_rowbuckets *bk, const double *d_M, const double *d_mu, double *d_q, double *d_t) {
    FRK_REQUIRE(ctx && bk && (d_rowptr || bk->implicit), "frk_quadform_spmv_bucketed: NULL argument");
    FRK_REQUIRE(bk->n == n && bk->ncols == r, "frk_quadform_spmv_bucketed: the row buckets belong to a %lld x %d matrix", (long long)bk->n, bk->ncols);
    if (n == 0) return 0;
    if (bk->dense) return quadform_dense(ctx, bk, n, r, d_val, d_M, d_mu, d_q, d_t);
    // warp-specialised kernel (one CTA per SM) when there is a matrix M and the buckets are long enough to keep its pipeline
    // busy; FRK_QUAD_MODE = ws | v2 forces one of the two (A/B timing, tests)
    const char *mode = getenv("FRK_QUAD_MODE");
    const bool force_ws = mode && !strcmp(mode, "ws"), force_v2 = mode && !strcmp(mode, "v2");
    static const int ws_env = getenv("FRK_QUAD_WS_MIN") ? atoi(getenv("FRK_QUAD_WS_MIN")) : 0;    // tiles per column (A/B timing)
    const int ws_min = ws_env > 0 ? ws_env : ctx->quad_ws_min;
    if (d_M && bk->T > 0 && bk->max_nU < UMAX && !force_v2 && (force_ws || (int64_t)bk->T >= (int64_t)ws_min * r)) {
        static bool attr_ws = false;
        if (!attr_ws) {
            FRK_CHECK_CUDA(cudaFuncSetAttribute(quadform_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QuadWsSmem)));
            attr_ws = true;
        }
        const unsigned grid = (unsigned)std::min<int64_t>(r, (int64_t)ctx->sm_count);
        QuadWsArgs A{r, r, bk->bptr, bk->perm, bk->ucol, bk->nU, bk->btile, bk->tile_K, bk->tile_off, bk->ell_val, bk->ell_lid,
                     d_rowptr, d_col, d_val, d_M, d_mu, d_q, d_t};
        FRK_LAUNCH(ctx, quadform_ws_kernel, grid, QW_T, sizeof(QuadWsSmem), A);
        return 0;
    }
    static bool attr = false;
    if (!attr) {
        FRK_CHECK_CUDA(cudaFuncSetAttribute(quadform_ell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QuadDmmaSmem)));
        attr = true;
    }
    unsigned grid = (unsigned)std::min<int64_t>(r, (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, quadform_ell_kernel, grid, QE_T, sizeof(QuadDmmaSmem), r, bk->bptr, bk->perm, bk->ucol, bk->nU, bk->btile, bk->tile_K,
               bk->tile_off, bk->ell_val, bk->ell_lid, r, d_rowptr, d_col, d_val, d_M, d_mu, d_q, d_t);
    return 0;
}

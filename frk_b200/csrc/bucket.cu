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

constexpr int UMAX = 64;          // columns per bucket handled through shared memory
constexpr int KMAX = 40;          // non-zeros per row handled by the thread-per-row quadratic form
constexpr int HSIZE = 1024;       // open-addressing table for column -> local index (> UMAX + threads: never fills)
constexpr int MAXP_B = 4;
struct AlphaB { double a[MAXP_B]; };

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

__global__ void bucket_scatter_kernel(int64_t n, const int *__restrict__ key, const int *__restrict__ bptr, int *__restrict__ fill,
                                      int *__restrict__ perm) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = key[i];
        perm[bptr[k] + atomicAdd(fill + k, 1)] = (int)i;
    }
}

constexpr int BU_T = 256, BU_CH = 256;
struct UnionSmem {
    int keys[HSIZE];
    int ids[HSIZE];
    int ucol[UMAX];
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
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&key, (size_t)(n + 2 * (size_t)ncols) * sizeof(int), ctx->stream));
    int *cnt = key + n, *fill = cnt + ncols;
    FRK_CHECK_CUDA(cudaMemsetAsync(cnt, 0, 2 * (size_t)ncols * sizeof(int), ctx->stream));
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, bucket_count_kernel, grid, 256, 0, n, d_rowptr, d_col, key, cnt);
    int64_t tot = 0;
    int rc = frk_exclusive_scan_i32(ctx, cnt, ncols, d_bptr, &tot);
    if (!rc) {
        FRK_LAUNCH(ctx, bucket_scatter_kernel, grid, 256, 0, n, key, d_bptr, fill, d_perm);
        FRK_LAUNCH(ctx, bucket_union_kernel, (unsigned)std::min<int64_t>(ncols, (int64_t)ctx->sm_count * 8), BU_T, 0, ncols, d_bptr,
                   d_perm, d_rowptr, d_col, d_ucol, d_nU, d_lid);
    }
    cudaFreeAsync(key, ctx->stream);
    return rc;
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
                                                           double *__restrict__ acc) {
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
            // direct path: one warp per row, one fp64 atomic per (row, pair)
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
                    for (int bb = a + lane; bb < k; bb += 32) atomicAdd(G + ca + (size_t)r * col[beg + bb], wa * val[beg + bb]);
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
        // flush: each unordered pair of local columns once, into the upper triangle (row <= col) of G
        if (active) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int la = ty * 4 + i, lb = tx * 4 + j;
                    if (la < nU && lb < nU && (ty < tx || la <= lb) && t[i][j] != 0.0) {
                        const int ca = S.ucol[la], cb = S.ucol[lb];
                        atomicAdd(G + min(ca, cb) + (size_t)r * max(ca, cb), t[i][j]);
                    }
                }
            }
        }
        if (threadIdx.x < nU) atomicAdd(rhs + S.ucol[threadIdx.x], rh);
    }
    s_quad = warp_sum(s_quad);
    s_logd = warp_sum(s_logd);
    __syncthreads();
    if (lane == 0) { S.red[wid][0] = s_quad; S.red[wid][1] = s_logd; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, bsum = 0;
        for (int i = 0; i < GR_T / 32; i++) { a += S.red[i][0]; bsum += S.red[i][1]; }
        atomicAdd(rhs + r, a);
        atomicAdd(rhs + r + 1, bsum);
    }
}

extern "C" int frk_gram_rhs_bucketed(frk_ctx *ctx, int64_t m, int r, const int *d_rowptr, const int *d_col, const double *d_val,
                                     const int *d_perm, const int *d_bptr, const int *d_ucol, const int *d_nU,
                                     const unsigned char *d_lid, const double *d_z, const double *d_X, int p, const double *h_alpha,
                                     const double *d_ve, const double *d_vfs, double sigma2fs, double *d_acc) {
    FRK_REQUIRE(ctx && d_rowptr && d_acc && d_perm && d_bptr && d_ucol && d_nU && d_lid, "frk_gram_rhs_bucketed: NULL argument");
    FRK_REQUIRE(p >= 0 && p <= MAXP_B, "frk_gram_rhs_bucketed: %d fixed effects; the device path handles at most %d", p, MAXP_B);
    if (m == 0) return 0;
    AlphaB al{};
    for (int q = 0; q < p; q++) al.a[q] = h_alpha[q];
    unsigned grid = (unsigned)std::min<int64_t>(r, (int64_t)ctx->sm_count * 8);
    FRK_LAUNCH(ctx, gram_bucket_kernel, grid, GR_T, 0, r, d_bptr, d_perm, d_ucol, d_nU, d_lid, m, r, d_rowptr, d_col, d_val, d_z, d_X,
               p, al, d_ve, d_vfs, sigma2fs, d_acc);
    return 0;
}

// ---------------------------------------------------------------------------------
// quadratic form + SpMV
// ---------------------------------------------------------------------------------
constexpr int QF_T = 128;
constexpr int QF_LD = UMAX + 1;

struct QuadSmem {
    ChunkMeta<QF_T> C;
    int ucol[UMAX];
    double MU[UMAX][QF_LD];
    double muU[UMAX];
    double sval[KMAX][QF_T];
    unsigned char slid[KMAX][QF_T];
};

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

__global__ void __launch_bounds__(QF_T) quadform_bucket_kernel(int nbuckets, const int *__restrict__ bptr, const int *__restrict__ perm,
                                                               const int *__restrict__ ucol_all, const int *__restrict__ nU_all,
                                                               const unsigned char *__restrict__ lid, int r,
                                                               const int *__restrict__ rowptr, const int *__restrict__ col,
                                                               const double *__restrict__ val, const double *__restrict__ M,
                                                               const double *__restrict__ mu, double *__restrict__ q,
                                                               double *__restrict__ t) {
    extern __shared__ __align__(16) unsigned char qsm_raw[];
    QuadSmem &S = *reinterpret_cast<QuadSmem *>(qsm_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int b = blockIdx.x; b < nbuckets; b += gridDim.x) {
        const int b0 = bptr[b], b1 = bptr[b + 1];
        if (b1 <= b0) continue;
        const int nU = nU_all[b];
        if (nU > UMAX) {
            for (int pr = b0 + wid; pr < b1; pr += QF_T / 32) quad_row_direct(perm[pr], r, rowptr, col, val, M, mu, q, t, lane);
            continue;
        }
        __syncthreads();
        if (threadIdx.x < nU) S.ucol[threadIdx.x] = ucol_all[(size_t)b * UMAX + threadIdx.x];
        __syncthreads();
        // stage M[U,U] and mu[U]
        if (M) {
            for (int idx = threadIdx.x; idx < nU * nU; idx += QF_T) {
                const int la = idx % nU, lb = idx / nU;
                S.MU[lb][la] = __ldg(M + S.ucol[la] + (size_t)r * S.ucol[lb]);
            }
        }
        if (threadIdx.x < nU) S.muU[threadIdx.x] = mu ? __ldg(mu + S.ucol[threadIdx.x]) : 0.0;
        for (int c0 = b0; c0 < b1; c0 += QF_T) {
            const int nr = min(QF_T, b1 - c0);
            __syncthreads();
            chunk_load<QF_T>(S.C, perm, c0, nr, rowptr);
            const int total = S.C.off[QF_T];
            for (int g0 = threadIdx.x; g0 < total; g0 += 4 * QF_T) {   // stage: flat walk, row-major -> [entry][row]
                int rl[4], e0[4];                                    // four independent loads in flight per thread
                double v[4];
                unsigned char l[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int g = g0 + u * QF_T;
                    rl[u] = -1;
                    if (g < total) {
                        rl[u] = chunk_row_of<QF_T>(S.C, g);
                        e0[u] = g - S.C.off[rl[u]];
                        if (S.C.off[rl[u] + 1] - S.C.off[rl[u]] > KMAX) rl[u] = -1;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (rl[u] >= 0) { const int e = S.C.beg[rl[u]] + e0[u]; v[u] = val[e]; l[u] = lid[e]; }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (rl[u] >= 0) { S.sval[e0[u]][rl[u]] = v[u]; S.slid[e0[u]][rl[u]] = l[u]; }
            }
            __syncthreads();
            if (threadIdx.x < nr) {
                const int tI = threadIdx.x, k = S.C.off[tI + 1] - S.C.off[tI];
                if (k <= KMAX) {
                    double qq = 0.0, tt = 0.0;
                    for (int a = 0; a < k; a++) {
                        const int la = S.slid[a][tI];
                        const double va = S.sval[a][tI];
                        const double *Mr = S.MU[la];
                        double p0 = 0.5 * Mr[la] * va, p1 = 0.0;
                        int bb = a + 1;
                        for (; bb + 1 < k; bb += 2) {
                            p0 += Mr[S.slid[bb][tI]] * S.sval[bb][tI];
                            p1 += Mr[S.slid[bb + 1][tI]] * S.sval[bb + 1][tI];
                        }
                        if (bb < k) p0 += Mr[S.slid[bb][tI]] * S.sval[bb][tI];
                        qq += va * (p0 + p1);
                        tt += va * S.muU[la];
                    }
                    const int row = S.C.row[tI];
                    if (q && M) q[row] = 2.0 * qq;
                    if (t && mu) t[row] = tt;
                }
            }
            // rows longer than KMAX: direct, one warp each
            for (int rl = wid; rl < nr; rl += QF_T / 32)
                if (S.C.off[rl + 1] - S.C.off[rl] > KMAX) quad_row_direct(S.C.row[rl], r, rowptr, col, val, M, mu, q, t, lane);
        }
    }
}

extern "C" int frk_quadform_spmv_bucketed(frk_ctx *ctx, int64_t n, int r, const int *d_rowptr, const int *d_col, const double *d_val,
                                          const int *d_perm, const int *d_bptr, const int *d_ucol, const int *d_nU,
                                          const unsigned char *d_lid, const double *d_M, const double *d_mu, double *d_q, double *d_t) {
    FRK_REQUIRE(ctx && d_rowptr && d_perm && d_bptr && d_ucol && d_nU && d_lid, "frk_quadform_spmv_bucketed: NULL argument");
    if (n == 0) return 0;
    static bool attr = false;
    if (!attr) {
        FRK_CHECK_CUDA(cudaFuncSetAttribute(quadform_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QuadSmem)));
        attr = true;
    }
    unsigned grid = (unsigned)std::min<int64_t>(r, (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, quadform_bucket_kernel, grid, QF_T, sizeof(QuadSmem), r, d_bptr, d_perm, d_ucol, d_nU, d_lid, r, d_rowptr, d_col,
               d_val, d_M, d_mu, d_q, d_t);
    return 0;
}

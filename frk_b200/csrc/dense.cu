// Dense fp64 linear algebra for the r x r systems of the EM algorithm: one tensor-core GEMM
// (C = beta*C + alpha*A*B', DMMA m8n8k4) and the blocked Cholesky / triangular inverse / chol2inv
// built from it, plus the small element-wise and reduction kernels the E/M-steps need.
//
// Reference semantics: Matrix::chol / chol2inv / determinant as used in R/SREfit.R:177-198
// (.loglik.ind), :252-254 (.SRE.Estep.ind), :273 (Khat_inv), :475-478,517-524 (block-exponential
// f_tau); R/linalgfns.R:21-36 (tr, diag2, logdet).
//
// There is no fp64 kind of tcgen05.mma, so fp64 tensor-core work on sm_100a is mma.sync DMMA.
// Every product of the blocked algorithms is arranged as "NT" (both operands have the non-reduced
// index contiguous in column-major storage) so one kernel with natural, coalesced 16-byte cp.async
// loads serves potrf (panel solve + trailing SYRK), the triangular inverse (computed as L^-T, i.e.
// in the transposed world) and the final L^-T L^-1 product.
#include "common.cuh"
#include <algorithm>
#include <cmath>

// ---------------------------------------------------------------------------------
// GEMM NT
// ---------------------------------------------------------------------------------
constexpr int BM = 128, BN = 64, BK = 16, GEMM_T = 256, STAGES = 3;
constexpr int LDSA = BM + 4, LDSB = BN + 4;                 // = 4 (mod 16): conflict-free 8-byte fragment loads
constexpr int STAGE_DOUBLES = BK * (LDSA + LDSB);
constexpr size_t GEMM_SMEM = (size_t)STAGES * STAGE_DOUBLES * sizeof(double);

struct GemmArgs {
    int M, N, K;
    double alpha, beta;
    const double *A; int lda;
    const double *B; int ldb;
    double *C; int ldc;
    int lower;        // only tiles touching the lower triangle; elements above the diagonal are not stored
    int kstart_m0;    // start the k loop at the tile's first row (A, B upper triangular: L^-T L^-1)
};

__device__ __forceinline__ void cp_async16(void *dst, const void *src, int src_bytes) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src, int src_bytes) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// rows x BK tile of a column-major operand (row index contiguous) -> smem [k][row]
template <int ROWS, int LDS, bool ALIGN16>
__device__ __forceinline__ void load_tile(double *s, const double *__restrict__ g, int ld, int row0, int nrows, int k0, int K) {
    if (ALIGN16) {
        constexpr int PER_COL = ROWS / 2;
#pragma unroll
        for (int it = 0; it < (PER_COL * BK) / GEMM_T; it++) {
            int idx = threadIdx.x + it * GEMM_T;
            int kk = idx / PER_COL, r2 = (idx % PER_COL) * 2;
            int row = row0 + r2, k = k0 + kk;
            int bytes = (k < K) ? max(0, min(16, (nrows - row) * 8)) : 0;
            const double *src = bytes ? g + row + (size_t)k * ld : g;
            cp_async16(s + kk * LDS + r2, src, bytes);
        }
    } else {
#pragma unroll
        for (int it = 0; it < (ROWS * BK) / GEMM_T; it++) {
            int idx = threadIdx.x + it * GEMM_T;
            int kk = idx / ROWS, r1 = idx % ROWS;
            int row = row0 + r1, k = k0 + kk;
            int bytes = (k < K && row < nrows) ? 8 : 0;
            const double *src = bytes ? g + row + (size_t)k * ld : g;
            cp_async8(s + kk * LDS + r1, src, bytes);
        }
    }
}

template <bool ALIGN16>
__global__ void __launch_bounds__(GEMM_T, 2) dgemm_nt_kernel(GemmArgs p) {
    extern __shared__ __align__(16) double gsm[];
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (p.lower && n0 >= m0 + BM) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int wm = wid & 3, wn = wid >> 2;
    const int g = lane >> 2, tg = lane & 3;

    int kbeg = p.kstart_m0 ? (m0 / BK) * BK : 0;
    const int nk = (p.K - kbeg + BK - 1) / BK;

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto stageA = [&](int s) { return gsm + (size_t)s * STAGE_DOUBLES; };
    auto stageB = [&](int s) { return gsm + (size_t)s * STAGE_DOUBLES + BK * LDSA; };

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nk) {
            load_tile<BM, LDSA, ALIGN16>(stageA(s), p.A, p.lda, m0, p.M, kbeg + s * BK, p.K);
            load_tile<BN, LDSB, ALIGN16>(stageB(s), p.B, p.ldb, n0, p.N, kbeg + s * BK, p.K);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nxt = kt + STAGES - 1;
            if (nxt < nk) {
                load_tile<BM, LDSA, ALIGN16>(stageA(nxt % STAGES), p.A, p.lda, m0, p.M, kbeg + nxt * BK, p.K);
                load_tile<BN, LDSB, ALIGN16>(stageB(nxt % STAGES), p.B, p.ldb, n0, p.N, kbeg + nxt * BK, p.K);
            }
            cp_async_commit();
        }
        const double *sA = stageA(kt % STAGES) + wm * 32 + g;
        const double *sB = stageB(kt % STAGES) + wn * 32 + g;
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; k4++) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = sA[(k4 * 4 + tg) * LDSA + i * 8];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = sB[(k4 * 4 + tg) * LDSB + j * 8];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread holds C[row = g (+8i), col = 2*tg + e (+8j)]
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int row = m0 + wm * 32 + i * 8 + g;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int colc = n0 + wn * 32 + j * 8 + tg * 2 + e;
                if (colc >= p.N) continue;
                if (p.lower && colc > row) continue;
                double *c = p.C + row + (size_t)colc * p.ldc;
                double v = p.alpha * acc[i][j][e];
                if (p.beta != 0.0) v += p.beta * (*c);
                *c = v;
            }
        }
    }
}

static int launch_gemm(frk_ctx *ctx, const GemmArgs &a) {
    if (a.M <= 0 || a.N <= 0) return 0;
    static bool attr = false;
    if (!attr) {
        FRK_CHECK_CUDA(cudaFuncSetAttribute(dgemm_nt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
        FRK_CHECK_CUDA(cudaFuncSetAttribute(dgemm_nt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
        attr = true;
    }
    dim3 grid((unsigned)cdiv(a.M, BM), (unsigned)cdiv(a.N, BN));
    // algorithmic flops: full product 2MNK; lower-triangle SYRK M(M+1)K; upper-triangular Y Y' (lower half) ~ M^3/3
    double fl = 2.0 * a.M * a.N * a.K;
    if (a.lower) fl = a.kstart_m0 ? (double)a.M * a.M * a.M / 3.0 : (double)a.M * (a.M + 1.0) * a.K;
    FRK_WORK(ctx, fl, 0);
    bool al = ((uintptr_t)a.A % 16 == 0) && ((uintptr_t)a.B % 16 == 0) && (a.lda % 2 == 0) && (a.ldb % 2 == 0);
    if (al) { FRK_LAUNCH(ctx, dgemm_nt_kernel<true>, grid, GEMM_T, GEMM_SMEM, a); }
    else    { double f2 = ctx->next_flops; FRK_WORK(ctx, f2, 0); FRK_LAUNCH(ctx, dgemm_nt_kernel<false>, grid, GEMM_T, GEMM_SMEM, a); }
    return 0;
}

extern "C" int frk_dgemm_nt(frk_ctx *ctx, int M, int N, int K, double alpha, const double *d_A, int lda, const double *d_B,
                            int ldb, double beta, double *d_C, int ldc) {
    FRK_REQUIRE(ctx && d_A && d_B && d_C, "frk_dgemm_nt: NULL argument");
    FRK_REQUIRE(lda >= M && ldb >= N && ldc >= M && K >= 0, "frk_dgemm_nt: bad leading dimension");
    GemmArgs a{M, N, K, alpha, beta, d_A, lda, d_B, ldb, d_C, ldc, 0, 0};
    return launch_gemm(ctx, a);
}

// ---------------------------------------------------------------------------------
// diagonal blocks: Cholesky of an NB x NB block and inverse of its factor, one CTA each
// ---------------------------------------------------------------------------------
constexpr int NB = 64;
constexpr int DLD = NB + 1;
constexpr size_t DIAG_SMEM = 2 * NB * DLD * sizeof(double);

static int diag_attrs() {
    static bool done = false;
    if (!done) {
        extern __global__ void potf2_inv_kernel(double *, int, int, int, double *, double *);
        extern __global__ void trinv_blocks_kernel(const double *, int, int, double *);
        FRK_CHECK_CUDA(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
        FRK_CHECK_CUDA(cudaFuncSetAttribute(trinv_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
        done = true;
    }
    return 0;
}

// status[0] += sum(log(diag L)); status[1] = first failing (1-based) column, 0 if ok
__global__ void __launch_bounds__(256) potf2_inv_kernel(double *__restrict__ A, int lda, int kb, int col_offset,
                                                        double *__restrict__ invL /*NB x NB, ld NB*/, double *__restrict__ status) {
    extern __shared__ __align__(16) double dsm[];
    double *a = dsm, *x = dsm + NB * DLD;
    __shared__ int bad;
    const int tid = threadIdx.x;
    if (tid == 0) bad = 0;
    for (int idx = tid; idx < NB * NB; idx += 256) {
        int i = idx % NB, j = idx / NB;
        a[i * DLD + j] = (i < kb && j < kb && j <= i) ? A[i + (size_t)j * lda] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int j = 0; j < kb; j++) {
        const double djj = a[j * DLD + j];
        if (!(djj > 0.0)) {          // not positive definite (also catches NaN)
            if (tid == 0 && !bad) bad = col_offset + j + 1;
            break;                    // uniform: every thread reads the same djj
        }
        const double d = sqrt(djj);
        __syncthreads();
        if (tid == 0) a[j * DLD + j] = d;
        for (int i = j + 1 + tid; i < kb; i += 256) a[i * DLD + j] = a[i * DLD + j] / d;
        __syncthreads();
        // trailing update of the lower triangle: a[i][k] -= a[i][j]*a[k][j], j<k<=i
        const int nrem = kb - j - 1;
        for (int idx = tid; idx < nrem * nrem; idx += 256) {
            int i = j + 1 + idx / nrem, k = j + 1 + idx % nrem;
            if (k <= i) a[i * DLD + k] -= a[i * DLD + j] * a[k * DLD + j];
        }
        __syncthreads();
    }
    __syncthreads();
    if (bad) {
        if (tid == 0 && status[1] == 0.0) status[1] = (double)bad;
        return;
    }
    // write L back (lower triangle only)
    for (int idx = tid; idx < NB * NB; idx += 256) {
        int i = idx % NB, j = idx / NB;
        if (i < kb && j < kb && j <= i) A[i + (size_t)j * lda] = a[i * DLD + j];
    }
    // inverse: thread t < NB solves L x = e_t by forward substitution; x[k][t]
    if (tid < NB) {
        const int t = tid;
        for (int i = 0; i < NB; i++) {
            double s = (i == t) ? 1.0 : 0.0;
            for (int k = 0; k < i; k++) s -= a[i * DLD + k] * x[k * DLD + t];
            x[i * DLD + t] = (i < t) ? 0.0 : s / a[i * DLD + i];
        }
    }
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += 256) {
        int i = idx % NB, j = idx / NB;
        invL[i + j * NB] = (i < kb && j < kb) ? x[i * DLD + j] : 0.0;
    }
    if (tid == 0) {
        double s = 0.0;
        for (int j = 0; j < kb; j++) s += log(a[j * DLD + j]);
        status[0] += s;
    }
}

// batched inverse of all diagonal blocks of a factor L (for frk_potri)
__global__ void __launch_bounds__(256) trinv_blocks_kernel(const double *__restrict__ L, int ld, int r, double *__restrict__ invL_all) {
    extern __shared__ __align__(16) double dsm[];
    double *a = dsm, *x = dsm + NB * DLD;
    const int tid = threadIdx.x, o = blockIdx.x * NB, kb = min(NB, r - o);
    for (int idx = tid; idx < NB * NB; idx += 256) {
        int i = idx % NB, j = idx / NB;
        a[i * DLD + j] = (i < kb && j < kb && j <= i) ? L[(o + i) + (size_t)(o + j) * ld] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    if (tid < NB) {
        const int t = tid;
        for (int i = 0; i < NB; i++) {
            double s = (i == t) ? 1.0 : 0.0;
            for (int k = 0; k < i; k++) s -= a[i * DLD + k] * x[k * DLD + t];
            x[i * DLD + t] = (i < t) ? 0.0 : s / a[i * DLD + i];
        }
    }
    __syncthreads();
    double *out = invL_all + (size_t)blockIdx.x * NB * NB;
    for (int idx = tid; idx < NB * NB; idx += 256) {
        int i = idx % NB, j = idx / NB;
        out[i + j * NB] = (i < kb && j < kb) ? x[i * DLD + j] : 0.0;
    }
}

__global__ void set_identity_kernel(int r, double *__restrict__ Y) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (int64_t)r * r; idx += (int64_t)gridDim.x * blockDim.x) {
        int i = (int)(idx % r), j = (int)(idx / r);
        Y[idx] = (i == j) ? 1.0 : 0.0;
    }
}

// mirror: A[j,i] = A[i,j] for i > j (32x32 tiles through shared memory, coalesced both ways)
__global__ void __launch_bounds__(256) mirror_lower_kernel(int r, double *__restrict__ A) {
    __shared__ double t[32][33];
    const int bi = blockIdx.x, bj = blockIdx.y;
    if (bj > bi) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int jj = ty; jj < 32; jj += 8) {
        int i = bi * 32 + tx, j = bj * 32 + jj;
        t[jj][tx] = (i < r && j < r) ? A[i + (size_t)j * r] : 0.0;
    }
    __syncthreads();
    for (int ii = ty; ii < 32; ii += 8) {
        int i = bi * 32 + ii, j = bj * 32 + tx;     // write element (j, i) <- (i, j)
        if (i < r && j < r && i > j) A[j + (size_t)i * r] = t[tx][ii];
    }
}

extern "C" int frk_potrf(frk_ctx *ctx, int r, double *d_A, double *h_logdet) {
    FRK_REQUIRE(ctx && d_A && r > 0, "frk_potrf: bad argument");
    const int nb = (int)cdiv(r, NB);
    if (int rc = diag_attrs()) return rc;
    // scratch: W (r x NB, ld r) | invL (NB x NB) ; status lives in d_small
    void *raw = nullptr;
    if (int rc = frk_scratch(ctx, ((size_t)r * NB + NB * NB) * sizeof(double), &raw)) return rc;
    double *W = (double *)raw, *invL = W + (size_t)r * NB;
    double *status = ctx->d_small + 4096 - 8;
    FRK_CHECK_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(double), ctx->stream));
    for (int k = 0; k < nb; k++) {
        const int o = k * NB, kb = std::min(NB, r - o), Mp = r - o - kb;
        FRK_LAUNCH(ctx, potf2_inv_kernel, 1, 256, DIAG_SMEM, d_A + o + (size_t)o * r, r, kb, o, invL, status);
        if (Mp > 0) {
            // W = P * invL'   (panel solve P L^-T)
            GemmArgs g1{Mp, kb, kb, 1.0, 0.0, d_A + (o + kb) + (size_t)o * r, r, invL, NB, W, r, 0, 0};
            if (int rc = launch_gemm(ctx, g1)) return rc;
            FRK_CHECK_CUDA(cudaMemcpy2DAsync(d_A + (o + kb) + (size_t)o * r, (size_t)r * sizeof(double), W, (size_t)r * sizeof(double),
                                             (size_t)Mp * sizeof(double), kb, cudaMemcpyDeviceToDevice, ctx->stream));
            // A22 -= W W'  (lower triangle)
            double *A22 = d_A + (o + kb) + (size_t)(o + kb) * r;
            GemmArgs g2{Mp, Mp, kb, -1.0, 1.0, W, r, W, r, A22, r, 1, 0};
            if (int rc = launch_gemm(ctx, g2)) return rc;
        }
    }
    FRK_CHECK_CUDA(cudaMemcpyAsync(ctx->h_pinned, status, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    FRK_REQUIRE(ctx->h_pinned[1] == 0.0, "the leading minor of order %d is not positive definite", (int)ctx->h_pinned[1]);
    if (h_logdet) *h_logdet = 2.0 * ctx->h_pinned[0];
    return 0;
}

extern "C" int frk_potri(frk_ctx *ctx, int r, const double *d_L, double *d_Ainv, double *d_work) {
    FRK_REQUIRE(ctx && d_L && d_Ainv && d_work && r > 0, "frk_potri: bad argument");
    const int nb = (int)cdiv(r, NB);
    if (int rc = diag_attrs()) return rc;
    void *raw = nullptr;
    if (int rc = frk_scratch(ctx, ((size_t)r * NB + (size_t)nb * NB * NB) * sizeof(double), &raw)) return rc;
    double *T = (double *)raw, *invL = T + (size_t)r * NB;
    double *Y = d_work;
    FRK_LAUNCH(ctx, trinv_blocks_kernel, nb, 256, DIAG_SMEM, d_L, r, r, invL);
    FRK_LAUNCH(ctx, set_identity_kernel, ctx->sm_count * 8, 256, 0, r, Y);
    // Y = L^-T by blocked forward substitution on the rows of L^-1, stored transposed (all products NT)
    for (int j = 0; j < nb; j++) {
        const int o = j * NB, kb = std::min(NB, r - o), Mp = o + kb, Nn = r - Mp;
        GemmArgs g1{Mp, kb, kb, 1.0, 0.0, Y + (size_t)o * r, r, invL + (size_t)j * NB * NB, NB, T, r, 0, 0};
        if (int rc = launch_gemm(ctx, g1)) return rc;
        if (Nn > 0) {
            GemmArgs g2{Mp, Nn, kb, -1.0, 1.0, T, r, d_L + Mp + (size_t)o * r, r, Y + (size_t)Mp * r, r, 0, 0};
            if (int rc = launch_gemm(ctx, g2)) return rc;
        }
        FRK_CHECK_CUDA(cudaMemcpy2DAsync(Y + (size_t)o * r, (size_t)r * sizeof(double), T, (size_t)r * sizeof(double),
                                         (size_t)Mp * sizeof(double), kb, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    // A^-1 = L^-T L^-1 = Y Y'   (Y upper triangular: sum over k >= max(i,j))
    GemmArgs g3{r, r, r, 1.0, 0.0, Y, r, Y, r, d_Ainv, r, 1, 1};
    if (int rc = launch_gemm(ctx, g3)) return rc;
    dim3 grid((unsigned)cdiv(r, 32), (unsigned)cdiv(r, 32));
    FRK_LAUNCH(ctx, mirror_lower_kernel, grid, 256, 0, r, d_Ainv);
    return 0;
}

// ---------------------------------------------------------------------------------
// matrix-vector, element-wise and reduction helpers
// ---------------------------------------------------------------------------------
constexpr int SYMV_SPLIT = 32;

__global__ void __launch_bounds__(128) symv_partial_kernel(int r, const double *__restrict__ A, const double *__restrict__ x,
                                                           double *__restrict__ part) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    const int jchunk = (r + SYMV_SPLIT - 1) / SYMV_SPLIT;
    const int j0 = blockIdx.y * jchunk, j1 = min(r, j0 + jchunk);
    if (i >= r) return;
    double s = 0.0;
    for (int j = j0; j < j1; j++) s += A[i + (size_t)j * r] * __ldg(x + j);
    part[(size_t)blockIdx.y * r + i] = s;
}

__global__ void symv_final_kernel(int r, const double *__restrict__ part, double *__restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r) return;
    double s = 0.0;
    for (int k = 0; k < SYMV_SPLIT; k++) s += part[(size_t)k * r + i];
    y[i] = s;
}

extern "C" int frk_symv(frk_ctx *ctx, int r, const double *d_A, const double *d_x, double *d_y) {
    FRK_REQUIRE(ctx && d_A && d_x && d_y && r > 0, "frk_symv: bad argument");
    void *raw = nullptr;
    if (int rc = frk_scratch(ctx, (size_t)SYMV_SPLIT * r * sizeof(double), &raw)) return rc;
    dim3 grid((unsigned)cdiv(r, 128), SYMV_SPLIT);
    FRK_LAUNCH(ctx, symv_partial_kernel, grid, 128, 0, r, d_A, d_x, (double *)raw);
    FRK_LAUNCH(ctx, symv_final_kernel, (unsigned)cdiv(r, 128), 128, 0, r, (const double *)raw, d_y);
    return 0;
}

__global__ void __launch_bounds__(256) sym_from_upper_add_kernel(int r, const double *__restrict__ G, const double *__restrict__ B,
                                                                 double *__restrict__ out) {
    __shared__ double t[32][33];
    const int bi = blockIdx.x, bj = blockIdx.y;       // tile rows bi, cols bj
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const bool up = bi <= bj;                          // tile lies in (or on) the upper triangle: read directly
    // for lower tiles read the transposed upper tile (bj, bi)
    for (int jj = ty; jj < 32; jj += 8) {
        int i = (up ? bi : bj) * 32 + tx, j = (up ? bj : bi) * 32 + jj;
        t[jj][tx] = (i < r && j < r) ? G[i + (size_t)j * r] : 0.0;
    }
    __syncthreads();
    for (int jj = ty; jj < 32; jj += 8) {
        int i = bi * 32 + tx, j = bj * 32 + jj;
        if (i < r && j < r) {
            double v;
            if (bi < bj) v = t[jj][tx];
            else if (bi > bj) v = t[tx][jj];
            else v = (i <= j) ? t[jj][tx] : t[tx][jj];
            out[i + (size_t)j * r] = B ? v + B[i + (size_t)j * r] : v;
        }
    }
}

extern "C" int frk_sym_from_upper_add(frk_ctx *ctx, int r, const double *d_Gupper, const double *d_B, double *d_out) {
    FRK_REQUIRE(ctx && d_Gupper && d_out && r > 0, "frk_sym_from_upper_add: bad argument");
    FRK_REQUIRE(d_out != d_Gupper, "frk_sym_from_upper_add: out must not alias G");
    dim3 grid((unsigned)cdiv(r, 32), (unsigned)cdiv(r, 32));
    FRK_LAUNCH(ctx, sym_from_upper_add_kernel, grid, 256, 0, r, d_Gupper, d_B, d_out);
    return 0;
}

__global__ void add_outer_kernel(int r, const double *__restrict__ A, const double *__restrict__ x, double *__restrict__ out) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (int64_t)r * r; idx += (int64_t)gridDim.x * blockDim.x) {
        int i = (int)(idx % r), j = (int)(idx / r);
        out[idx] = A[idx] + x[i] * x[j];
    }
}

extern "C" int frk_add_outer(frk_ctx *ctx, int r, const double *d_A, const double *d_x, double *d_out) {
    FRK_REQUIRE(ctx && d_A && d_x && d_out && r > 0, "frk_add_outer: bad argument");
    FRK_LAUNCH(ctx, add_outer_kernel, ctx->sm_count * 8, 256, 0, r, d_A, d_x, d_out);
    return 0;
}

__global__ void exp_kernel_kernel(int64_t n, const double *__restrict__ D, double tau, double scale, double *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = exp(-D[i] / tau) * scale;
}

extern "C" int frk_exp_kernel(frk_ctx *ctx, int64_t n, const double *d_D, double tau, double scale, double *d_out) {
    FRK_REQUIRE(ctx && d_D && d_out, "frk_exp_kernel: NULL argument");
    if (n == 0) return 0;
    FRK_LAUNCH(ctx, exp_kernel_kernel, ctx->sm_count * 8, 256, 0, n, d_D, tau, scale, d_out);
    return 0;
}

__global__ void __launch_bounds__(256) dot_partial_kernel(int64_t n, const double *__restrict__ A, const double *__restrict__ B,
                                                          double *__restrict__ part) {
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += A[i] * B[i];
    __shared__ double sh[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0;
        for (int w = 0; w < 8; w++) v += sh[w];
        part[blockIdx.x] = v;
    }
}

__global__ void dot_final_kernel(int nblocks, const double *__restrict__ part, double *__restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double v = 0;
        for (int b = 0; b < nblocks; b++) v += part[b];
        *out = v;
    }
}

extern "C" int frk_dot(frk_ctx *ctx, int64_t n, const double *d_A, const double *d_B, double *h_out) {
    FRK_REQUIRE(ctx && d_A && d_B && h_out, "frk_dot: NULL argument");
    if (n == 0) { *h_out = 0.0; return 0; }
    const int nblocks = (int)std::min<int64_t>(cdiv(n, 256), 1024);
    double *part = ctx->d_small, *out = ctx->d_small + 2048;
    FRK_LAUNCH(ctx, dot_partial_kernel, nblocks, 256, 0, n, d_A, d_B, part);
    FRK_LAUNCH(ctx, dot_final_kernel, 1, 32, 0, nblocks, part, out);
    FRK_CHECK_CUDA(cudaMemcpyAsync(ctx->h_pinned, out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    *h_out = ctx->h_pinned[0];
    return 0;
}

__global__ void block_gather_kernel(int r, const double *__restrict__ A, const double *__restrict__ x, int ni,
                                    const int *__restrict__ idx, double *__restrict__ out) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)ni * ni; t += (int64_t)gridDim.x * blockDim.x) {
        int i = (int)(t % ni), j = (int)(t / ni);
        int gi = idx[i], gj = idx[j];
        double v = A[gi + (size_t)gj * r];
        if (x) v += x[gi] * x[gj];
        out[t] = v;
    }
}

__global__ void block_scatter_kernel(int r, double *__restrict__ A, int ni, const int *__restrict__ idx, const double *__restrict__ in) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)ni * ni; t += (int64_t)gridDim.x * blockDim.x) {
        int i = (int)(t % ni), j = (int)(t / ni);
        A[idx[i] + (size_t)idx[j] * r] = in[t];
    }
}

extern "C" int frk_block_gather(frk_ctx *ctx, int r, const double *d_A, int ni, const int *d_idx, double *d_out) {
    FRK_REQUIRE(ctx && d_A && d_idx && d_out, "frk_block_gather: NULL argument");
    if (ni == 0) return 0;
    FRK_LAUNCH(ctx, block_gather_kernel, ctx->sm_count * 8, 256, 0, r, d_A, (const double *)nullptr, ni, d_idx, d_out);
    return 0;
}

extern "C" int frk_block_gather_outer(frk_ctx *ctx, int r, const double *d_A, const double *d_x, int ni, const int *d_idx,
                                      double *d_out) {
    FRK_REQUIRE(ctx && d_A && d_x && d_idx && d_out, "frk_block_gather_outer: NULL argument");
    if (ni == 0) return 0;
    FRK_LAUNCH(ctx, block_gather_kernel, ctx->sm_count * 8, 256, 0, r, d_A, d_x, ni, d_idx, d_out);
    return 0;
}

extern "C" int frk_block_scatter(frk_ctx *ctx, int r, double *d_A, int ni, const int *d_idx, const double *d_in) {
    FRK_REQUIRE(ctx && d_A && d_idx && d_in, "frk_block_scatter: NULL argument");
    if (ni == 0) return 0;
    FRK_LAUNCH(ctx, block_scatter_kernel, ctx->sm_count * 8, 256, 0, r, d_A, ni, d_idx, d_in);
    return 0;
}

// x = Y' b for upper-triangular Y (= L^-T left in d_work by frk_potri): x = L^-1 b, i.e.
// (rDinv %*% S) %*% solve(R) of .loglik.ind (R/SREfit.R:211).  One warp per column of Y.
__global__ void __launch_bounds__(256) upper_tmulv_kernel(int r, const double *__restrict__ Y, const double *__restrict__ b,
                                                          double *__restrict__ x) {
    const int i = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (i >= r) return;
    const double *y = Y + (size_t)i * r;
    double s = 0.0;
    for (int k = lane; k <= i; k += 32) s += y[k] * __ldg(b + k);
    s = warp_sum(s);
    if (lane == 0) x[i] = s;
}

extern "C" int frk_upper_tmulv(frk_ctx *ctx, int r, const double *d_Y, const double *d_b, double *d_x) {
    FRK_REQUIRE(ctx && d_Y && d_b && d_x && r > 0, "frk_upper_tmulv: bad argument");
    FRK_LAUNCH(ctx, upper_tmulv_kernel, (unsigned)cdiv((int64_t)r * 32, 256), 256, 0, r, d_Y, d_b, d_x);
    return 0;
}

// A = value * I   (.EM_initialise: S_eta = Q_eta = I, Khat = var(Z) I, Khat_inv = I / var(Z); R/SRE.R:710-722)
__global__ void scaled_identity_kernel(int r, double value, double *__restrict__ A) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (int64_t)r * r; idx += (int64_t)gridDim.x * blockDim.x) {
        int i = (int)(idx % r), j = (int)(idx / r);
        A[idx] = (i == j) ? value : 0.0;
    }
}

extern "C" int frk_scaled_identity(frk_ctx *ctx, int r, double value, double *d_A) {
    FRK_REQUIRE(ctx && d_A && r > 0, "frk_scaled_identity: bad argument");
    FRK_LAUNCH(ctx, scaled_identity_kernel, ctx->sm_count * 8, 256, 0, r, value, d_A);
    return 0;
}

// diag(A) += d   (ridge term of .regularise_K, R/SREfit.R:669-686)
__global__ void add_diag_kernel(int r, const double *__restrict__ d, double *__restrict__ A) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < r) A[i + (size_t)i * r] += d[i];
}

extern "C" int frk_add_diag(frk_ctx *ctx, int r, const double *d_diag, double *d_A) {
    FRK_REQUIRE(ctx && d_A && d_diag && r > 0, "frk_add_diag: bad argument");
    FRK_LAUNCH(ctx, add_diag_kernel, (unsigned)cdiv(r, 256), 256, 0, r, d_diag, d_A);
    return 0;
}

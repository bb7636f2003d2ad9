// Dense fp64 linear algebra for the r x r systems of the EM algorithm.
//
// Reference semantics: Matrix::chol / chol2inv / determinant as used in R/SREfit.R:177-198
// (.loglik.ind), :252-254 (.SRE.Estep.ind), :273 (Khat_inv), :475-478,517-524 (block-exponential
// f_tau); R/linalgfns.R:21-36 (tr, diag2, logdet).
//
// There is no fp64 kind of tcgen05.mma, so fp64 tensor-core work on sm_100a is mma.sync DMMA
// (m8n8k4).  One GEMM kernel, templated on which index of each operand is contiguous in memory,
// with triangular k-range hints and a batch dimension, carries all the flops:
//   potrf  two-level right-looking Cholesky: 64-wide inner steps (register-resident 64x64 potf2,
//          thread-per-row panel trsm, rank-64 update inside the 256-wide outer panel) and one
//          rank-256 lower SYRK per outer panel;
//   potri  L^-1 by recursive doubling over the diagonal (64 -> 128 -> ... ; two batched GEMMs per
//          level), then A^-1 = L^-T L^-1 as one triangular-aware product, mirrored.
#include "common.cuh"
#include <algorithm>
#include <cmath>

// ---------------------------------------------------------------------------------
// GEMM: C = beta*C + alpha * op(A) op(B),  op(A) M x K, op(B) K x N, C column-major.
//   A_KC = false: A stored M x K column-major (m contiguous);  true: stored K x M (k contiguous)
//   B_KC = false: B stored N x K column-major (n contiguous, "NT"); true: stored K x N (k contiguous)
// ---------------------------------------------------------------------------------
constexpr int BM = 128, BN = 64, BK = 16, GEMM_T = 256, STAGES = 3;
constexpr int LDSA = BM + 4, LDSB = BN + 4;   // [k][m] layouts: = 4 (mod 16) -> conflict-free 8-byte fragment loads
constexpr int LDSK = BK + 4;                  // [m][k] layouts: 20 -> the 32 fragment addresses cover the 16 bank pairs twice
constexpr int A_STAGE = (BK * LDSA > BM * LDSK) ? BK * LDSA : BM * LDSK;
constexpr int B_STAGE = (BK * LDSB > BN * LDSK) ? BK * LDSB : BN * LDSK;
constexpr int STAGE_DOUBLES = A_STAGE + B_STAGE;
constexpr size_t GEMM_SMEM = (size_t)STAGES * STAGE_DOUBLES * sizeof(double);

struct GemmArgs {
    int M, N, K;
    double alpha, beta;
    const double *A; int lda;
    const double *B; int ldb;
    double *C; int ldc;
    int lower;                       // only tiles touching the lower triangle; elements above the diagonal are not stored
    int lower_off;                   // the diagonal sits lower_off columns to the right of C's origin: keep col <= row + lower_off
    int k_from_m, k_from_n, k_to_m;  // triangular operands: start k at the tile's first row / column, stop at its last row
    long long sA, sB, sC;            // batch strides (elements) for blockIdx.z
    int nbatch, M_last, K_last;      // the last batch entry may be ragged: its M and K
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

// ROWS x BK tile -> shared memory.  KC = false: operand element (row, k) at g[row + k*ld], smem [k][row] (ld LDS);
// KC = true: element at g[k + row*ld], smem [row][k] (ld LDSK).
template <int ROWS, int LDS, bool KC, bool ALIGN16>
__device__ __forceinline__ void load_tile(double *s, const double *__restrict__ g, int ld, int row0, int nrows, int k0, int K) {
    if (!KC) {
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
    } else {
        if (ALIGN16) {
            constexpr int PER_ROW = BK / 2;
#pragma unroll
            for (int it = 0; it < (ROWS * PER_ROW) / GEMM_T; it++) {
                int idx = threadIdx.x + it * GEMM_T;
                int r1 = idx / PER_ROW, k2 = (idx % PER_ROW) * 2;
                int row = row0 + r1, k = k0 + k2;
                int bytes = (row < nrows) ? max(0, min(16, (K - k) * 8)) : 0;
                const double *src = bytes ? g + k + (size_t)row * ld : g;
                cp_async16(s + r1 * LDSK + k2, src, bytes);
            }
        } else {
#pragma unroll
            for (int it = 0; it < (ROWS * BK) / GEMM_T; it++) {
                int idx = threadIdx.x + it * GEMM_T;
                int r1 = idx / BK, kk = idx % BK;
                int row = row0 + r1, k = k0 + kk;
                int bytes = (k < K && row < nrows) ? 8 : 0;
                const double *src = bytes ? g + k + (size_t)row * ld : g;
                cp_async8(s + r1 * LDSK + kk, src, bytes);
            }
        }
    }
}

// BM_ = 128 (default) or 64: the small tile spreads skinny / small products (panel updates) over twice as many SMs
template <bool A_KC, bool B_KC, bool ALIGN16, int BM_>
__global__ void __launch_bounds__(GEMM_T, 2) dgemm_kernel(GemmArgs p) {
    constexpr int WM = BM_ / 32, WN = 8 / WM, WTN = BN / WN, NJ = WTN / 8;   // warps along m / n, warp tile width, n-fragments per warp
    extern __shared__ __align__(16) double gsm[];
    // tiles whose k-range ends at their last row (triangular A) are heaviest at the bottom: schedule those first
    const int m0 = (p.k_to_m ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * BM_, n0 = blockIdx.y * BN;
    if (p.lower && n0 >= m0 + BM_ + p.lower_off) return;
    const int z = blockIdx.z;
    const int M = (z == p.nbatch - 1) ? p.M_last : p.M;
    const int K = (z == p.nbatch - 1) ? p.K_last : p.K;
    if (m0 >= M) return;
    const double *__restrict__ A = p.A + (size_t)z * p.sA;
    const double *__restrict__ B = p.B + (size_t)z * p.sB;
    double *__restrict__ Cm = p.C + (size_t)z * p.sC;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int wm = wid % WM, wn = wid / WM;
    const int g = lane >> 2, tg = lane & 3;

    int kbeg = 0;
    if (p.k_from_m) kbeg = max(kbeg, m0);
    if (p.k_from_n) kbeg = max(kbeg, n0);
    kbeg = (kbeg / BK) * BK;
    const int kend = p.k_to_m ? min(K, m0 + BM_) : K;
    const int nk = kend > kbeg ? (kend - kbeg + BK - 1) / BK : 0;

    double acc[4][NJ][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto stageA = [&](int s) { return gsm + (size_t)s * STAGE_DOUBLES; };
    auto stageB = [&](int s) { return gsm + (size_t)s * STAGE_DOUBLES + A_STAGE; };

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nk) {
            load_tile<BM_, LDSA, A_KC, ALIGN16>(stageA(s), A, p.lda, m0, M, kbeg + s * BK, kend);
            load_tile<BN, LDSB, B_KC, ALIGN16>(stageB(s), B, p.ldb, n0, p.N, kbeg + s * BK, kend);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nxt = kt + STAGES - 1;
            if (nxt < nk) {
                load_tile<BM_, LDSA, A_KC, ALIGN16>(stageA(nxt % STAGES), A, p.lda, m0, M, kbeg + nxt * BK, kend);
                load_tile<BN, LDSB, B_KC, ALIGN16>(stageB(nxt % STAGES), B, p.ldb, n0, p.N, kbeg + nxt * BK, kend);
            }
            cp_async_commit();
        }
        const double *sA = stageA(kt % STAGES);
        const double *sB = stageB(kt % STAGES);
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; k4++) {
            double a[4], b[NJ];
#pragma unroll
            for (int i = 0; i < 4; i++)
                a[i] = A_KC ? sA[(wm * 32 + i * 8 + g) * LDSK + k4 * 4 + tg] : sA[(k4 * 4 + tg) * LDSA + wm * 32 + i * 8 + g];
#pragma unroll
            for (int j = 0; j < NJ; j++)
                b[j] = B_KC ? sB[(wn * WTN + j * 8 + g) * LDSK + k4 * 4 + tg] : sB[(k4 * 4 + tg) * LDSB + wn * WTN + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < NJ; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread holds C[row = g (+8i), col = 2*tg + e (+8j)]
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int row = m0 + wm * 32 + i * 8 + g;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < NJ; j++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int colc = n0 + wn * WTN + j * 8 + tg * 2 + e;
                if (colc >= p.N) continue;
                if (p.lower && colc > row + p.lower_off) continue;
                double *c = Cm + row + (size_t)colc * p.ldc;
                double v = p.alpha * acc[i][j][e];
                if (p.beta != 0.0) v += p.beta * (*c);
                *c = v;
            }
        }
    }
}

template <bool A_KC, bool B_KC>
static int launch_gemm_t(frk_ctx *ctx, const GemmArgs &a, bool al) {
    // 64-row tiles when 128-row tiles would leave SMs idle or make a launch as long as its longest tile (fewer than 4 tiles per SM):
    // skinny panel updates, small matrices, the doubling levels of frk_potri (stand-alone 1.40 -> 1.22 ms at n = 3276).  With one tile
    // per SM as the limit (FRK_GEMM_SMALL_PER_SM=1) the EM iteration at C3 is 2.6 % slower now that the E-step's chain runs alone.
    const int64_t ctas128 = cdiv(a.M, 128) * cdiv(a.N, BN) * (int64_t)a.nbatch;
    static const int small_per_sm = getenv("FRK_GEMM_SMALL_PER_SM") ? std::max(atoi(getenv("FRK_GEMM_SMALL_PER_SM")), 1) : 4;
    const bool small = ctas128 < (int64_t)small_per_sm * ctx->sm_count;
    const int bm = small ? 64 : 128;
    dim3 grid((unsigned)cdiv(a.M, bm), (unsigned)cdiv(a.N, BN), (unsigned)a.nbatch);
    const double fl = ctx->next_flops;
    const char *tg = ctx->next_tag;
    if (al && !small)       { FRK_LAUNCH(ctx, (dgemm_kernel<A_KC, B_KC, true, 128>), grid, GEMM_T, GEMM_SMEM, a); }
    else if (al && small)   { FRK_LAUNCH(ctx, (dgemm_kernel<A_KC, B_KC, true, 64>), grid, GEMM_T, GEMM_SMEM, a); }
    else if (!small)        { FRK_WORK(ctx, fl, 0); FRK_TAG(ctx, tg); FRK_LAUNCH(ctx, (dgemm_kernel<A_KC, B_KC, false, 128>), grid, GEMM_T, GEMM_SMEM, a); }
    else                    { FRK_WORK(ctx, fl, 0); FRK_TAG(ctx, tg); FRK_LAUNCH(ctx, (dgemm_kernel<A_KC, B_KC, false, 64>), grid, GEMM_T, GEMM_SMEM, a); }
    return 0;
}

static int gemm_set_attrs();

// flops: the ALGORITHMIC count of the operation this launch performs (passed by the caller)
static int launch_gemm(frk_ctx *ctx, GemmArgs a, bool a_kc, bool b_kc, double flops, const char *tag = "dgemm_kernel") {
    if (a.M <= 0 || a.N <= 0) return 0;
    if (int rc = gemm_set_attrs()) return rc;
    if (a.nbatch <= 0) { a.nbatch = 1; a.M_last = a.M; a.K_last = a.K; }
    FRK_WORK(ctx, flops, 0);
    FRK_TAG(ctx, tag);
    bool al = ((uintptr_t)a.A % 16 == 0) && ((uintptr_t)a.B % 16 == 0) && (a.lda % 2 == 0) && (a.ldb % 2 == 0) &&
              (a.sA % 2 == 0) && (a.sB % 2 == 0);
    if (!a_kc && !b_kc) return launch_gemm_t<false, false>(ctx, a, al);
    if (!a_kc && b_kc) return launch_gemm_t<false, true>(ctx, a, al);
    if (a_kc && !b_kc) return launch_gemm_t<true, false>(ctx, a, al);
    return launch_gemm_t<true, true>(ctx, a, al);
}

static GemmArgs gemm_args(int M, int N, int K, double alpha, const double *A, int lda, const double *B, int ldb, double beta,
                          double *C, int ldc) {
    GemmArgs a{};
    a.M = M; a.N = N; a.K = K; a.alpha = alpha; a.beta = beta; a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.C = C; a.ldc = ldc;
    a.nbatch = 1; a.M_last = M; a.K_last = K;
    return a;
}

extern "C" int frk_dgemm_nt(frk_ctx *ctx, int M, int N, int K, double alpha, const double *d_A, int lda, const double *d_B,
                            int ldb, double beta, double *d_C, int ldc) {
    FRK_REQUIRE(ctx && d_A && d_B && d_C, "frk_dgemm_nt: NULL argument");
    FRK_REQUIRE(lda >= M && ldb >= N && ldc >= M && K >= 0, "frk_dgemm_nt: bad leading dimension");
    return launch_gemm(ctx, gemm_args(M, N, K, alpha, d_A, lda, d_B, ldb, beta, d_C, ldc), false, false, 2.0 * M * N * K);
}

// general layouts (exported for the tests): transa/transb != 0 -> the operand is stored with k contiguous
__global__ void mirror_lower_kernel(int r, double *__restrict__ A);

// C(lower triangle) += A B'  for a product known to be symmetric (A diag(w) A' with B = diag(w)-scaled A): half the tiles
int frk_dsyrk_like_lower(frk_ctx *ctx, int n, int K, const double *d_A, int lda, const double *d_B, int ldb, double *d_C, int ldc) {
    GemmArgs g = gemm_args(n, n, K, 1.0, d_A, lda, d_B, ldb, 1.0, d_C, ldc);
    g.lower = 1;
    return launch_gemm(ctx, g, false, false, (double)n * (n + 1.0) * K, "dgemm_kernel[Gram A diag(w) A']");
}

// C (n x N) = X B with X lower triangular n x n (column-major) and B stored K x N with k contiguous: k <= the tile's last row
int frk_dtrmm_lower(frk_ctx *ctx, int n, int N, const double *d_X, int ldx, const double *d_B, int ldb, double *d_C, int ldc) {
    GemmArgs g = gemm_args(n, N, n, 1.0, d_X, ldx, d_B, ldb, 0.0, d_C, ldc);
    g.k_to_m = 1;
    return launch_gemm(ctx, g, false, true, (double)n * n * N, "dgemm_kernel[L^-1 A]");
}

int frk_mirror_lower(frk_ctx *ctx, int r, double *d_A) {
    FRK_LAUNCH(ctx, mirror_lower_kernel, dim3((unsigned)cdiv(r, 32), (unsigned)cdiv(r, 32)), 256, 0, r, d_A);
    return 0;
}

extern "C" int frk_dgemm(frk_ctx *ctx, int transa, int transb, int M, int N, int K, double alpha, const double *d_A, int lda,
                         const double *d_B, int ldb, double beta, double *d_C, int ldc) {
    FRK_REQUIRE(ctx && d_A && d_B && d_C, "frk_dgemm: NULL argument");
    FRK_REQUIRE(lda >= (transa ? K : M) && ldb >= (transb ? K : N) && ldc >= M && K >= 0, "frk_dgemm: bad leading dimension");
    return launch_gemm(ctx, gemm_args(M, N, K, alpha, d_A, lda, d_B, ldb, beta, d_C, ldc), transa != 0, transb != 0, 2.0 * M * N * K);
}

// ---------------------------------------------------------------------------------
// 64 x 64 building blocks (one thread per row, the row lives in registers)
// ---------------------------------------------------------------------------------
constexpr int NB = 64;     // inner block
constexpr int NBO = 256;   // outer panel of potrf

// Each row of a 64 x 64 block is held by SPLIT = 4 adjacent lanes, lane h owning the columns k = 4*kk + h
// (16 registers); 256 threads per block.  The interleaved ownership makes the four lanes of a row read four
// consecutive doubles of shared memory (one wavefront, broadcast across rows) and keeps all lanes busy.
constexpr int SPLIT = 4, NBQ = NB / SPLIT, BLK_T = NB * SPLIT;

// Cholesky of the kb x kb diagonal block at A (lower triangle, in place), kb <= 64.
// status[0] += sum(log(diag L)); status[1] = first failing (1-based) column, 0 if ok.
// The block lives in shared memory and is factored in four 16-column steps.  Only the 16 x 16 diagonal sub-block is
// column-sequential: one warp holds it in registers (lane = row) and exchanges columns by shuffles -- no barrier and no
// shared-memory round trip on the per-column critical path (pivot -> rsqrt -> scale -> update).  The rows below it are
// solved one thread per row, and the rest of the block is updated by all 256 threads.
constexpr int PB = 16;
constexpr size_t POTF2_SMEM = 2 * NB * (NB + 1) * sizeof(double);
// Optional pre-update (look-ahead): before factoring, a -= Ptop Ptop' with Ptop the kb x kprev block of the panel just solved
// (the rows that face this diagonal block), so that the bulk of the panel's update can run beside this kernel.
__global__ void __launch_bounds__(BLK_T) potf2_kernel(double *__restrict__ A, int lda, int kb, int col_offset, double *__restrict__ status,
                                                      const double *__restrict__ Ptop, int ldp, int kprev) {
    extern __shared__ __align__(16) double potf2_smem[];
    double (*a)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(potf2_smem);
    double (*pt)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(potf2_smem + NB * (NB + 1));     // pt[c][i] = Ptop[i][c]
    __shared__ double ld[PB][PB + 1];      // factor of the current 16 x 16 diagonal sub-block
    __shared__ double rdg[PB];             // reciprocals of its diagonal
    __shared__ int sbad;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    {   // load (lower triangle; identity padding beyond kb); all loads in flight before the stores
        double v[NB * NB / BLK_T];
#pragma unroll
        for (int q = 0; q < NB * NB / BLK_T; q++) {
            const int idx = tid + q * BLK_T, i = idx % NB, k = idx / NB;
            v[q] = (i < kb && k < kb && k <= i) ? A[i + (size_t)k * lda] : (i == k ? 1.0 : 0.0);
        }
#pragma unroll
        for (int q = 0; q < NB * NB / BLK_T; q++) {
            const int idx = tid + q * BLK_T, i = idx % NB, k = idx / NB;
            a[i][k] = v[q];
        }
    }
    if (tid == 0) sbad = 0;
    __syncthreads();
    if (Ptop) {                                                    // a[i][k] -= sum_c Ptop[i][c] Ptop[k][c], k <= i
        // Ptop is kb x kprev; staged 64 columns at a time with all loads in flight, then 16 x 16 threads, 4 x 4 blocks
        // (blocks above the diagonal skipped)
        const int ty = tid >> 4, tx = tid & 15;
        double acc[4][4];
#pragma unroll
        for (int ii = 0; ii < 4; ii++)
#pragma unroll
            for (int kk = 0; kk < 4; kk++) acc[ii][kk] = 0.0;
        for (int c0 = 0; c0 < kprev; c0 += NB) {
            const int kc = min(NB, kprev - c0);
            {
                double v[NB * NB / BLK_T];
#pragma unroll
                for (int q = 0; q < NB * NB / BLK_T; q++) {
                    const int idx = tid + q * BLK_T, i = idx % NB, c = idx / NB;
                    v[q] = (i < kb && c < kc) ? Ptop[i + (size_t)(c0 + c) * ldp] : 0.0;
                }
                __syncthreads();                                   // previous chunk fully consumed
#pragma unroll
                for (int q = 0; q < NB * NB / BLK_T; q++) {
                    const int idx = tid + q * BLK_T, i = idx % NB, c = idx / NB;
                    pt[c][i] = v[q];
                }
            }
            __syncthreads();
            if (tx <= ty) {
#pragma unroll 4
                for (int c = 0; c < kc; c++) {
                    double xi[4], xk[4];
#pragma unroll
                    for (int ii = 0; ii < 4; ii++) xi[ii] = pt[c][4 * ty + ii];
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) xk[kk] = pt[c][4 * tx + kk];
#pragma unroll
                    for (int ii = 0; ii < 4; ii++)
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) acc[ii][kk] += xi[ii] * xk[kk];
                }
            }
        }
        if (tx <= ty) {
#pragma unroll
            for (int ii = 0; ii < 4; ii++)
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    const int i = 4 * ty + ii, k = 4 * tx + kk;
                    if (k <= i) a[i][k] -= acc[ii][kk];
                }
        }
        __syncthreads();
    }
    double lg = 0.0;
    for (int J = 0; J < NB; J += PB) {
        // (a) 16 x 16 diagonal sub-block: warp 0, lane = row, registers + shuffles
        if (wid == 0) {
            const int i = lane & (PB - 1);
            double d[PB];
#pragma unroll
            for (int k = 0; k < PB; k++) d[k] = a[J + i][J + k];
            int bad = 0;
#pragma unroll
            for (int j = 0; j < PB; j++) {
                const double piv = __shfl_sync(0xffffffffu, d[j], j);
                if (!(piv > 0.0)) { if (!bad) bad = J + j + 1; }       // uniform; also catches NaN
                const double rs = rsqrt(piv);
                const double lij = (i == j) ? piv * rs : d[j] * rs;
                d[j] = lij;
#pragma unroll
                for (int k = j + 1; k < PB; k++) d[k] -= lij * __shfl_sync(0xffffffffu, lij, k);   // entries with k > i are never used
            }
            if (lane < PB) {
#pragma unroll
                for (int k = 0; k < PB; k++) { ld[i][k] = (k <= i) ? d[k] : 0.0; a[J + i][J + k] = (k <= i) ? d[k] : 0.0; }
                double dii = 1.0;
#pragma unroll
                for (int k = 0; k < PB; k++) dii = (k == i) ? d[k] : dii;      // select first: ONE reciprocal, not 16 predicated ones
                rdg[i] = 1.0 / dii;
            }
            if (lane == 0 && bad) sbad = bad;
        }
        __syncthreads();
        if (sbad) break;
        const int nbelow = NB - J - PB;
        if (nbelow <= 0) break;
        // (b) rows below the sub-block: x L_D' = p, one thread per row, right-looking
        if (tid < nbelow) {
            const int row = J + PB + tid;
            double x[PB];
#pragma unroll
            for (int c = 0; c < PB; c++) x[c] = a[row][J + c];
#pragma unroll
            for (int c = 0; c < PB; c++) {
                const double xc = x[c] * rdg[c];
                x[c] = xc;
#pragma unroll
                for (int k = c + 1; k < PB; k++) x[k] -= xc * ld[k][c];
            }
#pragma unroll
            for (int c = 0; c < PB; c++) a[row][J + c] = x[c];
        }
        __syncthreads();
        // (c) trailing update of the block: a[i][k] -= sum_c x[i][c] x[k][c], J+16 <= k <= i < 64; 16 x 16 threads x (<=3 x 3)
        {
            const int ty = tid >> 4, tx = tid & 15, base = J + PB, nt = nbelow / PB;   // nt = 3, 2, 1
            double acc[3][3];
#pragma unroll
            for (int ii = 0; ii < 3; ii++)
#pragma unroll
                for (int kk = 0; kk < 3; kk++) acc[ii][kk] = 0.0;
#pragma unroll 4
            for (int c = 0; c < PB; c++) {
                double xi[3], xk[3];
#pragma unroll
                for (int ii = 0; ii < 3; ii++) xi[ii] = ii < nt ? a[base + ty + PB * ii][J + c] : 0.0;
#pragma unroll
                for (int kk = 0; kk < 3; kk++) xk[kk] = kk < nt ? a[base + tx + PB * kk][J + c] : 0.0;
#pragma unroll
                for (int ii = 0; ii < 3; ii++)
#pragma unroll
                    for (int kk = 0; kk < 3; kk++) acc[ii][kk] += xi[ii] * xk[kk];
            }
#pragma unroll
            for (int ii = 0; ii < 3; ii++)
#pragma unroll
                for (int kk = 0; kk < 3; kk++) {
                    const int i = base + ty + PB * ii, k = base + tx + PB * kk;
                    if (ii < nt && kk < nt && k <= i) a[i][k] -= acc[ii][kk];
                }
        }
        __syncthreads();
    }
    if (sbad) {
        if (tid == 0 && status[1] == 0.0) status[1] = (double)(col_offset + sbad);
        return;
    }
    for (int idx = tid; idx < NB * NB; idx += BLK_T) {
        const int i = idx % NB, k = idx / NB;
        if (i < kb && k < kb && k <= i) A[i + (size_t)k * lda] = a[i][k];
    }
    // log-determinant: the 64 diagonal entries in parallel, off the factorisation's critical path
    __shared__ double slog[2];
    if (tid < NB) {
        lg = tid < kb ? log(a[tid][tid]) : 0.0;
        lg = warp_sum(lg);
        if (lane == 0) slog[wid] = lg;
    }
    __syncthreads();
    if (tid == 0) status[0] += slog[0] + slog[1];
}

// Panel solve X L' = P in place: P is Mp x kb (rows below the diagonal block), L the kb x kb lower factor.
// Right-looking per row: x_c *= 1/L[c][c] (broadcast to the row's lanes by shuffle); x_k -= x_c L[k][c] for k > c.
// IDENT = true solves against the identity instead (P = I): row t of the result is column t of L^-1, which is
// written transposed -> out = L^-1 (batched over blockIdx.x: the diagonal blocks of a factor).
template <bool IDENT>
__global__ void __launch_bounds__(BLK_T) trsm_rows_kernel(const double *__restrict__ L, int ldl, double *__restrict__ P, int ldp,
                                                          int Mp, int kb, int r) {
    __shared__ double Lt[NB][NB + 4];     // Lt[c][k] = L[k][c]
    __shared__ double rd[NB];
    const int t = threadIdx.x / SPLIT, h = threadIdx.x % SPLIT;
    if (IDENT) {   // diagonal block blockIdx.x of the factor
        const int o = blockIdx.x * NB;
        kb = min(NB, r - o);
        L += o + (size_t)o * ldl;
        P += o + (size_t)o * ldp;
    }
    {   // all 16 loads of L in flight before the first store
        double v[NB * NB / BLK_T];
#pragma unroll
        for (int q = 0; q < NB * NB / BLK_T; q++) {
            const int idx = threadIdx.x + q * BLK_T, k = idx % NB, c = idx / NB;
            v[q] = (k < kb && c < kb && c <= k) ? L[k + (size_t)c * ldl] : (k == c ? 1.0 : 0.0);
        }
#pragma unroll
        for (int q = 0; q < NB * NB / BLK_T; q++) {
            const int idx = threadIdx.x + q * BLK_T, k = idx % NB, c = idx / NB;
            Lt[c][k] = v[q];
            if (k == c) rd[c] = __drcp_rn(v[q]);
        }
    }
    __syncthreads();
    const int row = IDENT ? t : blockIdx.x * NB + t;
    const bool live = IDENT ? (t < kb) : (row < Mp);
    double x[NBQ];
#pragma unroll
    for (int kk = 0; kk < NBQ; kk++) {
        const int c = kk * SPLIT + h;
        x[kk] = IDENT ? (c == t ? 1.0 : 0.0) : ((live && c < kb) ? P[row + (size_t)c * ldp] : 0.0);
    }
    const int lane_base = (threadIdx.x & 31) & ~(SPLIT - 1);
#pragma unroll
    for (int cc = 0; cc < NBQ; cc++) {                 // columns c = 4*cc + hh
#pragma unroll 1
        for (int hh = 0; hh < SPLIT; hh++) {
            const int c = cc * SPLIT + hh;
            double xc = x[cc] * rd[c];
            if (h == hh) x[cc] = xc;
            xc = __shfl_sync(0xffffffffu, xc, lane_base + hh);
            const double *lt = &Lt[c][h];
            if (h > hh) x[cc] -= xc * lt[cc * SPLIT];
#pragma unroll
            for (int kk = cc + 1; kk < NBQ; kk++) x[kk] -= xc * lt[kk * SPLIT];
        }
    }
    if (IDENT) {
        // x_(row t)[c] = (L^-1)[c][t]: stage through shared memory so that row t is stored with coalesced accesses
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < NBQ; kk++) Lt[t][kk * SPLIT + h] = x[kk];       // Lt[t][c] = (L^-1)[c][t]
        __syncthreads();
        for (int idx = threadIdx.x; idx < NB * NB; idx += BLK_T) {
            const int rr = idx % NB, cc = idx / NB;                             // element (rr, cc) of L^-1
            if (rr < kb && cc < kb) P[rr + (size_t)cc * ldp] = Lt[cc][rr];
        }
    } else if (live) {
#pragma unroll
        for (int kk = 0; kk < NBQ; kk++) {
            const int c = kk * SPLIT + h;
            if (c < kb) P[row + (size_t)c * ldp] = x[kk];
        }
    }
}

// ---------------------------------------------------------------------------------
// Panel Cholesky, second generation: 128-wide panels, TWO or three kernels on the critical path per panel instead of eight.
//   D  diag128_kernel        one CTA factors the 128 x 128 diagonal block entirely in shared memory:
//                            potf2(64) -> W00 = L00^-1 -> L10 = A10 W00' -> A11 -= L10 L10' -> potf2(64) -> W11 = L11^-1
//   S  panel_solve128_kernel rows below the block, 64 rows per CTA:  X0 = P0 W00';  P1 -= X0 L10';  X1 = P1 W11'
//   U  dgemm_kernel          the next diagonal block takes its rank-128 update (critical path); the rest of the next panel's
//                            columns on a side stream beside D of the next panel
//   R  dgemm_kernel          everything beyond the next panel (lower SYRK, rank 128) on a low-priority side stream
// Only the 16 x 16 diagonal sub-blocks are factored column by column (one warp, registers + shuffles).  Everything else --
// including the triangular solves, which are products with explicit inverses of 16 x 16 / 64 x 64 triangular blocks -- is a
// small dense product on the fp64 tensor cores (DMMA m8n8k4) with both operands in shared memory.  Measured before this
// formulation: the substitution loops were bound by shared-memory bandwidth (one LDS per FMA), 26 us for the panel solve.
// ---------------------------------------------------------------------------------
// phase timestamps of the last diag / panel-solve launch (clock64 of CTA 0, thread 0): tools/dense_phase_probe.py
__device__ long long g_dbg_clk[32];
#define DBG_CLK(slot) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_dbg_clk[slot] = clock64(); } while (0)
extern "C" int frk_debug_clocks(long long *h_out32) {
    FRK_CHECK_CUDA(cudaDeviceSynchronize());
    FRK_CHECK_CUDA(cudaMemcpyFromSymbol(h_out32, g_dbg_clk, sizeof(long long) * 32));
    return 0;
}

constexpr int DP = 128;                       // panel width
// All blocks live COLUMN-MAJOR in shared memory: element (i, j) at j * LD + i, with LD = 4 (mod 16) doubles.  With that stride the
// 32 addresses of a DMMA fragment load -- 8 consecutive (g) in one direction, 4 strided (tg) in the other, whichever operand and
// orientation -- fall on every 8-byte bank pair exactly twice: two wavefronts, the minimum for 256 bytes.
constexpr int MLD = DP + 4;                   // 132: the 128 x 128 diagonal block
constexpr int NWARP = BLK_T / 32;

// 1 / sqrt(x) without a branch, so that the compiler can interleave it with independent work: the hardware approximation
// (MUFU.RSQ64H, ~2^-22) followed by ONE third-order step  y (1 + e/2 + 3 e^2 / 8),  e = 1 - x y^2  (remaining error ~ e^3 ~ 2^-66).
// Valid for normal positive x; callers treat pivots that are not > 1e-290 as "not positive definite" before using the result.
__device__ __forceinline__ double rsqrt_fast(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double ye = y * e;
    return fma(ye, p, y);
}

// Cholesky of a 16 x 16 block by one warp: lane (& 15) = row, d[k] = a[row][k] in registers, columns exchanged by shuffles.
// The column-sequential chain is  rsqrt(pivot_j) -> l_{j+1,j} -> pivot_{j+1} -> rsqrt(pivot_{j+1}):  every lane computes the NEXT
// pivot itself from two values of row j+1 fetched before column j is scaled, so the reciprocal square root of column j+1 is in
// flight while the rank-1 update of column j (15 shuffles + 15 FMAs) is still being applied.  The next-pivot expression is the same
// fma the owner lane applies, so the value is identical to the one the plain right-looking loop would produce.
// rdiag (if given) receives 1 / L[j][j].  Returns 0 or the (1-based) index of the first pivot that is not > 1e-290 (also catches NaN).
__device__ __forceinline__ int warp_potf2_16(double d[PB], int i, double *rdiag) {
    int bad = 0;
    double piv = __shfl_sync(0xffffffffu, d[0], 0);
    double rs = rsqrt_fast(piv), my_rs = 1.0;
#pragma unroll
    for (int j = 0; j < PB; j++) {
        if (!(piv > 1e-290)) { if (!bad) bad = j + 1; }
        double piv_next = 1.0, rs_next = 1.0;
        if (j + 1 < PB) {
            const double t1 = __shfl_sync(0xffffffffu, d[j], j + 1);          // a_{j+1,j} and a_{j+1,j+1}, updated through column j-1
            const double t2 = __shfl_sync(0xffffffffu, d[j + 1 < PB ? j + 1 : j], j + 1);
            const double l = t1 * rs;
            piv_next = fma(-l, l, t2);
            rs_next = rsqrt_fast(piv_next);
        }
        my_rs = (i == j) ? rs : my_rs;
        const double lij = (i == j) ? piv * rs : d[j] * rs;
        d[j] = lij;
#pragma unroll
        for (int k = j + 1; k < PB; k++) d[k] = fma(-lij, __shfl_sync(0xffffffffu, lij, k), d[k]);   // entries with k > i are never used
        piv = piv_next;
        rs = rs_next;
    }
    if (rdiag) rdiag[i] = my_rs;
    return bad;
}

// Inverse of a 16 x 16 lower-triangular block: lane j (< 16) computes column j of the inverse by forward substitution on e_j.
// L(i,k) at L[k * ldl + i] (all lanes read the same element: broadcast), rdiag[k] = 1 / L(k,k); W(i,j) -> W[j * ldw + i].
__device__ __forceinline__ void tri_inv16(const double *__restrict__ L, int ldl, const double *__restrict__ rdiag, double *__restrict__ W, int ldw, int j) {
    double s[PB];
#pragma unroll
    for (int i = 0; i < PB; i++) s[i] = (i == j) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < PB; k++) {
        const double xk = s[k] * rdiag[k];
        s[k] = xk;
#pragma unroll
        for (int i = k + 1; i < PB; i++) s[i] = fma(-L[k * ldl + i], xk, s[i]);
    }
#pragma unroll
    for (int i = 0; i < PB; i++) W[j * ldw + i] = s[i];                          // rows i < j come out as exact zeros
}

// ---- small dense products on the fp64 tensor cores, operands in shared memory (column-major, LD = 4 mod 16) ----
//   C(m, n) = beta * C(m, n) + alpha * sum_k A(m, k) * Bop(k, n)      A(m,k) at A[k*lda + m];  C(m,n) at C[n*ldc + m]
//   BT = 0:  Bop(k, n) = B(n, k), read at B[k*ldb + n]   ("A B'")          BT = 1:  Bop(k, n) = B(k, n), read at B[n*ldb + k]   ("A B")
// The M x N result is cut into (8 MF) x (8 NF) warp tiles dealt round-robin to the warps of the CTA.
// k-ranges per tile (multiples of 4 by construction: all block sizes are multiples of 8):
//   KEND_N: k < n0 + 8 NF  (B(n,k) = 0 for k > n: lower-triangular B in the A B' form)
//   KBEG_N: k >= n0         (B(k,n) = 0 for k < n: lower-triangular B in the A B form)
//   KEND_M: k < m0 + 8 MF  (A(m,k) = 0 for k > m: lower-triangular A)
//   C_LOWER: tiles entirely above the diagonal are skipped (entries above the diagonal inside a kept tile are still written)
// INPLACE: C may alias A or B -- every warp keeps its (single) tile in registers until all warps are done reading; requires
// #tiles <= #warps.  The function ends with a __syncthreads() in every case.
enum { KEND_N = 1, KBEG_N = 2, KEND_M = 4, C_LOWER = 8 };

template <int MF, int NF, int BT, int FLAGS, bool INPLACE, int KFIX = 0>
__device__ __forceinline__ void smem_gemm(int M, int N, int K, double alpha, const double *__restrict__ A, int lda, const double *__restrict__ B,
                                          int ldb, double beta, double *C, int ldc) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3;
    const int tm = M / (8 * MF), tn = N / (8 * NF), ntask = tm * tn;
    for (int task = wid; task < (INPLACE ? NWARP : ntask); task += NWARP) {
        const bool live = task < ntask;
        const int m0 = live ? (task % tm) * 8 * MF : 0, n0 = live ? (task / tm) * 8 * NF : 0;
        const bool skip = !live || ((FLAGS & C_LOWER) && n0 > m0 + 8 * MF - 1);
        double acc[MF][NF][2];
#pragma unroll
        for (int i = 0; i < MF; i++)
#pragma unroll
            for (int j = 0; j < NF; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        if (!skip) {
            int kbeg = 0, kend = K;
            if (FLAGS & KBEG_N) kbeg = n0;
            if (FLAGS & KEND_N) kend = min(kend, n0 + 8 * NF);
            if (FLAGS & KEND_M) kend = min(kend, m0 + 8 * MF);
            const double *pa = A + (size_t)tg * lda + m0 + g;
            const double *pb = BT ? B + (size_t)(n0 + g) * ldb + tg : B + (size_t)tg * ldb + n0 + g;
            if (KFIX > 0) {
                // short fixed k-range (K = KFIX, no triangular flags): every fragment load is issued before the first DMMA, so the
                // product costs one shared-memory latency plus the accumulation chain instead of one latency per k-step
                constexpr int KS = KFIX > 0 ? KFIX / 4 : 1;
                double a[KS][MF], b[KS][NF];
#pragma unroll
                for (int ks = 0; ks < KFIX / 4; ks++) {
#pragma unroll
                    for (int i = 0; i < MF; i++) a[ks][i] = pa[(size_t)(4 * ks) * lda + 8 * i];
#pragma unroll
                    for (int j = 0; j < NF; j++) b[ks][j] = BT ? pb[(size_t)(8 * j) * ldb + 4 * ks] : pb[(size_t)(4 * ks) * ldb + 8 * j];
                }
#pragma unroll
                for (int ks = 0; ks < KFIX / 4; ks++)
#pragma unroll
                    for (int i = 0; i < MF; i++)
#pragma unroll
                        for (int j = 0; j < NF; j++) dmma(acc[i][j][0], acc[i][j][1], a[ks][i], b[ks][j]);
            } else
#pragma unroll 2
            for (int k0 = kbeg; k0 < kend; k0 += 4) {
                double a[MF], b[NF];
#pragma unroll
                for (int i = 0; i < MF; i++) a[i] = pa[(size_t)k0 * lda + 8 * i];
#pragma unroll
                for (int j = 0; j < NF; j++) b[j] = BT ? pb[(size_t)(8 * j) * ldb + k0] : pb[(size_t)k0 * ldb + 8 * j];
#pragma unroll
                for (int i = 0; i < MF; i++)
#pragma unroll
                    for (int j = 0; j < NF; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
        if (INPLACE) __syncthreads();
        if (!skip) {
#pragma unroll
            for (int i = 0; i < MF; i++)
#pragma unroll
                for (int j = 0; j < NF; j++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        double *c = C + (size_t)(n0 + 8 * j + 2 * tg + e) * ldc + m0 + 8 * i + g;
                        const double v = alpha * acc[i][j][e];
                        *c = beta != 0.0 ? fma(beta, *c, v) : v;
                    }
        }
    }
    __syncthreads();
}

// Cholesky of the n x n block at a (column-major, lda; n a multiple of 16, n <= 128), in place (lower triangle; the strict upper
// triangle of each 16 x 16 diagonal sub-block is zeroed), plus the inverses of the 16 x 16 diagonal sub-blocks into Wd (block J / 16
// at Wd + (J / 16) * 16 * WDL, column-major with stride WDL).  16-column steps: warp 0 factors the diagonal sub-block and inverts it;
// the rows below are solved as a product with that inverse and the rest of the block is updated, both on the tensor cores.
// Returns (uniformly) 0 or the 1-based column of the first non-positive pivot.
constexpr int WDL = PB + 4;                   // 20 = 4 (mod 16)
__device__ __forceinline__ int potf2_blk_smem(double *a, int lda, int n, double *Wd, double *rdiag, int *sbad) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int J = 0; J < n; J += PB) {
        double *ajj = a + (size_t)J * lda + J;
        double *wj = Wd + (size_t)(J / PB) * PB * WDL;
        const bool stamp = J == 0;
        if (wid == 0) {
            const int i = lane & (PB - 1);
            double d[PB];
#pragma unroll
            for (int k = 0; k < PB; k++) d[k] = ajj[(size_t)k * lda + i];
            if (stamp) DBG_CLK(16);
            int bad = warp_potf2_16(d, i, lane < PB ? rdiag + J : nullptr);
            if (stamp) DBG_CLK(17);
            if (lane < PB) {
#pragma unroll
                for (int k = 0; k < PB; k++) ajj[(size_t)k * lda + i] = (k <= i) ? d[k] : 0.0;
            }
            if (lane == 0 && bad) *sbad = J + bad;
            __syncwarp();
            if (stamp) DBG_CLK(18);
            if (lane < PB) tri_inv16(ajj, lda, rdiag + J, wj, WDL, lane);
            if (stamp) DBG_CLK(19);
        }
        __syncthreads();
        if (stamp) DBG_CLK(20);
        if (*sbad) break;
        const int nbelow = n - J - PB;
        if (nbelow <= 0) break;
        // rows below: X = A_below W16'   (nbelow x 16 x 16, in place; at most 7 warp tiles)
        smem_gemm<2, 2, 0, 0, true, PB>(nbelow, PB, PB, 1.0, ajj + PB, lda, wj, WDL, 0.0, ajj + PB, lda);
        if (stamp) DBG_CLK(21);
        // rest of the block (right-looking; a left-looking variant -- one update per step with a growing k-range -- measured 5 % slower):
        //   A(i,k) -= sum_c X(i,c) X(k,c), lower
        smem_gemm<2, 2, 0, C_LOWER, false, PB>(nbelow, nbelow, PB, -1.0, ajj + PB, lda, ajj + PB, lda, 1.0, ajj + (size_t)PB * lda + PB, lda);
        if (stamp) DBG_CLK(22);
    }
    return *sbad;
}

constexpr size_t DIAG_SMEM = ((size_t)DP * MLD + (size_t)(DP / PB) * PB * WDL) * sizeof(double);
constexpr int WINV_DOUBLES = (DP / PB) * PB * PB;      // 8 blocks of 16 x 16, column-major, dense

// D: Cholesky of the nb x nb (nb <= 128) diagonal block at A, in place (lower triangle), and the inverses of its eight 16 x 16
// diagonal sub-blocks into Winv (8 x 16 x 16, column-major) for the panel solve.
// status[0] += sum(log(diag L)); status[1] = first failing (1-based, global) column, untouched if ok.
__global__ void __launch_bounds__(BLK_T) diag128_kernel(double *__restrict__ A, int lda, int nb, int col_offset, double *__restrict__ status,
                                                        double *__restrict__ Winv) {
    extern __shared__ __align__(16) double dsm[];
    double *a = dsm;                                   // a[j * MLD + i]
    double *Wd = dsm + (size_t)DP * MLD;               // eight 16 x 16 inverse blocks
    __shared__ double rdiag[DP];
    __shared__ int sbad;
    __shared__ double slog[NWARP];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) sbad = 0;
    DBG_CLK(0);
    const int n16 = ((nb + PB - 1) / PB) * PB;         // identity padding up to the next multiple of 16
    // load (lower triangle; zeros above): columns are contiguous on both sides; 16 loads in flight
    for (int q0 = 0; q0 < DP * DP / BLK_T; q0 += 16) {
        if (q0 * BLK_T >= n16 * DP) break;             // columns beyond the block are never touched
        double v[16];
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int idx = tid + (q0 + q) * BLK_T, i = idx % DP, k = idx / DP;
            v[q] = (i < nb && k < nb && k <= i) ? A[i + (size_t)k * lda] : (i == k ? 1.0 : 0.0);
        }
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int idx = tid + (q0 + q) * BLK_T, i = idx % DP, k = idx / DP;
            a[(size_t)k * MLD + i] = v[q];
        }
    }
    __syncthreads();
    DBG_CLK(1);
    const int bad = potf2_blk_smem(a, MLD, n16, Wd, rdiag, &sbad);
    DBG_CLK(2);
    if (bad) {
        if (tid == 0 && status[1] == 0.0) status[1] = (double)(col_offset + bad);
        return;
    }
    for (int idx = tid; idx < WINV_DOUBLES; idx += BLK_T) {
        const int blk = idx / (PB * PB), e = idx % (PB * PB);
        Winv[idx] = Wd[(size_t)blk * PB * WDL + (e / PB) * WDL + e % PB];
    }
    for (int idx = tid; idx < n16 * DP; idx += BLK_T) {
        const int i = idx % DP, k = idx / DP;
        if (i < nb && k <= i) A[i + (size_t)k * lda] = a[(size_t)k * MLD + i];
    }
    double lg = tid < nb ? log(a[(size_t)tid * MLD + tid]) : 0.0;                 // (BLK_T = 256 >= 128 diagonal entries)
    lg = warp_sum(lg);
    if (lane == 0) slog[wid] = lg;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < NWARP; w++) s += slog[w];
        status[0] += s;
    }
    DBG_CLK(6);
}

// S: the rows below a factored 128 x 128 diagonal block.  P (Mp x 128, column-major, ldp) <- P L^-T in eight 16-column steps:
//   X_J = P_J W_J'   (W_J = inverse of the J-th 16 x 16 diagonal block of L, from diag128_kernel)
//   P(:, c) -= X_J L(c, J)'   for the columns c to the right.
// SR rows per CTA; the row tile, the factor L and the inverse blocks sit in shared memory; every step is a DMMA product.
constexpr int SR = 32;                        // rows per CTA (32: ~100 CTAs at r ~ 3000, i.e. most of the SMs)
constexpr int XLD = SR + 4;                   // 36 = 4 (mod 16)
constexpr size_t SOLVE_SMEM = ((size_t)DP * XLD + (size_t)DP * MLD + (size_t)(DP / PB) * PB * WDL) * sizeof(double);
__global__ void __launch_bounds__(BLK_T) panel_solve128_kernel(const double *__restrict__ L, int ldl, const double *__restrict__ Winv,
                                                               double *__restrict__ P, int ldp, int Mp) {
    extern __shared__ __align__(16) double ssm[];
    double *X = ssm;                                   // X[c * XLD + row]: SR rows x 128 columns
    double *Ls = X + (size_t)DP * XLD;                 // Ls[c * MLD + i]: the factor (lower part; the rest is not read)
    double *Wd = Ls + (size_t)DP * MLD;
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * SR, nrow = min(SR, Mp - row0);
    DBG_CLK(8);
    // the row tile: 16 loads in flight per thread
    {
        double v[SR * DP / BLK_T];
#pragma unroll
        for (int q = 0; q < SR * DP / BLK_T; q++) {
            const int idx = tid + q * BLK_T, i = idx % SR, c = idx / SR;
            v[q] = i < nrow ? P[row0 + i + (size_t)c * ldp] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < SR * DP / BLK_T; q++) {
            const int idx = tid + q * BLK_T, i = idx % SR, c = idx / SR;
            X[c * XLD + i] = v[q];
        }
    }
    // the factor: only the rows below the 16 x 16 diagonal blocks are read (column c: rows >= 16 (c / 16 + 1))
    for (int q0 = 0; q0 < DP * DP / BLK_T; q0 += 16) {
        double v[16];
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int idx = tid + (q0 + q) * BLK_T, i = idx % DP, c = idx / DP;
            v[q] = i >= PB * (c / PB + 1) ? L[i + (size_t)c * ldl] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int idx = tid + (q0 + q) * BLK_T, i = idx % DP, c = idx / DP;
            Ls[(size_t)c * MLD + i] = v[q];
        }
    }
    for (int idx = tid; idx < WINV_DOUBLES; idx += BLK_T) {
        const int blk = idx / (PB * PB), e = idx % (PB * PB);
        Wd[(size_t)blk * PB * WDL + (e / PB) * WDL + e % PB] = Winv[idx];
    }
    __syncthreads();
    DBG_CLK(9);
    for (int J = 0; J < DP; J += PB) {
        double *XJ = X + (size_t)J * XLD;
        smem_gemm<2, 2, 0, 0, true, PB>(SR, PB, PB, 1.0, XJ, XLD, Wd + (size_t)(J / PB) * PB * WDL, WDL, 0.0, XJ, XLD);      // X_J = P_J W_J'
        const int nrest = DP - J - PB;
        if (nrest > 0)                                                                                                        // P(:, right) -= X_J L(right, J)'
            smem_gemm<2, 2, 0, 0, false, PB>(SR, nrest, PB, -1.0, XJ, XLD, Ls + (size_t)J * MLD + J + PB, MLD, 1.0, XJ + (size_t)PB * XLD, XLD);
    }
    DBG_CLK(11);
    for (int idx = tid; idx < SR * DP; idx += BLK_T) {
        const int i = idx % SR, c = idx / SR;
        if (i < nrow) P[row0 + i + (size_t)c * ldp] = X[c * XLD + i];
    }
    DBG_CLK(12);
}

// U_diag: the next diagonal block (np x np, np <= 128, lower) takes its rank-128 update from the rows of the panel that face it:
//   C(i, j) -= sum_k X(i, k) X(j, k),  X = P[0 : np, 0 : 128].
// One CTA per 32 x 32 tile of the lower triangle (10 for np = 128): the two 32-row slabs of X go to shared memory, the tile is
// updated in place in global memory.  On the critical path of the factorisation: ~6 us instead of ~20 for the general GEMM kernel.
constexpr int UT = 32, ULD = UT + 4;          // tile size; slab stride 36 = 4 (mod 16)
constexpr size_t UDIAG_SMEM = (size_t)2 * DP * ULD * sizeof(double);
__global__ void __launch_bounds__(BLK_T) next_diag_update_kernel(const double *__restrict__ X, int ldx, double *__restrict__ Cg, int ldc, int np, int nk) {
    extern __shared__ __align__(16) double usm[];
    double *Xa = usm, *Xb = usm + (size_t)DP * ULD;    // Xa[k * ULD + i] = X(i0 + i, k), Xb likewise for the column slab
    // tile index -> (bi, bj), bj <= bi
    int bi = 0, t = blockIdx.x;
    while (t > bi) { t -= bi + 1; bi++; }
    const int bj = t, i0 = bi * UT, j0 = bj * UT, tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5, g = lane >> 2, tg = lane & 3;
    const int m0 = (wid & 3) * 8, n0 = (wid >> 2) * 16;
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  for (int kc = 0; kc < nk; kc++, X += (size_t)DP * ldx) {          // K in chunks of 128 columns (nk = 2: the two panels of a pair)
    if (kc) __syncthreads();
    for (int q0 = 0; q0 < UT * DP / BLK_T; q0 += 8) {
        double va[8], vb[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int idx = tid + (q0 + q) * BLK_T, i = idx % UT, k = idx / UT;
            va[q] = i0 + i < np ? X[i0 + i + (size_t)k * ldx] : 0.0;
            vb[q] = j0 + i < np ? X[j0 + i + (size_t)k * ldx] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int idx = tid + (q0 + q) * BLK_T, i = idx % UT, k = idx / UT;
            Xa[k * ULD + i] = va[q];
            Xb[k * ULD + i] = vb[q];
        }
    }
    __syncthreads();
    // 32 x 32 x 128 on 8 warps: 8 x 16 warp tiles (1 x 2 fragments); rows / columns beyond np are zero-padded and not stored
    const double *pa = Xa + (size_t)tg * ULD + m0 + g, *pb = Xb + (size_t)tg * ULD + n0 + g;
#pragma unroll 8
    for (int k0 = 0; k0 < DP; k0 += 4) {
        const double a = pa[(size_t)k0 * ULD], b0 = pb[(size_t)k0 * ULD], b1 = pb[(size_t)k0 * ULD + 8];
        dmma(acc[0][0], acc[0][1], a, b0);
        dmma(acc[1][0], acc[1][1], a, b1);
    }
  }
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int row = i0 + m0 + g, colc = j0 + n0 + 8 * j + 2 * tg + e;
            if (row < np && colc <= row) Cg[row + (size_t)colc * ldc] -= acc[j][e];
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

// opt every GEMM instantiation into its dynamic shared memory once, outside any stream capture
static int gemm_set_attrs() {
    static bool done = false;
    if (done) return 0;
#define FRK_GEMM_ATTR(a, b, c) FRK_CHECK_CUDA(cudaFuncSetAttribute(dgemm_kernel<a, b, c, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM)); \
                               FRK_CHECK_CUDA(cudaFuncSetAttribute(dgemm_kernel<a, b, c, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM))
    FRK_GEMM_ATTR(false, false, true); FRK_GEMM_ATTR(false, false, false); FRK_GEMM_ATTR(false, true, true); FRK_GEMM_ATTR(false, true, false);
    FRK_GEMM_ATTR(true, false, true); FRK_GEMM_ATTR(true, false, false); FRK_GEMM_ATTR(true, true, true); FRK_GEMM_ATTR(true, true, false);
#undef FRK_GEMM_ATTR
    FRK_CHECK_CUDA(cudaFuncSetAttribute(potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTF2_SMEM));
    FRK_CHECK_CUDA(cudaFuncSetAttribute(diag128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
    FRK_CHECK_CUDA(cudaFuncSetAttribute(panel_solve128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVE_SMEM));
    FRK_CHECK_CUDA(cudaFuncSetAttribute(next_diag_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UDIAG_SMEM));
    done = true;
    return 0;
}

// Replay a launch sequence through a cached CUDA graph (captured on first use for this key).  The factorisations are chains of
// ~150 small dependent kernels; inside a graph the kernel-to-kernel gap drops from ~3.5 us to ~1 us.  Not used while profiling
// (per-launch events) or when disabled (frk_use_graphs).
template <class F>
static int run_graphed(frk_ctx *ctx, const std::array<uintptr_t, 5> &key, F enqueue) {
    if (!ctx->use_graphs || ctx->profiling || ctx->stream == nullptr) return enqueue();   // the legacy default stream cannot be captured
    if (int rc = gemm_set_attrs()) return rc;
    auto it = ctx->graphs.find(key);
    if (it == ctx->graphs.end()) {
        if (ctx->graphs.size() >= 256) {
            for (auto &kv : ctx->graphs) cudaGraphExecDestroy(kv.second.exec);
            ctx->graphs.clear();
        }
        const int64_t l0 = ctx->launches;
        FRK_CHECK_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue();
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        FRK_CHECK_CUDA(e);
        frk_ctx::GraphEntry ent{nullptr, ctx->launches - l0};
        ctx->launches = l0;
        const cudaError_t e2 = cudaGraphInstantiate(&ent.exec, g, 0);
        cudaGraphDestroy(g);
        FRK_CHECK_CUDA(e2);
        it = ctx->graphs.emplace(key, ent).first;
    }
    FRK_CHECK_CUDA(cudaGraphLaunch(it->second.exec, ctx->stream));
    ctx->launches += it->second.nodes;
    return 0;
}

// Launch sequence of the two-level right-looking Cholesky with look-ahead on three streams (one CUDA graph with parallel branches
// when captured).  Per 256-wide outer panel J:
//   main : potf2'(first diagonal block; its rank-256 update from panel J-1 happens inside the kernel)
//          then for each 64-wide inner block: trsm -> potf2'(next diagonal block, rank-64 update inside) ...
//   side : the rest of each inner rank-64 update, beside the next diagonal factorisation;
//          NP = update of the next panel's columns with the whole panel J, beside the first diagonal factorisation of panel J+1
//   side2: the rest of the trailing SYRK (columns beyond the next panel), beside the whole of panel J+1.
static int potrf_enqueue(frk_ctx *ctx, int r, double *d_A, double *status) {
    FRK_CHECK_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(double), ctx->stream));
    const size_t ld = (size_t)r;
    cudaStream_t main = ctx->stream, side = ctx->stream2, side2 = ctx->stream3;
    const bool fork = main != nullptr && side != nullptr && side2 != nullptr;   // (the legacy default stream cannot fork)
    // rest(p) = the part of panel p's trailing update beyond the next panel (side2); it records ev_rest[p & 1].  Panel p+2 (its
    // first potf2' and the NP that feeds it) is the first consumer, so two alternating events keep rest(p) overlapped with panel p+1.
    int pidx = 0;                              // index of the current outer panel
    bool have_rest[2] = {false, false};        // ev_rest[i] holds a recorded rest-SYRK
    const double *Pprev = nullptr;             // panel J-1 rows facing the first diagonal block of panel J (for its pre-update)
    int kprev = 0;
    for (int J = 0; J < r; J += NBO) {
        const int Jend = std::min(r, J + NBO);
        if (fork && have_rest[pidx & 1]) FRK_CHECK_CUDA(cudaStreamWaitEvent(main, ctx->ev_rest[pidx & 1], 0));   // rest(pidx-2) reached this panel
        FRK_LAUNCH(ctx, potf2_kernel, 1, BLK_T, POTF2_SMEM, d_A + J + J * ld, r, std::min(NB, r - J), J, status, Pprev, r, kprev);
        if (J > 0 && fork) FRK_CHECK_CUDA(cudaStreamWaitEvent(main, ctx->ev_next, 0));            // the panel's columns got NP(J-1)
        for (int o = J; o < Jend; o += NB) {
            const int kb = std::min(NB, r - o), below = r - o - kb;
            if (below <= 0) continue;
            double *Ajj = d_A + o + o * ld;
            double *P = d_A + (o + kb) + o * ld;                       // below x kb panel
            FRK_WORK(ctx, (double)below * kb * kb, 0);
            FRK_LAUNCH(ctx, trsm_rows_kernel<false>, (unsigned)cdiv(below, NB), BLK_T, 0, Ajj, r, P, r, below, kb, r);
            const int ncols = Jend - (o + kb);                         // remaining columns of the outer panel
            if (ncols <= 0) continue;
            // Look-ahead inside the panel.  The next diagonal block (o+kb) takes its rank-kb update inside its own factorisation
            // kernel (from the kb2 rows of P that face it); everything else of the update, rows >= o+kb+kb2, runs beside it.
            const int kb2 = std::min(NB, r - (o + kb)), rest = below - kb2;
            if (rest > 0) {
                if (fork) {
                    FRK_CHECK_CUDA(cudaEventRecord(ctx->ev_fork, main));
                    FRK_CHECK_CUDA(cudaStreamWaitEvent(side, ctx->ev_fork, 0));
                    ctx->stream = side;
                }
                // C[rows >= o+kb+kb2, cols o+kb .. Jend) -= P[rows >= o+kb+kb2, :] P[0:ncols, :]'; the diagonal is kb2 columns to the right
                GemmArgs g = gemm_args(rest, ncols, kb, -1.0, P + kb2, r, P, r, 1.0, d_A + (o + kb + kb2) + (o + kb) * ld, r);
                g.lower = 1; g.lower_off = kb2;
                const int rc = launch_gemm(ctx, g, false, false, 2.0 * rest * ncols * kb, "dgemm_kernel[potrf inner rank-64]");
                ctx->stream = main;
                if (rc) return rc;
                if (fork) FRK_CHECK_CUDA(cudaEventRecord(ctx->ev_join, side));
            }
            FRK_LAUNCH(ctx, potf2_kernel, 1, BLK_T, POTF2_SMEM, d_A + (o + kb) + (o + kb) * ld, r, kb2, o + kb, status, (const double *)P, r, kb);
            if (rest > 0 && fork) FRK_CHECK_CUDA(cudaStreamWaitEvent(main, ctx->ev_join, 0));
        }
        // trailing update with the whole outer panel (rows >= Jend): first diagonal block of the next panel -> inside its potf2';
        // the rest of the next panel's columns (NP) -> side; everything beyond the next panel -> side2
        const int below = r - Jend, kw = Jend - J;
        Pprev = nullptr; kprev = 0;
        if (below <= 0) { pidx++; continue; }
        const double *P = d_A + Jend + J * ld;
        if (!fork) {                                                   // sequential: one lower SYRK
            GemmArgs g = gemm_args(below, below, kw, -1.0, P, r, P, r, 1.0, d_A + Jend + Jend * ld, r);
            g.lower = 1;
            if (int rc = launch_gemm(ctx, g, false, false, (double)below * (below + 1.0) * kw, "dgemm_kernel[potrf outer SYRK]")) return rc;
            pidx++;
            continue;
        }
        Pprev = P; kprev = kw;
        const int kb0 = std::min(NB, below), npan = std::min(NBO, below), beyond = below - npan;
        FRK_CHECK_CUDA(cudaEventRecord(ctx->ev_panel, main));
        if (below - kb0 > 0) {                                         // NP: rows >= Jend+kb0, cols [Jend, Jend+npan), diagonal kb0 to the right
            FRK_CHECK_CUDA(cudaStreamWaitEvent(side, ctx->ev_panel, 0));
            if (have_rest[(pidx + 1) & 1]) FRK_CHECK_CUDA(cudaStreamWaitEvent(side, ctx->ev_rest[(pidx + 1) & 1], 0));   // rest(pidx-1) reached those columns
            GemmArgs g = gemm_args(below - kb0, npan, kw, -1.0, P + kb0, r, P, r, 1.0, d_A + (Jend + kb0) + Jend * ld, r);
            g.lower = 1; g.lower_off = kb0;
            ctx->stream = side;
            const int rc = launch_gemm(ctx, g, false, false, 2.0 * (below - kb0) * npan * kw, "dgemm_kernel[potrf next-panel update]");
            ctx->stream = main;
            if (rc) return rc;
        }
        FRK_CHECK_CUDA(cudaEventRecord(ctx->ev_next, side));          // (recorded even when NP is empty: main always waits on it)
        if (beyond > 0) {                                              // rest: rows, cols >= Jend+npan
            FRK_CHECK_CUDA(cudaStreamWaitEvent(side2, ctx->ev_panel, 0));
            const double *P2 = P + npan;
            GemmArgs g = gemm_args(beyond, beyond, kw, -1.0, P2, r, P2, r, 1.0, d_A + (Jend + npan) + (Jend + npan) * ld, r);
            g.lower = 1;
            ctx->stream = side2;
            const int rc = launch_gemm(ctx, g, false, false, (double)beyond * (beyond + 1.0) * kw, "dgemm_kernel[potrf outer SYRK]");
            ctx->stream = main;
            if (rc) return rc;
            FRK_CHECK_CUDA(cudaEventRecord(ctx->ev_rest[pidx & 1], side2));
            have_rest[pidx & 1] = true;
        } else {
            have_rest[pidx & 1] = false;
        }
        pidx++;
    }
    if (fork)                                                        // join side2 (in-order: the later event covers the earlier)
        for (int i = 0; i < 2; i++)
            if (have_rest[i]) FRK_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_rest[i], 0));
    return 0;
}

// Launch sequence of the 128-wide panel Cholesky (see diag128_kernel).  Per panel p (columns J .. J+128):
//   main : D(p) -> S(p) -> [wait R(p-1)] -> U(p): columns of panel p+1, rows >= J+128
//   side : [wait S(p)] R(p): rows and columns >= J+256 (lower SYRK, rank 128); R(p) runs beside D(p+1), S(p+1).
// U(p+1) and R(p) update overlapping regions (columns of panel p+2), hence the wait; everything else that runs concurrently
// touches disjoint blocks.  Replayed as one CUDA graph with two branches.
static int potrf2_enqueue(frk_ctx *ctx, int r, double *d_A, double *status, double *winv) {
    FRK_CHECK_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(double), ctx->stream));
    const size_t ld = (size_t)r;
    cudaStream_t main = ctx->stream, side = ctx->stream3;
    const bool fork = main != nullptr && side != nullptr && ctx->stream2 != nullptr;
    static const int dbg_skip = getenv("FRK_POTRF_SKIP") ? atoi(getenv("FRK_POTRF_SKIP")) : 0;   // timing experiments only: 1 = no R / U_rest, 2 = R / U_rest on the main stream
    bool have_rest[2] = {false, false};
    bool pending_join = false, rest_pending = false;
    int pidx = 0;
    for (int J = 0; J < r; J += DP, pidx++) {
        const int nb = std::min(DP, r - J), below = r - J - nb;
        FRK_WORK(ctx, (double)nb * nb * nb / 3.0, 0);
        FRK_LAUNCH(ctx, diag128_kernel, 1, BLK_T, DIAG_SMEM, d_A + J + J * ld, r, nb, J, status, winv);
        if (pending_join) { FRK_CHECK_CUDA(cudaStreamWaitEvent(main, ctx->ev_join, 0)); pending_join = false; }   // rows below got U(p-1)
        if (below <= 0) break;
        double *P = d_A + (J + nb) + J * ld;                           // below x 128
        FRK_WORK(ctx, (double)below * nb * nb, 0);
        FRK_LAUNCH(ctx, panel_solve128_kernel, (unsigned)cdiv(below, SR), BLK_T, SOLVE_SMEM, (const double *)(d_A + J + J * ld), r,
                   (const double *)winv, P, r, below);
        const int npan = std::min(DP, below), beyond = below - npan;
        if (!fork) {                                                   // sequential: one lower SYRK over everything below
            GemmArgs g = gemm_args(below, below, nb, -1.0, P, r, P, r, 1.0, d_A + (J + nb) + (J + nb) * ld, r);
            g.lower = 1;
            if (int rc = launch_gemm(ctx, g, false, false, (double)below * (below + 1.0) * nb, "dgemm_kernel[potrf trailing SYRK]")) return rc;
            continue;
        }
        // The trailing matrix takes its updates in PAIRS of panels (rank 256: fewer, longer GEMMs -- the beta = 1 epilogue of the GEMM
        // kernel costs as much as ~60 columns of k): an even panel only updates the next panel's columns (rank 128); an odd panel updates
        // the next panel's columns with both panels of the pair and then everything beyond (R, side stream).
        const bool odd = (pidx & 1) != 0;
        const int kw = odd ? 2 * DP : nb;                               // columns of X in this update
        const double *Pk = odd ? P - (size_t)DP * ld : P;              // ... starting one panel to the left for the pair
        if (dbg_skip == 1) {
            const int nt = (int)cdiv(npan, UT);
            FRK_LAUNCH(ctx, next_diag_update_kernel, (unsigned)(nt * (nt + 1) / 2), BLK_T, UDIAG_SMEM, Pk, r,
                       d_A + (J + nb) + (J + nb) * ld, r, npan, odd ? 2 : 1);
            continue;
        }
        if (dbg_skip == 2) { side = main; }
        if (rest_pending) { FRK_CHECK_CUDA(cudaStreamWaitEvent(main, ctx->ev_rest[0], 0)); rest_pending = false; }   // R of the previous pair reached these columns
        if (beyond > 0 && odd) {                                       // R(pair) on the side stream: rows, columns beyond the next panel
            FRK_CHECK_CUDA(cudaEventRecord(ctx->ev_panel, main));
            FRK_CHECK_CUDA(cudaStreamWaitEvent(side, ctx->ev_panel, 0));
            const double *P2 = Pk + npan;
            GemmArgs g = gemm_args(beyond, beyond, kw, -1.0, P2, r, P2, r, 1.0, d_A + (J + nb + npan) + (J + nb + npan) * ld, r);
            g.lower = 1;
            ctx->stream = side;
            const int rc = launch_gemm(ctx, g, false, false, (double)beyond * (beyond + 1.0) * kw, "dgemm_kernel[potrf trailing SYRK]");
            ctx->stream = main;
            if (rc) return rc;
            FRK_CHECK_CUDA(cudaEventRecord(ctx->ev_rest[0], side));
            rest_pending = true;
            have_rest[0] = true;
        }
        // U(p): the next panel's columns.  Only the next DIAGONAL block is on the critical path (D(p+1) needs it); the rows below it
        // are needed by S(p+1), so they are updated on the second side stream beside D(p+1).
        if (beyond > 0) {
            cudaStream_t s2 = dbg_skip == 2 ? main : ctx->stream2;
            FRK_CHECK_CUDA(cudaEventRecord(ctx->ev_fork, main));                     // S(p) done and the previous pair's R waited for
            FRK_CHECK_CUDA(cudaStreamWaitEvent(s2, ctx->ev_fork, 0));
            const double *P2 = Pk + npan;
            GemmArgs g = gemm_args(beyond, npan, kw, -1.0, P2, r, Pk, r, 1.0, d_A + (J + nb + npan) + (J + nb) * ld, r);
            ctx->stream = s2;
            const int rc = launch_gemm(ctx, g, false, false, 2.0 * beyond * npan * kw, "dgemm_kernel[potrf next-panel update]");
            ctx->stream = main;
            if (rc) return rc;
            FRK_CHECK_CUDA(cudaEventRecord(ctx->ev_join, s2));
        }
        {   // next diagonal block (npan x npan, lower)
            const int nt = (int)cdiv(npan, UT);
            FRK_WORK(ctx, (double)npan * (npan + 1.0) * kw, 0);
            FRK_LAUNCH(ctx, next_diag_update_kernel, (unsigned)(nt * (nt + 1) / 2), BLK_T, UDIAG_SMEM, Pk, r,
                       d_A + (J + nb) + (J + nb) * ld, r, npan, odd ? 2 : 1);
        }
        pending_join = beyond > 0;
    }
    if (fork)
        for (int i = 0; i < 2; i++)
            if (have_rest[i]) FRK_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_rest[i], 0));
    return 0;
}

// FRK_POTRF_V1=1 selects the first-generation launch sequence (64-wide steps, look-ahead on three branches) for A/B timing
static bool potrf_use_v1() {
    static const bool v1 = getenv("FRK_POTRF_V1") != nullptr;
    return v1;
}

// enqueue only: factor + {sum log diag, first bad column} -> ctx->h_pinned[4088..4089]; read them with frk_potrf_finish once the
// stream has been synchronised (lets several contexts factorise side by side from one host thread)
int frk_potrf_async(frk_ctx *ctx, int r, double *d_A) {
    FRK_REQUIRE(ctx && d_A && r > 0, "frk_potrf: bad argument");
    double *status = ctx->d_small + 4096 - 8;
    if (int rc = gemm_set_attrs()) return rc;
    if (int rc = run_graphed(ctx, {1, (uintptr_t)r, (uintptr_t)d_A, 0, 0},
                             [&]() { return potrf_use_v1() ? potrf_enqueue(ctx, r, d_A, status) : potrf2_enqueue(ctx, r, d_A, status, ctx->d_winv); })) return rc;
    FRK_CHECK_CUDA(cudaMemcpyAsync(ctx->h_pinned + 4096 - 8, status, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}

int frk_potrf_finish(frk_ctx *ctx, double *h_logdet) {
    const double *st = ctx->h_pinned + 4096 - 8;
    FRK_REQUIRE(st[1] == 0.0, "the leading minor of order %d is not positive definite", (int)st[1]);
    if (h_logdet) *h_logdet = 2.0 * st[0];
    return 0;
}

// same, without failing: returns the order of the first non-positive leading minor (0: positive definite)
int frk_potrf_status(frk_ctx *ctx, double *h_logdet) {
    const double *st = ctx->h_pinned + 4096 - 8;
    if (h_logdet) *h_logdet = 2.0 * st[0];
    return (int)st[1];
}

extern "C" int frk_potrf(frk_ctx *ctx, int r, double *d_A, double *h_logdet) {
    if (int rc = frk_potrf_async(ctx, r, d_A)) return rc;
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return frk_potrf_finish(ctx, h_logdet);
}

// chol2inv.  d_work receives X = L^-1 (lower triangular, zeros above the diagonal).
static int potri_enqueue(frk_ctx *ctx, int r, const double *d_L, double *d_Ainv, double *d_work);

extern "C" int frk_potri(frk_ctx *ctx, int r, const double *d_L, double *d_Ainv, double *d_work) {
    FRK_REQUIRE(ctx && d_L && d_Ainv && d_work && r > 0, "frk_potri: bad argument");
    FRK_REQUIRE(d_work != d_Ainv && d_work != d_L, "frk_potri: work must not alias L or the output");
    return run_graphed(ctx, {2, (uintptr_t)r, (uintptr_t)d_L, (uintptr_t)d_Ainv, (uintptr_t)d_work},
                       [&]() { return potri_enqueue(ctx, r, d_L, d_Ainv, d_work); });
}

// profile names per doubling level (interned: the profiler keys on the pointer's string)
static const char *level_tag(const char *prefix, long long b) {
    static std::map<std::string, std::string> names;
    const std::string key = std::string(prefix) + std::to_string(b) + "]";
    return names.emplace(key, key).first->second.c_str();
}

static int potri_enqueue(frk_ctx *ctx, int r, const double *d_L, double *d_Ainv, double *d_work) {
    const size_t ld = (size_t)r;
    double *X = d_work, *T = d_Ainv;
    FRK_CHECK_CUDA(cudaMemsetAsync(X, 0, ld * r * sizeof(double), ctx->stream));
    const int nblk = (int)cdiv(r, NB);
    FRK_WORK(ctx, (double)nblk * NB * NB * NB / 3.0, 0);
    FRK_LAUNCH(ctx, trsm_rows_kernel<true>, nblk, BLK_T, 0, d_L, r, X, r, 0, 0, r);
    // X21 = -X22 (L21 X11), doubling the block size; T (scratch) lives in the output buffer at X21's position
    for (long long b = NB; b < r; b *= 2) {
        const int npairs = (int)cdiv(r - b, 2 * b);                   // pairs whose second block is non-empty
        if (npairs <= 0) break;
        const int b2_last = (int)std::min<long long>(b, r - (2LL * (npairs - 1) + 1) * b);
        const long long stride = 2 * b * (1 + (long long)ld);
        const size_t off21 = (size_t)b, off22 = (size_t)b * (1 + ld);
        const double pairs_full = npairs - 1.0;
        {   // T = L21 X11   (X11 lower: k from the tile's first column)
            GemmArgs g = gemm_args((int)b, (int)b, (int)b, 1.0, d_L + off21, r, X, r, 0.0, T + off21, r);
            g.k_from_n = 1; g.sA = g.sB = g.sC = stride; g.nbatch = npairs; g.M_last = b2_last; g.K_last = (int)b;
            if (int rc = launch_gemm(ctx, g, false, true, (pairs_full * b + b2_last) * (double)b * b, level_tag("dgemm_kernel[potri L21 X11 b=", b))) return rc;
        }
        {   // X21 = -X22 T  (X22 lower: k up to the tile's last row)
            GemmArgs g = gemm_args((int)b, (int)b, (int)b, -1.0, X + off22, r, T + off21, r, 0.0, X + off21, r);
            g.k_to_m = 1; g.sA = g.sB = g.sC = stride; g.nbatch = npairs; g.M_last = b2_last; g.K_last = b2_last;
            if (int rc = launch_gemm(ctx, g, false, true, (pairs_full * b * b + (double)b2_last * b2_last) * (double)b, level_tag("dgemm_kernel[potri X22 T b=", b))) return rc;
        }
    }
    // A^-1 = X' X (lower triangle; X lower -> k >= max(row, col)), then mirror
    GemmArgs g3 = gemm_args(r, r, r, 1.0, X, r, X, r, 0.0, d_Ainv, r);
    g3.lower = 1; g3.k_from_m = 1; g3.k_from_n = 1;
    if (int rc = launch_gemm(ctx, g3, true, true, (double)r * r * r / 3.0, "dgemm_kernel[potri X'X]")) return rc;
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

__global__ void add_scaled_outer_kernel(int r, const double *__restrict__ A, const double *__restrict__ x, double alpha, double *__restrict__ out) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (int64_t)r * r; idx += (int64_t)gridDim.x * blockDim.x) {
        int i = (int)(idx % r), j = (int)(idx / r);
        out[idx] = A[idx] + alpha * (x[i] * x[j]);
    }
}

extern "C" int frk_add_scaled_outer(frk_ctx *ctx, int r, const double *d_A, const double *d_x, double alpha, double *d_out) {
    FRK_REQUIRE(ctx && d_A && d_x && d_out && r > 0, "frk_add_scaled_outer: bad argument");
    FRK_LAUNCH(ctx, add_scaled_outer_kernel, ctx->sm_count * 8, 256, 0, r, d_A, d_x, alpha, d_out);
    return 0;
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

// fixed-order tree over <= 1024 partial sums: 256 threads take 4 each (stride 256), then warp / block reduction
__global__ void __launch_bounds__(256) dot_final_kernel(int nblocks, const double *__restrict__ part, double *__restrict__ out) {
    __shared__ double sh[8];
    double v = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) v += part[b];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += sh[w];
        *out = t;
    }
}

// enqueue only: the result lands in ctx->h_pinned[4080] (valid once the stream has been synchronised)
int frk_dot_async(frk_ctx *ctx, int64_t n, const double *d_A, const double *d_B) {
    FRK_REQUIRE(ctx && d_A && d_B, "frk_dot: NULL argument");
    if (n == 0) { ctx->h_pinned[4080] = 0.0; return 0; }
    const int nblocks = (int)std::min<int64_t>(cdiv(n, 256), 1024);
    double *part = ctx->d_small, *out = ctx->d_small + 2048;
    FRK_LAUNCH(ctx, dot_partial_kernel, nblocks, 256, 0, n, d_A, d_B, part);
    FRK_LAUNCH(ctx, dot_final_kernel, 1, 256, 0, nblocks, part, out);
    FRK_CHECK_CUDA(cudaMemcpyAsync(ctx->h_pinned + 4080, out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}

extern "C" int frk_dot(frk_ctx *ctx, int64_t n, const double *d_A, const double *d_B, double *h_out) {
    FRK_REQUIRE(h_out, "frk_dot: NULL argument");
    if (int rc = frk_dot_async(ctx, n, d_A, d_B)) return rc;
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    *h_out = ctx->h_pinned[4080];
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

__global__ void block_scatter_kernel(int r, double *__restrict__ A, int ni, const int *__restrict__ idx, const double *__restrict__ in, double alpha) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)ni * ni; t += (int64_t)gridDim.x * blockDim.x) {
        int i = (int)(t % ni), j = (int)(t / ni);
        A[idx[i] + (size_t)idx[j] * r] = alpha * in[t];          // (alpha = 1: exact copy)
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
    FRK_LAUNCH(ctx, block_scatter_kernel, ctx->sm_count * 8, 256, 0, r, d_A, ni, d_idx, d_in, 1.0);
    return 0;
}

// A[idx, idx] = alpha * in: the scaled block is written straight into place and `in` stays as it is (model.cu keeps the unscaled
// exp(-D/tau)^-1 of the last K update for the next M-step)
int frk_block_scatter_scaled(frk_ctx *ctx, int r, double *d_A, int ni, const int *d_idx, const double *d_in, double alpha) {
    FRK_REQUIRE(ctx && d_A && d_idx && d_in, "frk_block_scatter_scaled: NULL argument");
    if (ni == 0) return 0;
    FRK_LAUNCH(ctx, block_scatter_kernel, ctx->sm_count * 8, 256, 0, r, d_A, ni, d_idx, d_in, alpha);
    return 0;
}

// x = X b for lower-triangular X (= L^-1 left in d_work by frk_potri), i.e. x = L^-1 b:
// (rDinv %*% S) %*% solve(R) of .loglik.ind (R/SREfit.R:211).  Column-split partial sums, then a fixed-order sum.
__global__ void __launch_bounds__(128) lower_mulv_partial_kernel(int r, const double *__restrict__ X, const double *__restrict__ b,
                                                                 double *__restrict__ part) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    const int jchunk = (r + SYMV_SPLIT - 1) / SYMV_SPLIT;
    const int j0 = blockIdx.y * jchunk, j1 = min(r, j0 + jchunk);
    if (i >= r) return;
    double s = 0.0;
    const int jend = min(j1, i + 1);
    for (int j = j0; j < jend; j++) s += X[i + (size_t)j * r] * __ldg(b + j);
    part[(size_t)blockIdx.y * r + i] = s;
}

extern "C" int frk_lower_mulv(frk_ctx *ctx, int r, const double *d_X, const double *d_b, double *d_x) {
    FRK_REQUIRE(ctx && d_X && d_b && d_x && r > 0, "frk_lower_mulv: bad argument");
    void *raw = nullptr;
    if (int rc = frk_scratch(ctx, (size_t)SYMV_SPLIT * r * sizeof(double), &raw)) return rc;
    dim3 grid((unsigned)cdiv(r, 128), SYMV_SPLIT);
    FRK_LAUNCH(ctx, lower_mulv_partial_kernel, grid, 128, 0, r, d_X, d_b, (double *)raw);
    FRK_LAUNCH(ctx, symv_final_kernel, (unsigned)cdiv(r, 128), 128, 0, r, (const double *)raw, d_x);
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

// ---------------------------------------------------------------------------------
// Kronecker helpers for the block-exponential K of a tensor-product (space x time) basis, R/SREfit.R:496-514,627-649:
// the (n_t n_s) x (n_t n_s) matrices kronecker(At, As) (time slowest, space fastest) are never formed.
// ---------------------------------------------------------------------------------
// partial sums of sum_{a,b} At[ta,tb] As[sa,sb] E[a,b],  a = sa + ns*ta   (sum(diag2(kronecker(Qi1,Qi2), eta2, symm=TRUE)))
__global__ void __launch_bounds__(256) kron_dot_partial_kernel(int nt, int ns, const double *__restrict__ At, const double *__restrict__ As,
                                                               const double *__restrict__ E, double *__restrict__ part) {
    const int64_t n = (int64_t)nt * ns, nn = n * n;
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = (int)(i % n), b = (int)(i / n);
        s += __ldg(At + a / ns + (size_t)nt * (b / ns)) * __ldg(As + a % ns + (size_t)ns * (b % ns)) * E[i];
    }
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

extern "C" int frk_kron_dot(frk_ctx *ctx, int nt, int ns, const double *d_At, const double *d_As, const double *d_E, double *h_out) {
    FRK_REQUIRE(ctx && d_At && d_As && d_E && h_out && nt > 0 && ns > 0, "frk_kron_dot: bad argument");
    const int64_t nn = (int64_t)nt * ns * nt * ns;
    const int nblocks = (int)std::min<int64_t>(cdiv(nn, 256), 1024);
    double *part = ctx->d_small, *out = ctx->d_small + 2048;
    FRK_LAUNCH(ctx, kron_dot_partial_kernel, nblocks, 256, 0, nt, ns, d_At, d_As, d_E, part);
    FRK_LAUNCH(ctx, dot_final_kernel, 1, 256, 0, nblocks, part, out);
    FRK_CHECK_CUDA(cudaMemcpyAsync(ctx->h_pinned, out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    *h_out = ctx->h_pinned[0];
    return 0;
}

// A[idx[a], idx[b]] = scale * At[ta,tb] * As[sa,sb]   (K_i = kronecker(exp(-Dt/tau2), exp(-Ds/tau1)) / omega_i placed by bdiag +
// reverse_permute, R/SREfit.R:627-649; idx = the ascending tensor columns of the resolution: space fastest, time slowest)
__global__ void kron_scatter_kernel(int r, int nt, int ns, const double *__restrict__ At, const double *__restrict__ As, double scale,
                                    const int *__restrict__ idx, double *__restrict__ A) {
    const int64_t n = (int64_t)nt * ns, nn = n * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = (int)(i % n), b = (int)(i / n);
        A[idx[a] + (size_t)idx[b] * r] = scale * __ldg(At + a / ns + (size_t)nt * (b / ns)) * __ldg(As + a % ns + (size_t)ns * (b % ns));
    }
}

extern "C" int frk_kron_scatter(frk_ctx *ctx, int r, int nt, int ns, const double *d_At, const double *d_As, double scale,
                                const int *d_idx, double *d_A) {
    FRK_REQUIRE(ctx && d_At && d_As && d_idx && d_A && nt > 0 && ns > 0, "frk_kron_scatter: bad argument");
    FRK_LAUNCH(ctx, kron_scatter_kernel, ctx->sm_count * 8, 256, 0, r, nt, ns, d_At, d_As, scale, d_idx, d_A);
    return 0;
}

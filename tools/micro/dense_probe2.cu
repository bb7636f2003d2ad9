// Probe: back-to-back launch cost of the small dense kernels (includes dense.cu directly).
#include "../../frk_b200/csrc/dense.cu"
__global__ void empty_k() {}
__global__ void __launch_bounds__(BLK_T) potf2_timed(double *__restrict__ A, int lda, int kb, int col_offset, double *__restrict__ status, long long *ts) {
    __shared__ double colbuf[2][NB];
    __shared__ double rdiag[2];
    __shared__ double rsq[NB];
    __shared__ double slog[BLK_T / 32];
    long long t0 = clock64();
    const int i = threadIdx.x / SPLIT, h = threadIdx.x % SPLIT;
    double a[NBQ];
#pragma unroll
    for (int kk = 0; kk < NBQ; kk++) {
        const int k = kk * SPLIT + h;
        a[kk] = (i < kb && k < kb && k <= i) ? A[i + (size_t)k * lda] : (k == i ? 1.0 : 0.0);
    }
    __syncthreads(); long long t1 = clock64();
    int bad = 0;
#pragma unroll
    for (int jj = 0; jj < NBQ; jj++) {                 // columns j = 4*jj + hh; a[jj] is lane hh's entry of column j
#pragma unroll 1
        for (int hh = 0; hh < SPLIT; hh++) {
            const int j = jj * SPLIT + hh;
            if (h == hh) {
                const double v = a[jj];                // unscaled column j (rows i < j hold don't-care values)
                colbuf[j & 1][i] = v;
                if (i == j) rdiag[j & 1] = (v > 0.0) ? __drcp_rn(v) : -1.0;     // also catches NaN
            }
            __syncthreads();
            const double rd = rdiag[j & 1];
            if (rd < 0.0) { bad = j + 1; break; }      // uniform: not positive definite
            const double c = colbuf[j & 1][i] * rd;
            const double *cb = &colbuf[j & 1][h];
            if (h > hh) a[jj] -= c * cb[jj * SPLIT];
#pragma unroll
            for (int kk = jj + 1; kk < NBQ; kk++) a[kk] -= c * cb[kk * SPLIT];   // entries with k > i are never read
        }
        if (bad) break;
    }
    long long t2 = clock64();
    if (bad) {
        if (threadIdx.x == 0 && status[1] == 0.0) status[1] = (double)(col_offset + bad);
        return;
    }
    // d_i sits on the diagonal (lane h = i % 4 of row i): scale column j by 1/sqrt(d_j)
    double lg = 0.0;
#pragma unroll
    for (int kk = 0; kk < NBQ; kk++) {
        if (kk * SPLIT + h == i) {
            const double sq = sqrt(a[kk]);
            rsq[i] = 1.0 / sq;
            a[kk] = sq;
            lg = (i < kb) ? log(sq) : 0.0;
        }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < NBQ; kk++) {
        const int k = kk * SPLIT + h;
        if (i < kb && k < kb && k <= i) A[i + (size_t)k * lda] = (k == i) ? a[kk] : a[kk] * rsq[k];
    }
    lg = warp_sum(lg);
    if ((threadIdx.x & 31) == 0) slog[threadIdx.x >> 5] = lg;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < BLK_T / 32; w++) t += slog[w];
        status[0] += t;
    }
    if (threadIdx.x == 0) { ts[0] = t1 - t0; ts[1] = t2 - t1; ts[2] = clock64() - t2; }
}

int main() {
    const int r = 3276;
    double *A, *st; cudaMalloc(&A, (size_t)r * r * 8); cudaMalloc(&st, 64);
    cudaMemset(A, 0, (size_t)r * r * 8); cudaMemset(st, 0, 64);
    // identity-ish SPD so potf2 succeeds: fill the diagonal with 4
    double *h = (double *)malloc((size_t)r * r * 8);
    for (size_t i = 0; i < (size_t)r * r; i++) h[i] = 0.001;
    for (int i = 0; i < r; i++) h[i + (size_t)i * r] = 4.0;
    cudaMemcpy(A, h, (size_t)r * r * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms; const int N = 200;
    auto report = [&](const char *nm) { cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("%-40s %.2f us per launch (%s)\n", nm, 1e3 * ms / N, cudaGetErrorString(cudaGetLastError())); };
    for (int w = 0; w < 2; w++) {
        cudaEventRecord(e0); for (int i = 0; i < N; i++) empty_k<<<1, 32>>>(); report("empty kernel");
        cudaEventRecord(e0); for (int i = 0; i < N; i++) potf2_kernel<<<1, BLK_T>>>(A + 64 + 64 * (size_t)r, r, 64, 0, st); report("potf2 (same block, refactoring L: fine)");
        cudaEventRecord(e0); for (int i = 0; i < N; i++) trsm_rows_kernel<false><<<20, BLK_T>>>(A, r, A + 2048, r, 1280, 64, r); report("trsm_rows<false> 20 CTAs");
        cudaEventRecord(e0); for (int i = 0; i < N; i++) trsm_rows_kernel<false><<<50, BLK_T>>>(A, r, A + 64, r, 3200, 64, r); report("trsm_rows<false> 50 CTAs");
    }
    { long long *ts; cudaMallocManaged(&ts, 64); for (int w = 0; w < 3; w++) { potf2_timed<<<1, BLK_T>>>(A + 64 + 64 * (size_t)r, r, 64, 0, st, ts); cudaDeviceSynchronize(); printf("potf2 cycles: load %lld  loop %lld  epilogue %lld\n", ts[0], ts[1], ts[2]); } }
    cudaFuncSetAttribute(dgemm_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    for (int K : {64, 256}) for (int gx : {9, 25}) for (int gy : {1, 3, 50}) {
        GemmArgs g = gemm_args(gx * 128, gy * 64, K, -1.0, A, r, A + 64, r, 1.0, A + 1024 * (size_t)r, r);
        if (gy * 64 + 1024 > r || gx * 128 > r) continue;
        dim3 grid(gx, gy, 1);
        cudaEventRecord(e0); for (int i = 0; i < N; i++) dgemm_kernel<false, false, true><<<grid, GEMM_T, GEMM_SMEM>>>(g);
        char nm[64]; snprintf(nm, 64, "dgemm NT K=%d grid %dx%d", K, gx, gy); report(nm);
    }
    return 0;
}

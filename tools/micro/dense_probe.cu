// Probe: back-to-back launch cost of the small dense kernels (includes dense.cu directly).
#include "../../frk_b200/csrc/dense.cu"
__global__ void empty_k() {}
int main() {
    const int r = 3276;
    double *A, *st; cudaMalloc(&A, (size_t)r * r * 8); cudaMalloc(&st, 64);
    cudaMemset(A, 0, (size_t)r * r * 8); cudaMemset(st, 0, 64);
    // identity-ish SPD so potf2 succeeds: fill the diagonal with 4
    double *h = (double *)malloc((size_t)r * r * 8);
    for (size_t i = 0; i < (size_t)r * r; i++) h[i] = 0.001;
    for (int i = 0; i < r; i++) h[i + (size_t)i * r] = 4.0;
    cudaMemcpy(A, h, (size_t)r * r * 8, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTF2_SMEM);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms; const int N = 200;
    auto report = [&](const char *nm) { cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("%-40s %.2f us per launch (%s)\n", nm, 1e3 * ms / N, cudaGetErrorString(cudaGetLastError())); };
    for (int w = 0; w < 2; w++) {
        cudaEventRecord(e0); for (int i = 0; i < N; i++) empty_k<<<1, 32>>>(); report("empty kernel");
        cudaEventRecord(e0); for (int i = 0; i < N; i++) potf2_kernel<<<1, BLK_T, POTF2_SMEM>>>(A + 64 + 64 * (size_t)r, r, 64, 0, st, (const double *)nullptr, 0, 0); report("potf2 (same block, refactoring L: fine)");
        cudaEventRecord(e0); for (int i = 0; i < N; i++) trsm_rows_kernel<false><<<20, BLK_T>>>(A, r, A + 2048, r, 1280, 64, r); report("trsm_rows<false> 20 CTAs");
        cudaEventRecord(e0); for (int i = 0; i < N; i++) trsm_rows_kernel<false><<<50, BLK_T>>>(A, r, A + 64, r, 3200, 64, r); report("trsm_rows<false> 50 CTAs");
    }
    cudaFuncSetAttribute(dgemm_kernel<false, false, true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    for (int K : {64, 256}) for (int gx : {9, 25}) for (int gy : {1, 3, 50}) {
        GemmArgs g = gemm_args(gx * 128, gy * 64, K, -1.0, A, r, A + 64, r, 1.0, A + 1024 * (size_t)r, r);
        if (gy * 64 + 1024 > r || gx * 128 > r) continue;
        dim3 grid(gx, gy, 1);
        cudaEventRecord(e0); for (int i = 0; i < N; i++) dgemm_kernel<false, false, true, 128><<<grid, GEMM_T, GEMM_SMEM>>>(g);
        char nm[64]; snprintf(nm, 64, "dgemm NT K=%d grid %dx%d", K, gx, gy); report(nm);
    }
    return 0;
}

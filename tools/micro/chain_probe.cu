// Latency of the per-column dependency chain of a 16x16 register Cholesky (one warp): which op dominates?
#include <cstdio>
#include <cuda_runtime.h>
template <int VAR>
__global__ void k(double *out, const double *in, long long *ts) {
    const int lane = threadIdx.x & 31, i = lane & 15;
    double d[16];
#pragma unroll
    for (int c = 0; c < 16; c++) d[c] = in[i * 16 + c];
    long long t0 = clock64();
#pragma unroll
    for (int j = 0; j < 16; j++) {
        double piv = __shfl_sync(0xffffffffu, d[j], j);
        double rs;
        if (VAR == 0) rs = rsqrt(piv);
        else if (VAR == 1) rs = 1.0 / sqrt(piv);
        else if (VAR == 2) rs = piv * 0.25;                  // no special function at all
        else if (VAR == 3) { float f = rsqrtf((float)piv); double y = (double)f; y = y * (1.5 - 0.5 * piv * y * y); y = y * (1.5 - 0.5 * piv * y * y); rs = y; }
        else rs = __drcp_rn(piv);
        const double lij = (i == j) ? piv * rs : d[j] * rs;
        d[j] = lij;
#pragma unroll
        for (int c = j + 1; c < 16; c++) {
            if (VAR == 5) d[c] -= lij * d[c - 1];            // no shuffles in the update
            else d[c] -= lij * __shfl_sync(0xffffffffu, lij, c);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < 16; c++) s += d[c];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) ts[VAR] = t1 - t0;
}
int main() {
    double h[256]; for (int i = 0; i < 16; i++) for (int c = 0; c < 16; c++) h[i * 16 + c] = (i == c) ? 20.0 : 0.01 * (i + c);
    double *in, *out; long long *ts; cudaMalloc(&in, sizeof(h)); cudaMalloc(&out, 4096); cudaMallocManaged(&ts, 64);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int w = 0; w < 2; w++) {
        k<0><<<1, 32>>>(out, in, ts); k<1><<<1, 32>>>(out, in, ts); k<2><<<1, 32>>>(out, in, ts); k<3><<<1, 32>>>(out, in, ts); k<4><<<1, 32>>>(out, in, ts); k<5><<<1, 32>>>(out, in, ts);
        cudaDeviceSynchronize();
        printf("16 column steps: rsqrt %lld | 1/sqrt %lld | no special fn %lld | f32 rsqrt + 2 Newton %lld | drcp %lld | rsqrt, no shuffles in update %lld cycles\n", ts[0], ts[1], ts[2], ts[3], ts[4], ts[5]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}

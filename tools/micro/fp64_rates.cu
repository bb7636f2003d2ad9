// Microbenchmark: fp64 vector (DFMA) and tensor (DMMA m8n8k4) issue rates per SM on B200; LDS broadcast; bar.sync.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_k(double *out, int iters, double a, double b) {
    double x[8];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0; for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma_k(double *out, int iters, double a, double b) {
    double c[8][2];
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0; for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void chain_k(double *out, int iters, double a, double b) {   // dependent DFMA chain: latency
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) x = fma(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) printf("dependent DFMA latency: %.1f cycles\n", double(t1 - t0) / iters);
}
__global__ void div_k(double *out, int iters, double a) {
    double x = threadIdx.x + 2.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) x = a / x + 1.5;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) printf("dependent DDIV+DADD latency: %.1f cycles\n", double(t1 - t0) / iters);
}
__global__ void bar_k(double *out, int iters) {
    __shared__ double s[256];
    long long t0 = clock64();
    double x = 0;
    for (int it = 0; it < iters; it++) { s[threadIdx.x] = x; __syncthreads(); x += s[(threadIdx.x + 1) & 255]; __syncthreads(); }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) printf("STS + bar + LDS + DADD + bar (256 thr): %.1f cycles\n", double(t1 - t0) / iters);
}
int main() {
    double *out; cudaMalloc(&out, 1 << 24);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 148;
    for (int warps = 4; warps <= 32; warps *= 2) {
        int iters = 20000; float ms;
        dfma_k<<<sms, warps * 32>>>(out, 10, 1.0000001, 1e-9);
        cudaEventRecord(e0); dfma_k<<<sms, warps * 32>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * iters * warps * 32 * sms;
        printf("DFMA  %2d warps/SM: %.2f TFLOP/s\n", warps, fl / ms / 1e9);
        dmma_k<<<sms, warps * 32>>>(out, 10, 1.0000001, 1e-9);
        cudaEventRecord(e0); dmma_k<<<sms, warps * 32>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        fl = 512.0 * 8 * iters * warps * sms;
        printf("DMMA  %2d warps/SM: %.2f TFLOP/s\n", warps, fl / ms / 1e9);
    }
    chain_k<<<1, 32>>>(out, 100000, 1.0000001, 1e-9);
    div_k<<<1, 32>>>(out, 100000, 3.0);
    bar_k<<<1, 256>>>(out, 100000);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

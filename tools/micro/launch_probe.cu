#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_empty() {}
__global__ void k_smem(double *o) { __shared__ double s[4400]; s[threadIdx.x] = threadIdx.x; __syncthreads(); if (o) o[threadIdx.x] = s[(threadIdx.x + 1) % 256]; }
__global__ void __launch_bounds__(256) k_regs(double *o, int n) { double x[40]; for (int i = 0; i < 40; i++) x[i] = threadIdx.x + i; for (int it = 0; it < n; it++) for (int i = 0; i < 40; i++) x[i] = x[i] * 1.0001 + x[(i + 1) % 40]; double s = 0; for (int i = 0; i < 40; i++) s += x[i]; if (o) o[threadIdx.x] = s; }
__global__ void k_gload(const double *a, double *o, int ld) { double s = 0; for (int c = 0; c < 16; c++) s += a[threadIdx.x + blockIdx.x * 64 + (size_t)c * ld]; o[threadIdx.x + blockIdx.x * 256] = s; }
__global__ void k_dyn(double *o) { extern __shared__ double d[]; d[threadIdx.x] = 1; __syncthreads(); if (o) o[threadIdx.x] = d[(threadIdx.x + 1) % 256]; }
int main() {
    double *o, *a; cudaMalloc(&o, 1 << 22); cudaMalloc(&a, (size_t)3276 * 3276 * 8); cudaMemset(a, 0, (size_t)3276 * 3276 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms; const int N = 500;
    auto rep = [&](const char *nm) { cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); printf("%-44s %.2f us/launch\n", nm, 1e3 * ms / N); };
    cudaFuncSetAttribute(k_dyn, cudaFuncAttributeMaxDynamicSharedMemorySize, 92160);
    for (int w = 0; w < 2; w++) {
        cudaEventRecord(e0); for (int i = 0; i < N; i++) k_empty<<<1, 32>>>(); rep("empty 1x32");
        cudaEventRecord(e0); for (int i = 0; i < N; i++) k_empty<<<50, 256>>>(); rep("empty 50x256");
        cudaEventRecord(e0); for (int i = 0; i < N; i++) k_smem<<<50, 256>>>(o); rep("35KB static smem 50x256");
        cudaEventRecord(e0); for (int i = 0; i < N; i++) k_regs<<<50, 256>>>(o, 1); rep("regs 50x256");
        cudaEventRecord(e0); for (int i = 0; i < N; i++) k_gload<<<50, 256>>>(a, o, 3276); rep("16 strided global loads 50x256");
        cudaEventRecord(e0); for (int i = 0; i < N; i++) k_dyn<<<50, 256, 92160>>>(o); rep("92KB dynamic smem 50x256");
        cudaEventRecord(e0); for (int i = 0; i < N; i++) { k_dyn<<<50, 256, 92160>>>(o); k_smem<<<50, 256>>>(o); k_empty<<<1, 256>>>(); } rep("alternating dyn/static/empty (x3 launches)");
    }
    cudaStream_t st; cudaStreamCreate(&st); cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < N; i++) k_smem<<<50, 256, 0, st>>>(o);
    cudaStreamEndCapture(st, &g); cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
    cudaEventRecord(e0, st); cudaGraphLaunch(ge, st); cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %.2f us/launch\n", "graph of 500 x (35KB static smem 50x256)", 1e3 * ms / N);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}

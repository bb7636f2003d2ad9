// Timing of the 16 x 16 register Cholesky variants of dense.cu (one warp, clock64): the column-sequential chain.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o potf2_variants potf2_variants.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int PB = 16;
__device__ __forceinline__ double rsqrt_fast(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double ye = y * e;
    return fma(ye, p, y);
}

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

__device__ __forceinline__ int warp_potf2_inv_16(double x[PB], int lane) {
    const bool factor = lane < PB;
    int bad = 0;
    double rs = rsqrt_fast(__shfl_sync(0xffffffffu, x[0], 0));
#pragma unroll
    for (int j = 0; j < PB; j++) {
        if (factor && lane == j && !(x[j] > 1e-290)) bad = j + 1;     // x[j] on lane j is the pivot, updated through column j-1
        const double mine = x[j] * rs;
        x[j] = mine;
        double rs_next = 1.0;
        if (j + 1 < PB) {
            const double pn = fma(-mine, mine, x[j + 1 < PB ? j + 1 : j]);      // lane j+1: its next pivot (others: unused)
            rs_next = __shfl_sync(0xffffffffu, rsqrt_fast(pn), j + 1);
        }
        const double lij = factor ? mine : 0.0;                          // what the shuffles read comes from the factor lanes only
#pragma unroll
        for (int k = j + 1; k < PB; k++) x[k] = fma(-mine, __shfl_sync(0xffffffffu, lij, k), x[k]);
        rs = rs_next;
    }
    // first failing column over the factor lanes (a failed pivot poisons everything after it; only the first index is reported)
    unsigned v = bad ? (unsigned)bad : 0xffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v == 0xffffu ? 0 : (int)v;
}


// V2: the new chain (lane j+1 forms its own next pivot, only the rsqrt travels) WITHOUT the inverse lanes
__device__ __forceinline__ int warp_potf2_v2(double x[PB], int lane) {
    const int i = lane & 15;
    int bad = 0;
    double rs = rsqrt_fast(__shfl_sync(0xffffffffu, x[0], 0));
#pragma unroll
    for (int j = 0; j < PB; j++) {
        if (i == j && !(x[j] > 1e-290)) bad = j + 1;
        const double mine = x[j] * rs;
        x[j] = mine;
        double rs_next = 1.0;
        if (j + 1 < PB) {
            const double pn = fma(-mine, mine, x[j + 1 < PB ? j + 1 : j]);
            rs_next = __shfl_sync(0xffffffffu, rsqrt_fast(pn), j + 1);
        }
#pragma unroll
        for (int k = j + 1; k < PB; k++) x[k] = fma(-mine, __shfl_sync(0xffffffffu, mine, k), x[k]);
        rs = rs_next;
    }
    return bad;
}
// V3: as V2 but the next pivot is broadcast and every lane takes its own rsqrt (shuffle before the special function)
__device__ __forceinline__ int warp_potf2_v3(double x[PB], int lane) {
    const int i = lane & 15;
    int bad = 0;
    double rs = rsqrt_fast(__shfl_sync(0xffffffffu, x[0], 0));
#pragma unroll
    for (int j = 0; j < PB; j++) {
        if (i == j && !(x[j] > 1e-290)) bad = j + 1;
        const double mine = x[j] * rs;
        x[j] = mine;
        double rs_next = 1.0;
        if (j + 1 < PB) {
            const double pn = fma(-mine, mine, x[j + 1 < PB ? j + 1 : j]);
            rs_next = rsqrt_fast(__shfl_sync(0xffffffffu, pn, j + 1));
        }
#pragma unroll
        for (int k = j + 1; k < PB; k++) x[k] = fma(-mine, __shfl_sync(0xffffffffu, mine, k), x[k]);
        rs = rs_next;
    }
    return bad;
}
// V4: V2 with the update loop issued BEFORE the next-pivot chain in program order
__device__ __forceinline__ int warp_potf2_v4(double x[PB], int lane) {
    const int i = lane & 15;
    int bad = 0;
    double rs = rsqrt_fast(__shfl_sync(0xffffffffu, x[0], 0));
#pragma unroll
    for (int j = 0; j < PB; j++) {
        if (i == j && !(x[j] > 1e-290)) bad = j + 1;
        const double mine = x[j] * rs;
        x[j] = mine;
        const double pn = fma(-mine, mine, x[j + 1 < PB ? j + 1 : j]);
        const double rn = rsqrt_fast(pn);
#pragma unroll
        for (int k = j + 1; k < PB; k++) x[k] = fma(-mine, __shfl_sync(0xffffffffu, mine, k), x[k]);
        rs = __shfl_sync(0xffffffffu, rn, j + 1 < PB ? j + 1 : j);
    }
    return bad;
}
template <int VAR>
__global__ void k(double *out, const double *in, long long *ts) {
    const int lane = threadIdx.x & 31, i = lane & 15;
    double d[16];
#pragma unroll
    for (int c = 0; c < 16; c++) d[c] = (VAR == 1 && lane >= 16) ? (c == i ? 1.0 : 0.0) : in[i + 16 * c];
    __syncwarp();
    long long t0 = clock64();
    int bad;
    if (VAR == 0) bad = warp_potf2_16(d, i, nullptr);
    else if (VAR == 1) bad = warp_potf2_inv_16(d, lane);
    else if (VAR == 2) bad = warp_potf2_v2(d, lane);
    else if (VAR == 3) bad = warp_potf2_v3(d, lane);
    else bad = warp_potf2_v4(d, lane);
    long long t1 = clock64();
    double s = bad;
#pragma unroll
    for (int c = 0; c < 16; c++) s += (c <= i || (VAR == 1 && lane >= 16)) ? d[c] * (c + 1) : 0.0;
    out[VAR * 32 + threadIdx.x] = s;
    if (threadIdx.x == 0) ts[VAR] = t1 - t0;
}
int main() {
    double h[256];
    for (int i = 0; i < 16; i++) for (int c = 0; c < 16; c++) h[i + 16 * c] = (i == c) ? 20.0 + i : 0.01 * (i + c) + 0.3 / (1 + abs(i - c));
    double *in, *out; long long *ts; cudaMalloc(&in, sizeof(h)); cudaMallocManaged(&out, 5 * 32 * 8); cudaMallocManaged(&ts, 64);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int w = 0; w < 3; w++) {
        k<0><<<1, 32>>>(out, in, ts); k<1><<<1, 32>>>(out, in, ts); k<2><<<1, 32>>>(out, in, ts); k<3><<<1, 32>>>(out, in, ts); k<4><<<1, 32>>>(out, in, ts);
        cudaDeviceSynchronize();
        printf("16 columns: old (all lanes take the next pivot) %lld | fused factor + inverse %lld | own-pivot chain %lld | pivot broadcast %lld | chain after update %lld cycles\n",
               ts[0], ts[1], ts[2], ts[3], ts[4]);
    }
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    for (int l = 0; l < 16; l++) { c0 += out[l]; c1 += out[32 + l]; c2 += out[64 + l]; c3 += out[96 + l]; c4 += out[128 + l]; }
    printf("checksums (factor lanes): %.17g %.17g %.17g %.17g %.17g\n", c0, c1, c2, c3, c4);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}

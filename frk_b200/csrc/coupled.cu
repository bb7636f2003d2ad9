// Observations coupled through shared BAUs: the non-diagonal fine-scale covariance.
//
// Reference: when two observation footprints share a BAU (overlapping areal supports, or several point observations in one BAU with
// average_in_BAU = FALSE) Vfs = C_Z diag(fs) C_Z' (R/SRE.R:498) has off-diagonal entries and D = sigma2fs Vfs + Ve is no longer
// diagonal.  The reference then takes its general branches:  chol(D) / solve(cholD) in .loglik.ind (R/SREfit.R:186-189,194,206-211)
// and .SRE.Estep.ind (:240-252), and the general J(sigma2fs) of the M-step (:296-328) with Dinv %*% Vfs %*% Dinv.
//
// D is block diagonal: its blocks are the connected COMPONENTS of the graph "observations i, j share a BAU".  Everything here works
// per component (one CTA each, blocks of at most NCMAX observations held in shared memory):
//  * whiten:  L_c = chol(D_c);  S~ = L^-1 S, Z~ = L^-1 Z, X~ = L^-1 X.  In the whitened variables D is the identity, so the existing
//    bucketed Gram / right-hand-side kernels produce  S'D^-1 S,  S'D^-1 rho,  rho'D^-1 rho  unchanged; logdet(cholD) comes from the
//    factors.  The rows of a component share the union of their column patterns (fixed; built once on the host), the values are
//    recomputed whenever sigma2fs changes.
//  * cross:   B_c = S_c Sigma S_c'  (the quadratic forms s_i' Sigma s_j for all pairs inside a component), once per M-step.
//  * msums:   for one sigma2fs the sums behind J and the GLS estimate of alpha,
//             uxx = X'D^-1 X, uxz = X'D^-1 (Z - t), sv = tr(D^-1 V), and with W = D^-1 V D^-1, e = Z - t:
//             w0 = tr(W B) + e'W e, wx = -X'W e, wxx = X'W X   --  the same vector frk_mstep_sums returns in the diagonal case, so
//             the root search and the GLS step of the M-step (model.cu) are shared.
#include "common.cuh"
#include "coupled.cuh"
#include <algorithm>
#include <numeric>

constexpr int NCMAX = FRK_COUPLED_NCMAX;     // observations per component held in shared memory
constexpr int CLD = NCMAX + 1;
constexpr int CP_T = 128;
constexpr int MAXP_C = 4;

struct frk_coupling {
    frk_ctx *ctx = nullptr;
    int64_t m = 0;
    int r = 0, p = 0, ncomp = 0, npairs = 0;
    int *cptr = nullptr, *cmem = nullptr;          // component -> member observations (ascending)
    int64_t *voff = nullptr;                       // component -> offset of its n_c x n_c block (column-major) in Vc / Bc
    double *Vc = nullptr, *Bc = nullptr;
    int *pair_i = nullptr, *pair_j = nullptr;      // all pairs a <= b inside a component: observation indices ...
    int64_t *pair_dst = nullptr, *pair_dst2 = nullptr;   // ... and where (a,b) / (b,a) sit in Bc
    // whitened copy of S: every row of a component holds the union of the component's columns
    int *rowptr_w = nullptr, *col_w = nullptr, *dst = nullptr;
    int64_t nnz = 0, nnz_w = 0;
    double *val_w = nullptr, *Zw = nullptr, *Xw = nullptr, *ones = nullptr, *zeros = nullptr;
    frk_rowbuckets *bk_w = nullptr;
    double *part = nullptr;                        // per-CTA partial sums
    int grid = 0;
    std::vector<void *> allocs;
};

namespace {

template <typename T>
int cp_alloc(frk_coupling *C, T **out, size_t n) {
    void *p = nullptr;
    FRK_CHECK_CUDA(cudaMalloc(&p, std::max<size_t>(n * sizeof(T), 16)));
    C->allocs.push_back(p);
    *out = (T *)p;
    return 0;
}

template <typename T>
int cp_upload(frk_coupling *C, T **out, const std::vector<T> &v) {
    if (int rc = cp_alloc(C, out, v.size())) return rc;
    if (!v.empty()) FRK_CHECK_CUDA(cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

// D_c = s2 V_c + diag(ve) into shared memory (lower triangle used), then the Cholesky factor in place.  Returns false if a pivot is
// not positive.  All threads of the CTA.
__device__ bool load_and_factor(double *Dm, int n, const double *__restrict__ V, const int *__restrict__ mem, const double *__restrict__ ve,
                                double s2, int *flag) {
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int i = idx % n, j = idx / n;
        double v = s2 * V[idx];
        if (i == j) v += ve[mem[i]];
        Dm[i * CLD + j] = v;
    }
    if (threadIdx.x == 0) *flag = 0;
    __syncthreads();
    for (int j = 0; j < n; j++) {
        if (threadIdx.x == 0) {
            const double d = Dm[j * CLD + j];
            if (!(d > 0.0)) *flag = 1;
            Dm[j * CLD + j] = sqrt(d);
        }
        __syncthreads();
        if (*flag) return false;
        const double dj = Dm[j * CLD + j];
        for (int i = j + 1 + threadIdx.x; i < n; i += blockDim.x) Dm[i * CLD + j] /= dj;
        __syncthreads();
        const int nb = n - j - 1;
        for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) {
            const int i = j + 1 + idx % nb, k = j + 1 + idx / nb;
            if (k <= i) Dm[i * CLD + k] -= Dm[i * CLD + j] * Dm[k * CLD + j];
        }
        __syncthreads();
    }
    return true;
}

__global__ void scatter_kernel(int64_t nnz, const int *__restrict__ dst, const double *__restrict__ val, double *__restrict__ val_w) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) val_w[dst[e]] = val[e];
}

// one CTA per component (grid-stride): factor D_c, forward-substitute every column of the component's rows of [S~ | Z~ | X~] in place.
// part[2 blockIdx.x] += sum log diag L (x2 = logdet(cholD) in the reference's sense), part[.. + 1] = failure flag
__global__ void __launch_bounds__(CP_T) whiten_kernel(int ncomp, const int *__restrict__ cptr, const int *__restrict__ cmem,
                                                      const int64_t *__restrict__ voff, const double *__restrict__ Vc,
                                                      const double *__restrict__ ve, double s2, int64_t m, int p,
                                                      const int *__restrict__ rowptr_w, double *__restrict__ val_w,
                                                      double *__restrict__ Zw, double *__restrict__ Xw, double *__restrict__ part) {
    extern __shared__ __align__(16) double csm[];
    double *Dm = csm;
    __shared__ int flag;
    __shared__ int smem_mem[NCMAX];
    double lsum = 0.0, bad = 0.0;
    for (int c = blockIdx.x; c < ncomp; c += gridDim.x) {
        const int c0 = cptr[c], n = cptr[c + 1] - c0;
        __syncthreads();
        if (n == 1) {                                      // the common case: a diagonal entry
            const int i = cmem[c0];
            const double d = s2 * Vc[voff[c]] + ve[i];
            if (!(d > 0.0)) { bad = 1.0; continue; }
            const double l = sqrt(d);
            if (threadIdx.x == 0) lsum += log(l);
            const int b = rowptr_w[i], k = rowptr_w[i + 1] - b;
            for (int u = threadIdx.x; u < k + 1 + p; u += blockDim.x) {
                double *x = u < k ? val_w + b + u : (u == k ? Zw + i : Xw + i + (int64_t)(u - k - 1) * m);
                *x = *x / l;
            }
            continue;
        }
        for (int a = threadIdx.x; a < n; a += blockDim.x) smem_mem[a] = cmem[c0 + a];
        __syncthreads();
        if (!load_and_factor(Dm, n, Vc + voff[c], smem_mem, ve, s2, &flag)) { bad = 1.0; continue; }
        if (threadIdx.x == 0) for (int a = 0; a < n; a++) lsum += log(Dm[a * CLD + a]);
        const int k = rowptr_w[smem_mem[0] + 1] - rowptr_w[smem_mem[0]];       // all rows of the component hold the same columns
        for (int u = threadIdx.x; u < k + 1 + p; u += blockDim.x) {
            for (int a = 0; a < n; a++) {
                const int ia = smem_mem[a];
                double *xa = u < k ? val_w + rowptr_w[ia] + u : (u == k ? Zw + ia : Xw + ia + (int64_t)(u - k - 1) * m);
                double acc = *xa;
                for (int j = 0; j < a; j++) {
                    const int ij = smem_mem[j];
                    const double yj = u < k ? val_w[rowptr_w[ij] + u] : (u == k ? Zw[ij] : Xw[ij + (int64_t)(u - k - 1) * m]);
                    acc -= Dm[a * CLD + j] * yj;
                }
                *xa = acc / Dm[a * CLD + a];
            }
        }
    }
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = lsum; part[2 * blockIdx.x + 1] = bad; }      // (bad is uniform over the CTA)
}

// out[0] += 2 * sum of the per-CTA log sums (fixed order); out[1] = any failure
__global__ void whiten_finish_kernel(int grid, const double *__restrict__ part, double *__restrict__ logdet_slot, double *__restrict__ status) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0, bad = 0.0;
        for (int b = 0; b < grid; b++) { s += part[2 * b]; bad += part[2 * b + 1]; }
        *logdet_slot += 2.0 * s;
        *status = bad;
    }
}

// one warp per pair (i, j) of a component: s_i' M s_j
__global__ void __launch_bounds__(256) cross_kernel(int npairs, const int *__restrict__ pi, const int *__restrict__ pj,
                                                    const int64_t *__restrict__ d1, const int64_t *__restrict__ d2, int r,
                                                    const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
                                                    const double *__restrict__ M, double *__restrict__ Bc) {
    const int lane = threadIdx.x & 31;
    for (int q = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); q < npairs; q += (int)(((int64_t)gridDim.x * blockDim.x) >> 5)) {
        const int i = pi[q], j = pj[q];
        const int bi = rowptr[i], ki = rowptr[i + 1] - bi, bj = rowptr[j], kj = rowptr[j + 1] - bj;
        double tq = 0.0;
        for (int a = 0; a < ki; a++) {
            const double *Mc = M + (size_t)r * col[bi + a];
            double part = 0.0;
            for (int b = lane; b < kj; b += 32) part += __ldg(Mc + col[bj + b]) * val[bj + b];
            tq += val[bi + a] * part;
        }
        tq = warp_sum(tq);
        if (lane == 0) { Bc[d1[q]] = tq; Bc[d2[q]] = tq; }
    }
}

// the sums behind J(sigma2fs) and the GLS step; one CTA per component (grid-stride), per-CTA partials of NS values
template <int P>
__global__ void __launch_bounds__(CP_T) msums_kernel(int ncomp, const int *__restrict__ cptr, const int *__restrict__ cmem,
                                                     const int64_t *__restrict__ voff, const double *__restrict__ Vc,
                                                     const double *__restrict__ Bc, int64_t m, const double *__restrict__ X,
                                                     const double *__restrict__ z, const double *__restrict__ t,
                                                     const double *__restrict__ ve, double s2, double *__restrict__ part) {
    constexpr int NS = 2 * P * P + 2 * P + 2;
    extern __shared__ __align__(16) double csm[];
    double *Dm = csm, *Di = csm + NCMAX * CLD, *Wm = csm + 2 * NCMAX * CLD;      // L then scratch; D^-1; W = D^-1 V D^-1
    __shared__ int flag;
    __shared__ int smem_mem[NCMAX];
    __shared__ double e_s[NCMAX], x_s[NCMAX][P];
    __shared__ double red[CP_T / 32][NS];
    double s[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) s[i] = 0.0;
    double bad = 0.0;
    for (int c = blockIdx.x; c < ncomp; c += gridDim.x) {
        const int c0 = cptr[c], n = cptr[c + 1] - c0;
        __syncthreads();
        if (n == 1) {
            if (threadIdx.x == 0) {
                const int j = cmem[c0];
                const double v = Vc[voff[c]], d = s2 * v + ve[j], u = 1.0 / d, w = v / (d * d);
                const double e = z[j] - t[j];
                double x[P];
#pragma unroll
                for (int a = 0; a < P; a++) x[a] = X[j + (int64_t)a * m];
#pragma unroll
                for (int a = 0; a < P; a++) {
#pragma unroll
                    for (int b = 0; b < P; b++) {
                        s[a + P * b] += u * x[a] * x[b];
                        s[P * P + 2 * P + 2 + a + P * b] += w * x[a] * x[b];
                    }
                    s[P * P + a] += u * x[a] * e;
                    s[P * P + P + 2 + a] -= w * x[a] * e;
                }
                s[P * P + P] += v * u;
                s[P * P + P + 1] += w * (Bc[voff[c]] + e * e);
            }
            continue;
        }
        for (int a = threadIdx.x; a < n; a += blockDim.x) {
            const int j = cmem[c0 + a];
            smem_mem[a] = j;
            e_s[a] = z[j] - t[j];
#pragma unroll
            for (int q = 0; q < P; q++) x_s[a][q] = X[j + (int64_t)q * m];
        }
        __syncthreads();
        const double *V = Vc + voff[c], *B = Bc + voff[c];
        if (!load_and_factor(Dm, n, V, smem_mem, ve, s2, &flag)) { bad = 1.0; continue; }
        // Linv (lower) into Wm: thread j solves L x = e_j
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            for (int i = 0; i < n; i++) {
                double acc = (i == j) ? 1.0 : 0.0;
                for (int k = j; k < i; k++) acc -= Dm[i * CLD + k] * Wm[k * CLD + j];
                Wm[i * CLD + j] = i < j ? 0.0 : acc / Dm[i * CLD + i];
            }
        }
        __syncthreads();
        // D^-1 = Linv' Linv
        for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
            const int i = idx % n, j = idx / n;
            double acc = 0.0;
            for (int k = max(i, j); k < n; k++) acc += Wm[k * CLD + i] * Wm[k * CLD + j];
            Di[i * CLD + j] = acc;
        }
        __syncthreads();
        // T = D^-1 V into Dm (the factor is no longer needed), then W = T D^-1
        for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
            const int i = idx % n, j = idx / n;
            double acc = 0.0;
            for (int k = 0; k < n; k++) acc += Di[i * CLD + k] * V[k + (size_t)j * n];
            Dm[i * CLD + j] = acc;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
            const int i = idx % n, j = idx / n;
            double acc = 0.0;
            for (int k = 0; k < n; k++) acc += Dm[i * CLD + k] * Di[k * CLD + j];
            Wm[i * CLD + j] = acc;
        }
        __syncthreads();
        // the sums: every thread takes entries (i, j) of the n x n blocks
        for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
            const int i = idx % n, j = idx / n;
            const double u = Di[i * CLD + j], w = Wm[i * CLD + j];
#pragma unroll
            for (int a = 0; a < P; a++) {
#pragma unroll
                for (int b = 0; b < P; b++) {
                    s[a + P * b] += u * x_s[i][a] * x_s[j][b];
                    s[P * P + 2 * P + 2 + a + P * b] += w * x_s[i][a] * x_s[j][b];
                }
                s[P * P + a] += u * x_s[i][a] * e_s[j];
                s[P * P + P + 2 + a] -= w * x_s[i][a] * e_s[j];
            }
            if (i == j) s[P * P + P] += Dm[i * CLD + i];                                   // tr(D^-1 V)
            s[P * P + P + 1] += w * (B[j + (size_t)i * n] + e_s[i] * e_s[j]);              // tr(W B) + e'W e
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NS; i++) {
        const double v = warp_sum(s[i]);
        if (lane == 0) red[wid][i] = v;
    }
    bad = warp_sum(bad);
    __shared__ double bad_s[CP_T / 32];
    if (lane == 0) bad_s[wid] = bad;
    __syncthreads();
    if (threadIdx.x < NS) {
        double v = 0.0;
        for (int w8 = 0; w8 < CP_T / 32; w8++) v += red[w8][threadIdx.x];
        part[(size_t)blockIdx.x * (NS + 1) + threadIdx.x] = v;
    }
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int w8 = 0; w8 < CP_T / 32; w8++) v += bad_s[w8];
        part[(size_t)blockIdx.x * (NS + 1) + NS] = v;
    }
}

__global__ void msums_finish_kernel(int grid, int ns1, const double *__restrict__ part, double *__restrict__ out) {
    const int k = threadIdx.x;
    if (k < ns1) {
        double v = 0.0;
        for (int b = 0; b < grid; b++) v += part[(size_t)b * ns1 + k];
        out[k] = v;
    }
}


// D_c^-1 of every component at sigma2fs into out (layout of Vc); part[blockIdx.x] = failure flag
__global__ void __launch_bounds__(CP_T) dinv_kernel(int ncomp, const int *__restrict__ cptr, const int *__restrict__ cmem,
                                                    const int64_t *__restrict__ voff, const double *__restrict__ Vc,
                                                    const double *__restrict__ ve, double s2, double *__restrict__ out,
                                                    double *__restrict__ part) {
    extern __shared__ __align__(16) double csm[];
    double *Dm = csm, *Wm = csm + NCMAX * CLD;
    __shared__ int flag;
    __shared__ int smem_mem[NCMAX];
    double bad = 0.0;
    for (int c = blockIdx.x; c < ncomp; c += gridDim.x) {
        const int c0 = cptr[c], n = cptr[c + 1] - c0;
        __syncthreads();
        if (n == 1) {
            const double d = s2 * Vc[voff[c]] + ve[cmem[c0]];
            if (!(d > 0.0)) bad = 1.0;
            if (threadIdx.x == 0) out[voff[c]] = 1.0 / d;
            continue;
        }
        for (int a = threadIdx.x; a < n; a += blockDim.x) smem_mem[a] = cmem[c0 + a];
        __syncthreads();
        if (!load_and_factor(Dm, n, Vc + voff[c], smem_mem, ve, s2, &flag)) { bad = 1.0; continue; }
        for (int j = threadIdx.x; j < n; j += blockDim.x)
            for (int i = 0; i < n; i++) {
                double acc = (i == j) ? 1.0 : 0.0;
                for (int k = j; k < i; k++) acc -= Dm[i * CLD + k] * Wm[k * CLD + j];
                Wm[i * CLD + j] = i < j ? 0.0 : acc / Dm[i * CLD + i];
            }
        __syncthreads();
        double *o = out + voff[c];
        for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
            const int i = idx % n, j = idx / n;
            double acc = 0.0;
            for (int k = max(i, j); k < n; k++) acc += Wm[k * CLD + i] * Wm[k * CLD + j];
            o[idx] = acc;
        }
    }
    if (threadIdx.x == 0) part[blockIdx.x] = bad;
}

// one warp per (touched BAU, observation of its component): s_i' M s0_k with the two rows in different CSR matrices
__global__ void __launch_bounds__(256) cross_bau_kernel(int64_t npairs, const int *__restrict__ pk, const int *__restrict__ pi, int r,
                                                        const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
                                                        const int *__restrict__ rowptr0, const int *__restrict__ col0,
                                                        const double *__restrict__ val0, const double *__restrict__ M, double *__restrict__ H) {
    const int lane = threadIdx.x & 31;
    for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < npairs; q += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int i = pi[q], k = pk[q];
        const int bi = rowptr[i], ki = rowptr[i + 1] - bi, bk = rowptr0[k], kk = rowptr0[k + 1] - bk;
        double tq = 0.0;
        for (int a = 0; a < ki; a++) {
            const double *Mc = M + (size_t)r * col[bi + a];
            double part = 0.0;
            for (int b = lane; b < kk; b += 32) part += __ldg(Mc + col0[bk + b]) * val0[bk + b];
            tq += val[bi + a] * part;
        }
        tq = warp_sum(tq);
        if (lane == 0) H[q] = tq;
    }
}

// Case 2 of .SRE.predict_EM (R/SREpredict.R:308-333) for the BAUs inside observation footprints, one thread per touched BAU t:
//   u = D_c^-1 c_k,  g = sigma2fs fs_k u  (c_k: the BAU's column of Cmat restricted to its component)
//   mu_k  += g'(Z_c - X_c alpha - S_c mu~)
//   var_k  = s_k'Sigma~ s_k - 2 g'h_k + g'B_c g + sigma2fs fs_k - (sigma2fs fs_k)^2 c_k'u,   h_k[a] = s_{i_a}'Sigma~ s_k
__global__ void __launch_bounds__(128) case2_kernel(int nt, const int *__restrict__ tb_bau, const int *__restrict__ tb_comp,
                                                    const int *__restrict__ tb_ptr, const int *__restrict__ tb_a, const double *__restrict__ tb_w,
                                                    const int64_t *__restrict__ tb_hoff, const int *__restrict__ cptr,
                                                    const int *__restrict__ cmem, const int64_t *__restrict__ voff,
                                                    const double *__restrict__ Dinv, const double *__restrict__ Bc, const double *__restrict__ H,
                                                    const double *__restrict__ res, const double *__restrict__ smuS, const double *__restrict__ q,
                                                    const double *__restrict__ fs, double s2, double *__restrict__ mu, double *__restrict__ var, double *__restrict__ sd) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const int k = tb_bau[t], c = tb_comp[t], c0 = cptr[c], n = cptr[c + 1] - c0;
    const double *Di = Dinv + voff[c], *B = Bc + voff[c], *h = H + tb_hoff[t];
    const double sf = s2 * fs[k];
    double u[NCMAX];
    for (int a = 0; a < n; a++) {
        double acc = 0.0;
        for (int x = tb_ptr[t]; x < tb_ptr[t + 1]; x++) acc += Di[a + (size_t)tb_a[x] * n] * tb_w[x];
        u[a] = acc;
    }
    double cu = 0.0;
    for (int x = tb_ptr[t]; x < tb_ptr[t + 1]; x++) cu += tb_w[x] * u[tb_a[x]];
    double ge = 0.0, gh = 0.0, gBg = 0.0;
    for (int a = 0; a < n; a++) {
        const double ga = sf * u[a];
        ge += ga * (res[cmem[c0 + a]] - smuS[cmem[c0 + a]]);                 // e = Z - X alpha - S mu~
        gh += ga * h[a];
        double row = 0.0;
        for (int b = 0; b < n; b++) row += B[a + (size_t)b * n] * u[b];
        gBg += ga * sf * row;
    }
    const double v = q[k] - 2.0 * gh + gBg + sf - sf * sf * cu;
    mu[k] += ge;
    var[k] = v;
    sd[k] = sqrt(v);
}

}  // namespace

int frk_coupling_destroy(frk_coupling *C) {
    if (!C) return 0;
    cudaSetDevice(C->ctx->device);
    cudaStreamSynchronize(C->ctx->stream);
    if (C->bk_w) frk_rowbuckets_destroy(C->bk_w);
    for (void *p : C->allocs) cudaFree(p);
    delete C;
    return 0;
}

int frk_coupling_create(frk_ctx *ctx, int64_t m, int r, int p, const int *h_vrowptr, const int *h_vcol, const double *h_vval,
                        const int *d_rowptr, const int *d_col, frk_coupling **out) {
    FRK_REQUIRE(ctx && h_vrowptr && h_vcol && h_vval && d_rowptr && d_col && out, "frk_coupling_create: NULL argument");
    FRK_REQUIRE(m > 0 && m < 2147483647LL && p >= 1 && p <= MAXP_C, "frk_coupling_create: bad sizes");
    FRK_CHECK_CUDA(cudaSetDevice(ctx->device));
    const int mi = (int)m;
    // ---- connected components of the pattern of Vfs (union-find) ----
    std::vector<int> parent(mi);
    std::iota(parent.begin(), parent.end(), 0);
    auto find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    for (int i = 0; i < mi; i++)
        for (int e = h_vrowptr[i]; e < h_vrowptr[i + 1]; e++) {
            const int j = h_vcol[e];
            FRK_REQUIRE(j >= 0 && j < mi, "frk_coupling_create: column index out of range in Vfs");
            if (j != i && h_vval[e] != 0.0) { const int a = find(i), b = find(j); if (a != b) parent[std::max(a, b)] = std::min(a, b); }
        }
    std::vector<int> comp_of(mi, -1), cptr{0}, cmem(mi);
    int ncomp = 0;
    for (int i = 0; i < mi; i++) { const int root = find(i); if (comp_of[root] < 0) comp_of[root] = ncomp++; comp_of[i] = comp_of[root]; }
    std::vector<int> csize(ncomp, 0);
    for (int i = 0; i < mi; i++) csize[comp_of[i]]++;
    cptr.resize(ncomp + 1);
    for (int c = 0; c < ncomp; c++) cptr[c + 1] = cptr[c] + csize[c];
    {
        std::vector<int> fill(cptr.begin(), cptr.end() - 1);
        for (int i = 0; i < mi; i++) cmem[fill[comp_of[i]]++] = i;             // ascending inside every component
    }
    const int nmax = *std::max_element(csize.begin(), csize.end());
    FRK_REQUIRE(nmax <= NCMAX, "%d observations are coupled through shared BAUs in one connected group; the device path handles groups of up "
                "to %d (no CPU fallback)", nmax, NCMAX);
    std::vector<int> local(mi);
    for (int c = 0; c < ncomp; c++) for (int a = cptr[c]; a < cptr[c + 1]; a++) local[cmem[a]] = a - cptr[c];
    std::vector<int64_t> voff(ncomp + 1, 0);
    for (int c = 0; c < ncomp; c++) voff[c + 1] = voff[c] + (int64_t)csize[c] * csize[c];
    std::vector<double> Vc((size_t)voff[ncomp], 0.0);
    for (int i = 0; i < mi; i++)
        for (int e = h_vrowptr[i]; e < h_vrowptr[i + 1]; e++) {
            const int j = h_vcol[e], c = comp_of[i];
            if (comp_of[j] != c) continue;                                     // (explicit zeros across components)
            Vc[(size_t)voff[c] + local[i] + (size_t)local[j] * csize[c]] = h_vval[e];
        }
    std::vector<int> pi, pj;
    std::vector<int64_t> pd1, pd2;
    for (int c = 0; c < ncomp; c++) {
        const int n = csize[c];
        for (int b = 0; b < n; b++)
            for (int a = 0; a <= b; a++) {
                pi.push_back(cmem[cptr[c] + a]); pj.push_back(cmem[cptr[c] + b]);
                pd1.push_back(voff[c] + a + (int64_t)b * n); pd2.push_back(voff[c] + b + (int64_t)a * n);
            }
    }
    FRK_REQUIRE(pi.size() < 2147483647ULL, "frk_coupling_create: too many coupled pairs");
    // ---- the whitened pattern: every row of a component carries the union of the component's columns ----
    std::vector<int> rp(mi + 1);
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    FRK_CHECK_CUDA(cudaMemcpy(rp.data(), d_rowptr, (size_t)(mi + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    const int64_t nnz = rp[mi];
    std::vector<int> cl((size_t)std::max<int64_t>(nnz, 1));
    if (nnz) FRK_CHECK_CUDA(cudaMemcpy(cl.data(), d_col, (size_t)nnz * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int> rpw(mi + 1, 0);
    std::vector<std::vector<int>> ucols(ncomp);
    for (int c = 0; c < ncomp; c++) {
        if (csize[c] == 1) continue;
        std::vector<int> &u = ucols[c];
        for (int a = cptr[c]; a < cptr[c + 1]; a++) { const int i = cmem[a]; u.insert(u.end(), cl.begin() + rp[i], cl.begin() + rp[i + 1]); }
        std::sort(u.begin(), u.end());
        u.erase(std::unique(u.begin(), u.end()), u.end());
    }
    int64_t nnz_w = 0;
    for (int i = 0; i < mi; i++) {
        const int c = comp_of[i];
        const int k = csize[c] == 1 ? rp[i + 1] - rp[i] : (int)ucols[c].size();
        rpw[i] = (int)nnz_w;
        nnz_w += k;
        FRK_REQUIRE(nnz_w < 2147483647LL, "frk_coupling_create: the whitened basis matrix exceeds 2^31 entries");
    }
    rpw[mi] = (int)nnz_w;
    std::vector<int> clw((size_t)std::max<int64_t>(nnz_w, 1)), dst((size_t)std::max<int64_t>(nnz, 1));
    for (int i = 0; i < mi; i++) {
        const int c = comp_of[i];
        if (csize[c] == 1) {
            for (int e = rp[i]; e < rp[i + 1]; e++) { clw[rpw[i] + e - rp[i]] = cl[e]; dst[e] = rpw[i] + e - rp[i]; }
        } else {
            const std::vector<int> &u = ucols[c];
            std::copy(u.begin(), u.end(), clw.begin() + rpw[i]);
            for (int e = rp[i]; e < rp[i + 1]; e++)
                dst[e] = rpw[i] + (int)(std::lower_bound(u.begin(), u.end(), cl[e]) - u.begin());
        }
    }
    frk_coupling *C = new frk_coupling();
    C->ctx = ctx; C->m = m; C->r = r; C->p = p; C->ncomp = ncomp; C->npairs = (int)pi.size(); C->nnz = nnz; C->nnz_w = nnz_w;
    auto fail = [&](int rc) { frk_coupling_destroy(C); return rc; };
    std::vector<double> one((size_t)mi, 1.0);
    if (cp_upload(C, &C->cptr, cptr) || cp_upload(C, &C->cmem, cmem) || cp_upload(C, &C->voff, voff) || cp_upload(C, &C->Vc, Vc) ||
        cp_alloc(C, &C->Bc, Vc.size()) || cp_upload(C, &C->pair_i, pi) || cp_upload(C, &C->pair_j, pj) || cp_upload(C, &C->pair_dst, pd1) ||
        cp_upload(C, &C->pair_dst2, pd2) || cp_upload(C, &C->rowptr_w, rpw) || cp_upload(C, &C->col_w, clw) || cp_upload(C, &C->dst, dst) ||
        cp_alloc(C, &C->val_w, (size_t)std::max<int64_t>(nnz_w, 1)) || cp_alloc(C, &C->Zw, (size_t)mi) || cp_alloc(C, &C->Xw, (size_t)mi * p) ||
        cp_upload(C, &C->ones, one) || cp_alloc(C, &C->zeros, (size_t)mi))
        return fail(1);
    FRK_CHECK_CUDA(cudaMemset(C->zeros, 0, (size_t)mi * sizeof(double)));
    FRK_CHECK_CUDA(cudaMemset(C->val_w, 0, (size_t)std::max<int64_t>(nnz_w, 1) * sizeof(double)));
    FRK_CHECK_CUDA(cudaMemset(C->Bc, 0, std::max<size_t>(Vc.size(), 1) * sizeof(double)));
    C->grid = (int)std::min<int64_t>(ncomp, (int64_t)ctx->sm_count * 8);
    if (cp_alloc(C, &C->part, (size_t)C->grid * 64)) return fail(1);
    if (int rc = frk_rowbuckets_create(ctx, m, r, C->rowptr_w, C->col_w, C->val_w, &C->bk_w)) return fail(rc);
    *out = C;
    return 0;
}

int frk_coupling_ncomp(const frk_coupling *C) { return C ? C->ncomp : -1; }

// whiten at sigma2fs: afterwards the Gram / rhs of (rowptr_w, col_w, val_w, bk_w, Zw, Xw, Ve = 1, Vfs = 0) are S'D^-1 S etc.
// *d_logdet_slot += logdet(cholD) (the reference's helper: 2 sum log diag chol D)
int frk_coupling_whiten(frk_coupling *C, frk_ctx *ctx, double sigma2fs, const double *d_val, const double *d_Z, const double *d_X,
                        const double *d_Ve, double *d_logdet_slot) {
    const int64_t m = C->m;
    FRK_CHECK_CUDA(cudaMemsetAsync(C->val_w, 0, (size_t)std::max<int64_t>(C->nnz_w, 1) * sizeof(double), ctx->stream));
    if (C->nnz) FRK_LAUNCH(ctx, scatter_kernel, (unsigned)std::min<int64_t>(cdiv(C->nnz, 256), (int64_t)ctx->sm_count * 16), 256, 0, C->nnz,
                           C->dst, d_val, C->val_w);
    FRK_CHECK_CUDA(cudaMemcpyAsync(C->Zw, d_Z, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    FRK_CHECK_CUDA(cudaMemcpyAsync(C->Xw, d_X, (size_t)m * C->p * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    static bool attr = false;
    const size_t smem = (size_t)NCMAX * CLD * sizeof(double);
    if (!attr) {
        FRK_CHECK_CUDA(cudaFuncSetAttribute(whiten_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    FRK_LAUNCH(ctx, whiten_kernel, C->grid, CP_T, smem, C->ncomp, C->cptr, C->cmem, C->voff, C->Vc, d_Ve, sigma2fs, m, C->p, C->rowptr_w,
               C->val_w, C->Zw, C->Xw, C->part);
    double *status = ctx->d_small + 4096 - 16;
    FRK_LAUNCH(ctx, whiten_finish_kernel, 1, 32, 0, C->grid, C->part, d_logdet_slot, status);
    double h_bad = 0.0;
    if (int rc = frk_d2h(ctx, &h_bad, status, sizeof(double))) return rc;
    FRK_REQUIRE(h_bad == 0.0, "sigma2fs * Vfs + Ve is not positive definite (R: chol(D) fails, R/SREfit.R:188)");
    return frk_rowbuckets_refresh(C->bk_w, ctx, C->rowptr_w, C->val_w);
}

const int *frk_coupling_rowptr(const frk_coupling *C) { return C->rowptr_w; }
const int *frk_coupling_col(const frk_coupling *C) { return C->col_w; }
const double *frk_coupling_val(const frk_coupling *C) { return C->val_w; }
const double *frk_coupling_Z(const frk_coupling *C) { return C->Zw; }
const double *frk_coupling_X(const frk_coupling *C) { return C->Xw; }
const double *frk_coupling_ones(const frk_coupling *C) { return C->ones; }
const double *frk_coupling_zeros(const frk_coupling *C) { return C->zeros; }
const frk_rowbuckets *frk_coupling_buckets(const frk_coupling *C) { return C->bk_w; }

// B_c = S_c M S_c' for every component (M = S_eta in the M-step)
int frk_coupling_cross(frk_coupling *C, frk_ctx *ctx, const int *d_rowptr, const int *d_col, const double *d_val, const double *d_M) {
    if (C->npairs == 0) return 0;
    const unsigned grid = (unsigned)std::min<int64_t>(cdiv((int64_t)C->npairs * 32, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, cross_kernel, grid, 256, 0, C->npairs, C->pair_i, C->pair_j, C->pair_dst, C->pair_dst2, C->r, d_rowptr, d_col, d_val, d_M,
               C->Bc);
    return 0;
}

// h_out: [uxx p^2 | uxz p | sv | w0 | wx p | wxx p^2], the layout of frk_mstep_sums
int frk_coupling_msums(frk_coupling *C, frk_ctx *ctx, const double *d_X, const double *d_Z, const double *d_t, const double *d_Ve,
                       double sigma2fs, double *h_out) {
    const int p = C->p, ns = 2 * p * p + 2 * p + 2;
    const size_t smem = (size_t)3 * NCMAX * CLD * sizeof(double);
    static bool attr = false;
    if (!attr) {
        FRK_CHECK_CUDA(cudaFuncSetAttribute(msums_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FRK_CHECK_CUDA(cudaFuncSetAttribute(msums_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FRK_CHECK_CUDA(cudaFuncSetAttribute(msums_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FRK_CHECK_CUDA(cudaFuncSetAttribute(msums_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    switch (p) {
        case 1: FRK_LAUNCH(ctx, msums_kernel<1>, C->grid, CP_T, smem, C->ncomp, C->cptr, C->cmem, C->voff, C->Vc, C->Bc, C->m, d_X, d_Z, d_t, d_Ve, sigma2fs, C->part); break;
        case 2: FRK_LAUNCH(ctx, msums_kernel<2>, C->grid, CP_T, smem, C->ncomp, C->cptr, C->cmem, C->voff, C->Vc, C->Bc, C->m, d_X, d_Z, d_t, d_Ve, sigma2fs, C->part); break;
        case 3: FRK_LAUNCH(ctx, msums_kernel<3>, C->grid, CP_T, smem, C->ncomp, C->cptr, C->cmem, C->voff, C->Vc, C->Bc, C->m, d_X, d_Z, d_t, d_Ve, sigma2fs, C->part); break;
        default: FRK_LAUNCH(ctx, msums_kernel<4>, C->grid, CP_T, smem, C->ncomp, C->cptr, C->cmem, C->voff, C->Vc, C->Bc, C->m, d_X, d_Z, d_t, d_Ve, sigma2fs, C->part); break;
    }
    double *d_out = ctx->d_small + 4096 - 128;
    FRK_LAUNCH(ctx, msums_finish_kernel, 1, 64, 0, C->grid, ns + 1, C->part, d_out);
    double h[64];
    if (int rc = frk_d2h(ctx, h, d_out, (size_t)(ns + 1) * sizeof(double))) return rc;
    FRK_REQUIRE(h[ns] == 0.0, "sigma2fs * Vfs + Ve is not positive definite (R: chol(D) fails, R/SREfit.R:304)");
    for (int k = 0; k < ns; k++) h_out[k] = h[k];
    return 0;
}

// Case-2 prediction (obs_fs = FALSE, sigma2fs > 0) at the BAUs that lie inside observation footprints; every other BAU keeps the
// "unobserved" values the caller has already written.  cz_*: Cmat as device CSR over the observations (rows normalised);
// d_Sig = Sigma~ (posterior covariance at the final parameters), d_res = Z - X alpha, d_smuS = S mu~, d_q[k] = s_k'Sigma~ s_k.
int frk_coupling_case2(frk_coupling *C, frk_ctx *ctx, double sigma2fs, const double *d_Ve, int64_t nbau, const int *d_cz_rowptr,
                       const int *d_cz_members, const double *d_cz_weights, const int *d_rowptr, const int *d_col, const double *d_val,
                       const int *d_rowptr0, const int *d_col0, const double *d_val0, const double *d_Sig, const double *d_res,
                       const double *d_smuS, const double *d_q, const double *d_fs, double *d_mu, double *d_var, double *d_sd) {
    const int m = (int)C->m;
    // ---- host: the touched BAUs, their components and their columns of Cmat ----
    std::vector<int> crp(m + 1), cptr(C->ncomp + 1), cmem(m);
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    FRK_CHECK_CUDA(cudaMemcpy(crp.data(), d_cz_rowptr, (size_t)(m + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    FRK_CHECK_CUDA(cudaMemcpy(cptr.data(), C->cptr, (size_t)(C->ncomp + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    FRK_CHECK_CUDA(cudaMemcpy(cmem.data(), C->cmem, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost));
    const int nz = crp[m];
    std::vector<int> cmemb((size_t)std::max(nz, 1));
    std::vector<double> cw((size_t)std::max(nz, 1));
    if (nz) {
        FRK_CHECK_CUDA(cudaMemcpy(cmemb.data(), d_cz_members, (size_t)nz * sizeof(int), cudaMemcpyDeviceToHost));
        FRK_CHECK_CUDA(cudaMemcpy(cw.data(), d_cz_weights, (size_t)nz * sizeof(double), cudaMemcpyDeviceToHost));
    }
    std::vector<int> comp_of(m), local(m);
    for (int c = 0; c < C->ncomp; c++) for (int a = cptr[c]; a < cptr[c + 1]; a++) { comp_of[cmem[a]] = c; local[cmem[a]] = a - cptr[c]; }
    // entries (bau, obs, weight) sorted by bau (stable: observations ascending within a BAU)
    std::vector<int> order((size_t)nz), eobs((size_t)nz);
    for (int i = 0; i < m; i++) for (int x = crp[i]; x < crp[i + 1]; x++) eobs[x] = i;
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cmemb[a] < cmemb[b]; });
    std::vector<int> tb_bau, tb_comp, tb_ptr{0}, tb_a, pk, pi;
    std::vector<double> tb_w;
    std::vector<int64_t> tb_hoff{0};
    for (size_t x = 0; x < order.size();) {
        const int k = cmemb[order[x]];
        FRK_REQUIRE(k >= 0 && k < nbau, "frk_coupling_case2: BAU index out of range in the incidence matrix");
        const int c = comp_of[eobs[order[x]]];
        for (; x < order.size() && cmemb[order[x]] == k; x++) {
            FRK_REQUIRE(comp_of[eobs[order[x]]] == c, "frk_coupling_case2: the incidence matrix does not match the model's Vfs");
            tb_a.push_back(local[eobs[order[x]]]);
            tb_w.push_back(cw[order[x]]);
        }
        tb_bau.push_back(k); tb_comp.push_back(c); tb_ptr.push_back((int)tb_a.size());
        for (int a = cptr[c]; a < cptr[c + 1]; a++) { pk.push_back(k); pi.push_back(cmem[a]); }
        tb_hoff.push_back((int64_t)pk.size());
    }
    const int nt = (int)tb_bau.size();
    if (nt == 0) return 0;
    // ---- device ----
    const size_t npairs = pk.size(), nv = 0;
    (void)nv;
    int64_t vtot = 0;
    FRK_CHECK_CUDA(cudaMemcpy(&vtot, C->voff + C->ncomp, sizeof(int64_t), cudaMemcpyDeviceToHost));
    char *buf = nullptr;
    auto al8 = [](size_t b) { return (b + 15) & ~(size_t)15; };
    const size_t b_bau = al8(nt * sizeof(int)), b_ptr = al8((nt + 1) * sizeof(int)), b_a = al8(tb_a.size() * sizeof(int)),
                 b_w = al8(tb_w.size() * sizeof(double)), b_hoff = al8((nt + 1) * sizeof(int64_t)), b_pk = al8(npairs * sizeof(int)),
                 b_H = al8(npairs * sizeof(double)), b_Di = al8((size_t)vtot * sizeof(double));
    FRK_CHECK_CUDA(cudaMalloc((void **)&buf, 2 * b_bau + b_ptr + b_a + b_w + b_hoff + 2 * b_pk + b_H + b_Di));
    char *o = buf;
    int *d_bau = (int *)o; o += b_bau; int *d_comp = (int *)o; o += b_bau; int *d_ptr = (int *)o; o += b_ptr; int *d_a = (int *)o; o += b_a;
    double *d_w = (double *)o; o += b_w; int64_t *d_hoff = (int64_t *)o; o += b_hoff; int *d_pk = (int *)o; o += b_pk; int *d_pi = (int *)o; o += b_pk;
    double *d_H = (double *)o; o += b_H; double *d_Di = (double *)o;
    auto run = [&]() -> int {
        FRK_CHECK_CUDA(cudaMemcpy(d_bau, tb_bau.data(), nt * sizeof(int), cudaMemcpyHostToDevice));
        FRK_CHECK_CUDA(cudaMemcpy(d_comp, tb_comp.data(), nt * sizeof(int), cudaMemcpyHostToDevice));
        FRK_CHECK_CUDA(cudaMemcpy(d_ptr, tb_ptr.data(), (nt + 1) * sizeof(int), cudaMemcpyHostToDevice));
        FRK_CHECK_CUDA(cudaMemcpy(d_a, tb_a.data(), tb_a.size() * sizeof(int), cudaMemcpyHostToDevice));
        FRK_CHECK_CUDA(cudaMemcpy(d_w, tb_w.data(), tb_w.size() * sizeof(double), cudaMemcpyHostToDevice));
        FRK_CHECK_CUDA(cudaMemcpy(d_hoff, tb_hoff.data(), (nt + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
        FRK_CHECK_CUDA(cudaMemcpy(d_pk, pk.data(), npairs * sizeof(int), cudaMemcpyHostToDevice));
        FRK_CHECK_CUDA(cudaMemcpy(d_pi, pi.data(), npairs * sizeof(int), cudaMemcpyHostToDevice));
        const size_t smem = (size_t)2 * NCMAX * CLD * sizeof(double);
        static bool attr = false;
        if (!attr) {
            FRK_CHECK_CUDA(cudaFuncSetAttribute(dinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr = true;
        }
        FRK_LAUNCH(ctx, dinv_kernel, C->grid, CP_T, smem, C->ncomp, C->cptr, C->cmem, C->voff, C->Vc, d_Ve, sigma2fs, d_Di, C->part);
        if (int rc = frk_coupling_cross(C, ctx, d_rowptr, d_col, d_val, d_Sig)) return rc;                 // B_c = S_c Sigma~ S_c'
        const unsigned grid = (unsigned)std::min<int64_t>(cdiv((int64_t)npairs * 32, 256), (int64_t)ctx->sm_count * 16);
        FRK_LAUNCH(ctx, cross_bau_kernel, grid, 256, 0, (int64_t)npairs, d_pk, d_pi, C->r, d_rowptr, d_col, d_val, d_rowptr0, d_col0, d_val0,
                   d_Sig, d_H);
        FRK_LAUNCH(ctx, case2_kernel, (unsigned)cdiv(nt, 128), 128, 0, nt, d_bau, d_comp, d_ptr, d_a, d_w, d_hoff, C->cptr, C->cmem, C->voff,
                   d_Di, C->Bc, d_H, d_res, d_smuS, d_q, d_fs, sigma2fs, d_mu, d_var, d_sd);
        FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        return 0;
    };
    const int rc = run();
    cudaFree(buf);
    return rc;
}

// Row-streaming kernels over the CSR basis matrix: the weighted Gram / right-hand side of the
// E-step and log-likelihood, the gather-based sparse quadratic form + SpMV used by the M-step and
// by predict, the M-step row sums, and the per-BAU prediction epilogue.
//
// Reference: R/SREfit.R:174-220 (.loglik.ind), :232-260 (.SRE.Estep.ind), :263-438
// (.SRE.Mstep.ind), R/SREutils.R:245-305 (.batch_compute_var), R/SREpredict.R:243-375.
#include "common.cuh"

constexpr int MAXP = 4;             // fixed effects handled on the device path
struct Alpha { double a[MAXP]; };

// ---------------------------------------------------------------------------------
// Gram + rhs: one warp per row, lanes over the (a<=b) pairs of the row's non-zeros.
// acc layout: [r*r upper-tri G | r rhs | sum rho^2/d | sum log d]
// ---------------------------------------------------------------------------------
constexpr int GRAM_WARPS = 8;
constexpr int ROW_STAGE = 64;       // row entries staged in shared memory per warp

__global__ void __launch_bounds__(GRAM_WARPS * 32) gram_rhs_kernel(
    int64_t m, int r, const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
    const double *__restrict__ z, const double *__restrict__ X, int p, Alpha alpha, const double *__restrict__ ve,
    const double *__restrict__ vfs, double sigma2fs, double *__restrict__ acc) {
    __shared__ int scol[GRAM_WARPS][ROW_STAGE];
    __shared__ double sval[GRAM_WARPS][ROW_STAGE];
    __shared__ double sred[GRAM_WARPS][2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double *G = acc, *rhs = acc + (size_t)r * r;
    double s_quad = 0.0, s_logd = 0.0;
    const int64_t nwarps = (int64_t)gridDim.x * GRAM_WARPS;
    for (int64_t row = (int64_t)blockIdx.x * GRAM_WARPS + wid; row < m; row += nwarps) {
        const int beg = rowptr[row], k = rowptr[row + 1] - beg;
        double rho = z[row];
        for (int q = 0; q < p; q++) rho -= X[row + (int64_t)q * m] * alpha.a[q];
        const double d = sigma2fs * vfs[row] + ve[row];
        const double w = 1.0 / d;
        if (lane == 0) { s_quad += rho * rho * w; s_logd += log(d); }
        const bool staged = k <= ROW_STAGE;
        if (staged) {
            for (int t = lane; t < k; t += 32) { scol[wid][t] = col[beg + t]; sval[wid][t] = val[beg + t]; }
            __syncwarp();
        }
        const double wr = w * rho;
        for (int t = lane; t < k; t += 32) {
            int c = staged ? scol[wid][t] : col[beg + t];
            double v = staged ? sval[wid][t] : val[beg + t];
            atomicAdd(rhs + c, wr * v);
        }
        for (int a = 0; a < k; a++) {
            const int ca = staged ? scol[wid][a] : col[beg + a];
            const double wa = w * (staged ? sval[wid][a] : val[beg + a]);
            double *Gc = G + ca;   // element (ca, cb), ca <= cb, column-major: ca + r*cb
            for (int b = a + lane; b < k; b += 32) {
                const int cb = staged ? scol[wid][b] : col[beg + b];
                const double vb = staged ? sval[wid][b] : val[beg + b];
                atomicAdd(Gc + (size_t)r * cb, wa * vb);
            }
        }
        __syncwarp();
    }
    if (lane == 0) { sred[wid][0] = s_quad; sred[wid][1] = s_logd; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int i = 0; i < GRAM_WARPS; i++) { a += sred[i][0]; b += sred[i][1]; }
        atomicAdd(rhs + r, a);
        atomicAdd(rhs + r + 1, b);
    }
}

extern "C" int frk_gram_rhs(frk_ctx *ctx, int64_t m, int r, const int *d_rowptr, const int *d_col, const double *d_val,
                            const double *d_z, const double *d_X, int p, const double *h_alpha, const double *d_ve,
                            const double *d_vfs, double sigma2fs, double *d_acc) {
    FRK_REQUIRE(ctx && d_rowptr && d_acc, "frk_gram_rhs: NULL argument");
    FRK_REQUIRE(p >= 0 && p <= MAXP, "frk_gram_rhs: %d fixed effects; the device path handles at most %d", p, MAXP);
    if (m == 0) return 0;
    Alpha al{};
    for (int q = 0; q < p; q++) al.a[q] = h_alpha[q];
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(m, GRAM_WARPS), (int64_t)ctx->sm_count * 8);
    FRK_LAUNCH(ctx, gram_rhs_kernel, grid, GRAM_WARPS * 32, 0, m, r, d_rowptr, d_col, d_val, d_z, d_X, p, al, d_ve, d_vfs,
               sigma2fs, d_acc);
    return 0;
}

// ---------------------------------------------------------------------------------
// q_i = s_i' M s_i and t_i = s_i' mu: one warp per row, M gathered through L1/L2.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(GRAM_WARPS * 32) quadform_spmv_kernel(
    int64_t n, int r, const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
    const double *__restrict__ M, const double *__restrict__ mu, double *__restrict__ q, double *__restrict__ t) {
    __shared__ int scol[GRAM_WARPS][ROW_STAGE];
    __shared__ double sval[GRAM_WARPS][ROW_STAGE];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * GRAM_WARPS;
    for (int64_t row = (int64_t)blockIdx.x * GRAM_WARPS + wid; row < n; row += nwarps) {
        const int beg = rowptr[row], k = rowptr[row + 1] - beg;
        const bool staged = k <= ROW_STAGE;
        if (staged) {
            for (int x = lane; x < k; x += 32) { scol[wid][x] = col[beg + x]; sval[wid][x] = val[beg + x]; }
            __syncwarp();
        }
        double tq = 0.0, tt = 0.0;
        if (mu) {
            for (int b = lane; b < k; b += 32) {
                int cb = staged ? scol[wid][b] : col[beg + b];
                double vb = staged ? sval[wid][b] : val[beg + b];
                tt += vb * __ldg(mu + cb);
            }
        }
        if (M) {
            for (int a = 0; a < k; a++) {
                const int ca = staged ? scol[wid][a] : col[beg + a];
                const double va = staged ? sval[wid][a] : val[beg + a];
                const double *Mc = M + (size_t)r * ca;
                double part = 0.0;
                for (int b = lane; b < k; b += 32) {
                    int cb = staged ? scol[wid][b] : col[beg + b];
                    double vb = staged ? sval[wid][b] : val[beg + b];
                    part += __ldg(Mc + cb) * vb;
                }
                tq += va * part;
            }
        }
        tq = warp_sum(tq);
        tt = warp_sum(tt);
        if (lane == 0) {
            if (q && M) q[row] = tq;
            if (t && mu) t[row] = tt;
        }
        __syncwarp();
    }
}

extern "C" int frk_quadform_spmv(frk_ctx *ctx, int64_t n, int r, const int *d_rowptr, const int *d_col, const double *d_val,
                                 const double *d_M, const double *d_mu, double *d_q, double *d_t) {
    FRK_REQUIRE(ctx && d_rowptr, "frk_quadform_spmv: NULL argument");
    if (n == 0) return 0;
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, GRAM_WARPS), (int64_t)ctx->sm_count * 8);
    FRK_LAUNCH(ctx, quadform_spmv_kernel, grid, GRAM_WARPS * 32, 0, n, r, d_rowptr, d_col, d_val, d_M, d_mu, d_q, d_t);
    return 0;
}

// ---------------------------------------------------------------------------------
// M-step sums (deterministic two-stage reduction)
// ---------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(256) mstep_sums_kernel(int64_t m, const double *__restrict__ X, const double *__restrict__ z,
                                                         const double *__restrict__ t, const double *__restrict__ om1,
                                                         const double *__restrict__ ve, const double *__restrict__ vfs,
                                                         double sigma2fs, int mode, double *__restrict__ partial) {
    constexpr int NS = 2 * P * P + 2 * P + 2;
    double s[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) s[i] = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (int64_t)gridDim.x * blockDim.x) {
        double u = 1.0, w = 1.0, v = vfs[j];
        if (mode == 0) {
            double d = sigma2fs * v + ve[j];
            u = 1.0 / d;
            w = v / (d * d);
        }
        double x[P];
#pragma unroll
        for (int a = 0; a < P; a++) x[a] = X[j + (int64_t)a * m];
        const double zj = z[j], tj = t[j];
#pragma unroll
        for (int a = 0; a < P; a++) {
#pragma unroll
            for (int b = 0; b < P; b++) {
                s[a + P * b] += u * x[a] * x[b];
                s[P * P + 2 * P + 2 + a + P * b] += w * x[a] * x[b];
            }
            s[P * P + a] += u * x[a] * (zj - tj);
            s[P * P + P + 2 + a] += w * x[a] * (tj - zj);
        }
        s[P * P + P] += v * u;
        s[P * P + P + 1] += w * ((om1[j] + tj * tj) - 2.0 * tj * zj + zj * zj);   // Omega1_j = q_j + t_j^2
    }
    __shared__ double sh[8][NS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NS; i++) {
        double v = warp_sum(s[i]);
        if (lane == 0) sh[wid][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        double v = 0;
        for (int w8 = 0; w8 < 8; w8++) v += sh[w8][threadIdx.x];
        partial[(size_t)blockIdx.x * NS + threadIdx.x] = v;
    }
}

__global__ void final_sum_kernel(int nblocks, int ns, const double *__restrict__ partial, double *__restrict__ out) {
    int i = threadIdx.x;
    if (i < ns) {
        double v = 0;
        for (int b = 0; b < nblocks; b++) v += partial[(size_t)b * ns + i];
        out[i] = v;
    }
}

extern "C" int frk_mstep_sums(frk_ctx *ctx, int64_t m, int p, const double *d_X, const double *d_z, const double *d_t,
                              const double *d_omega1, const double *d_ve, const double *d_vfs, double sigma2fs, int mode,
                              double *h_out) {
    FRK_REQUIRE(ctx && h_out, "frk_mstep_sums: NULL argument");
    FRK_REQUIRE(p >= 1 && p <= MAXP, "frk_mstep_sums: %d fixed effects; the device path handles 1..%d", p, MAXP);
    const int ns = 2 * p * p + 2 * p + 2;
    if (m == 0) { for (int i = 0; i < ns; i++) h_out[i] = 0.0; return 0; }
    const int nblocks = (int)std::min<int64_t>(cdiv(m, 256), 64);
    double *partial = ctx->d_small;            // 64 * 42 <= 4096 - 64
    double *out = ctx->d_small + 4096 - 64;
    switch (p) {
        case 1: FRK_LAUNCH(ctx, mstep_sums_kernel<1>, nblocks, 256, 0, m, d_X, d_z, d_t, d_omega1, d_ve, d_vfs, sigma2fs, mode, partial); break;
        case 2: FRK_LAUNCH(ctx, mstep_sums_kernel<2>, nblocks, 256, 0, m, d_X, d_z, d_t, d_omega1, d_ve, d_vfs, sigma2fs, mode, partial); break;
        case 3: FRK_LAUNCH(ctx, mstep_sums_kernel<3>, nblocks, 256, 0, m, d_X, d_z, d_t, d_omega1, d_ve, d_vfs, sigma2fs, mode, partial); break;
        default: FRK_LAUNCH(ctx, mstep_sums_kernel<4>, nblocks, 256, 0, m, d_X, d_z, d_t, d_omega1, d_ve, d_vfs, sigma2fs, mode, partial); break;
    }
    FRK_LAUNCH(ctx, final_sum_kernel, 1, 64, 0, nblocks, ns, partial, out);
    FRK_CHECK_CUDA(cudaMemcpyAsync(ctx->h_pinned, out, ns * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < ns; i++) h_out[i] = ctx->h_pinned[i];
    return 0;
}

// ---------------------------------------------------------------------------------
// prediction epilogue, scatter, residual
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) predict_finalize_kernel(int64_t n, int mode, double sigma2fs, const double *__restrict__ q,
                                                               const double *__restrict__ smu, const double *__restrict__ xa,
                                                               const double *__restrict__ A, const double *__restrict__ ytil,
                                                               const double *__restrict__ fs, double *__restrict__ mu,
                                                               double *__restrict__ var, double *__restrict__ sd) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double m_, v_;
        if (mode == 0) {
            m_ = xa[i] + smu[i];
            v_ = q[i];
        } else {
            const double Ai = A ? A[i] : 0.0, yt = ytil ? ytil[i] : 0.0;
            const double Qf = 1.0 / (sigma2fs * fs[i]);
            const double den = Ai + Qf;
            const double a = Ai / den;
            v_ = (1.0 - a) * (1.0 - a) * q[i] + 1.0 / den;
            m_ = xa[i] + smu[i] + (yt - Ai * smu[i]) / den;
        }
        mu[i] = m_;
        var[i] = v_;
        sd[i] = sqrt(v_);
    }
}

extern "C" int frk_predict_finalize(frk_ctx *ctx, int64_t n, int mode, double sigma2fs, const double *d_q, const double *d_smu,
                                    const double *d_xa, const double *d_A, const double *d_ytil, const double *d_fs,
                                    double *d_mu, double *d_var, double *d_sd) {
    FRK_REQUIRE(ctx && d_q && d_smu && d_xa && d_mu && d_var && d_sd, "frk_predict_finalize: NULL argument");
    FRK_REQUIRE(mode == 0 || d_fs, "frk_predict_finalize: mode 1 needs the BAU fs field");
    if (n == 0) return 0;
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, predict_finalize_kernel, grid, 256, 0, n, mode, sigma2fs, d_q, d_smu, d_xa, d_A, d_ytil, d_fs, d_mu, d_var, d_sd);
    return 0;
}

__global__ void scatter_add_kernel(int64_t m, const int *__restrict__ idx, const double *__restrict__ src, double *__restrict__ dst) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(dst + idx[j], src[j]);
}

extern "C" int frk_scatter_add(frk_ctx *ctx, int64_t m, const int *d_idx, const double *d_src, double *d_dst) {
    if (m == 0) return 0;
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(m, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, scatter_add_kernel, grid, 256, 0, m, d_idx, d_src, d_dst);
    return 0;
}

__global__ void resid_kernel(int64_t m, int p, const double *__restrict__ X, Alpha alpha, const double *__restrict__ z,
                             const double *__restrict__ dv, double *__restrict__ out) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (int64_t)gridDim.x * blockDim.x) {
        double rho = z ? z[j] : 0.0;
        for (int q = 0; q < p; q++) rho -= X[j + (int64_t)q * m] * alpha.a[q];
        out[j] = dv ? rho / dv[j] : rho;
    }
}

extern "C" int frk_resid(frk_ctx *ctx, int64_t m, int p, const double *d_X, const double *h_alpha, const double *d_z,
                         const double *d_div, double *d_out) {
    FRK_REQUIRE(p >= 0 && p <= MAXP, "frk_resid: %d fixed effects; the device path handles at most %d", p, MAXP);
    if (m == 0) return 0;
    Alpha al{};
    for (int q = 0; q < p; q++) al.a[q] = h_alpha[q];
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(m, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, resid_kernel, grid, 256, 0, m, p, d_X, al, d_z, d_div, d_out);
    return 0;
}

__global__ void xalpha_kernel(int64_t n, int p, const double *__restrict__ X, Alpha alpha, double *__restrict__ out) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < p; q++) s += X[j + (int64_t)q * n] * alpha.a[q];
        out[j] = s;
    }
}

extern "C" int frk_xalpha(frk_ctx *ctx, int64_t n, int p, const double *d_X, const double *h_alpha, double *d_out) {
    FRK_REQUIRE(p >= 0 && p <= MAXP, "frk_xalpha: %d fixed effects; the device path handles at most %d", p, MAXP);
    if (n == 0) return 0;
    Alpha al{};
    for (int q = 0; q < p; q++) al.a[q] = h_alpha[q];
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, xalpha_kernel, grid, 256, 0, n, p, d_X, al, d_out);
    return 0;
}

// ---------------------------------------------------------------------------------
// vector statistics: sum(x - shift), sum((x - shift)^2), min, max  (var(Z) and mean(Ve) of
// .EM_initialise R/SRE.R:704-731; the homoscedasticity test all(a == a[1]) R/SREfit.R:278-280)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vec_stats_kernel(int64_t n, const double *__restrict__ x, double shift,
                                                        double *__restrict__ partial) {
    double s = 0.0, s2 = 0.0, mn = INFINITY, mx = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = x[i], c = v - shift;
        s += c; s2 += c * c;
        mn = fmin(mn, v); mx = fmax(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __shared__ double sh[8][4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sh[wid][0] = s; sh[wid][1] = s2; sh[wid][2] = mn; sh[wid][3] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) { s += sh[w][0]; s2 += sh[w][1]; mn = fmin(mn, sh[w][2]); mx = fmax(mx, sh[w][3]); }
        double *o = partial + (size_t)blockIdx.x * 4;
        o[0] = s; o[1] = s2; o[2] = mn; o[3] = mx;
    }
}

__global__ void vec_stats_final_kernel(int nblocks, const double *__restrict__ partial, double *__restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0, s2 = 0, mn = INFINITY, mx = -INFINITY;
        for (int b = 0; b < nblocks; b++) {
            s += partial[b * 4]; s2 += partial[b * 4 + 1];
            mn = fmin(mn, partial[b * 4 + 2]); mx = fmax(mx, partial[b * 4 + 3]);
        }
        out[0] = s; out[1] = s2; out[2] = mn; out[3] = mx;
    }
}

extern "C" int frk_vec_stats(frk_ctx *ctx, int64_t n, const double *d_x, double shift, double *h_out) {
    FRK_REQUIRE(ctx && h_out && (n == 0 || d_x), "frk_vec_stats: NULL argument");
    if (n == 0) { h_out[0] = h_out[1] = 0.0; h_out[2] = INFINITY; h_out[3] = -INFINITY; return 0; }
    const int nblocks = (int)std::min<int64_t>(cdiv(n, 256), 512);
    double *part = ctx->d_small, *out = ctx->d_small + 2048;
    FRK_LAUNCH(ctx, vec_stats_kernel, nblocks, 256, 0, n, d_x, shift, part);
    FRK_LAUNCH(ctx, vec_stats_final_kernel, 1, 32, 0, nblocks, part, out);
    FRK_CHECK_CUDA(cudaMemcpyAsync(ctx->h_pinned, out, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 4; i++) h_out[i] = ctx->h_pinned[i];
    return 0;
}

// ---------------------------------------------------------------------------------
// element-wise helpers for SRE(): Ve = std^2 (R/SRE.R:468), 1/Ve (R/SREpredict.R:286-293), and the
// row gather behind Vfs = diag(fs[bau]) (R/SRE.R:498) and X = model matrix at the observed BAUs (:464-465)
// ---------------------------------------------------------------------------------
__global__ void vec_unary_kernel(int64_t n, int op, const double *__restrict__ a, double *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = a[i];
        out[i] = op == 0 ? v * v : 1.0 / v;
    }
}

extern "C" int frk_square(frk_ctx *ctx, int64_t n, const double *d_a, double *d_out) {
    if (n == 0) return 0;
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, vec_unary_kernel, grid, 256, 0, n, 0, d_a, d_out);
    return 0;
}

extern "C" int frk_recip(frk_ctx *ctx, int64_t n, const double *d_a, double *d_out) {
    if (n == 0) return 0;
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, vec_unary_kernel, grid, 256, 0, n, 1, d_a, d_out);
    return 0;
}

__global__ void gather_rows_kernel(int64_t m, int ncols, const int *__restrict__ idx, const double *__restrict__ src, int64_t ld_src,
                                   double *__restrict__ dst, int64_t ld_dst) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = idx[j];
        for (int c = 0; c < ncols; c++) dst[j + c * ld_dst] = src[s + c * ld_src];
    }
}

extern "C" int frk_gather_rows(frk_ctx *ctx, int64_t m, int ncols, const int *d_idx, const double *d_src, int64_t ld_src,
                               double *d_dst, int64_t ld_dst) {
    FRK_REQUIRE(ctx && (m == 0 || (d_idx && d_src && d_dst)), "frk_gather_rows: NULL argument");
    if (m == 0 || ncols == 0) return 0;
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(m, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, gather_rows_kernel, grid, 256, 0, m, ncols, d_idx, d_src, ld_src, d_dst, ld_dst);
    return 0;
}

// x *= alpha   (K_i^-1 = omega_i exp(-D_i/tau_i)^-1 in the block-exponential update, R/SREfit.R:625-637)
__global__ void scale_kernel(int64_t n, double alpha, double *__restrict__ x) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] *= alpha;
}

extern "C" int frk_scale(frk_ctx *ctx, int64_t n, double alpha, double *d_x) {
    FRK_REQUIRE(ctx && (n == 0 || d_x), "frk_scale: NULL argument");
    if (n == 0) return 0;
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, scale_kernel, grid, 256, 0, n, alpha, d_x);
    return 0;
}

// x[i] = value   (the intercept column of the model matrix at the BAUs, R/SRE.R:464-465)
__global__ void fill_kernel(int64_t n, double value, double *__restrict__ x) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = value;
}

extern "C" int frk_fill(frk_ctx *ctx, int64_t n, double value, double *d_x) {
    FRK_REQUIRE(ctx && (n == 0 || d_x), "frk_fill: NULL argument");
    if (n == 0) return 0;
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, fill_kernel, grid, 256, 0, n, value, d_x);
    return 0;
}

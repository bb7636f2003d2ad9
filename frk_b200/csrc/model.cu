// frk_model: the device-resident SRE object and the EM / prediction control flow above the kernels.
//
// Reference: R/SRE.R:704-731 (.EM_initialise), R/SREfit.R:78-160 (.EM_fit), :174-220 (.loglik.ind), :232-260
// (.SRE.Estep.ind), :263-438 (.SRE.Mstep.ind), :441-694 (.update_K / .regularise_K), R/SREpredict.R:243-375
// (.SRE.predict_EM), R/SREutils.R:245-305 (.batch_compute_var).  stats::optim(method = "Nelder-Mead") and
// stats::uniroot, which the reference runs on the host around the data-parallel objective functions
// (R/SREfit.R:386-388, :618-621), are restated here from the published algorithms (R src/appl/optim.c nmmin,
// R src/library/stats/src/zeroin.c) so that one C-ABI call -- what the R `.Call` shim binds -- runs a whole
// SRE.fit(method = "EM") with every m-, N- or r^2-sized operation on the GPU and nothing but scalars on the host.
//
// Arithmetic that the reference computes twice per iteration with identical inputs is computed once:
// .loglik.ind and .SRE.Estep.ind both form S'D^-1 S and factor K^-1 + S'D^-1 S (the reference's two K^-1 are
// the same chol2inv(chol(Khat))), and the M-step's Khat_inv factorisation supplies the next log|K|.
#include "common.cuh"
#include "coupled.cuh"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <limits>
#include <memory>
#include <string>
#include <thread>

namespace {

constexpr int MAXP_M = 4;

struct Block {                    // one resolution of the block-exponential K
    int ni = 0, ix0 = 0;
    int *d_ix = nullptr;
    double *D = nullptr, *eta2 = nullptr, *Ki = nullptr, *Kinv = nullptr, *work = nullptr, *KinvBest = nullptr;
    double D01 = 0.0;
    // what KinvBest holds: exp(-D/kinv_tau0)^-1 (UNSCALED) and log|exp(-D/kinv_tau0)|; kinv_tau0 <= 0: nothing usable
    double kinv_tau0 = -1.0, kinv_ld = 0.0;
    // memo of recent evaluations: exp(-D/tau0)^-1 and log|exp(-D/tau0)| are functions of tau alone, and the Nelder-Mead runs of
    // consecutive M-steps ask for the same points again once a resolution's tau has settled (f_tau_lookup); LRU, buffers allocated
    // with the block
    struct Memo { double tau0 = -1.0, ld = 0.0; double *buf = nullptr; unsigned long long stamp = 0; };
    std::vector<Memo> memo;
    unsigned long long memo_clock = 0;
    // tensor-product basis: n_i = ns * nt, K_i = kronecker(exp(-Dt/tau_t), exp(-D/tau_s)) (D = spatial distances, ns x ns)
    int ns = 0, nt = 0;
    double *Ks = nullptr, *Qs = nullptr, *ws = nullptr, *QsBest = nullptr;     // ns x ns
    double *Kt = nullptr, *Qt = nullptr, *wt = nullptr, *QtBest = nullptr;     // nt x nt
};

}  // namespace

struct frk_model {
    frk_ctx *ctx = nullptr;
    int64_t m = 0;
    int r = 0, p = 0;
    // S (m x r CSR), data vectors: device; owned when created from host buffers
    const int *rowptr = nullptr, *col = nullptr;
    const double *val = nullptr, *Z = nullptr, *X = nullptr, *Ve = nullptr, *Vfs = nullptr;
    bool owned = false;
    frk_rowbuckets *bk = nullptr;
    double *Linv = nullptr;         // dense / implicit S: L^-1 of the last E-step into the slots (S_eta = L^-T L^-1), for the triangular quadratic forms
    bool linv_valid = false;
    frk_coupling *cp = nullptr;     // non-diagonal Vfs (observations coupled through shared BAUs): see coupled.cu; M->Vfs then holds diag(Vfs)
    // r-sized slots (device), replicated on every rank
    double *mu_eta = nullptr, *S_eta = nullptr, *Q_eta = nullptr, *Khat = nullptr, *Khat_inv = nullptr;
    // scratch
    double *acc = nullptr, *w1 = nullptr, *w2 = nullptr, *vec = nullptr, *t = nullptr, *q = nullptr;
    // host state
    double alpha[MAXP_M] = {0, 0, 0, 0};
    double sigma2fs = 0.0, logdetK = 0.0, logdetQ = 0.0;   // logdetQ: log|Q_eta| of the last E-step into the slots
    double m_total = 0.0, Ve0 = 0.0, Vfs0 = 0.0;
    bool homoscedastic = false, all_vfs_zero = false, initialised = false;
    bool khat_block_diag = false;   // Khat is block diagonal by resolution AND Khat_inv = chol2inv(chol(Khat)) (init / block-exponential update)
    // block-exponential K
    std::vector<Block> blocks;
    double max_l = 0.0;
    bool tensor = false;
    double *Dt = nullptr;          // nt x nt temporal distances (tensor bases)
    double Dt01 = 0.0;
    std::vector<double> taus, omegas;
    double nm_requests = 0, nm_rounds = 0, nm_evals = 0, nm_reevals = 0;   // Nelder-Mead statistics of the last M-step (distinct points asked, batches, points evaluated)
    double nm_prefetch_hits = 0;     // ... of which had their factorisation enqueued before the E-step (prefetch_f_tau)
    double nm_reuse_hits = 0;        // ... of which were served by a saved inverse instead of a factorisation (f_tau_lookup)
    // Prefetch: K_i(tau)^-1 and log|K_i(tau)| depend on tau only.  The first two points the NEXT M-step's Nelder-Mead will ask for
    // (its start value and start + step) are known as soon as this M-step has assembled K, so their factorisations are enqueued on the
    // lanes right away and run beside the next E-step's factorisation of Q instead of after it.  pre[j][c]: (active, tau) of lane c of block j.
    struct PreSlot { bool on = false; double tau = 0.0; bool reused = false; double tau0 = 0.0, ld = 0.0; };
    std::vector<std::array<PreSlot, 8>> pre;
    bool prefetch = true;
    // multi-rank reduction hook
    int rank = 0, world = 1;
    bool spec_emulate = false;     // evaluate every speculative candidate on this rank (single-GPU test of the multi-rank logic)
    bool nm_distribute = false;    // spread the Nelder-Mead candidates over the ranks (else: every rank evaluates them on its own lanes,
                                   // replicated and deterministic like the rest of the r-sized work, with no exchange at all)
    frk_allreduce_fn allreduce = nullptr;
    void *user = nullptr;
    std::vector<void *> allocs;
    // lanes: extra contexts on this GPU that evaluate speculative Nelder-Mead candidates concurrently (a factorisation is latency
    // bound and fills a fraction of the SMs, so a few of them side by side cost little more than one)
    int lanes = 1;
    cudaEvent_t ev_lane = nullptr;
    double t_phase[4] = {0, 0, 0, 0};   // FRK_PHASE_DEBUG: host seconds in posterior / update_K / rest of the M-step / Nelder-Mead rounds
    frk_ctx *aux = nullptr;          // low-priority lane: the sigma2fs update and the next Gram run beside the (latency-bound) K update
    cudaEvent_t ev_aux0 = nullptr, ev_aux1 = nullptr;
    bool gram_early = false;         // acc already holds the Gram for (gram_alpha, gram_s2): enqueued on aux, completion = ev_aux1
    double gram_alpha[MAXP_M] = {0, 0, 0, 0}, gram_s2 = 0.0;
};

namespace {

#define RC(expr) do { if (int _rc = (expr)) return _rc; } while (0)

inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <typename T>
int dalloc(frk_model *M, T **out, size_t n) {
    void *p = nullptr;
    FRK_CHECK_CUDA(cudaMallocAsync(&p, std::max<size_t>(n * sizeof(T), 16), M->ctx->stream));   // pooled: re-creating a model is cheap
    M->allocs.push_back(p);
    *out = (T *)p;
    return 0;
}

// Row-partitioned runs: the caller's reduction hook if one was given, else the context's own NCCL communicator (comm.cu)
inline bool multi_rank(const frk_model *M) { return M->allreduce != nullptr || M->ctx->comm != nullptr; }

int reduce_dev(frk_model *M, double *d_buf, int64_t n) {
    if (M->allreduce) {
        FRK_REQUIRE(M->allreduce(M->user, d_buf, n, 1, 0) == 0, "all-reduce callback failed (device buffer)");
        return 0;
    }
    return frk_comm_allreduce(M->ctx, d_buf, n, 0);
}

int reduce_host(frk_model *M, double *h_buf, int64_t n, int op) {
    if (M->allreduce) {
        FRK_REQUIRE(M->allreduce(M->user, h_buf, n, 0, op) == 0, "all-reduce callback failed (host buffer)");
        return 0;
    }
    return frk_comm_allreduce_host(M->ctx, h_buf, n, op);
}

// d_buf (valid on rank `root` of the Nelder-Mead world) to every rank: ncclBroadcast on the communicator; through a
// reduction hook it is a sum with zeros elsewhere
int bcast_dev(frk_model *M, double *d_buf, int64_t n, int root, bool mine) {
    if (!M->allreduce && M->ctx->comm) return frk_comm_broadcast(M->ctx, d_buf, n, root);
    if (!mine) FRK_CHECK_CUDA(cudaMemsetAsync(d_buf, 0, (size_t)n * sizeof(double), M->ctx->stream));
    return reduce_dev(M, d_buf, n);
}

// acc = [upper(G) full r x r storage | rhs r | 2 scalars] summed over the ranks; what travels is the packed upper triangle
int reduce_gram(frk_model *M) {
    if (!multi_rank(M)) return 0;
    const int r = M->r;
    const size_t rr = (size_t)r * r, np = (size_t)r * (r + 1) / 2, tail = (size_t)r + 2;
    void *raw = nullptr;
    RC(frk_scratch(M->ctx, (np + tail) * sizeof(double), &raw));
    double *P = (double *)raw;
    RC(frk_pack_upper(M->ctx, r, M->acc, P));
    FRK_CHECK_CUDA(cudaMemcpyAsync(P + np, M->acc + rr, tail * sizeof(double), cudaMemcpyDeviceToDevice, M->ctx->stream));
    RC(reduce_dev(M, P, (int64_t)(np + tail)));
    RC(frk_unpack_upper(M->ctx, r, P, M->acc));
    FRK_CHECK_CUDA(cudaMemcpyAsync(M->acc + rr, P + np, tail * sizeof(double), cudaMemcpyDeviceToDevice, M->ctx->stream));
    return 0;
}

// ranks that share the Nelder-Mead candidates (1: replicated evaluation on local lanes)
inline int nm_world(const frk_model *M) { return (M->nm_distribute || M->spec_emulate) ? std::max(M->world, 1) : 1; }
inline int nm_rank(const frk_model *M) { return (M->nm_distribute && !M->spec_emulate) ? M->rank : 0; }

// sum, sum of squares about `shift`, min, max over ALL ranks' rows
int global_stats(frk_model *M, const double *d_x, double shift, double out[4]) {
    RC(frk_vec_stats(M->ctx, M->m, d_x, shift, out));
    if (multi_rank(M)) {
        double s[2] = {out[0], out[1]}, mn[1] = {-out[2]}, mx[1] = {out[3]};
        RC(reduce_host(M, s, 2, 0));
        RC(reduce_host(M, mn, 1, 1));
        RC(reduce_host(M, mx, 1, 1));
        out[0] = s[0]; out[1] = s[1]; out[2] = -mn[0]; out[3] = mx[0];
    }
    return 0;
}

// p x p solve A x = b (Gaussian elimination with partial pivoting; A column-major); p <= 4
int solve_small(int p, const double *A, const double *b, double *x) {
    double a[MAXP_M][MAXP_M + 1];
    for (int i = 0; i < p; i++) { for (int j = 0; j < p; j++) a[i][j] = A[i + p * j]; a[i][p] = b[i]; }
    for (int c = 0; c < p; c++) {
        int piv = c;
        for (int i = c + 1; i < p; i++) if (std::fabs(a[i][c]) > std::fabs(a[piv][c])) piv = i;
        FRK_REQUIRE(a[piv][c] != 0.0, "the fixed-effects normal equations are singular (X'X not invertible)");
        if (piv != c) for (int j = 0; j <= p; j++) std::swap(a[c][j], a[piv][j]);
        for (int i = c + 1; i < p; i++) {
            const double f = a[i][c] / a[c][c];
            for (int j = c; j <= p; j++) a[i][j] -= f * a[c][j];
        }
    }
    for (int i = p - 1; i >= 0; i--) {
        double s = a[i][p];
        for (int j = i + 1; j < p; j++) s -= a[i][j] * x[j];
        x[i] = s / a[i][i];
    }
    return 0;
}

struct MSums { double uxx[16], uxz[4], sv, w0, wx[4], wxx[16]; };

// R/SREfit.R:335-356,391-405,408-417: the row sums behind J(sigma2fs), alpha and the closed-form sigma2fs
int mstep_sums(frk_model *M, frk_ctx *c, double sigma2fs, int mode, MSums &s) {
    const int p = M->p, ns = 2 * p * p + 2 * p + 2;
    double out[2 * 16 + 2 * 4 + 2];
    if (M->cp && mode == 0) RC(frk_coupling_msums(M->cp, c, M->X, M->Z, M->t, M->Ve, sigma2fs, out));   // general J, R/SREfit.R:296-328
    else RC(frk_mstep_sums(c, M->m, p, M->X, M->Z, M->t, M->q, M->Ve, M->Vfs, sigma2fs, mode, out));
    RC(reduce_host(M, out, ns, 0));                                  // O(p^2) doubles
    int o = 0;
    for (int i = 0; i < p * p; i++) s.uxx[i] = out[o++];
    for (int i = 0; i < p; i++) s.uxz[i] = out[o++];
    s.sv = out[o++]; s.w0 = out[o++];
    for (int i = 0; i < p; i++) s.wx[i] = out[o++];
    for (int i = 0; i < p * p; i++) s.wxx[i] = out[o++];
    return 0;
}

// alpha = (sum u x x')^-1 sum u x (z - t);  sumwOm = sum w Omega(alpha)
int gls_and_sums(frk_model *M, frk_ctx *c, double s2, int mode, double *al, MSums &s, double &sumwOm) {
    RC(mstep_sums(M, c, s2, mode, s));
    RC(solve_small(M->p, s.uxx, s.uxz, al));
    const int p = M->p;
    double lin = 0.0, quad = 0.0;
    for (int i = 0; i < p; i++) {
        lin += al[i] * s.wx[i];
        double row = 0.0;
        for (int j = 0; j < p; j++) row += s.wxx[i + p * j] * al[j];
        quad += al[i] * row;
    }
    sumwOm = s.w0 + 2.0 * lin + quad;
    return 0;
}

// ---------------------------------------------------------------------------------
// stats::optim(method = "Nelder-Mead")  (R src/appl/optim.c: nmmin)
// ---------------------------------------------------------------------------------
struct NMResult { std::vector<double> x; double f; int count, fail; };

int nmmin(const std::function<int(const double *, double &)> &fn, std::vector<double> B, NMResult &res, int maxit = 100,
          double reltol = 1e-4, double abstol = -std::numeric_limits<double>::infinity(), double alpha = 1.0, double bet = 0.5,
          double gamm = 2.0) {
    const double big = 1.0e35;
    const int n = (int)B.size(), n1 = n + 1, C = n + 2;
    std::vector<double> P((size_t)n1 * C, 0.0);
    auto Pm = [&](int i, int j) -> double & { return P[(size_t)i * C + j]; };   // P[i][j], 0-based
    double f;
    RC(fn(B.data(), f));
    FRK_REQUIRE(std::isfinite(f), "function cannot be evaluated at initial parameters");
    int funcount = 1;
    const double convtol = reltol * (std::fabs(f) + reltol);
    Pm(n1 - 1, 0) = f;
    for (int i = 0; i < n; i++) Pm(i, 0) = B[i];
    int L = 1;
    double size = 0.0, step = 0.0;
    for (int i = 0; i < n; i++) if (0.1 * std::fabs(B[i]) > step) step = 0.1 * std::fabs(B[i]);
    if (step == 0.0) step = 0.1;
    for (int j = 2; j <= n1; j++) {
        for (int i = 0; i < n; i++) Pm(i, j - 1) = B[i];
        double trystep = step;
        while (Pm(j - 2, j - 1) == B[j - 2]) { Pm(j - 2, j - 1) = B[j - 2] + trystep; trystep *= 10; }
        size += trystep;
    }
    double oldsize = size;
    bool calcvert = true;
    int fail = 0, H;
    double VL, VH, VR;
    while (true) {
        if (calcvert) {
            for (int j = 0; j < n1; j++) {
                if (j + 1 != L) {
                    for (int i = 0; i < n; i++) B[i] = Pm(i, j);
                    RC(fn(B.data(), f));
                    if (!std::isfinite(f)) f = big;
                    funcount++;
                    Pm(n1 - 1, j) = f;
                }
            }
            calcvert = false;
        }
        VL = Pm(n1 - 1, L - 1); VH = VL; H = L;
        for (int j = 1; j <= n1; j++) {
            if (j != L) {
                f = Pm(n1 - 1, j - 1);
                if (f < VL) { L = j; VL = f; }
                if (f > VH) { H = j; VH = f; }
            }
        }
        if (VH <= VL + convtol || VL <= abstol) break;
        for (int i = 0; i < n; i++) {
            double temp = -Pm(i, H - 1);
            for (int j = 0; j < n1; j++) temp += Pm(i, j);
            Pm(i, C - 1) = temp / n;
        }
        for (int i = 0; i < n; i++) B[i] = (1.0 + alpha) * Pm(i, C - 1) - alpha * Pm(i, H - 1);
        RC(fn(B.data(), f));
        if (!std::isfinite(f)) f = big;
        funcount++;
        VR = f;
        if (VR < VL) {
            Pm(n1 - 1, C - 1) = f;
            for (int i = 0; i < n; i++) {
                f = gamm * B[i] + (1 - gamm) * Pm(i, C - 1);
                Pm(i, C - 1) = B[i];
                B[i] = f;
            }
            RC(fn(B.data(), f));
            if (!std::isfinite(f)) f = big;
            funcount++;
            if (f < VR) {
                for (int i = 0; i < n; i++) Pm(i, H - 1) = B[i];
                Pm(n1 - 1, H - 1) = f;
            } else {
                for (int i = 0; i < n; i++) Pm(i, H - 1) = Pm(i, C - 1);
                Pm(n1 - 1, H - 1) = VR;
            }
        } else {
            if (VR < VH) {
                for (int i = 0; i < n; i++) Pm(i, H - 1) = B[i];
                Pm(n1 - 1, H - 1) = VR;
            }
            for (int i = 0; i < n; i++) B[i] = (1 - bet) * Pm(i, H - 1) + bet * Pm(i, C - 1);
            RC(fn(B.data(), f));
            if (!std::isfinite(f)) f = big;
            funcount++;
            if (f < Pm(n1 - 1, H - 1)) {
                for (int i = 0; i < n; i++) Pm(i, H - 1) = B[i];
                Pm(n1 - 1, H - 1) = f;
            } else if (VR >= VH) {
                calcvert = true;
                size = 0.0;
                for (int j = 0; j < n1; j++) {
                    if (j + 1 != L) {
                        for (int i = 0; i < n; i++) {
                            Pm(i, j) = bet * (Pm(i, j) - Pm(i, L - 1)) + Pm(i, L - 1);
                            size += std::fabs(Pm(i, j) - Pm(i, L - 1));
                        }
                    }
                }
                if (size < oldsize) oldsize = size;
                else { fail = 10; break; }
            }
        }
        if (!(funcount <= maxit)) break;
    }
    res.f = Pm(n1 - 1, L - 1);
    res.x.resize(n);
    for (int i = 0; i < n; i++) res.x[i] = Pm(i, L - 1);
    if (funcount > maxit) fail = 1;
    res.count = funcount; res.fail = fail;
    return 0;
}

// ---------------------------------------------------------------------------------
// stats::uniroot -> R_zeroin2 (Brent), default tol = .Machine$double.eps^0.25
// ---------------------------------------------------------------------------------
int zeroin(const std::function<int(double, double &)> &f, double lower, double upper, double &root,
           double tol = std::pow(std::numeric_limits<double>::epsilon(), 0.25), int maxiter = 1000) {
    const double EPS = std::numeric_limits<double>::epsilon(), xmax = std::numeric_limits<double>::max();
    double fa, fb;
    RC(f(lower, fa));
    RC(f(upper, fb));
    fa = std::min(std::max(fa, -xmax), xmax);
    fb = std::min(std::max(fb, -xmax), xmax);
    FRK_REQUIRE(!(fa * fb > 0), "f() values at end points not of opposite sign");
    double a = lower, b = upper, c = a, fc = fa;
    if (fa == 0.0) { root = a; return 0; }
    if (fb == 0.0) { root = b; return 0; }
    int maxit = maxiter + 1;
    while (maxit-- > 0) {
        const double prev_step = b - a;
        if (std::fabs(fc) < std::fabs(fb)) { a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
        const double tol_act = 2 * EPS * std::fabs(b) + tol / 2;
        double new_step = (c - b) / 2;
        if (std::fabs(new_step) <= tol_act || fb == 0.0) { root = b; return 0; }
        if (std::fabs(prev_step) >= tol_act && std::fabs(fa) > std::fabs(fb)) {
            double t1, t2, pp, qq;
            const double cb = c - b;
            if (a == c) { t1 = fb / fa; pp = cb * t1; qq = 1.0 - t1; }
            else {
                qq = fa / fc; t1 = fb / fc; t2 = fb / fa;
                pp = t2 * (cb * qq * (qq - t1) - (b - a) * (t1 - 1.0));
                qq = (qq - 1.0) * (t1 - 1.0) * (t2 - 1.0);
            }
            if (pp > 0) qq = -qq; else pp = -pp;
            if (pp < (0.75 * cb * qq - std::fabs(tol_act * qq) / 2) && pp < std::fabs(prev_step * qq / 2)) new_step = pp / qq;
        }
        if (std::fabs(new_step) < tol_act) new_step = new_step > 0 ? tol_act : -tol_act;
        a = b; fa = fb;
        b += new_step;
        RC(f(b, fb));
        if ((fb > 0 && fc > 0) || (fb < 0 && fc < 0)) { c = a; fc = fa; }
    }
    root = b;
    return 0;
}

// ---------------------------------------------------------------------------------
// Speculative Nelder-Mead over the ranks of a row-partitioned run.
// Every f_tau evaluation is a replicated r_i x r_i factorisation, so on W > 1 GPUs the ranks would all compute the same thing.
// Instead, whenever Nelder-Mead asks for a point that has not been evaluated, the points it may ask for NEXT are predicted by
// re-running nmmin on the host against the values known so far plus hypothetical outcomes ("lower than everything" / "higher
// than everything") for the unknown ones; rank k evaluates candidate k, one small all-reduce shares the values, and nmmin continues
// from the cache.  The iterate path -- hence every result -- is identical to the sequential one; only wall time changes.
// ---------------------------------------------------------------------------------
struct SpecNM {
    frk_model *M;
    // evaluates nb <= lanes points on THIS rank, concurrently on the lanes (and tracks the rank-local best)
    std::function<int(int nb, const double *const *, double *)> eval;
    int dim, maxit;
    double reltol;
    std::vector<double> par0;
    int lanes = 1;
    // called for every point Nelder-Mead actually asks for (cache hit or not), in request order: the caller keeps the by-products
    // (inverse, log-determinant) of the best REQUESTED point -- a speculative point that is never requested cannot be the result
    std::function<int(const double *, double)> requested;
    std::vector<std::vector<double>> pts;
    std::vector<double> fv;
    std::vector<int> owner;

    // A value of NaN marks a point whose K_i(tau) could not be factorised.  Speculative candidates are evaluated as a matter of
    // course, so such a failure must not abort the fit: it is an error only if Nelder-Mead actually ASKS for the point -- exactly
    // when the reference's chol() inside optim()'s objective would have stopped (R/SREfit.R:503,517).
    int check_requested(const double *x, double f) const {
        FRK_REQUIRE(!std::isnan(f), "block-exponential update: K_i(tau = %g) is not positive definite (chol failed at a requested point)", x[0]);
        return 0;
    }

    int find(const double *x) const {
        for (size_t i = 0; i < pts.size(); i++) {
            bool eq = true;
            for (int d = 0; d < dim; d++) eq = eq && pts[i][d] == x[d];
            if (eq) return (int)i;
        }
        return -1;
    }

    // Points nmmin may request after `x` (at most `want` in all, `x` included), for the lanes that would otherwise idle.
    // (1) One dimension, at least two known values: f_tau is a smooth function of tau, so the outcome of the next comparisons is
    //     PREDICTED from a local interpolant (the quadratic through the three known points nearest to the new one; a line through
    //     two) and nmmin is replayed along that predicted path: its next unknown points, in the order it would ask for them.
    //     With four or more evaluators and only a line to go by, one place is kept for an alternative (scheme 2).
    // (2) Otherwise, and for that alternative: hypothetical outcomes for the unknown points -- M = between the best and the second
    //     best known value (inside the simplex whatever its other vertices are), L = better than everything, H = worse than
    //     everything -- breadth first, the commonest outcome first (M, then L, then H: tools/spec_predictor_eval.py).
    // A wrong guess costs lane time only: values are used by nmmin only for points it asks for itself (replay).
    void candidates(const double *x, size_t want, std::vector<std::vector<double>> &batch) const {
        auto in_batch = [&](const std::vector<double> &c) {
            for (auto &b : batch) { bool eq = true; for (int d = 0; d < dim; d++) eq = eq && b[d] == c[d]; if (eq) return true; }
            return false;
        };
        static const bool no_model = getenv("FRK_SPEC_NO_MODEL") != nullptr;       // A/B: hypotheses only
        // ---- (1) the predicted path ----
        if (dim == 1 && !no_model && want > batch.size()) {
            std::vector<std::pair<double, double>> known;
            for (size_t i = 0; i < pts.size(); i++) if (std::isfinite(fv[i])) known.emplace_back(pts[i][0], fv[i]);
            if (known.size() >= 2) {
                auto predict = [&](double y) {
                    std::vector<std::pair<double, double>> k = known;
                    std::sort(k.begin(), k.end(), [y](const std::pair<double, double> &a, const std::pair<double, double> &b) {
                        return std::fabs(a.first - y) < std::fabs(b.first - y); });
                    const double x0 = k[0].first, f0 = k[0].second, x1 = k[1].first, f1 = k[1].second;
                    if (k.size() == 2) return f0 + (f1 - f0) * (y - x0) / (x1 - x0);
                    const double x2 = k[2].first, f2 = k[2].second;
                    return f0 * (y - x1) * (y - x2) / ((x0 - x1) * (x0 - x2)) + f1 * (y - x0) * (y - x2) / ((x1 - x0) * (x1 - x2)) +
                           f2 * (y - x0) * (y - x1) / ((x2 - x0) * (x2 - x1));
                };
                const size_t keep = (want >= 4 && known.size() < 3) ? 1 : 0;   // places left for scheme (2) while the model is a line
                std::vector<std::vector<double>> hp;
                std::vector<double> hf;
                auto sim = [&](const double *y, double &f) -> int {
                    const int i = find(y);
                    if (i >= 0) { f = fv[i]; return std::isnan(f) ? 99 : 0; }
                    for (size_t j = 0; j < hp.size(); j++) if (hp[j][0] == y[0]) { f = hf[j]; return 0; }
                    if (batch.size() + keep >= want && !in_batch(std::vector<double>(y, y + 1))) return 99;
                    f = predict(y[0]);
                    if (!std::isfinite(f)) return 99;
                    hp.emplace_back(y, y + 1); hf.push_back(f);
                    if (!in_batch(hp.back())) batch.push_back(hp.back());
                    return 0;
                };
                NMResult tmp;
                (void)nmmin(sim, par0, tmp, maxit, reltol);
            }
        }
        // ---- (2) hypothetical outcomes ----
        std::vector<std::string> queue = {"M", "L", "H"};
        for (size_t qi = 0; qi < queue.size() && batch.size() < want && qi < 400; qi++) {
            const std::string hyp = queue[qi];
            std::vector<std::vector<double>> hp;
            std::vector<double> hf;
            size_t used = 0;
            std::vector<double> cand;
            auto sim = [&](const double *y, double &f) -> int {
                const int i = find(y);
                if (i >= 0) { f = fv[i]; return 0; }
                for (size_t j = 0; j < hp.size(); j++) {
                    bool eq = true;
                    for (int d = 0; d < dim; d++) eq = eq && hp[j][d] == y[d];
                    if (eq) { f = hf[j]; return 0; }
                }
                if (used < hyp.size()) {
                    double lo = 0.0, lo2 = 0.0, hi = 0.0;                // best, second best, worst of everything known or assumed
                    int n = 0;
                    auto upd = [&](double v) {
                        if (!std::isfinite(v)) return;
                        if (n == 0) { lo = lo2 = hi = v; }
                        else {
                            if (v < lo) { lo2 = lo; lo = v; } else if (n == 1 || v < lo2) lo2 = v;
                            hi = std::max(hi, v);
                        }
                        n++;
                    };
                    for (double v : fv) upd(v);
                    for (double v : hf) upd(v);
                    if (n == 1) lo2 = hi = lo;
                    f = hyp[used] == 'L' ? lo - 1.0 - 1e-3 * std::fabs(lo) : hyp[used] == 'H' ? hi + 1.0 + 1e-3 * std::fabs(hi) : 0.5 * (lo + lo2);
                    used++;
                    hp.emplace_back(y, y + dim); hf.push_back(f);
                    return 0;
                }
                cand.assign(y, y + dim);
                return 99;                                  // the next point the real run would ask for under this hypothesis
            };
            NMResult tmp;
            const int rc = nmmin(sim, par0, tmp, maxit, reltol);
            if (rc == 99) {
                if (find(cand.data()) < 0 && !in_batch(cand)) batch.push_back(cand);
                queue.push_back(hyp + "M");
                queue.push_back(hyp + "L");
                queue.push_back(hyp + "H");
            }
        }
        (void)x;
    }

    // Replay nmmin against the values known so far.  0: Nelder-Mead finished (res); 99: `need` is the first point it asks for that
    // has no value yet.  Every point it asks for is reported to `requested` (repeats included, in request order).
    std::vector<char> seen;
    int replay(std::vector<double> &need, NMResult &res) {
        auto sim = [&](const double *y, double &f) -> int {
            const int i = find(y);
            if (i < 0) { need.assign(y, y + dim); return 99; }
            f = fv[i];
            RC(check_requested(y, f));
            if ((int)seen.size() <= i) seen.resize(i + 1, 0);
            if (!seen[i]) { seen[i] = 1; M->nm_requests += 1; }
            return requested ? requested(y, f) : 0;
        };
        return nmmin(sim, par0, res, maxit, reltol);
    }

    int operator()(const double *x, double &f) {
        int i = find(x);
        M->nm_requests += 1;
        if (i >= 0) { f = fv[i]; RC(check_requested(x, f)); return requested ? requested(x, f) : 0; }
        M->nm_rounds += 1;
        static const bool dbg = getenv("FRK_SPEC_DEBUG") != nullptr;
        // candidate k goes to rank k % world (the likeliest ones spread over the GPUs first), lane k / world
        const int world = nm_world(M);
        const int L = std::max(lanes, 1), W = world * L;
        std::vector<std::vector<double>> batch = {std::vector<double>(x, x + dim)};
        if (W > 1) candidates(x, (size_t)W, batch);
        std::vector<double> fl(batch.size(), 0.0);
        std::vector<size_t> mine;
        for (size_t k = 0; k < batch.size(); k++)
            if (M->spec_emulate || (int)(k % world) == nm_rank(M)) mine.push_back(k);
        for (size_t c0 = 0; c0 < mine.size(); c0 += (size_t)L) {
            const int nb = (int)std::min<size_t>((size_t)L, mine.size() - c0);
            const double *ptsv[16];
            double fvv[16];
            for (int j = 0; j < nb; j++) ptsv[j] = batch[mine[c0 + j]].data();
            RC(eval(nb, ptsv, fvv));
            for (int j = 0; j < nb; j++) fl[mine[c0 + j]] = fvv[j];
        }
        if (world > 1 && !M->spec_emulate) RC(reduce_host(M, fl.data(), (int64_t)fl.size(), 0));
        if (dbg && M->rank == 0) {
            fprintf(stderr, "[spec] miss at %.10g (known %zu) batch:", x[0], pts.size());
            for (size_t k = 0; k < batch.size(); k++) fprintf(stderr, " %.10g->%.8g", batch[k][0], fl[k]);
            fprintf(stderr, "\n");
        }
        for (size_t k = 0; k < batch.size(); k++) { pts.push_back(batch[k]); fv.push_back(fl[k]); owner.push_back((int)(k % world)); }
        M->nm_evals += (double)batch.size();
        i = find(x);
        f = fv[i];
        RC(check_requested(x, f));
        return requested ? requested(x, f) : 0;
    }
};

// ---------------------------------------------------------------------------------
// E-step + log-likelihood
// ---------------------------------------------------------------------------------
// acc = [upper(S'D^-1 S) | S'D^-1 rho | rho'D^-1 rho | sum log d], summed over ranks
int gram(frk_model *M, const double *alpha, double sigma2fs) {
    const size_t n = (size_t)M->r * M->r + M->r + 2;
    if (M->gram_early) {              // computed beside the previous K update for exactly these parameters?
        M->gram_early = false;
        bool same = sigma2fs == M->gram_s2;
        for (int i = 0; i < M->p; i++) same = same && alpha[i] == M->gram_alpha[i];
        FRK_CHECK_CUDA(cudaStreamWaitEvent(M->ctx->stream, M->ev_aux1, 0));       // either way acc is not touched before aux is done
        if (same) return reduce_gram(M);                                            // (this rank's rows: the exchange happens here)
    }
    FRK_CHECK_CUDA(cudaMemsetAsync(M->acc, 0, n * sizeof(double), M->ctx->stream));
    if (M->cp) {
        // D = sigma2fs Vfs + Ve is block diagonal, not diagonal (R/SREfit.R:186-189,240-247): whiten with chol(D) per block, then the
        // same kernels on (S~, Z~, X~) with unit weights; logdet(cholD) goes into the slot the weights' sum of logs would fill
        frk_coupling *C = M->cp;
        RC(frk_coupling_whiten(C, M->ctx, sigma2fs, M->val, M->Z, M->X, M->Ve, M->acc + (size_t)M->r * M->r + M->r + 1));
        return frk_gram_rhs_bucketed(M->ctx, M->m, M->r, frk_coupling_rowptr(C), frk_coupling_col(C), frk_coupling_val(C),
                                     frk_coupling_buckets(C), frk_coupling_Z(C), frk_coupling_X(C), M->p, alpha, frk_coupling_ones(C),
                                     frk_coupling_zeros(C), 0.0, M->acc);
    }
    RC(frk_gram_rhs_bucketed(M->ctx, M->m, M->r, M->rowptr, M->col, M->val, M->bk,
                             M->Z, M->X, M->p, alpha, M->Ve, M->Vfs, sigma2fs, M->acc));
    return reduce_gram(M);                                            // the ONE large exchange per EM iteration
}

// Q = S'D^-1 S + Khat_inv; Sigma = chol2inv(chol(Q)); mu = Sigma S'D^-1 rho   (R/SREfit.R:252-254)
// out = {logdet(R) = 2 sum log diag chol Q, rho'D^-1 rho, sum log d, ||L^-1 S'D^-1 rho||^2}
int posterior(frk_model *M, const double *alpha, double sigma2fs, double *Q_out, double *Sig_out, double *mu_out, double out[4]) {
    frk_ctx *ctx = M->ctx;
    const int r = M->r;
    const size_t rr = (size_t)r * r;
    RC(gram(M, alpha, sigma2fs));
    const double *b = M->acc + rr;
    RC(frk_sym_from_upper_add(ctx, r, M->acc, M->Khat_inv, Q_out));
    FRK_CHECK_CUDA(cudaMemcpyAsync(M->w1, Q_out, rr * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    RC(frk_potrf(ctx, r, M->w1, &out[0]));
    if (Q_out == M->Q_eta) M->logdetQ = out[0];
    RC(frk_potri(ctx, r, M->w1, Sig_out, M->w2));
    if (Q_out == M->Q_eta && frk_rowbuckets_is_dense(M->bk)) {     // keep L^-1 beside the S_eta slot (w2 is scratch for the K update)
        if (!M->Linv) RC(dalloc(M, &M->Linv, rr));
        FRK_CHECK_CUDA(cudaMemcpyAsync(M->Linv, M->w2, rr * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        M->linv_valid = true;
    }
    RC(frk_symv(ctx, r, Sig_out, b, mu_out));
    // || (rDinv S) solve(R) ||^2 with solve(R)' = L^-1 left in w2 by potri   (R/SREfit.R:211)
    RC(frk_lower_mulv(ctx, r, M->w2, b, M->vec));
    RC(frk_dot(ctx, r, M->vec, M->vec, &out[3]));
    double sc[2];
    RC(frk_d2h(ctx, sc, M->acc + rr + r, 2 * sizeof(double)));
    out[1] = sc[0]; out[2] = sc[1];
    return 0;
}

double llk_from(const frk_model *M, const double o[4]) {
    const double log_det_SigmaZ = o[0] + M->logdetK + o[2];          // logdet(R) + log|K| + logdet(cholD)
    const double quad = o[1] - o[3];
    return -0.5 * M->m_total * std::log(2.0 * M_PI) - 0.5 * log_det_SigmaZ - 0.5 * quad;
}

int chol2inv_logdet(frk_ctx *ctx, int n, double *A, double *Ainv, double *work, double *logdet) {
    RC(frk_potrf(ctx, n, A, logdet));
    return frk_potri(ctx, n, A, Ainv, work);
}

// same, but a matrix that is not positive definite is reported through *bad (order of the failing minor) instead of an error
int chol2inv_logdet_soft(frk_ctx *ctx, int n, double *A, double *Ainv, double *work, double *logdet, int *bad) {
    RC(frk_potrf_async(ctx, n, A));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    *bad = frk_potrf_status(ctx, logdet);
    if (*bad) return 0;
    return frk_potri(ctx, n, A, Ainv, work);
}

// ---------------------------------------------------------------------------------
// Block-exponential K, plain bases: the per-resolution Nelder-Mead runs (optim(), R/SREfit.R:618-621) are independent, and every
// f_tau evaluation is a latency-bound r_i x r_i factorisation that fills a fraction of the GPU.  So the runs advance TOGETHER:
// an unfinished run is replayed against the values known so far until it asks for a point without a value; that point plus the
// points it may ask for next (SpecNM::candidates) are evaluated side by side on lane contexts of this GPU.  On one rank (or with
// the candidates replicated on every rank, the default) each run is fed as soon as its values are in; when the candidates are
// distributed over the ranks the runs move in lockstep and ONE all-reduce per round shares the values of all resolutions.
// Each run sees exactly the values the sequential algorithm would: the iterates are identical.
// ---------------------------------------------------------------------------------
using Model_PreSlot = frk_model::PreSlot;

struct NMSlot {                    // one concurrent evaluation: context + buffers, and what its Kinv currently holds
    frk_ctx *ctx = nullptr;
    double *Ki = nullptr, *Kinv = nullptr, *work = nullptr;
    double pt = 0.0, ld = 0.0;
    double tau0 = 0.0;             // the tau its Kinv was actually factorised for (== pt unless the inverse was reused, see f_tau_lookup)
    bool reused = false;           // Kinv is a copy of the block's saved inverse: no factorisation status to read, ld is known
    bool valid = false, used = false;
    int k = -1;
};

struct BlkNM {
    Block *B = nullptr;
    double omega = 0.0;
    int rounds = 0;
    cudaEvent_t ev_copy = nullptr;      // "the best inverse has been copied out of the slots" (recorded on slot 0's stream)
    ~BlkNM() { if (ev_copy) cudaEventDestroy(ev_copy); }
    double bestF = std::numeric_limits<double>::infinity(), bestTau = -1.0, bestLogdet = 0.0;
    bool done = false;
    NMResult nm;
    std::unique_ptr<SpecNM> spec;
    std::vector<std::vector<double>> batch;
    std::vector<double> fl;
    std::vector<NMSlot> slot;
};

// the tau-only part of an f_tau evaluation: K_i(tau), its factor and inverse on the slot's context (enqueue only; log-determinant and
// status land in the context's pinned staging buffer; on a non-PD matrix the inverse is garbage that is discarded when the status is read)
int f_tau_factor_enqueue(const Block &B, NMSlot &s, double tau) {
    const size_t nn = (size_t)B.ni * B.ni;
    RC(frk_exp_kernel(s.ctx, (int64_t)nn, B.D, tau, 1.0, s.Ki));
    RC(frk_potrf_async(s.ctx, B.ni, s.Ki));
    return frk_potri(s.ctx, B.ni, s.Ki, s.Kinv, s.work);
}

// ... and tr(K_i^-1 eta2_i), which needs the current E-step
int f_tau_enqueue(const Block &B, NMSlot &s, double tau) {
    RC(f_tau_factor_enqueue(B, s, tau));
    s.tau0 = tau; s.reused = false;
    return frk_dot_async(s.ctx, (int64_t)B.ni * B.ni, s.Kinv, B.eta2);
}

// Reuse of inverses.  K_i(tau)^-1 = exp(-D_i/tau)^-1 and log|K_i(tau)| depend on tau alone, so an evaluation whose tau was
// factorised before only needs the trace against the new E-step.  Two sources:
//  (1) KinvBest, the inverse the last K update kept (for the winner tau).  optim() starts every M-step from
//      tau' = -D01 / log(K_i[1,2] / K_i[1,1]) read back from that K (R/SREfit.R:596-614): tau' is tau up to the rounding of exp, the
//      division and log -- a few tens of ulps, often exactly tau.
//  (2) a small LRU memo of recent evaluations (Block::memo): once a resolution's tau has settled, the Nelder-Mead runs of
//      consecutive M-steps ask for the same points bit for bit (start, start + 10 %, ...).
// A hit needs |tau - tau0| <= 1e-13 tau0 with tau0 the tau that WAS factorised (so drift cannot accumulate): the reused inverse then
// agrees with exp(-D/tau)^-1 to ~1e-13 relative or exactly, below the rounding error of a factorisation of these matrices (condition
// 1e3 - 1e6).  At C3 this removes both r_i^3 evaluations of the 2916-block per M-step (50 of ~85 GFLOP per EM iteration).
// FRK_NO_KINV_REUSE=1 switches it off (A/B timing).
static bool kinv_reuse_on() { return getenv("FRK_NO_KINV_REUSE") == nullptr; }   // (read at every call: the tests switch it within a process)
static const double *f_tau_lookup(Block &B, double tau, double &tau0, double &ld) {
    if (!kinv_reuse_on() || B.ns != 0) return nullptr;
    auto near = [tau](double t0) { return t0 > 0.0 && std::fabs(tau - t0) <= 1e-13 * t0; };
    if (near(B.kinv_tau0)) { tau0 = B.kinv_tau0; ld = B.kinv_ld; return B.KinvBest; }
    for (auto &e : B.memo)
        if (near(e.tau0)) { e.stamp = ++B.memo_clock; tau0 = e.tau0; ld = e.ld; return e.buf; }
    return nullptr;
}
// the slot's Kinv := a saved inverse (the caller has ordered this stream after the last write of the source)
static int f_tau_reuse_enqueue(const Block &B, NMSlot &s, const double *src, double tau0, double ld) {
    FRK_CHECK_CUDA(cudaMemcpyAsync(s.Kinv, src, (size_t)B.ni * B.ni * sizeof(double), cudaMemcpyDeviceToDevice, s.ctx->stream));
    s.tau0 = tau0; s.ld = ld; s.reused = true;
    return 0;
}
// a freshly factorised, valid evaluation into the memo (least recently used entry); `stream` is the block's copy stream (slot 0):
// every lane of the block has been drained, and the lanes wait for that stream before they overwrite their buffers
static int memo_insert(Block &B, const NMSlot &s, cudaStream_t stream) {
    if (!kinv_reuse_on() || B.ns != 0 || B.memo.empty() || s.reused || !s.valid) return 0;
    Block::Memo *lru = &B.memo[0];
    for (auto &e : B.memo) {
        if (e.tau0 == s.tau0) return 0;
        if (e.stamp < lru->stamp) lru = &e;
    }
    FRK_CHECK_CUDA(cudaMemcpyAsync(lru->buf, s.Kinv, (size_t)B.ni * B.ni * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    lru->tau0 = s.tau0; lru->ld = s.ld; lru->stamp = ++B.memo_clock;
    return 0;
}

constexpr int LCAP = 8;            // slot (block j, lane c) -> lane context j*LCAP + c - 1: stable, so graph caches hit

// context and buffers of evaluation slot (block j, lane c)
int nm_slot_setup(frk_model *M, int j, int c, Block &B, bool one_stream, NMSlot &s) {
    frk_ctx *ctx = M->ctx;
    const size_t nn = (size_t)B.ni * B.ni;
    const int q = j * LCAP + c;
    if (one_stream || q == 0) {
        s.ctx = ctx; s.Ki = B.Ki; s.Kinv = B.Kinv; s.work = B.work;
    } else {
        // a lane context serves one (resolution, lane) pair and keeps its buffers in its own grow-only scratch: the
        // pointers -- hence the cached CUDA graphs -- survive from one model to the next
        RC(frk_ctx_lane(ctx, q - 1, true, &s.ctx));
        void *buf = nullptr;
        RC(frk_scratch(s.ctx, 3 * nn * sizeof(double), &buf));
        s.Ki = (double *)buf; s.Kinv = s.Ki + nn; s.work = s.Kinv + nn;
    }
    return 0;
}

// Lanes of block j: the configured number for the large resolutions; four for blocks of <= 1024 functions, whose factorisations are
// short latency-bound chains (0.3 ms a round at n_i = 324, and 5-7 rounds per M-step while their tau is still moving): deeper
// speculation turns 5 rounds into 4.  Measured at C3 (5 + 20 iterations): 5.63 -> 5.41 ms per EM iteration; 8 lanes evaluate twice as
// many points for one round less and are slower.  (While the large block's own factorisations still ran beside the E-step -- before
// the memo of f_tau_lookup -- the extra evaluations cost more than they saved; FRK_SMALL_LANES overrides.)
inline int block_lanes(const frk_model *M, const Block &B, bool one_stream) {
    if (one_stream) return 1;
    const int L = std::max(M->lanes, 1);
    if (M->spec_emulate || L == 1) return L;                        // (L = 1: the strictly sequential reference run of the tests)
    static const int small_lanes = getenv("FRK_SMALL_LANES") ? std::min(std::max(atoi(getenv("FRK_SMALL_LANES")), 1), LCAP) : 4;
    return B.ni <= 1024 ? std::max(small_lanes, L) : L;
}

// Enqueue, on the lanes, the factorisations the next M-step will need first (see frk_model::pre).  Called when K has just been
// assembled: the start value of optim() is read from the new K exactly as the next M-step will read it (R/SREfit.R:596-614), the
// second point is the one nmmin builds its initial simplex with (start + 0.1 |start|).  A wrong guess costs nothing but lane time:
// the next M-step compares tau bit for bit and evaluates normally if it differs.
int prefetch_f_tau(frk_model *M) {
    frk_ctx *ctx = M->ctx;
    const int nblk = (int)M->blocks.size();
    M->pre.assign(nblk, {});
    // candidate k of a round goes to rank k % world, lane k / world (blockexp_nelder_mead): with the candidates distributed over the
    // ranks every rank prefetches its own share of the first two points
    const int world = nm_world(M), rank = nm_rank(M);
    const bool ok = M->prefetch && !M->tensor && world * std::max(M->lanes, 1) >= 2 && !M->spec_emulate && !ctx->profiling &&
                    ctx->stream != nullptr && ctx->use_graphs;
    if (!ok) return 0;
    for (int j = 0; j < nblk; j++) {
        Block &B = M->blocks[j];
        if (B.ni < 128) continue;                                   // tiny blocks: nothing to hide
        const int L = block_lanes(M, B, false);
        double k2[2] = {0, 0};
        RC(frk_d2h(ctx, k2, B.Ki, 2 * sizeof(double)));             // K_i[1,1], K_i[2,1] of the K just assembled (B.Ki = exp(-D/tau)/omega)
        const double k00 = k2[0], k01 = k2[1];
        double par0 = 1e-9;
        { const double v = -B.D01 / std::log(k01 / k00); par0 = (v > 1e-9) ? v : 1e-9; }
        if (!std::isfinite(par0) || par0 == 1e-9) par0 = M->max_l / 10.0;
        // the batch of the first round, exactly as blockexp_nelder_mead will build it: the start value, then the points nmmin may
        // ask for next under the three-outcome hypotheses (the second simplex vertex first; with more than two evaluators also the
        // reflections that follow either outcome) -- SpecNM::candidates is deterministic and needs no function value for this
        SpecNM spec{M, nullptr, 1, 100, 1e-4, std::vector<double>{par0}};
        std::vector<std::vector<double>> batch = {std::vector<double>{par0}};
        spec.candidates(&par0, (size_t)(world * L), batch);
        for (int c = 0; c < std::min(L, LCAP); c++) {
            const int k = c * world + rank;                         // the candidate this lane of this rank will be given
            if (k >= (int)batch.size()) continue;
            if (j * LCAP + c == 0) continue;                        // that slot is the model's own stream and buffers
            const double tau = batch[k][0];
            if (!(tau > 1e-10)) continue;
            NMSlot s;
            RC(nm_slot_setup(M, j, c, B, false, s));
            Model_PreSlot &pf = M->pre[j][c];
            double tau0 = 0.0, ld = 0.0;
            if (const double *src = f_tau_lookup(B, tau, tau0, ld)) {  // factorised before: a copy serves (see f_tau_lookup)
                if (!M->ev_lane) FRK_CHECK_CUDA(cudaEventCreateWithFlags(&M->ev_lane, cudaEventDisableTiming));
                FRK_CHECK_CUDA(cudaEventRecord(M->ev_lane, ctx->stream));       // the saved inverses are final once this stream has assembled K
                FRK_CHECK_CUDA(cudaStreamWaitEvent(s.ctx->stream, M->ev_lane, 0));
                RC(f_tau_reuse_enqueue(B, s, src, tau0, ld));
                pf.reused = true; pf.tau0 = tau0; pf.ld = ld;
            } else {
                RC(f_tau_factor_enqueue(B, s, tau));
                pf.reused = false; pf.tau0 = tau;
            }
            pf.on = true;
            pf.tau = tau;
        }
    }
    return 0;
}

int blockexp_nelder_mead(frk_model *M, std::vector<std::unique_ptr<BlkNM>> &nmb, double *Knew, double *Kinvnew, double &logdetK,
                      const std::function<int()> &side_work) {
    frk_ctx *ctx = M->ctx;
    const int r = M->r, nblk = (int)nmb.size();
    const int world = nm_world(M);
    const bool one_stream = ctx->profiling || ctx->stream == nullptr;   // per-kernel profiling times one stream: keep everything on it
    const int Lmax = one_stream ? 1 : std::max(M->lanes, 1);        // (the pretend-world walk of spec_emulate uses this one for every block)
    const bool shared = world > 1 && !M->spec_emulate;
    if ((int)M->pre.size() != nblk) M->pre.assign(nblk, {});
    if (!M->ev_lane) FRK_CHECK_CUDA(cudaEventCreateWithFlags(&M->ev_lane, cudaEventDisableTiming));

    for (int j = 0; j < nblk; j++) {
        BlkNM &S = *nmb[j];
        Block &B = *S.B;
        const size_t nn = (size_t)B.ni * B.ni;
        const int L = block_lanes(M, B, one_stream);
        S.slot.resize(L);
        for (int c = 0; c < L; c++) RC(nm_slot_setup(M, j, c, B, one_stream, S.slot[c]));
        (void)nn;
        // Nelder-Mead returns its lowest vertex, a REQUESTED point (a speculative point that is never requested cannot be the result):
        // keep the inverse of the best requested point while it is still in a slot (slots are drained before any replay; the next
        // round starts after this copy)
        if (!S.ev_copy) FRK_CHECK_CUDA(cudaEventCreateWithFlags(&S.ev_copy, cudaEventDisableTiming));
        // a prefetched COPY of a saved inverse on another lane reads a buffer that this run's first replay may overwrite (on slot 0's
        // stream): order slot 0's stream after it (the copy finished long ago; this only makes the order explicit)
        if (!one_stream)
            for (int c = 1; c < L; c++)
                if (M->pre[j][c].on && M->pre[j][c].reused && S.slot[c].ctx != S.slot[0].ctx) {
                    FRK_CHECK_CUDA(cudaEventRecord(S.ev_copy, S.slot[c].ctx->stream));
                    FRK_CHECK_CUDA(cudaStreamWaitEvent(S.slot[0].ctx->stream, S.ev_copy, 0));
                }
        BlkNM *Sp = &S;
        S.spec->requested = [Sp](const double *x, double f) -> int {
            if (!(f < Sp->bestF)) return 0;
            Sp->bestF = f; Sp->bestTau = -1.0;
            const size_t nn2 = (size_t)Sp->B->ni * Sp->B->ni;
            for (NMSlot &s : Sp->slot) {
                if (!s.valid || s.pt != x[0]) continue;
                FRK_CHECK_CUDA(cudaMemcpyAsync(Sp->B->KinvBest, s.Kinv, nn2 * sizeof(double), cudaMemcpyDeviceToDevice,
                                               Sp->slot[0].ctx->stream));
                Sp->bestTau = x[0]; Sp->bestLogdet = s.ld;
                Sp->B->kinv_tau0 = s.tau0; Sp->B->kinv_ld = s.ld;     // (host-side tags of KinvBest, updated in enqueue order like the buffer)
                break;
            }
            return 0;
        };
    }

    // ---- building blocks of a round, per resolution j ----
    // what does run j ask for next?  false: it has finished
    auto begin_round = [&](int j, bool &active) -> int {
        BlkNM &S = *nmb[j];
        S.batch.clear();
        S.fl.clear();
        active = false;
        if (S.done) return 0;
        std::vector<double> need;
        const int rc = S.spec->replay(need, S.nm);        // (also saves the best requested inverse out of the slots, on slot 0's stream)
        if (rc == 0) { S.done = true; return 0; }
        if (rc != 99) return rc;
        S.batch.push_back(need);
        const int W = world * (int)S.slot.size();
        if (W > 1) S.spec->candidates(need.data(), (size_t)W, S.batch);
        S.fl.assign(S.batch.size(), 0.0);
        S.rounds += 1;
        active = true;
        return 0;
    };
    // enqueue this rank's share of run j's batch: candidate k -> rank k % world, lane k / world (a pretend world is walked in chunks)
    auto enqueue = [&](int j, int chunk) -> int {
        BlkNM &S = *nmb[j];
        // the slots' buffers may be overwritten once the copies of the last replay (slot 0's stream) have been made
        FRK_CHECK_CUDA(cudaEventRecord(S.ev_copy, S.slot[0].ctx->stream));
        const int L = (int)S.slot.size();
        for (int c = 0; c < L; c++) {
            NMSlot &s = S.slot[c];
            s.used = false;
            const int k = M->spec_emulate ? chunk * L + c : c * world + nm_rank(M);
            if (k >= (int)S.batch.size()) continue;
            s.k = k; s.valid = false;
            const double tau = S.batch[k][0];
            if (tau <= 1e-10) { S.fl[k] = std::numeric_limits<double>::infinity(); continue; }
            if (c > 0) FRK_CHECK_CUDA(cudaStreamWaitEvent(s.ctx->stream, S.ev_copy, 0));
            s.used = true; s.pt = tau;
            Model_PreSlot &pf = M->pre[j][c];
            double rt0 = 0.0, rld = 0.0;
            if (pf.on && pf.tau == tau && !one_stream && !M->spec_emulate) {
                // this slot's buffers already hold K_i(tau)^-1 (or it is on its way): enqueued before the E-step.  Only the trace is left.
                s.reused = pf.reused; s.tau0 = pf.tau0;
                if (pf.reused) { s.ld = pf.ld; M->nm_reuse_hits += 1; }
                RC(frk_dot_async(s.ctx, (int64_t)S.B->ni * S.B->ni, s.Kinv, S.B->eta2));
                M->nm_prefetch_hits += 1;
            } else if (const double *src = f_tau_lookup(*S.B, tau, rt0, rld)) {   // (this stream is ordered after every write of the saved inverses: ev_copy / ev_lane)
                RC(f_tau_reuse_enqueue(*S.B, s, src, rt0, rld));
                RC(frk_dot_async(s.ctx, (int64_t)S.B->ni * S.B->ni, s.Kinv, S.B->eta2));
                M->nm_reuse_hits += 1;
            } else {
                RC(f_tau_enqueue(*S.B, s, tau));
            }
            pf.on = false;
        }
        for (int c = L; c < 8; c++) M->pre[j][c].on = false;
        return 0;
    };
    auto finished = [&](int j) -> bool {                 // non-blocking: have all of run j's evaluations completed?
        for (NMSlot &s : nmb[j]->slot)
            if (s.used && cudaStreamQuery(s.ctx->stream) == cudaErrorNotReady) return false;
        return true;                                     // completed -- or failed, which drain() reports (never spin on an error)
    };
    auto drain = [&](int j) -> int {
        int rc = 0;
        for (NMSlot &s : nmb[j]->slot) {
            if (!s.used) continue;
            if (cudaStreamSynchronize(s.ctx->stream) != cudaSuccess && !rc) { frk_set_error("f_tau lane: stream synchronisation failed"); rc = 1; }
            if (s.ctx != ctx) { ctx->launches += s.ctx->launches; s.ctx->launches = 0; }
        }
        return rc;
    };
    auto collect = [&](int j) -> int {                   // after drain: read the values out of the contexts' pinned staging
        BlkNM &S = *nmb[j];
        for (NMSlot &s : S.slot) {
            if (!s.used) continue;
            s.used = false;
            double logdetKi;
            // a candidate that cannot be factorised is only a failed guess unless Nelder-Mead asks for it (SpecNM::check_requested)
            const int bad = s.reused ? 0 : frk_potrf_status(s.ctx, &logdetKi);
            if (s.reused) logdetKi = s.ld;
            const double t = s.ctx->h_pinned[FRK_PINNED_DOT];
            const double f = bad ? std::numeric_limits<double>::quiet_NaN() : -(0.5 * (-logdetKi) - S.omega / 2.0 * t);
            S.fl[s.k] = f; s.ld = logdetKi; s.valid = std::isfinite(f);
            RC(memo_insert(*S.B, s, S.slot[0].ctx->stream));
        }
        return 0;
    };
    auto commit = [&](int j) {                           // the round's values into run j's cache
        BlkNM &S = *nmb[j];
        for (size_t k = 0; k < S.batch.size(); k++) {
            S.spec->pts.push_back(S.batch[k]); S.spec->fv.push_back(S.fl[k]); S.spec->owner.push_back((int)(k % world));
        }
        M->nm_evals += (double)S.batch.size();
    };

    const double tr0 = now_s();
    // every lane starts once this stream has produced eta2 of all resolutions
    FRK_CHECK_CUDA(cudaEventRecord(M->ev_lane, ctx->stream));
    for (int j = 0; j < nblk; j++)
        for (NMSlot &s : nmb[j]->slot)
            if (s.ctx != ctx) FRK_CHECK_CUDA(cudaStreamWaitEvent(s.ctx->stream, M->ev_lane, 0));

    if (!shared && !M->spec_emulate && !one_stream) {
        // One rank: the runs advance INDEPENDENTLY -- a small resolution goes through all of its rounds while one evaluation of the
        // largest is in flight.  The host polls the lanes and feeds whichever run has its values.
        std::vector<char> inflight(nblk, 0);
        int live = 0, rc = 0;
        for (int j = 0; j < nblk && !rc; j++) {
            bool active = false;
            rc = begin_round(j, active);
            if (!rc && active) { rc = enqueue(j, 0); inflight[j] = 1; live++; }
        }
        // with the first evaluations in flight the host is free: the rest of the M-step (sigma2fs, alpha) and the next iteration's
        // Gram matrix are driven from here, on their own low-priority lane
        if (!rc && side_work) rc = side_work();
        while (live > 0 && !rc) {
            bool progressed = false;
            for (int j = 0; j < nblk && !rc; j++) {
                if (!inflight[j] || !finished(j)) continue;
                progressed = true;
                inflight[j] = 0; live--;
                if ((rc = drain(j)) || (rc = collect(j))) break;
                commit(j);
                bool active = false;
                rc = begin_round(j, active);
                if (!rc && active) { rc = enqueue(j, 0); inflight[j] = 1; live++; }
            }
            if (!progressed) std::this_thread::yield();
        }
        for (int j = 0; j < nblk; j++) if (inflight[j]) drain(j);     // error path: leave nothing running
        if (rc) return rc;
        for (int j = 0; j < nblk; j++) M->nm_rounds = std::max(M->nm_rounds, (double)nmb[j]->rounds);
    } else {
        // Candidates distributed over the ranks (or a pretend world): lockstep, so that every rank issues the same all-reduces in the same order
        for (;;) {
            size_t maxbatch = 0;
            for (int j = 0; j < nblk; j++) {
                bool active = false;
                RC(begin_round(j, active));
                if (active) maxbatch = std::max(maxbatch, nmb[j]->batch.size());
            }
            if (maxbatch == 0) break;
            M->nm_rounds += 1;
            const int nchunks = M->spec_emulate ? (int)((maxbatch + Lmax - 1) / Lmax) : 1;
            for (int ch = 0; ch < nchunks; ch++) {
                int rc = 0;
                if (one_stream) {      // every slot is this context (one pinned staging buffer): one resolution at a time
                    for (int j = 0; j < nblk && !rc; j++) {
                        if (nmb[j]->batch.empty()) continue;
                        rc = enqueue(j, ch);
                        const int rd = drain(j);
                        if (!rc) rc = rd;
                        if (!rc) rc = collect(j);
                    }
                    if (rc) return rc;
                    continue;
                }
                for (int j = 0; j < nblk && !rc; j++) if (!nmb[j]->batch.empty()) rc = enqueue(j, ch);
                for (int j = 0; j < nblk; j++) { const int rd = drain(j); if (!rc) rc = rd; }    // drained error or not
                if (rc) return rc;
                for (int j = 0; j < nblk; j++) RC(collect(j));
            }
            if (shared) {     // the values of all resolutions in one all-reduce (a value is set on its owner and 0 elsewhere)
                std::vector<double> all;
                for (int j = 0; j < nblk; j++) all.insert(all.end(), nmb[j]->fl.begin(), nmb[j]->fl.end());
                RC(reduce_host(M, all.data(), (int64_t)all.size(), 0));
                for (int j = 0, o = 0; j < nblk; j++)
                    for (size_t k = 0; k < nmb[j]->fl.size(); k++, o++) nmb[j]->fl[k] = all[o];
            }
            for (int j = 0; j < nblk; j++) commit(j);
        }
    }
    // the saved inverses (slot 0's stream of each resolution) and the slots themselves are consumed on this context's stream below
    for (int j = 0; j < nblk; j++) {
        FRK_CHECK_CUDA(cudaEventRecord(nmb[j]->ev_copy, nmb[j]->slot[0].ctx->stream));
        FRK_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, nmb[j]->ev_copy, 0));
    }
    M->t_phase[3] += now_s() - tr0;
    static const bool rounds_dbg = getenv("FRK_PHASE_DEBUG") != nullptr;
    if (rounds_dbg && M->rank == 0) {
        fprintf(stderr, "[frk nm] %.3f ms;", 1e3 * (now_s() - tr0));
        for (int j = 0; j < nblk; j++) fprintf(stderr, " block %d (n=%d): %d rounds, %zu points;", j, nmb[j]->B->ni, nmb[j]->rounds, nmb[j]->spec->pts.size());
        fprintf(stderr, "\n");
    }
    // K_i = exp(-D_i/tau_i)/omega_i and its inverse into place, :625-637
    for (int j = 0; j < nblk; j++) {
        BlkNM &S = *nmb[j];
        Block &B = *S.B;
        const int ni = B.ni;
        const size_t nn = (size_t)ni * ni;
        const double tau = S.nm.x[0], omega = S.omega;
        M->taus.push_back(tau);
        const int own = shared ? S.spec->owner[S.spec->find(&tau)] : nm_rank(M);
        if (own == nm_rank(M)) {
            if (S.bestTau != tau) {        // its inverse is no longer in a slot (ties, or a point evaluated rounds before it was requested)
                S.bestF = std::numeric_limits<double>::infinity();
                NMSlot &s = S.slot[0];
                s.valid = false;
                RC(f_tau_enqueue(B, s, tau));
                FRK_CHECK_CUDA(cudaStreamSynchronize(s.ctx->stream));
                if (s.ctx != ctx) { ctx->launches += s.ctx->launches; s.ctx->launches = 0; }
                double logdetKi;
                RC(frk_potrf_finish(s.ctx, &logdetKi));
                const double f = -(0.5 * (-logdetKi) - omega / 2.0 * s.ctx->h_pinned[FRK_PINNED_DOT]);
                s.pt = tau; s.ld = logdetKi; s.valid = std::isfinite(f);
                RC(S.spec->requested(&tau, f));
                FRK_CHECK_CUDA(cudaStreamSynchronize(s.ctx->stream));            // (the copy above went to this slot's stream)
                M->nm_reevals += 1;
                FRK_REQUIRE(S.bestTau == tau, "block-exponential update: tau = %g cannot be evaluated", tau);
            }
        } else {
            S.bestLogdet = 0.0;
        }
        if (shared) {                                                // from the owner to everybody
            RC(bcast_dev(M, B.KinvBest, (int64_t)nn, own, own == nm_rank(M)));
            const bool mine = own == nm_rank(M);
            double tags[3] = {S.bestLogdet, mine ? B.kinv_tau0 : 0.0, mine ? B.kinv_ld : 0.0};
            RC(reduce_host(M, tags, 3, 0));
            S.bestLogdet = tags[0]; B.kinv_tau0 = tags[1]; B.kinv_ld = tags[2];
        }
        RC(frk_exp_kernel(ctx, (int64_t)nn, B.D, tau, 1.0 / omega, B.Ki));
        RC(frk_block_scatter(ctx, r, Knew, ni, B.d_ix, B.Ki));
        RC(frk_block_scatter_scaled(ctx, r, Kinvnew, ni, B.d_ix, B.KinvBest, omega));   // K_i^-1 = omega_i exp(-D_i/tau_i)^-1 (KinvBest stays unscaled)
        logdetK += S.bestLogdet - ni * std::log(omega);
    }
    return 0;
}

// ---------------------------------------------------------------------------------
// M-step
// ---------------------------------------------------------------------------------
int update_K(frk_model *M, int K_type, int nlambda, const double *lambda, const int *h_res, bool &have_inverse,
             const std::function<int()> &side_work) {
    have_inverse = false;
    frk_ctx *ctx = M->ctx;
    const int r = M->r;
    const size_t rr = (size_t)r * r;
    if (K_type == FRK_K_UNSTRUCTURED) {
        RC(frk_add_outer(ctx, r, M->S_eta, M->mu_eta, M->Khat));     // K = S_eta + mu mu', R/SREfit.R:689
        bool any = false;
        for (int i = 0; i < nlambda; i++) any = any || lambda[i] > 0;
        if (any) {                                                   // K = ((K)^-1 + Lambda)^-1, :669-686
            std::vector<double> reg(r);
            for (int j = 0; j < r; j++) reg[j] = nlambda == 1 ? lambda[0] : lambda[(h_res ? h_res[j] : 1) - 1];
            RC(frk_h2d(ctx, M->vec, reg.data(), r * sizeof(double)));
            double ld;
            RC(chol2inv_logdet(ctx, r, M->Khat, M->w1, M->w2, &ld));
            RC(frk_add_diag(ctx, r, M->vec, M->w1));
            RC(chol2inv_logdet(ctx, r, M->w1, M->Khat, M->w2, &ld));
            return 0;
        }
        // Khat_inv = chol2inv(chol(K)) (:273) without a factorisation: S_eta = Q_eta^-1 was just computed by the E-step, so by
        // Sherman-Morrison (S_eta + mu mu')^-1 = Q - (Q mu)(Q mu)' / (1 + mu'Q mu) and log|K| = -log|Q| + log(1 + mu'Q mu).
        double s1;
        RC(frk_symv(ctx, r, M->Q_eta, M->mu_eta, M->vec));
        RC(frk_dot(ctx, r, M->vec, M->mu_eta, &s1));
        RC(frk_add_scaled_outer(ctx, r, M->Q_eta, M->vec, -1.0 / (1.0 + s1), M->Khat_inv));
        M->logdetK = -M->logdetQ + std::log1p(s1);
        M->khat_block_diag = false;
        have_inverse = true;
        return 0;
    }
    FRK_REQUIRE(!M->blocks.empty(), "K_type = \"block-exponential\" needs the per-resolution distance matrices (frk_model_set_blocks)");
    // K_new is assembled in w1 and K_new^-1 in w2 (both free here); the old Khat / Khat_inv are read for omega and the start values
    M->taus.clear(); M->omegas.clear();
    M->nm_requests = M->nm_rounds = M->nm_evals = M->nm_reevals = M->nm_prefetch_hits = M->nm_reuse_hits = 0;
    double *Knew = M->w1, *Kinvnew = M->w2;
    FRK_CHECK_CUDA(cudaMemsetAsync(Knew, 0, rr * sizeof(double), ctx->stream));
    FRK_CHECK_CUDA(cudaMemsetAsync(Kinvnew, 0, rr * sizeof(double), ctx->stream));
    double logdetK = 0.0;
    std::vector<std::unique_ptr<BlkNM>> nmb;     // per-resolution Nelder-Mead state (plain bases)
    for (Block &B : M->blocks) {
        const int ni = B.ni;
        const size_t nn = (size_t)ni * ni;
        // eta2_i = (S_eta + mu mu')[idx, idx]                                                      :459-464
        RC(frk_block_gather_outer(ctx, r, M->S_eta, M->mu_eta, ni, B.d_ix, B.eta2));
        // omega_i = n_i / tr(C_i^-1 eta2_i),  C_i = Khat[idx,idx] / Khat[idx1,idx1]                 :466-482
        RC(frk_block_gather(ctx, r, M->Khat, ni, B.d_ix, B.Ki));
        double k2[2] = {0, 0};
        RC(frk_d2h(ctx, k2, B.Ki, sizeof(double) * std::min<size_t>(2, nn)));
        const double k00 = k2[0], k01 = ni > 1 ? k2[1] : 0.0;
        double ld, trc;
        if (M->khat_block_diag) {
            // Khat is block diagonal, so (Khat[idx,idx])^-1 is the idx-block of Khat_inv = chol2inv(chol(Khat)) already held
            RC(frk_block_gather(ctx, r, M->Khat_inv, ni, B.d_ix, B.Kinv));
        } else {
            RC(chol2inv_logdet(ctx, ni, B.Ki, B.Kinv, B.work, &ld));
        }
        RC(frk_dot(ctx, (int64_t)nn, B.Kinv, B.eta2, &trc));
        const double omega = ni / (k00 * trc);                       // (K/k00)^-1 = k00 K^-1
        M->omegas.push_back(omega);
        if (M->tensor) {
            // f_tau(tau_s, tau_t), :496-514: K_i^-1 = kronecker(Kt^-1, Ks^-1), log|K_i| = ns log|Kt| + nt log|Ks|
            const int ns = B.ns, nt = B.nt;
            const size_t nss = (size_t)ns * ns, ntt = (size_t)nt * nt;
            double bestF = std::numeric_limits<double>::infinity(), bestTs = -1.0, bestTt = -1.0, bestLd = 0.0;
            auto f_tau2 = [&](const double *tau, double &fval) -> int {
                if (tau[0] <= 1e-10 || tau[1] <= 1e-10) { fval = std::numeric_limits<double>::infinity(); return 0; }
                double ldt, lds, trk;
                int bad = 0;
                RC(frk_exp_kernel(ctx, (int64_t)ntt, M->Dt, tau[1], 1.0, B.Kt));
                RC(chol2inv_logdet_soft(ctx, nt, B.Kt, B.Qt, B.wt, &ldt, &bad));
                if (bad) { fval = std::numeric_limits<double>::quiet_NaN(); return 0; }   // an error only if requested (SpecNM)
                RC(frk_exp_kernel(ctx, (int64_t)nss, B.D, tau[0], 1.0, B.Ks));
                RC(chol2inv_logdet_soft(ctx, ns, B.Ks, B.Qs, B.ws, &lds, &bad));
                if (bad) { fval = std::numeric_limits<double>::quiet_NaN(); return 0; }
                RC(frk_kron_dot(ctx, nt, ns, B.Qt, B.Qs, B.eta2, &trk));
                const double det_part = 0.5 * (ns * (-ldt) + nt * (-lds));
                fval = -(det_part - omega / 2.0 * trk);
                if (fval < bestF) {
                    bestF = fval; bestTs = tau[0]; bestTt = tau[1]; bestLd = ns * ldt + nt * lds;
                    FRK_CHECK_CUDA(cudaMemcpyAsync(B.QsBest, B.Qs, nss * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
                    FRK_CHECK_CUDA(cudaMemcpyAsync(B.QtBest, B.Qt, ntt * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
                }
                return 0;
            };
            // start values, :596-617: Ki[1,2] (space) and Ki[1, 1 + ns] (time) of the current correlation matrix
            double kt = 0.0;
            RC(frk_block_gather(ctx, r, M->Khat, ni, B.d_ix, B.Ki));           // (Kinv above may have overwritten nothing of Ki, but be explicit)
            RC(frk_d2h(ctx, &kt, B.Ki + (size_t)ns * ni, sizeof(double)));       // element (1, 1 + ns)
            double p0 = 1e-9, p1;
            if (ns > 1) { const double v = -B.D01 / std::log(k01 / k00); p0 = (v > 1e-9) ? v : 1e-9; }
            { const double v = -M->Dt01 / std::log(kt / k00); p1 = (v > 1e-9) ? v : 1e-9; }
            if (!std::isfinite(p1) || p1 == 1e-9) p1 = 1.0;
            if (!std::isfinite(p0) || p0 == 1e-9) p0 = M->max_l / 10.0;
            NMResult nm;
            auto f_tau2_batch = [&](int nb, const double *const *pts, double *fv) -> int {
                for (int j = 0; j < nb; j++) RC(f_tau2(pts[j], fv[j]));
                return 0;
            };
            SpecNM spec{M, f_tau2_batch, 2, 100, 1e-4, std::vector<double>{p0, p1}};
            RC(nmmin(std::ref(spec), spec.par0, nm, 100, 1e-4));
            M->taus.push_back(nm.x[0]); M->taus.push_back(nm.x[1]);
            const bool shared = nm_world(M) > 1 && !M->spec_emulate;
            const int own = shared ? spec.owner[spec.find(nm.x.data())] : nm_rank(M);
            if (own == nm_rank(M)) {
                if (bestTs != nm.x[0] || bestTt != nm.x[1]) {
                    bestF = std::numeric_limits<double>::infinity();
                    double f;
                    RC(f_tau2(nm.x.data(), f));
                    FRK_REQUIRE(bestTs == nm.x[0] && bestTt == nm.x[1], "block-exponential update: tau cannot be evaluated");
                }
            } else {
                bestLd = 0.0;
            }
            if (shared) {
                RC(bcast_dev(M, B.QsBest, (int64_t)nss, own, own == nm_rank(M)));
                RC(bcast_dev(M, B.QtBest, (int64_t)ntt, own, own == nm_rank(M)));
                RC(reduce_host(M, &bestLd, 1, 0));
            }
            RC(frk_exp_kernel(ctx, (int64_t)ntt, M->Dt, nm.x[1], 1.0, B.Kt));
            RC(frk_exp_kernel(ctx, (int64_t)nss, B.D, nm.x[0], 1.0, B.Ks));
            RC(frk_kron_scatter(ctx, r, nt, ns, B.Kt, B.Ks, 1.0 / omega, B.d_ix, Knew));         // :627-649
            RC(frk_kron_scatter(ctx, r, nt, ns, B.QtBest, B.QsBest, omega, B.d_ix, Kinvnew));
            logdetK += bestLd - ni * std::log(omega);
            continue;
        }
        // f_tau, :485-529: deferred -- the Nelder-Mead runs of all resolutions advance together below
        double par0 = 1e-9;                                          // :608-614
        if (ni > 1) {
            const double v = -B.D01 / std::log(k01 / k00);
            par0 = (v > 1e-9) ? v : 1e-9;                            // max(v, 1e-9); NaN -> 1e-9
        }
        if (!std::isfinite(par0) || par0 == 1e-9) par0 = M->max_l / 10.0;
        nmb.emplace_back(new BlkNM());
        BlkNM &S = *nmb.back();
        S.B = &B; S.omega = omega;
        S.spec.reset(new SpecNM{M, nullptr, 1, 100, 1e-4, std::vector<double>{par0}});
    }
    if (!M->tensor) RC(blockexp_nelder_mead(M, nmb, Knew, Kinvnew, logdetK, side_work));
    FRK_CHECK_CUDA(cudaMemcpyAsync(M->Khat, Knew, rr * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    FRK_CHECK_CUDA(cudaMemcpyAsync(M->Khat_inv, Kinvnew, rr * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    M->logdetK = logdetK;
    M->khat_block_diag = true;
    have_inverse = true;                                             // Khat_inv = chol2inv(chol(K)) (:273) block by block
    RC(prefetch_f_tau(M));                                           // the next M-step's first factorisations start now, beside the next E-step
    return 0;
}

int mstep(frk_model *M, int K_type, int nlambda, const double *lambda, const int *h_res, int known, double known_sigma2fs) {
    frk_ctx *ctx = M->ctx;
    const int r = M->r, p = M->p;
    const size_t rr = (size_t)r * r;
    const double sigma2fs = M->sigma2fs;
    bool have_inverse = false;
    // Omega_diag1 = rowSums((S R_eta')^2) = s_j'(S_eta + mu mu')s_j = q_j + t_j^2,  t = S mu_eta     :332-334
    // It reads only the E-step's S_eta / mu_eta, so it is enqueued on a low-priority lane first and fills the SMs
    // the block-exponential Nelder-Mead (one latency-bound factorisation chain after another) leaves idle.
    const bool beside = K_type == FRK_K_BLOCK_EXPONENTIAL && !ctx->profiling && ctx->stream != nullptr;
    if (beside) {
        if (!M->aux) {
            RC(frk_ctx_lane(ctx, 0, false, &M->aux));
            FRK_CHECK_CUDA(cudaEventCreateWithFlags(&M->ev_aux0, cudaEventDisableTiming));
            FRK_CHECK_CUDA(cudaEventCreateWithFlags(&M->ev_aux1, cudaEventDisableTiming));
        }
        FRK_CHECK_CUDA(cudaEventRecord(M->ev_aux0, ctx->stream));
        FRK_CHECK_CUDA(cudaStreamWaitEvent(M->aux->stream, M->ev_aux0, 0));
        if (M->linv_valid && frk_rowbuckets_is_dense(M->bk))        // dense rows: q_j = |L^-1 s_j|^2, half the flops of S_eta s_j
            RC(frk_quadform_dense_factor(M->aux, M->m, r, M->val, M->bk, M->Linv, M->mu_eta, M->q, M->t));
        else
        {   // (runs beside the Nelder-Mead lanes: measured 1 % per EM iteration in favour of the kernel with the smaller footprint at C3)
            M->aux->quad_ws_min = 24;
            RC(frk_quadform_spmv_bucketed(M->aux, M->m, r, M->rowptr, M->col, M->val, M->bk, M->S_eta, M->mu_eta, M->q, M->t));
        }
        FRK_CHECK_CUDA(cudaEventRecord(M->ev_aux1, M->aux->stream));
        ctx->launches += M->aux->launches;
        M->aux->launches = 0;
    }
    const bool est = !known;
    bool sigma_done = false;
    // the early Gram needs this function to run BEFORE the K update finishes, i.e. as the asynchronous driver's side work
    const bool early_gram = beside && !M->tensor && nm_world(M) == 1 && !M->spec_emulate && !M->cp;
    // sigma2fs and alpha (:275-430).  Everything here depends on the E-step only, not on K: with `beside` it runs on the aux lane
    // (after the quadratic form) while the Nelder-Mead evaluations are in flight, and is followed there by the NEXT iteration's Gram.
    auto sigma_update = [&]() -> int {
        sigma_done = true;
        frk_ctx *c = beside ? M->aux : ctx;
        double s2new = est ? sigma2fs : known_sigma2fs;
        if (!beside) {
            if (M->linv_valid && frk_rowbuckets_is_dense(M->bk))
                RC(frk_quadform_dense_factor(ctx, M->m, r, M->val, M->bk, M->Linv, M->mu_eta, M->q, M->t));
            else
                RC(frk_quadform_spmv_bucketed(ctx, M->m, r, M->rowptr, M->col, M->val, M->bk, M->S_eta, M->mu_eta, M->q, M->t));
        }
        if (M->cp && !M->all_vfs_zero) RC(frk_coupling_cross(M->cp, c, M->rowptr, M->col, M->val, M->S_eta));   // S_c S_eta S_c' per coupled group
        double al[MAXP_M];
        MSums s;
        double swo;
        auto J = [&](double s2, double &val) -> int {                    // :335-356
            if (s2 < 0) { val = std::numeric_limits<double>::infinity(); return 0; }
            double a2[MAXP_M]; MSums ss; double sw;
            RC(gls_and_sums(M, c, s2, 0, a2, ss, sw));
            val = -(-0.5 * ss.sv + 0.5 * sw);
            return 0;
        };
        auto sgn = [](double v) { return (v > 0) - (v < 0); };
        if (!M->all_vfs_zero) {
            if (!M->homoscedastic) {
                if (est) {
                    double amp = 10.0;
                    bool ok = false;
                    while (!ok) {                                        // :368-380
                        amp *= 10;
                        double j1, j2;
                        RC(J(sigma2fs / amp, j1));
                        RC(J(sigma2fs * amp, j2));
                        if (sgn(j1) != sgn(j2) || std::isnan(j1) || std::isnan(j2)) ok = true;
                        if (amp > 1e9) ok = true;
                    }
                    if (amp > 1e9) s2new = 0.0;
                    else RC(zeroin(J, sigma2fs / amp, sigma2fs * amp, s2new));   // uniroot, :386-388
                }
                RC(gls_and_sums(M, c, s2new, 0, al, s, swo));            // :391-400
            } else {
                RC(gls_and_sums(M, c, 0.0, 1, al, s, swo));              // :404-405
                if (est) {
                    s2new = 1.0 / M->Vfs0 * (swo / M->m_total - M->Ve0); // :408-417
                    if (s2new < 0) s2new = 0.0;
                }
            }
        } else {                                                         // :424-428
            RC(gls_and_sums(M, c, 0.0, 0, al, s, swo));
            if (est) s2new = 0.0;
        }
        for (int i = 0; i < p; i++) M->alpha[i] = al[i];
        M->sigma2fs = s2new;
        if (beside && early_gram) {
            // the next E-step's Gram matrix (R/SREfit.R:252) needs only these two: enqueue it behind the sums, beside the K update
            const size_t n = rr + r + 2;
            FRK_CHECK_CUDA(cudaMemsetAsync(M->acc, 0, n * sizeof(double), c->stream));
            RC(frk_gram_rhs_bucketed(c, M->m, r, M->rowptr, M->col, M->val, M->bk, M->Z, M->X, p, M->alpha, M->Ve, M->Vfs,
                                     M->sigma2fs, M->acc));
            M->gram_early = true;
            M->gram_s2 = M->sigma2fs;
            for (int i = 0; i < p; i++) M->gram_alpha[i] = M->alpha[i];
        }
        if (beside) {
            FRK_CHECK_CUDA(cudaEventRecord(M->ev_aux1, c->stream));
            ctx->launches += c->launches;
            c->launches = 0;
        }
        return 0;
    };
    const double tk0 = now_s();
    // (the callback is only used where the host is otherwise idle: one rank, plain basis, lanes available)
    RC(update_K(M, K_type, nlambda, lambda, h_res, have_inverse, beside ? std::function<int()>(sigma_update) : std::function<int()>()));   // :270-272
    const double tk1 = now_s();
    M->t_phase[1] += tk1 - tk0;
    struct Rest { frk_model *M; double t0; ~Rest() { M->t_phase[2] += now_s() - t0; } } rest_timer{M, tk1};
    if (!have_inverse) {
        FRK_CHECK_CUDA(cudaMemcpyAsync(M->w1, M->Khat, rr * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        RC(chol2inv_logdet(ctx, r, M->w1, M->Khat_inv, M->w2, &M->logdetK));   // Khat_inv = chol2inv(chol(K)), :273
        M->khat_block_diag = false;
    }
    if (!sigma_done) RC(sigma_update());
    if (beside && !M->gram_early) FRK_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, M->ev_aux1, 0));
    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------
static int model_common_init(frk_model *M) {
    const int r = M->r;
    const size_t rr = (size_t)r * r;
    RC(dalloc(M, &M->mu_eta, r)); RC(dalloc(M, &M->S_eta, rr)); RC(dalloc(M, &M->Q_eta, rr));
    RC(dalloc(M, &M->Khat, rr)); RC(dalloc(M, &M->Khat_inv, rr));
    RC(dalloc(M, &M->acc, rr + r + 2)); RC(dalloc(M, &M->w1, rr)); RC(dalloc(M, &M->w2, rr)); RC(dalloc(M, &M->vec, r));
    RC(dalloc(M, &M->t, (size_t)std::max<int64_t>(M->m, 1))); RC(dalloc(M, &M->q, (size_t)std::max<int64_t>(M->m, 1)));
    FRK_CHECK_CUDA(cudaMemsetAsync(M->q, 0, std::max<int64_t>(M->m, 1) * sizeof(double), M->ctx->stream));
    FRK_CHECK_CUDA(cudaMemsetAsync(M->t, 0, std::max<int64_t>(M->m, 1) * sizeof(double), M->ctx->stream));
    // row buckets of S
    if (!M->bk) RC(frk_rowbuckets_create(M->ctx, M->m, r, M->rowptr, M->col, M->val, &M->bk));   // (frk_model_create_implicit brings its own)
    if (!M->allreduce && M->ctx->comm) { M->rank = M->ctx->comm_rank; M->world = M->ctx->comm_world; }
    // totals and the homoscedasticity test all(a == a[1]) & all(b == b[1])   (R/SREfit.R:278-280)
    double mt[1] = {(double)M->m};
    RC(reduce_host(M, mt, 1, 0));
    M->m_total = mt[0];
    FRK_REQUIRE(M->m_total > 0, "no observation falls inside a BAU");
    double ves[4], vfs[4];
    RC(global_stats(M, M->Ve, 0.0, ves));
    RC(global_stats(M, M->Vfs, 0.0, vfs));
    M->homoscedastic = (ves[2] == ves[3]) && (vfs[2] == vfs[3]);
    M->Ve0 = ves[2]; M->Vfs0 = vfs[2];
    M->all_vfs_zero = (vfs[2] == 0.0 && vfs[3] == 0.0);
    return 0;
}

extern "C" int frk_model_create(frk_ctx *ctx, int64_t m, int r, int p, const int *d_rowptr, const int *d_col, const double *d_val,
                                const double *d_Z, const double *d_X, const double *d_Ve, const double *d_Vfs,
                                frk_allreduce_fn allreduce, void *user, frk_model **out) {
    FRK_REQUIRE(ctx && out && d_rowptr && d_Z && d_X && d_Ve && d_Vfs && r > 0 && m >= 0, "frk_model_create: bad argument");
    FRK_REQUIRE(p >= 1 && p <= MAXP_M, "frk_model_create: %d fixed effects; the device path handles 1..%d", p, MAXP_M);
    frk_model *M = new frk_model();
    M->ctx = ctx; M->m = m; M->r = r; M->p = p;
    M->rowptr = d_rowptr; M->col = d_col; M->val = d_val; M->Z = d_Z; M->X = d_X; M->Ve = d_Ve; M->Vfs = d_Vfs;
    M->allreduce = allreduce; M->user = user;
    if (int rc = model_common_init(M)) { frk_model_destroy(M); return rc; }
    *out = M;
    return 0;
}

// S is never stored: row j = eval_basis(basis, coords[j, ]) (normalised if asked), evaluated chunk by chunk inside every Gram /
// quadratic-form pass.  For bases without compact support at scale (BASELINE config 5: 1e7 observations x r ~ 4k Matern functions).
extern "C" int frk_model_create_implicit(frk_ctx *ctx, int64_t m, int p, const frk_basis *basis, const double *d_coords, int64_t ld_coords,
                                         int normalise, const double *d_Z, const double *d_X, const double *d_Ve, const double *d_Vfs,
                                         frk_allreduce_fn allreduce, void *user, frk_model **out) {
    FRK_REQUIRE(ctx && out && basis && d_coords && d_Z && d_X && d_Ve && d_Vfs && m >= 0, "frk_model_create_implicit: bad argument");
    FRK_REQUIRE(p >= 1 && p <= MAXP_M, "frk_model_create_implicit: %d fixed effects; the device path handles 1..%d", p, MAXP_M);
    frk_model *M = new frk_model();
    M->ctx = ctx; M->m = m; M->r = frk_basis_n(basis); M->p = p;
    M->Z = d_Z; M->X = d_X; M->Ve = d_Ve; M->Vfs = d_Vfs;
    M->allreduce = allreduce; M->user = user;
    if (int rc = frk_rowbuckets_create_implicit(ctx, m, basis, d_coords, ld_coords, normalise, &M->bk)) { delete M; return rc; }
    if (int rc = model_common_init(M)) { frk_model_destroy(M); return rc; }
    *out = M;
    return 0;
}

extern "C" int frk_model_create_host(frk_ctx *ctx, int64_t m, int r, int p, const int *h_rowptr, const int *h_col,
                                     const double *h_val, const double *h_Z, const double *h_X, const double *h_Ve,
                                     const double *h_Vfs, frk_model **out) {
    FRK_REQUIRE(ctx && out && h_rowptr && h_Z && h_X && h_Ve && h_Vfs && r > 0 && m >= 0, "frk_model_create_host: bad argument");
    FRK_REQUIRE(p >= 1 && p <= MAXP_M, "frk_model_create_host: %d fixed effects; the device path handles 1..%d", p, MAXP_M);
    frk_model *M = new frk_model();
    M->ctx = ctx; M->m = m; M->r = r; M->p = p; M->owned = true;
    const int64_t nnz = h_rowptr[m];
    int *rp, *cl; double *vl, *z, *x, *ve, *vfs;
    int rc = 0;
    rc = rc || dalloc(M, &rp, (size_t)m + 1) || dalloc(M, &cl, (size_t)std::max<int64_t>(nnz, 1)) || dalloc(M, &vl, (size_t)std::max<int64_t>(nnz, 1)) ||
         dalloc(M, &z, (size_t)std::max<int64_t>(m, 1)) || dalloc(M, &x, (size_t)std::max<int64_t>(m * p, 1)) ||
         dalloc(M, &ve, (size_t)std::max<int64_t>(m, 1)) || dalloc(M, &vfs, (size_t)std::max<int64_t>(m, 1));
    if (rc) { frk_model_destroy(M); return 1; }
    cudaStream_t st = ctx->stream;
    cudaMemcpyAsync(rp, h_rowptr, (m + 1) * sizeof(int), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(cl, h_col, nnz * sizeof(int), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(vl, h_val, nnz * sizeof(double), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(z, h_Z, m * sizeof(double), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(x, h_X, m * p * sizeof(double), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(ve, h_Ve, m * sizeof(double), cudaMemcpyHostToDevice, st);
    cudaError_t e = cudaMemcpyAsync(vfs, h_Vfs, m * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { frk_set_error("frk_model_create_host: upload failed: %s", cudaGetErrorString(e)); frk_model_destroy(M); return 1; }
    M->rowptr = rp; M->col = cl; M->val = vl; M->Z = z; M->X = x; M->Ve = ve; M->Vfs = vfs;
    if (int rc2 = model_common_init(M)) { frk_model_destroy(M); return rc2; }
    *out = M;
    return 0;
}

extern "C" int frk_model_destroy(frk_model *M) {
    if (!M) return 0;
    cudaSetDevice(M->ctx->device);
    if (M->aux) cudaStreamSynchronize(M->aux->stream);          // nothing of this model may still be running on the lanes when its
    cudaStreamSynchronize(M->ctx->stream);                      // buffers go back to the pool (the lane contexts belong to M->ctx)
    for (size_t j = 0; j < M->pre.size(); j++)                  // prefetched factorisations read this model's distance matrices
        for (int c = 0; c < 8; c++)
            if (M->pre[j][c].on) {
                frk_ctx *lc = nullptr;
                if ((int)j * 8 + c > 0 && frk_ctx_lane(M->ctx, (int)j * 8 + c - 1, true, &lc) == 0 && lc) cudaStreamSynchronize(lc->stream);
            }
    for (void *p : M->allocs) cudaFreeAsync(p, M->ctx->stream);
    frk_rowbuckets_destroy(M->bk);
    frk_coupling_destroy(M->cp);
    if (M->ev_lane) cudaEventDestroy(M->ev_lane);
    if (M->ev_aux0) cudaEventDestroy(M->ev_aux0);
    if (M->ev_aux1) cudaEventDestroy(M->ev_aux1);
    delete M;
    return 0;
}

// Non-diagonal Vfs (R/SRE.R:498: tcrossprod(Cmat %*% Diagonal(sqrt(fs))) when observation footprints share BAUs, or several point
// observations sit in one BAU with average_in_BAU = FALSE): the full symmetric m x m matrix as host CSR.  The model then takes the
// reference's general branches (R/SREfit.R:186-189, 240-252, 296-328); see coupled.cu.  Call before frk_model_init.
extern "C" int frk_model_set_vfs_csr(frk_model *M, const int *h_rowptr, const int *h_col, const double *h_val) {
    FRK_REQUIRE(M && h_rowptr && h_col && h_val, "frk_model_set_vfs_csr: NULL argument");
    FRK_REQUIRE(!M->cp, "frk_model_set_vfs_csr: already set");
    FRK_REQUIRE(M->rowptr, "frk_model_set_vfs_csr: not available for an implicit (never stored) basis matrix");
    FRK_REQUIRE(!multi_rank(M) && M->world == 1, "a non-diagonal Vfs couples observations across the row partition: not on the device path (one rank only)");
    bool offdiag = false;
    for (int64_t i = 0; i < M->m && !offdiag; i++)
        for (int e = h_rowptr[i]; e < h_rowptr[i + 1]; e++)
            if (h_col[e] != i && h_val[e] != 0.0) { offdiag = true; break; }
    if (!offdiag) return 0;                                          // isDiagonal(Vfs): the diagonal paths apply as they are
    RC(frk_coupling_create(M->ctx, M->m, M->r, M->p, h_rowptr, h_col, h_val, M->rowptr, M->col, &M->cp));
    M->homoscedastic = false;                                        // homoscedastic <- ... & isDiagonal(Sm@Vfs), R/SREfit.R:278-280
    return 0;
}

extern "C" int frk_model_coupled_groups(const frk_model *M) { return (M && M->cp) ? frk_coupling_ncomp(M->cp) : 0; }

// this rank's position in a row-partitioned run (enables the speculative Nelder-Mead); emulate != 0: evaluate all of a pretend
// world's candidates on this rank (tests the logic on one GPU)
extern "C" int frk_model_set_parallel(frk_model *M, int rank, int world, int emulate) {
    FRK_REQUIRE(M && world >= 1 && rank >= 0 && rank < world, "frk_model_set_parallel: bad argument");
    FRK_REQUIRE(emulate || world == 1 || multi_rank(M), "frk_model_set_parallel: world > 1 needs the all-reduce hook or frk_comm_init");
    M->rank = emulate ? 0 : rank; M->world = world; M->spec_emulate = emulate != 0;
    return 0;
}

// distribute != 0: spread the Nelder-Mead candidates of a round over the ranks (one small all-reduce per round, the winner's inverse
// broadcast); 0 (default): every rank evaluates them on its own lanes -- replicated, deterministic, no exchange
extern "C" int frk_model_set_nm_distribute(frk_model *M, int distribute) {
    FRK_REQUIRE(M, "frk_model_set_nm_distribute: NULL model");
    FRK_REQUIRE(!distribute || M->world == 1 || multi_rank(M), "frk_model_set_nm_distribute: needs the all-reduce hook or frk_comm_init");
    M->nm_distribute = distribute != 0;
    return 0;
}

// number of concurrent evaluation lanes on this GPU for the block-exponential Nelder-Mead (1 = strictly sequential)
extern "C" int frk_model_set_lanes(frk_model *M, int lanes) {
    FRK_REQUIRE(M && lanes >= 1 && lanes <= 8, "frk_model_set_lanes: lanes must be in 1..8");
    M->lanes = lanes;
    return 0;
}

// per-resolution index sets and centroid distance matrices (BuildD, R/basisfns.R:1131-1153) for K_type = "block-exponential"
static int set_blocks_impl(frk_model *M, int nres, const int *h_res, const double *const *h_D, const frk_basis *basis) {
    FRK_REQUIRE(M && nres > 0 && h_res, "frk_model_set_blocks: bad argument");
    FRK_REQUIRE(M->blocks.empty(), "frk_model_set_blocks: already set");
    const int r = M->r;
    for (int i = 1; i <= nres; i++) {
        std::vector<int> ix;
        for (int j = 0; j < r; j++) if (h_res[j] == i) ix.push_back(j);
        FRK_REQUIRE(!ix.empty(), "frk_model_set_blocks: resolution %d has no basis function", i);
        Block B;
        B.ni = (int)ix.size(); B.ix0 = ix[0];
        const size_t nn = (size_t)B.ni * B.ni;
        RC(dalloc(M, &B.d_ix, ix.size())); RC(dalloc(M, &B.D, nn)); RC(dalloc(M, &B.eta2, nn)); RC(dalloc(M, &B.Ki, nn));
        RC(dalloc(M, &B.Kinv, nn)); RC(dalloc(M, &B.work, nn)); RC(dalloc(M, &B.KinvBest, nn));
        B.memo.resize(B.ni >= 1024 ? 3 : 12);                        // (3 x 68 MB at n_i = 2916; 12 x 0.8 MB at 324)
        for (auto &e : B.memo) RC(dalloc(M, &e.buf, nn));
        FRK_CHECK_CUDA(cudaMemcpyAsync(B.d_ix, ix.data(), ix.size() * sizeof(int), cudaMemcpyHostToDevice, M->ctx->stream));
        FRK_CHECK_CUDA(cudaStreamSynchronize(M->ctx->stream));       // ix is a temporary
        if (h_D) {
            FRK_CHECK_CUDA(cudaMemcpyAsync(B.D, h_D[i - 1], nn * sizeof(double), cudaMemcpyHostToDevice, M->ctx->stream));
            B.D01 = B.ni > 1 ? h_D[i - 1][B.ni] : 0.0;               // D[1,2] (column-major, symmetric)
            if (i == 1) { double mx = 0; for (size_t k = 0; k < nn; k++) mx = std::max(mx, h_D[0][k]); M->max_l = mx; }
        } else {                                                     // BuildD on the device from the basis' centroid table
            RC(frk_basis_dist_matrix(M->ctx, basis, B.ni, B.d_ix, B.D));
            if (B.ni > 1) RC(frk_d2h(M->ctx, &B.D01, B.D + B.ni, sizeof(double)));
            if (i == 1) { double st[4]; RC(frk_vec_stats(M->ctx, (int64_t)nn, B.D, 0.0, st)); M->max_l = st[3]; }
        }
        M->blocks.push_back(B);
    }
    return 0;
}

extern "C" int frk_model_set_blocks(frk_model *M, int nres, const int *h_res, const double *const *h_D) {
    FRK_REQUIRE(h_D, "frk_model_set_blocks: bad argument");
    return set_blocks_impl(M, nres, h_res, h_D, nullptr);
}

// BuildD on the device: the distance matrices come from the basis' own centroid table (same arithmetic as eval_basis)
extern "C" int frk_model_set_blocks_basis(frk_model *M, int nres, const int *h_res, const frk_basis *basis) {
    FRK_REQUIRE(basis && frk_basis_n(basis) == (M ? M->r : -1), "frk_model_set_blocks_basis: the basis does not match the model");
    return set_blocks_impl(M, nres, h_res, nullptr, basis);
}

extern "C" int frk_model_set_blocks_tensor(frk_model *M, int nres, int r_s, int r_t, const int *h_res_s, const double *const *h_Ds,
                                           const double *h_Dt) {
    FRK_REQUIRE(M && nres > 0 && h_res_s && h_Ds && h_Dt && r_s > 0 && r_t > 1, "frk_model_set_blocks_tensor: bad argument");
    FRK_REQUIRE((int64_t)r_s * r_t == M->r, "frk_model_set_blocks_tensor: r_s * r_t = %d * %d does not match r = %d", r_s, r_t, M->r);
    FRK_REQUIRE(M->blocks.empty(), "frk_model_set_blocks_tensor: already set");
    M->tensor = true;
    const size_t ntt = (size_t)r_t * r_t;
    RC(dalloc(M, &M->Dt, ntt));
    FRK_CHECK_CUDA(cudaMemcpy(M->Dt, h_Dt, ntt * sizeof(double), cudaMemcpyHostToDevice));
    M->Dt01 = h_Dt[r_t];                                             // D[1,2]
    for (int i = 1; i <= nres; i++) {
        std::vector<int> is, ix;
        for (int j = 0; j < r_s; j++) if (h_res_s[j] == i) is.push_back(j);
        FRK_REQUIRE(!is.empty(), "frk_model_set_blocks_tensor: resolution %d has no basis function", i);
        for (int t = 0; t < r_t; t++) for (int j : is) ix.push_back(j + r_s * t);   // ascending tensor columns: space fastest
        Block B;
        B.ns = (int)is.size(); B.nt = r_t; B.ni = (int)ix.size(); B.ix0 = ix[0];
        const size_t nn = (size_t)B.ni * B.ni, nss = (size_t)B.ns * B.ns;
        RC(dalloc(M, &B.d_ix, ix.size())); RC(dalloc(M, &B.D, nss)); RC(dalloc(M, &B.eta2, nn)); RC(dalloc(M, &B.Ki, nn));
        RC(dalloc(M, &B.Kinv, nn)); RC(dalloc(M, &B.work, nn));
        RC(dalloc(M, &B.Ks, nss)); RC(dalloc(M, &B.Qs, nss)); RC(dalloc(M, &B.ws, nss)); RC(dalloc(M, &B.QsBest, nss));
        RC(dalloc(M, &B.Kt, ntt)); RC(dalloc(M, &B.Qt, ntt)); RC(dalloc(M, &B.wt, ntt)); RC(dalloc(M, &B.QtBest, ntt));
        FRK_CHECK_CUDA(cudaMemcpy(B.d_ix, ix.data(), ix.size() * sizeof(int), cudaMemcpyHostToDevice));
        FRK_CHECK_CUDA(cudaMemcpy(B.D, h_Ds[i - 1], nss * sizeof(double), cudaMemcpyHostToDevice));
        B.D01 = B.ns > 1 ? h_Ds[i - 1][B.ns] : 0.0;
        if (i == 1) { double mx = 0; for (size_t k = 0; k < nss; k++) mx = std::max(mx, h_Ds[0][k]); M->max_l = mx; }
        M->blocks.push_back(B);
    }
    return 0;
}

// .EM_initialise, R/SRE.R:704-731
extern "C" int frk_model_init(frk_model *M) {
    FRK_REQUIRE(M, "frk_model_init: NULL model");
    frk_ctx *ctx = M->ctx;
    const int r = M->r;
    double zs[4], zs2[4], ves[4];
    RC(global_stats(M, M->Z, 0.0, zs));
    const double zbar = zs[0] / M->m_total;
    RC(global_stats(M, M->Z, zbar, zs2));
    const double varZ = M->m_total > 1 ? zs2[1] / (M->m_total - 1) : std::numeric_limits<double>::quiet_NaN();
    RC(global_stats(M, M->Ve, 0.0, ves));
    M->sigma2fs = (ves[0] / M->m_total) / 4.0;
    RC(frk_scaled_identity(ctx, r, 1.0, M->S_eta));
    RC(frk_scaled_identity(ctx, r, 1.0, M->Q_eta));
    RC(frk_scaled_identity(ctx, r, 1.0 / (1.0 / varZ), M->Khat));
    RC(frk_scaled_identity(ctx, r, 1.0 / (1.0 / (1.0 / varZ)), M->Khat_inv));
    FRK_CHECK_CUDA(cudaMemsetAsync(M->mu_eta, 0, r * sizeof(double), ctx->stream));
    M->logdetK = r * std::log(varZ);
    M->khat_block_diag = true;                                       // diagonal K, Khat_inv its inverse
    // alpha = (X'X)^-1 X'Z: the mode-1 moments with t = 0
    FRK_CHECK_CUDA(cudaMemsetAsync(M->t, 0, std::max<int64_t>(M->m, 1) * sizeof(double), ctx->stream));
    MSums s;
    RC(mstep_sums(M, M->ctx, 0.0, 1, s));
    RC(solve_small(M->p, s.uxx, s.uxz, M->alpha));
    M->initialised = true;
    return 0;
}

extern "C" int frk_model_em_fit(frk_model *M, int n_EM, double tol, int K_type, int nlambda, const double *h_lambda,
                                const int *h_res, int known_sigma2fs_given, double known_sigma2fs, double *h_llk, int *h_niter) {
    FRK_REQUIRE(M && h_llk && h_niter && n_EM >= 0, "frk_model_em_fit: bad argument");
    FRK_REQUIRE(M->initialised, "frk_model_em_fit: call frk_model_init (or frk_model_set for every slot) first");
    FRK_REQUIRE(K_type == FRK_K_UNSTRUCTURED || K_type == FRK_K_BLOCK_EXPONENTIAL,
                "K_type needs to be \"block-exponential\" or \"unstructured\" for method = \"EM\"");
    if (known_sigma2fs_given) M->sigma2fs = known_sigma2fs;          // R/SREfit.R:19
    int i = 0;
    static const bool phase_dbg = getenv("FRK_PHASE_DEBUG") != nullptr;
    for (int k = 0; k < 4; k++) M->t_phase[k] = 0.0;
    for (i = 0; i < n_EM; i++) {
        double o[4];
        const double t0 = now_s();
        RC(posterior(M, M->alpha, M->sigma2fs, M->Q_eta, M->S_eta, M->mu_eta, o));   // logLik + E-step, :105-106
        h_llk[i] = llk_from(M, o);
        M->t_phase[0] += now_s() - t0;
        RC(mstep(M, K_type, nlambda, h_lambda, h_res, known_sigma2fs_given, known_sigma2fs));   // :107
        if (i > 0 && std::fabs(h_llk[i] - h_llk[i - 1]) < tol) { i++; break; }    // :110-114
    }
    *h_niter = std::min(i, n_EM);
    if (M->aux) FRK_CHECK_CUDA(cudaStreamSynchronize(M->aux->stream));   // (an early Gram for the next call may still be in flight)
    FRK_CHECK_CUDA(cudaStreamSynchronize(M->ctx->stream));
    if (phase_dbg && M->rank == 0 && *h_niter > 0)
        fprintf(stderr, "[frk phases] per iteration over %d: posterior %.3f ms, update_K %.3f ms (Nelder-Mead rounds %.3f), rest of M-step %.3f ms\n",
                *h_niter, 1e3 * M->t_phase[0] / *h_niter, 1e3 * M->t_phase[1] / *h_niter, 1e3 * M->t_phase[3] / *h_niter,
                1e3 * M->t_phase[2] / *h_niter);
    return 0;
}

// logLik(SRE) at the current parameters (R/SREutils.R:22-32 -> .loglik.ind); the slots are left untouched
extern "C" int frk_model_loglik(frk_model *M, double *h_out) {
    FRK_REQUIRE(M && h_out && M->initialised, "frk_model_loglik: bad argument");
    double *Q, *Sg, *mu;
    const size_t rr = (size_t)M->r * M->r;
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&Q, (2 * rr + M->r) * sizeof(double), M->ctx->stream));
    Sg = Q + rr; mu = Sg + rr;
    double o[4];
    int rc = posterior(M, M->alpha, M->sigma2fs, Q, Sg, mu, o);
    cudaFreeAsync(Q, M->ctx->stream);
    if (rc) return rc;
    *h_out = llk_from(M, o);
    return 0;
}

static double *slot_ptr(frk_model *M, const char *name, size_t *n) {
    const size_t rr = (size_t)M->r * M->r;
    if (!strcmp(name, "mu_eta")) { *n = M->r; return M->mu_eta; }
    if (!strcmp(name, "S_eta")) { *n = rr; return M->S_eta; }
    if (!strcmp(name, "Q_eta")) { *n = rr; return M->Q_eta; }
    if (!strcmp(name, "Khat")) { *n = rr; return M->Khat; }
    if (!strcmp(name, "Khat_inv")) { *n = rr; return M->Khat_inv; }
    return nullptr;
}

extern "C" int frk_model_get(frk_model *M, const char *slot, double *h_out) {
    FRK_REQUIRE(M && slot && h_out, "frk_model_get: bad argument");
    if (!strcmp(slot, "alphahat")) { for (int i = 0; i < M->p; i++) h_out[i] = M->alpha[i]; return 0; }
    if (!strcmp(slot, "sigma2fshat")) { h_out[0] = M->sigma2fs; return 0; }
    if (!strcmp(slot, "m_total")) { h_out[0] = M->m_total; return 0; }
    if (!strcmp(slot, "homoscedastic")) { h_out[0] = M->homoscedastic ? 1.0 : 0.0; return 0; }
    if (!strcmp(slot, "tau")) { for (size_t i = 0; i < M->taus.size(); i++) h_out[i] = M->taus[i]; return 0; }
    if (!strcmp(slot, "nm_stats")) { h_out[0] = M->nm_requests; h_out[1] = M->nm_rounds; h_out[2] = M->nm_evals; h_out[3] = M->nm_reevals; return 0; }
    if (!strcmp(slot, "nm_prefetch_hits")) { h_out[0] = M->nm_prefetch_hits; return 0; }
    if (!strcmp(slot, "nm_reuse_hits")) { h_out[0] = M->nm_reuse_hits; return 0; }
    if (!strcmp(slot, "omega")) { for (size_t i = 0; i < M->omegas.size(); i++) h_out[i] = M->omegas[i]; return 0; }
    size_t n = 0;
    double *p = slot_ptr(M, slot, &n);
    FRK_REQUIRE(p, "frk_model_get: unknown slot '%s'", slot);
    return frk_d2h(M->ctx, h_out, p, n * sizeof(double));
}

extern "C" int frk_model_set(frk_model *M, const char *slot, const double *h_in) {
    FRK_REQUIRE(M && slot && h_in, "frk_model_set: bad argument");
    if (!strcmp(slot, "alphahat")) { for (int i = 0; i < M->p; i++) M->alpha[i] = h_in[i]; return 0; }
    if (!strcmp(slot, "sigma2fshat")) { M->sigma2fs = h_in[0]; return 0; }
    size_t n = 0;
    double *p = slot_ptr(M, slot, &n);
    FRK_REQUIRE(p, "frk_model_set: unknown slot '%s'", slot);
    RC(frk_h2d(M->ctx, p, h_in, n * sizeof(double)));
    FRK_CHECK_CUDA(cudaStreamSynchronize(M->ctx->stream));
    if (!strcmp(slot, "Khat") || !strcmp(slot, "Khat_inv")) M->khat_block_diag = false;   // caller-supplied: assume nothing
    if (!strcmp(slot, "S_eta") || !strcmp(slot, "Q_eta")) M->linv_valid = false;          // ... nor that L^-1 still belongs to S_eta
    if (!strcmp(slot, "Khat")) {                                     // keep log|K| consistent with the slot (.loglik.ind :177,197)
        const size_t rr = (size_t)M->r * M->r;
        FRK_CHECK_CUDA(cudaMemcpyAsync(M->w1, M->Khat, rr * sizeof(double), cudaMemcpyDeviceToDevice, M->ctx->stream));
        RC(frk_potrf(M->ctx, M->r, M->w1, &M->logdetK));
        M->initialised = true;
    }
    return 0;
}

extern "C" int frk_model_slot_device(frk_model *M, const char *slot, double **d_out) {
    FRK_REQUIRE(M && slot && d_out, "frk_model_slot_device: bad argument");
    size_t n = 0;
    *d_out = slot_ptr(M, slot, &n);
    FRK_REQUIRE(*d_out, "frk_model_slot_device: unknown slot '%s'", slot);
    return 0;
}

// predict at n BAUs (this rank's rows): .SRE.predict_EM with predict_BAUs = TRUE, covariances = FALSE.
// S0: n x r device CSR with its row buckets; d_XBAU n x p column-major; d_fs n; d_obs_bau m (BAU row of every observation).
extern "C" int frk_model_predict(frk_model *M, int64_t n, const int *d_rowptr0, const int *d_col0, const double *d_val0,
                                 const frk_rowbuckets *bk0, const double *d_XBAU, const double *d_fs, const int *d_obs_bau,
                                 int obs_fs, double *d_mu, double *d_var, double *d_sd) {
    FRK_REQUIRE(M && bk0 && d_XBAU && d_mu && d_var && d_sd && n >= 0, "frk_model_predict: bad argument");   // (d_rowptr0 may be NULL for implicit row buckets)
    frk_ctx *ctx = M->ctx;
    const int r = M->r;
    const size_t rr = (size_t)r * r;
    const int64_t n1 = std::max<int64_t>(n, 1), m1 = std::max<int64_t>(M->m, 1);
    double *buf = nullptr;
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&buf, (size_t)(5 * n1 + 2 * m1) * sizeof(double), ctx->stream));
    double *xa = buf, *q = xa + n1, *smu = q + n1, *A = smu + n1, *ytil = A + n1, *tmp = ytil + n1, *inv_ve = tmp + m1;
    int rc = 0;
    auto run = [&]() -> int {
        RC(frk_xalpha(ctx, n, M->p, d_XBAU, M->alpha, xa));
        if (obs_fs || M->sigma2fs == 0.0) {                           // the stored posterior (R/SREpredict.R:334-361)
            if (M->linv_valid && frk_rowbuckets_is_dense(bk0))
                RC(frk_quadform_dense_factor(ctx, n, r, d_val0, bk0, M->Linv, M->mu_eta, q, smu));
            else
                RC(frk_quadform_spmv_bucketed(ctx, n, r, d_rowptr0, d_col0, d_val0, bk0, M->S_eta, M->mu_eta, q, smu));
            return frk_predict_finalize(ctx, n, 0, M->sigma2fs, q, smu, xa, nullptr, nullptr, nullptr, d_mu, d_var, d_sd);
        }
        // joint (eta, xi) solve at the final parameters (:308-333); for point data: one more E-step + the same quadratic form
        FRK_REQUIRE(d_fs && d_obs_bau, "frk_model_predict: obs_fs = FALSE needs the BAU fs field and the observation -> BAU map");
        FRK_REQUIRE(!M->cp, "frk_model_predict: observations coupled through shared BAUs (non-diagonal Vfs) -- obs_fs = FALSE goes through "
                    "frk_model_predict_areal with the incidence matrix Cmat (the one-BAU-per-observation closed form does not apply)");
        double *Q, o[4];
        FRK_CHECK_CUDA(cudaMallocAsync((void **)&Q, (2 * rr + r) * sizeof(double), ctx->stream));
        double *Sig = Q + rr, *mue = Sig + rr;
        int rc2 = posterior(M, M->alpha, M->sigma2fs, Q, Sig, mue, o);
        if (!rc2 && frk_rowbuckets_is_dense(bk0))                   // (posterior left L^-1 of THIS Sig in w2)
            rc2 = frk_quadform_dense_factor(ctx, n, r, d_val0, bk0, M->w2, mue, q, smu);
        else if (!rc2) rc2 = frk_quadform_spmv_bucketed(ctx, n, r, d_rowptr0, d_col0, d_val0, bk0, Sig, mue, q, smu);
        cudaFreeAsync(Q, ctx->stream);
        if (rc2) return rc2;
        // A = diag(C' Ve^-1 C), ytil = C' Ve^-1 (Z - X alpha)
        FRK_CHECK_CUDA(cudaMemsetAsync(A, 0, 2 * n1 * sizeof(double), ctx->stream));
        RC(frk_resid(ctx, M->m, M->p, M->X, M->alpha, M->Z, M->Ve, tmp));
        RC(frk_scatter_add(ctx, M->m, d_obs_bau, tmp, ytil));
        RC(frk_recip(ctx, M->m, M->Ve, inv_ve));
        RC(frk_scatter_add(ctx, M->m, d_obs_bau, inv_ve, A));
        return frk_predict_finalize(ctx, n, 1, M->sigma2fs, q, smu, xa, A, ytil, d_fs, d_mu, d_var, d_sd);
    };
    rc = run();
    cudaFreeAsync(buf, ctx->stream);
    return rc;
}

// ---------------------------------------------------------------------------------
// predict(obs_fs = FALSE), areal data with supports that share no BAU.  The joint (eta, xi) solve of .SRE.predict_EM
// (R/SREpredict.R:308-333) has a closed form here: with d_j = sigma2fs * Vfs_j + Ve_j, the posterior (mu~, Sigma~) of eta at the final
// parameters (one more E-step), S_j = row j of S = Cmat S0 and, for BAU k inside support j with weight c_jk,
//   g_k   = sigma2fs * fs_k * c_jk / d_j
//   E     = x_k'alpha + s_k'mu~ + g_k (Z_j - X_j alpha - S_j mu~)
//   Var   = s_k'Sigma~ s_k - 2 g_k s_k'Sigma~ S_j' + g_k^2 S_j Sigma~ S_j' + sigma2fs fs_k - g_k^2 d_j
// (BAUs in no support: g = 0).  One CTA per support forms v = Sigma~ S_j' in shared memory, one warp per member BAU takes s_k'v.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) areal_case2_kernel(int64_t m, int r, const int *__restrict__ crp, const int *__restrict__ mem,
                                                          const double *__restrict__ wgt, const int *__restrict__ rpS,
                                                          const int *__restrict__ colS, const double *__restrict__ valS,
                                                          const int *__restrict__ rp0, const int *__restrict__ col0,
                                                          const double *__restrict__ val0, const double *__restrict__ Sig,
                                                          const double *__restrict__ QS, const double *__restrict__ res,
                                                          const double *__restrict__ dj, const double *__restrict__ q,
                                                          const double *__restrict__ fs, double sigma2fs, double *__restrict__ mu,
                                                          double *__restrict__ var, double *__restrict__ sd) {
    extern __shared__ __align__(16) double a2_v[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int64_t j = blockIdx.x; j < m; j += gridDim.x) {
        for (int l = threadIdx.x; l < r; l += blockDim.x) {        // v[l] = sum_i S_ji Sigma~[l, i]  (Sigma~ symmetric: column i is contiguous)
            double acc = 0.0;
            for (int e = rpS[j]; e < rpS[j + 1]; e++) acc += valS[e] * Sig[(size_t)colS[e] * r + l];
            a2_v[l] = acc;
        }
        __syncthreads();
        const double Qj = QS[j], rj = res[j], d = dj[j];
        for (int k = crp[j] + wid; k < crp[j + 1]; k += nwarp) {
            const int b = mem[k];
            double cr = 0.0;
            for (int e = rp0[b] + lane; e < rp0[b + 1]; e += 32) cr += val0[e] * a2_v[col0[e]];
            cr = warp_sum(cr);
            if (lane == 0) {
                const double sf = sigma2fs * fs[b], g = sf * wgt[k] / d;
                const double v_ = q[b] - 2.0 * g * cr + g * g * Qj + sf - g * g * d;
                mu[b] += g * rj;
                var[b] = v_;
                sd[b] = sqrt(v_);
            }
        }
        __syncthreads();
    }
}

__global__ void areal_dj_kernel(int64_t m, double sigma2fs, const double *__restrict__ vfs, const double *__restrict__ ve,
                                const double *__restrict__ smuS, double *__restrict__ res, double *__restrict__ dj) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (int64_t)gridDim.x * blockDim.x) {
        dj[j] = sigma2fs * vfs[j] + ve[j];
        res[j] -= smuS[j];                                  // res came in as Z - X alpha
    }
}

extern "C" int frk_model_predict_areal(frk_model *M, int64_t n, const int *d_rowptr0, const int *d_col0, const double *d_val0,
                                       const frk_rowbuckets *bk0, const double *d_XBAU, const double *d_fs, const int *d_cz_rowptr,
                                       const int *d_cz_members, const double *d_cz_weights, double *d_mu, double *d_var, double *d_sd) {
    FRK_REQUIRE(M && d_rowptr0 && bk0 && d_XBAU && d_fs && d_cz_rowptr && d_cz_members && d_cz_weights && d_mu && d_var && d_sd && n > 0,
                "frk_model_predict_areal: bad argument");
    FRK_REQUIRE(M->sigma2fs > 0.0, "frk_model_predict_areal: sigma2fs = 0 -- use frk_model_predict (stored posterior)");
    FRK_REQUIRE(!multi_rank(M), "frk_model_predict_areal: row-partitioned runs are not supported for areal data");
    frk_ctx *ctx = M->ctx;
    const int r = M->r;
    const size_t rr = (size_t)r * r;
    const int64_t m = M->m;
    FRK_REQUIRE((size_t)r * sizeof(double) <= 200 * 1024, "frk_model_predict_areal: r = %d does not fit the shared-memory vector", r);
    double *buf = nullptr;
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&buf, (size_t)(3 * n + 4 * m + 2 * rr + r) * sizeof(double), ctx->stream));
    double *xa = buf, *q = xa + n, *smu = q + n, *QS = smu + n, *smuS = QS + m, *res = smuS + m, *dj = res + m, *Q = dj + m,
           *Sig = Q + rr, *mue = Sig + rr;
    auto run = [&]() -> int {
        double o[4];
        RC(frk_xalpha(ctx, n, M->p, d_XBAU, M->alpha, xa));
        RC(posterior(M, M->alpha, M->sigma2fs, Q, Sig, mue, o));                                   // eta | Z at the final parameters
        RC(frk_quadform_spmv_bucketed(ctx, n, r, d_rowptr0, d_col0, d_val0, bk0, Sig, mue, q, smu));           // s_k'Sigma~ s_k, s_k'mu~
        RC(frk_quadform_spmv_bucketed(ctx, m, r, M->rowptr, M->col, M->val, M->bk, Sig, mue, QS, smuS));       // S_j Sigma~ S_j', S_j mu~
        RC(frk_resid(ctx, m, M->p, M->X, M->alpha, M->Z, nullptr, res));
        // every BAU as if unobserved (g = 0): mu = x'alpha + s'mu~, var = s'Sigma~ s + sigma2fs fs; the supports' BAUs are then corrected
        RC(frk_predict_finalize(ctx, n, 1, M->sigma2fs, q, smu, xa, nullptr, nullptr, d_fs, d_mu, d_var, d_sd));
        if (M->cp) {
            // supports that share BAUs (or several observations per BAU): the correction couples the observations of a connected group
            return frk_coupling_case2(M->cp, ctx, M->sigma2fs, M->Ve, n, d_cz_rowptr, d_cz_members, d_cz_weights, M->rowptr, M->col, M->val,
                                      d_rowptr0, d_col0, d_val0, Sig, res, smuS, q, d_fs, d_mu, d_var, d_sd);
        }
        FRK_LAUNCH(ctx, areal_dj_kernel, (unsigned)cdiv(m, 256), 256, 0, m, M->sigma2fs, M->Vfs, M->Ve, smuS, res, dj);
        static bool attr = false;
        if (!attr) {
            FRK_CHECK_CUDA(cudaFuncSetAttribute(areal_case2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr = true;
        }
        const unsigned grid = (unsigned)std::min<int64_t>(m, (int64_t)ctx->sm_count * 4);
        FRK_LAUNCH(ctx, areal_case2_kernel, grid, 256, (size_t)r * sizeof(double), m, r, d_cz_rowptr, d_cz_members, d_cz_weights, M->rowptr,
                   M->col, M->val, d_rowptr0, d_col0, d_val0, Sig, QS, res, dj, q, d_fs, M->sigma2fs, d_mu, d_var, d_sd);
        return 0;
    };
    const int rc = run();
    cudaFreeAsync(buf, ctx->stream);
    return rc;
}

// ---------------------------------------------------------------------------------
// predict(newdata = <polygons>): linear combinations C_P of the BAU-level process (R/SREpredict.R:365-419).
//   mu_P  = C_P mu_BAU
//   obs_fs or sigma2fs == 0:  var_P = diag(C_P S0 Cov S0' C_P'),  Cov = S_eta                       (:385-391)
//   else (Case 2, point data): the reference solves the joint system for every polygon (:392-404).  For point-referenced
//   data xi_k | eta, Z are independent over BAUs with variance v_k = 1/(A_k + Q_k) and Y_k = (1 - a_k) s_k'eta + const + noise_k,
//   so  var_P = s~_P' Sigma~ s~_P + sum_k c_Pk^2 v_k  with  s~_P = sum_k c_Pk (1 - a_k) s_k  (same closed form as the BAU-level Case 2).
// ---------------------------------------------------------------------------------
__global__ void cp_case2_kernel(int64_t mP, const int *__restrict__ rp, const int *__restrict__ mem, const double *__restrict__ w,
                                const double *__restrict__ A, const double *__restrict__ fs, double sigma2fs, double *__restrict__ w2,
                                double *__restrict__ extra) {
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= mP) return;
    double e = 0.0;
    for (int k = rp[j] + lane; k < rp[j + 1]; k += 32) {
        const int b = mem[k];
        const double den = A[b] + 1.0 / (sigma2fs * fs[b]);
        w2[k] = w[k] * (1.0 - A[b] / den);
        e += w[k] * w[k] / den;
    }
    e = warp_sum(e);
    if (lane == 0) extra[j] = e;
}

__global__ void cp_finish_kernel(int64_t mP, const double *__restrict__ q, const double *__restrict__ extra, double *__restrict__ var,
                                 double *__restrict__ sd) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < mP; j += (int64_t)gridDim.x * blockDim.x) {
        const double v = q[j] + (extra ? extra[j] : 0.0);
        var[j] = v;
        sd[j] = sqrt(v);
    }
}

extern "C" int frk_model_predict_polygons(frk_model *M, int64_t n, const int *d_rowptr0, const int *d_col0, const double *d_val0,
                                          const frk_rowbuckets *bk0, const double *d_XBAU, const double *d_fs, const int *d_obs_bau,
                                          int obs_fs, int64_t mP, const int *d_cp_rowptr, const int *d_cp_members,
                                          const double *d_cp_weights, int64_t cp_nnz, double *d_mu, double *d_var, double *d_sd) {
    FRK_REQUIRE(M && d_rowptr0 && bk0 && d_XBAU && d_cp_rowptr && d_cp_members && d_cp_weights && d_mu && d_var && d_sd && n > 0 && mP > 0 &&
                cp_nnz > 0, "frk_model_predict_polygons: bad argument");
    FRK_REQUIRE(!multi_rank(M), "frk_model_predict_polygons: row-partitioned runs are not supported (polygons may straddle ranks)");
    frk_ctx *ctx = M->ctx;
    const int r = M->r;
    const size_t rr = (size_t)r * r;
    const bool case1 = obs_fs || M->sigma2fs == 0.0;
    FRK_REQUIRE(case1 || (d_fs && d_obs_bau), "frk_model_predict_polygons: obs_fs = FALSE needs the BAU fs field and the observation -> BAU map");
    FRK_REQUIRE(case1 || !M->cp, "frk_model_predict_polygons: obs_fs = FALSE over polygons for observations coupled through shared BAUs is not on "
                "the device path (the polygon variance needs the xi covariances inside a coupled group); use obs_fs = TRUE or predict at the BAUs");
    const int64_t m1 = std::max<int64_t>(M->m, 1);
    double *buf = nullptr;
    int *irp = nullptr;
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&buf, (size_t)(8 * n + 2 * m1 + 3 * mP + cp_nnz + (case1 ? 0 : 2 * rr + r)) * sizeof(double), ctx->stream));
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&irp, (size_t)(mP + 1) * sizeof(int), ctx->stream));
    double *xa = buf, *q = xa + n, *smu = q + n, *A = smu + n, *ytil = A + n, *muB = ytil + n, *varB = muB + n, *sdB = varB + n,
           *tmp = sdB + n, *inv_ve = tmp + m1, *qP = inv_ve + m1, *extra = qP + mP, *spare = extra + mP, *w2 = spare + mP, *Q = w2 + cp_nnz;
    int *pcol = nullptr;
    double *pval = nullptr;
    auto run = [&]() -> int {
        const double *Sig = M->S_eta, *mue = M->mu_eta;
        RC(frk_xalpha(ctx, n, M->p, d_XBAU, M->alpha, xa));
        if (case1) {
            RC(frk_quadform_spmv_bucketed(ctx, n, r, d_rowptr0, d_col0, d_val0, bk0, Sig, mue, q, smu));
            RC(frk_predict_finalize(ctx, n, 0, M->sigma2fs, q, smu, xa, nullptr, nullptr, nullptr, muB, varB, sdB));
        } else {
            double o[4];
            double *SigT = Q + rr, *mueT = SigT + rr;
            RC(posterior(M, M->alpha, M->sigma2fs, Q, SigT, mueT, o));
            Sig = SigT; mue = mueT;
            RC(frk_quadform_spmv_bucketed(ctx, n, r, d_rowptr0, d_col0, d_val0, bk0, Sig, mue, q, smu));
            FRK_CHECK_CUDA(cudaMemsetAsync(A, 0, 2 * n * sizeof(double), ctx->stream));
            RC(frk_resid(ctx, M->m, M->p, M->X, M->alpha, M->Z, M->Ve, tmp));
            RC(frk_scatter_add(ctx, M->m, d_obs_bau, tmp, ytil));
            RC(frk_recip(ctx, M->m, M->Ve, inv_ve));
            RC(frk_scatter_add(ctx, M->m, d_obs_bau, inv_ve, A));
            RC(frk_predict_finalize(ctx, n, 1, M->sigma2fs, q, smu, xa, A, ytil, d_fs, muB, varB, sdB));
            FRK_LAUNCH(ctx, cp_case2_kernel, (unsigned)cdiv(mP * 32, 256), 256, 0, mP, d_cp_rowptr, d_cp_members, d_cp_weights, A, d_fs,
                       M->sigma2fs, w2, extra);
        }
        RC(frk_areal_aggregate(ctx, mP, d_cp_rowptr, d_cp_members, d_cp_weights, 1, 1, muB, n, d_mu));       // CP %*% BAUs$mu
        // CPM = CP %*% S0 (Case 2: rows weighted by c (1 - a)), then diag2(CPM %*% Cov, t(CPM))
        int64_t nnz = 0;
        RC(frk_csr_aggregate_count(ctx, mP, r, d_cp_rowptr, d_cp_members, d_rowptr0, d_col0, irp, &nnz));
        FRK_CHECK_CUDA(cudaMallocAsync((void **)&pcol, (size_t)std::max<int64_t>(nnz, 1) * sizeof(int), ctx->stream));
        FRK_CHECK_CUDA(cudaMallocAsync((void **)&pval, (size_t)std::max<int64_t>(nnz, 1) * sizeof(double), ctx->stream));
        RC(frk_csr_aggregate_fill(ctx, mP, r, d_cp_rowptr, d_cp_members, case1 ? d_cp_weights : w2, d_rowptr0, d_col0, d_val0, irp, pcol, pval));
        RC(frk_quadform_spmv(ctx, mP, r, irp, pcol, pval, Sig, nullptr, qP, nullptr));
        FRK_LAUNCH(ctx, cp_finish_kernel, (unsigned)std::min<int64_t>(cdiv(mP, 256), 4096), 256, 0, mP, qP, case1 ? (const double *)nullptr : extra,
                   d_var, d_sd);
        return 0;
    };
    const int rc = run();
    if (pcol) cudaFreeAsync(pcol, ctx->stream);
    if (pval) cudaFreeAsync(pval, ctx->stream);
    cudaFreeAsync(irp, ctx->stream);
    cudaFreeAsync(buf, ctx->stream);
    return rc;
}

// ---------------------------------------------------------------------------------
// predict() from HOST buffers: what the R `.Call` shim binds for .SRE.predict_EM (R/SREpredict.R:243-420).  Uploads S0 (host CSR,
// the dgRMatrix slots of object@S0), builds its row buckets, runs the matching device prediction and downloads mu / var / sd.
//   h_cp_* given (mP > 0)            -> prediction over polygons  (frk_model_predict_polygons, :378-419), outputs have length mP
//   h_cz_* given, !obs_fs, s2fs > 0  -> areal data, Case 2        (frk_model_predict_areal, :308-333), length n
//   else                             -> BAU level                 (frk_model_predict), length n; h_obs_bau needed for Case 2
// ---------------------------------------------------------------------------------
extern "C" int frk_model_predict_host(frk_model *M, int64_t n, const int *h_rowptr0, const int *h_col0, const double *h_val0,
                                      const double *h_XBAU, const double *h_fs, const int *h_obs_bau, const int *h_cz_rowptr,
                                      const int *h_cz_members, const double *h_cz_weights, int obs_fs, int64_t mP,
                                      const int *h_cp_rowptr, const int *h_cp_members, const double *h_cp_weights, double *h_mu,
                                      double *h_var, double *h_sd) {
    FRK_REQUIRE(M && h_rowptr0 && h_col0 && h_val0 && h_XBAU && h_mu && h_var && h_sd && n > 0, "frk_model_predict_host: bad argument");
    frk_ctx *ctx = M->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t nnz = h_rowptr0[n], m = M->m;
    const bool polygons = mP > 0 && h_cp_rowptr && h_cp_members && h_cp_weights;
    const bool areal = h_cz_rowptr && h_cz_members && h_cz_weights;
    const int64_t cz_nnz = areal ? h_cz_rowptr[m] : 0, cp_nnz = polygons ? h_cp_rowptr[mP] : 0, nout = polygons ? mP : n;
    std::vector<void *> bufs;
    frk_rowbuckets *bk = nullptr;
    auto up = [&](const void *h, size_t bytes, void **d) -> int {
        *d = nullptr;
        if (!h) return 0;
        FRK_CHECK_CUDA(cudaMallocAsync(d, std::max<size_t>(bytes, 16), st));
        bufs.push_back(*d);
        FRK_CHECK_CUDA(cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, st));
        return 0;
    };
    auto run = [&]() -> int {
        void *rp, *cl, *vl, *xb, *fs, *ob, *czr, *czm, *czw, *cpr, *cpm, *cpw, *out;
        RC(up(h_rowptr0, (size_t)(n + 1) * sizeof(int), &rp));
        RC(up(h_col0, (size_t)nnz * sizeof(int), &cl));
        RC(up(h_val0, (size_t)nnz * sizeof(double), &vl));
        RC(up(h_XBAU, (size_t)n * M->p * sizeof(double), &xb));
        RC(up(h_fs, (size_t)n * sizeof(double), &fs));
        RC(up(h_obs_bau, (size_t)std::max<int64_t>(m, 1) * sizeof(int), &ob));
        RC(up(areal ? h_cz_rowptr : nullptr, (size_t)(m + 1) * sizeof(int), &czr));
        RC(up(areal ? h_cz_members : nullptr, (size_t)cz_nnz * sizeof(int), &czm));
        RC(up(areal ? h_cz_weights : nullptr, (size_t)cz_nnz * sizeof(double), &czw));
        RC(up(polygons ? h_cp_rowptr : nullptr, (size_t)(mP + 1) * sizeof(int), &cpr));
        RC(up(polygons ? h_cp_members : nullptr, (size_t)cp_nnz * sizeof(int), &cpm));
        RC(up(polygons ? h_cp_weights : nullptr, (size_t)cp_nnz * sizeof(double), &cpw));
        FRK_CHECK_CUDA(cudaMallocAsync(&out, (size_t)3 * nout * sizeof(double), st));
        bufs.push_back(out);
        double *mu = (double *)out, *var = mu + nout, *sd = var + nout;
        RC(frk_rowbuckets_create(ctx, n, M->r, (const int *)rp, (const int *)cl, (const double *)vl, &bk));
        if (polygons) {
            RC(frk_model_predict_polygons(M, n, (const int *)rp, (const int *)cl, (const double *)vl, bk, (const double *)xb, (const double *)fs,
                                          (const int *)ob, obs_fs, mP, (const int *)cpr, (const int *)cpm, (const double *)cpw, cp_nnz, mu, var, sd));
        } else if (areal && !obs_fs && M->sigma2fs > 0.0) {
            RC(frk_model_predict_areal(M, n, (const int *)rp, (const int *)cl, (const double *)vl, bk, (const double *)xb, (const double *)fs,
                                       (const int *)czr, (const int *)czm, (const double *)czw, mu, var, sd));
        } else {
            RC(frk_model_predict(M, n, (const int *)rp, (const int *)cl, (const double *)vl, bk, (const double *)xb, (const double *)fs,
                                 (const int *)ob, obs_fs, mu, var, sd));
        }
        FRK_CHECK_CUDA(cudaMemcpyAsync(h_mu, mu, (size_t)nout * sizeof(double), cudaMemcpyDeviceToHost, st));
        FRK_CHECK_CUDA(cudaMemcpyAsync(h_var, var, (size_t)nout * sizeof(double), cudaMemcpyDeviceToHost, st));
        FRK_CHECK_CUDA(cudaMemcpyAsync(h_sd, sd, (size_t)nout * sizeof(double), cudaMemcpyDeviceToHost, st));
        FRK_CHECK_CUDA(cudaStreamSynchronize(st));
        return 0;
    };
    const int rc = run();
    if (rc) cudaStreamSynchronize(st);
    if (bk) frk_rowbuckets_destroy(bk);
    for (void *p : bufs) cudaFreeAsync(p, st);
    return rc;
}

// ---- the host optimisers, exported so that the CPU test-suite can pin them without a GPU --------------------
extern "C" int frk_optim_nelder_mead(frk_objective_fn fn, void *user, int n, const double *par, int maxit, double reltol,
                                     double *x_out, double *f_out, int *count_out, int *fail_out) {
    FRK_REQUIRE(fn && par && x_out && n > 0, "frk_optim_nelder_mead: bad argument");
    NMResult res;
    auto f = [&](const double *x, double &v) -> int { v = fn(user, n, x); return 0; };
    RC(nmmin(f, std::vector<double>(par, par + n), res, maxit, reltol));
    for (int i = 0; i < n; i++) x_out[i] = res.x[i];
    if (f_out) *f_out = res.f;
    if (count_out) *count_out = res.count;
    if (fail_out) *fail_out = res.fail;
    return 0;
}

// The speculative driver of the block-exponential M-step with a host objective: Nelder-Mead is replayed against the values known
// so far; whenever it asks for a new point, that point and up to width-1 points it may ask for next are evaluated as one round.
// Same result as frk_optim_nelder_mead by construction -- exported so that the CPU test-suite can check exactly that.
extern "C" int frk_optim_nelder_mead_speculative(frk_objective_fn fn, void *user, int n, const double *par, int maxit, double reltol,
                                                 int width, double *x_out, double *f_out, int *requests_out, int *rounds_out,
                                                 int *evals_out) {
    FRK_REQUIRE(fn && par && x_out && n > 0 && width >= 1 && width <= 64, "frk_optim_nelder_mead_speculative: bad argument");
    frk_model dummy;                                   // (statistics sink only; no device work)
    SpecNM spec{&dummy, nullptr, n, maxit, reltol, std::vector<double>(par, par + n)};
    NMResult res;
    int rounds = 0, evals = 0;
    for (;;) {
        std::vector<double> need;
        const int rc = spec.replay(need, res);
        if (rc == 0) break;
        if (rc != 99) return rc;
        std::vector<std::vector<double>> batch = {need};
        if (width > 1) spec.candidates(need.data(), (size_t)width, batch);
        for (auto &b : batch) {
            spec.pts.push_back(b);
            spec.fv.push_back(fn(user, n, b.data()));
            spec.owner.push_back(0);
        }
        rounds++;
        evals += (int)batch.size();
    }
    for (int i = 0; i < n; i++) x_out[i] = res.x[i];
    if (f_out) *f_out = res.f;
    if (requests_out) *requests_out = (int)dummy.nm_requests;
    if (rounds_out) *rounds_out = rounds;
    if (evals_out) *evals_out = evals;
    return 0;
}

extern "C" int frk_optim_zeroin(frk_objective_fn fn, void *user, double lower, double upper, double *root_out) {
    FRK_REQUIRE(fn && root_out, "frk_optim_zeroin: bad argument");
    auto f = [&](double x, double &v) -> int { v = fn(user, 1, &x); return 0; };
    return zeroin(f, lower, upper, *root_out);
}

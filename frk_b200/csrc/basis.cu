// eval_basis -> CSR on the device.
//
// Reference semantics (R/basisfns.R:709-748,1157-1161,1213-1250; R/geometryfns.R:202-219,
// 1717-1740): S[i,j] = phi_j(dist(s_i, c_j)), pattern = {phi != 0}, every arithmetic op
// individually rounded in the reference's order (no FMA contraction: the __d*_rn intrinsics
// below are never fused by nvcc).
//
// Design: the reference does N*r distance evaluations.  Here each resolution of the basis gets
// a uniform cell grid built once on the host at frk_basis_create(): for every cell, the ascending
// list of centroids whose support can reach the cell (a conservative superset).  A point looks
// up ONE cell per resolution and walks its candidate list, so work is O(nnz) and the row comes
// out in ascending column order.  One thread owns one point; a block stages its rows' (col,val)
// pairs in shared memory and then writes the block's contiguous CSR segment fully coalesced.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>

constexpr int MAXG = 8;

struct GroupDev {
    double x0, y0, hx, hy;
    double rhx, rhy;         // 1 / hx, 1 / hy: the cell of a point is floor((u - x0) * rhx); the lists carry a margin that covers the rounding
    int nx, ny;
    const int *cell_start;   // ncell+1
    const int *cand;         // ascending centroid ids per cell
    const double4 *cpk;      // fast bases: pk[cand[k]], sk[cand[k]] stored beside the candidate list (sequential, independent loads)
    const double2 *csk;
    int all;                 // 1: single cell holding every centroid of the group (non-compact types)
};

struct BasisDev {
    int r, ngroups, manifold, type, d;
    double radius;
    const double *c0, *c1, *c2;   // plane/line/STplane: x,(y),(t); sphere/STsphere: unit normal of the centroid
    const double *c3;             // STsphere: time coordinate of the centroid
    const double *scale;
    // fast path (bisquare on a Euclidean manifold): per centroid {x, y, t, T} and {scale, RN(1/scale)}, T = the smallest double
    // q with sqrt_rn(q) >= scale, so that  phi != 0  <=>  sqrt_rn(q) < scale  <=>  q < T  is decided without the square root
    const double4 *pk;
    const double2 *sk;
    int fast;
    GroupDev g[MAXG];
};

struct frk_basis {
    frk_ctx *ctx;
    BasisDev dev;
    int cols_sorted;               // candidate order == ascending column order across groups
    std::vector<void *> allocs;
};

// ---------------------------------------------------------------------------------
// host: cell lists
// ---------------------------------------------------------------------------------
namespace {

const double DEG = M_PI / 180.0;

struct HostGroup {
    double x0 = 0, y0 = 0, hx = 1, hy = 1;
    int nx = 1, ny = 1, all = 0;
    std::vector<int> start, cand;
};

void build_all(const std::vector<int> &ids, HostGroup &g) {
    g.all = 1;
    g.nx = g.ny = 1;
    g.start = {0, (int)ids.size()};
    g.cand = ids;
}

// plane / real line
void build_euclid(const std::vector<int> &ids, int d, const double *loc, int r, const double *scale, HostGroup &g) {
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300, smax = 0, smin = 1e300;
    for (int j : ids) {
        double R = scale[j], x = loc[j], y = d > 1 ? loc[j + r] : 0.0;
        xmin = std::min(xmin, x - R); xmax = std::max(xmax, x + R);
        ymin = std::min(ymin, y - R); ymax = std::max(ymax, y + R);
        smax = std::max(smax, R); smin = std::min(smin, R);
    }
    double w = std::max(xmax - xmin, 1e-300), hgt = d > 1 ? std::max(ymax - ymin, 1e-300) : 1.0;
    double h = smax / 4.0;
    const double cap = 4.0e6;   // cells per resolution
    if (d > 1) { if ((w / h) * (hgt / h) > cap) h = std::sqrt(w * hgt / cap); }
    else       { if (w / h > cap) h = w / cap; }
    g.x0 = xmin; g.y0 = d > 1 ? ymin : 0.0; g.hx = h; g.hy = d > 1 ? h : 1.0;
    g.nx = std::max(1, (int)std::ceil(w / h));
    g.ny = d > 1 ? std::max(1, (int)std::ceil(hgt / h)) : 1;
    const int64_t ncell = (int64_t)g.nx * g.ny;
    std::vector<int> cnt(ncell + 1, 0);
    const double m = 1e-7 * h;   // margin: absorbs rounding in the point's cell assignment (floor of a product with 1 / h, up to 4e6 cells)
    for (int pass = 0; pass < 2; pass++) {
        for (int j : ids) {
            double R = scale[j], x = loc[j], y = d > 1 ? loc[j + r] : 0.0;
            int ix0 = std::max(0, (int)std::floor((x - R - g.x0) / h) - 1);
            int ix1 = std::min(g.nx - 1, (int)std::floor((x + R - g.x0) / h) + 1);
            int iy0 = 0, iy1 = 0;
            if (d > 1) {
                iy0 = std::max(0, (int)std::floor((y - R - g.y0) / h) - 1);
                iy1 = std::min(g.ny - 1, (int)std::floor((y + R - g.y0) / h) + 1);
            }
            for (int iy = iy0; iy <= iy1; iy++) {
                double dy = 0;
                if (d > 1) {
                    double lo = g.y0 + iy * h - m, hi = g.y0 + (iy + 1) * h + m;
                    dy = y < lo ? lo - y : (y > hi ? y - hi : 0.0);
                }
                for (int ix = ix0; ix <= ix1; ix++) {
                    double lo = g.x0 + ix * h - m, hi = g.x0 + (ix + 1) * h + m;
                    double dx = x < lo ? lo - x : (x > hi ? x - hi : 0.0);
                    if (dx * dx + dy * dy <= R * R * (1 + 1e-12)) {
                        int64_t c = (int64_t)iy * g.nx + ix;
                        if (pass == 0) cnt[c + 1]++;
                        else g.cand[g.start[c] + cnt[c]++] = j;
                    }
                }
            }
        }
        if (pass == 0) {
            g.start.assign(ncell + 1, 0);
            for (int64_t c = 0; c < ncell; c++) g.start[c + 1] = g.start[c] + cnt[c + 1];
            g.cand.resize(g.start[ncell]);
            std::fill(cnt.begin(), cnt.end(), 0);
        }
    }
}

inline double ang_between(double lon1, double lat1, double lon2, double lat2) {   // degrees in, radians out
    double p1 = lat1 * DEG, p2 = lat2 * DEG, dl = (lon1 - lon2) * DEG;
    double c = std::sin(p1) * std::sin(p2) + std::cos(p1) * std::cos(p2) * std::cos(dl);
    return std::acos(std::max(-1.0, std::min(1.0, c)));
}

// sphere: lon/lat cell grid over [-180,180) x [-90,90]
void build_sphere(const std::vector<int> &ids, const double *loc, int r, const double *scale, double radius, HostGroup &g) {
    double tmax = 0;
    for (int j : ids) tmax = std::max(tmax, scale[j] / radius);
    double h = std::max(std::min(tmax, M_PI) / DEG / 4.0, 0.25);
    g.nx = std::max(1, (int)std::ceil(360.0 / h));
    g.ny = std::max(1, (int)std::ceil(180.0 / h));
    g.hx = 360.0 / g.nx; g.hy = 180.0 / g.ny; g.x0 = -180.0; g.y0 = -90.0;
    const int64_t ncell = (int64_t)g.nx * g.ny;
    // circumradius of a cell (angle centre -> farthest corner) depends on the row only
    std::vector<double> rho(g.ny);
    for (int iy = 0; iy < g.ny; iy++) {
        double la0 = g.y0 + iy * g.hy, la1 = la0 + g.hy, lam = 0.5 * (la0 + la1);
        rho[iy] = std::max(ang_between(0, lam, 0.5 * g.hx, la0), ang_between(0, lam, 0.5 * g.hx, la1));
    }
    std::vector<int> cnt(ncell + 1, 0);
    const double margin = 1e-7;
    for (int pass = 0; pass < 2; pass++) {
        for (int j : ids) {
            double th = scale[j] / radius;
            double lon = loc[j], lat = loc[j + r];
            lon = lon - 360.0 * std::floor((lon + 180.0) / 360.0);
            double thd = th / DEG;
            int iy0 = std::max(0, (int)std::floor((lat - thd - g.y0) / g.hy) - 1);
            int iy1 = std::min(g.ny - 1, (int)std::floor((lat + thd - g.y0) / g.hy) + 1);
            bool all_lon = (std::fabs(lat) + thd + 2 * g.hy >= 90.0) || th >= M_PI / 2;
            int ixc = (int)std::floor((lon - g.x0) / g.hx), half = g.nx;
            if (!all_lon) {
                double s = std::sin(th) / std::cos(lat * DEG);
                if (s >= 1.0) all_lon = true;
                else half = (int)std::ceil(std::asin(s) / DEG / g.hx) + 2;
            }
            if (all_lon || 2 * half + 1 >= g.nx) { half = g.nx / 2 + 1; }
            for (int iy = iy0; iy <= iy1; iy++) {
                double lam = g.y0 + (iy + 0.5) * g.hy;
                int nvis = std::min(2 * half + 1, g.nx);
                for (int k = 0; k < nvis; k++) {
                    int ix = ((ixc - half + k) % g.nx + g.nx) % g.nx;
                    double lom = g.x0 + (ix + 0.5) * g.hx;
                    if (ang_between(lon, lat, lom, lam) <= th + rho[iy] + margin) {
                        int64_t c = (int64_t)iy * g.nx + ix;
                        if (pass == 0) cnt[c + 1]++;
                        else g.cand[g.start[c] + cnt[c]++] = j;
                    }
                }
            }
        }
        if (pass == 0) {
            g.start.assign(ncell + 1, 0);
            for (int64_t c = 0; c < ncell; c++) g.start[c + 1] = g.start[c] + cnt[c + 1];
            g.cand.resize(g.start[ncell]);
            std::fill(cnt.begin(), cnt.end(), 0);
        }
    }
    // candidates were appended in ascending j only if no wrap-around duplicates: sort+unique per cell
    for (int64_t c = 0; c < ncell; c++) std::sort(g.cand.begin() + g.start[c], g.cand.begin() + g.start[c + 1]);
}

template <typename T>
int upload(frk_basis *b, const std::vector<T> &v, const T **out) {
    void *p = nullptr;
    FRK_CHECK_CUDA(cudaMalloc(&p, std::max<size_t>(v.size() * sizeof(T), 8)));
    b->allocs.push_back(p);
    if (!v.empty())
        FRK_CHECK_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T *)p;
    return 0;
}

}  // namespace

extern "C" int frk_basis_create(frk_ctx *ctx, int manifold, double radius, int type, int d, int r,
                                const double *h_loc, const double *h_scale, const int *h_res, frk_basis **out) {
    FRK_REQUIRE(ctx && out && h_loc && h_scale, "frk_basis_create: NULL argument");
    FRK_REQUIRE(manifold >= FRK_REAL_LINE && manifold <= FRK_STSPHERE,
                "frk_basis_create: unsupported manifold %d (only real_line, plane, sphere, STplane, STsphere; custom measures are refused, no CPU fallback)", manifold);
    FRK_REQUIRE(type >= FRK_BISQUARE && type <= FRK_MATERN32, "frk_basis_create: unknown basis type %d", type);
    const bool on_sphere = manifold == FRK_SPHERE || manifold == FRK_STSPHERE;
    int dexp = manifold == FRK_REAL_LINE ? 1 : (manifold >= FRK_STPLANE ? 3 : 2);
    FRK_REQUIRE(d == dexp, "number of columns in loc (%d) needs to be the same as the number of manifold dimensions (%d)", d, dexp);
    FRK_REQUIRE(r > 0, "frk_basis_create: need at least one basis function");
    for (int j = 0; j < r; j++)
        FRK_REQUIRE(h_scale[j] > 0, "The scale parameter needs to be greater than zero (basis %d)", j + 1);
    FRK_CHECK_CUDA(cudaSetDevice(ctx->device));

    frk_basis *b = new frk_basis();
    b->ctx = ctx;
    BasisDev &D = b->dev;
    D.r = r; D.manifold = manifold; D.type = type; D.d = d; D.radius = radius;

    // centroid table
    std::vector<double> c0(r), c1(r, 0.0), c2(r, 0.0), c3(r, 0.0), sc(h_scale, h_scale + r);
    if (on_sphere) {
        for (int j = 0; j < r; j++) {
            // x2 * pi/180; n2 = (cos(lat)cos(lon), cos(lat)sin(lon), sin(lat))   R/geometryfns.R:1727,1732
            double lon = (h_loc[j] * M_PI) / 180.0, lat = (h_loc[j + r] * M_PI) / 180.0;
            double cl = std::cos(lat);
            c0[j] = cl * std::cos(lon); c1[j] = cl * std::sin(lon); c2[j] = std::sin(lat);
            if (d > 2) c3[j] = h_loc[j + 2 * (size_t)r];
        }
    } else {
        for (int j = 0; j < r; j++) {
            c0[j] = h_loc[j];
            if (d > 1) c1[j] = h_loc[j + r];
            if (d > 2) c2[j] = h_loc[j + 2 * (size_t)r];
        }
    }
    int rc = 0;
    rc |= upload(b, c0, &D.c0); rc |= upload(b, c1, &D.c1); rc |= upload(b, c2, &D.c2); rc |= upload(b, c3, &D.c3);
    rc |= upload(b, sc, &D.scale);
    D.fast = (type == FRK_BISQUARE && !on_sphere) ? 1 : 0;
    D.pk = nullptr; D.sk = nullptr;
    std::vector<double4> h_pk;
    std::vector<double2> h_sk;
    if (D.fast) {
        std::vector<double4> &pk = h_pk;
        std::vector<double2> &sk = h_sk;
        pk.resize(r); sk.resize(r);
        for (int j = 0; j < r; j++) {
            const double R = sc[j];
            double T = R * R;
            while (std::sqrt(T) >= R && T > 0) T = std::nextafter(T, 0.0);
            while (std::sqrt(T) < R) T = std::nextafter(T, INFINITY);
            pk[j] = make_double4(c0[j], c1[j], c2[j], T);
            sk[j] = make_double2(R, 1.0 / R);
        }
        rc |= upload(b, pk, &D.pk);
        rc |= upload(b, sk, &D.sk);
    }
    if (rc) { frk_basis_destroy(b); return rc; }

    // groups = resolutions, in order of first appearance
    std::vector<int> keys;
    std::vector<std::vector<int>> members;
    for (int j = 0; j < r; j++) {
        int key = h_res ? h_res[j] : 1;
        size_t g = 0;
        for (; g < keys.size(); g++) if (keys[g] == key) break;
        if (g == keys.size()) { keys.push_back(key); members.emplace_back(); }
        members[g].push_back(j);
    }
    if ((int)keys.size() > MAXG || type != FRK_BISQUARE) {   // non-compact types: every centroid is a candidate
        std::vector<int> all(r);
        for (int j = 0; j < r; j++) all[j] = j;
        members.assign(1, all);
    }
    D.ngroups = (int)members.size();
    b->cols_sorted = 1;
    for (size_t g = 1; g < members.size(); g++)
        if (members[g].front() < members[g - 1].back()) b->cols_sorted = 0;
    for (size_t g = 0; g < members.size(); g++) {
        HostGroup hg;
        if (type != FRK_BISQUARE) build_all(members[g], hg);
        // space-time: dist >= spatial dist, so the spatial cell list (time ignored) is a valid superset
        else if (on_sphere) build_sphere(members[g], h_loc, r, h_scale, radius, hg);
        else build_euclid(members[g], std::min(d, 2), h_loc, r, h_scale, hg);
        GroupDev &G = D.g[g];
        G.x0 = hg.x0; G.y0 = hg.y0; G.hx = hg.hx; G.hy = hg.hy; G.rhx = 1.0 / hg.hx; G.rhy = 1.0 / hg.hy; G.nx = hg.nx; G.ny = hg.ny; G.all = hg.all;
        rc |= upload(b, hg.start, &G.cell_start);
        rc |= upload(b, hg.cand, &G.cand);
        G.cpk = nullptr; G.csk = nullptr;
        if (D.fast) {
            std::vector<double4> cpk(hg.cand.size());
            std::vector<double2> csk(hg.cand.size());
            for (size_t k = 0; k < hg.cand.size(); k++) { cpk[k] = h_pk[hg.cand[k]]; csk[k] = h_sk[hg.cand[k]]; }
            rc |= upload(b, cpk, &G.cpk);
            rc |= upload(b, csk, &G.csk);
        }
        if (rc) { frk_basis_destroy(b); return rc; }
    }
    *out = b;
    return 0;
}

extern "C" int frk_basis_destroy(frk_basis *b) {
    if (!b) return 0;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    for (void *p : b->allocs) cudaFree(p);
    delete b;
    return 0;
}

extern "C" int frk_basis_n(const frk_basis *b) { return b ? b->dev.r : -1; }
extern "C" int frk_basis_dim(const frk_basis *b) { return b ? b->dev.d : -1; }

// ---------------------------------------------------------------------------------
// device: distance + basis function in the reference's operation order
// ---------------------------------------------------------------------------------
struct PointCoord { double a, b, c, t; double u, v; };   // (a,b,c,t) distance operand; (u,v) cell-lookup coords

template <int MAN> struct man_traits {
    static constexpr bool sphere = MAN == FRK_SPHERE || MAN == FRK_STSPHERE;
    static constexpr bool line = MAN == FRK_REAL_LINE;
};

template <int MAN>
__device__ __forceinline__ PointCoord load_point(const double *__restrict__ s, int64_t ld, int64_t i) {
    PointCoord p;
    p.t = 0;
    if (man_traits<MAN>::sphere) {
        double lon = s[i], lat = s[i + ld];
        double lo = __ddiv_rn(__dmul_rn(lon, M_PI), 180.0);    // x1 * pi/180, R/geometryfns.R:1726
        double la = __ddiv_rn(__dmul_rn(lat, M_PI), 180.0);
        double sl, cl, so, co;
        sincos(la, &sl, &cl);
        sincos(lo, &so, &co);
        p.a = __dmul_rn(cl, co); p.b = __dmul_rn(cl, so); p.c = sl;   // :1731
        p.u = lon - 360.0 * floor((lon + 180.0) / 360.0);
        p.v = lat;
        if (MAN == FRK_STSPHERE) p.t = s[i + 2 * ld];
    } else if (MAN == FRK_STPLANE) {
        p.a = s[i]; p.b = s[i + ld]; p.c = s[i + 2 * ld]; p.u = p.a; p.v = p.b;
    } else if (MAN == FRK_PLANE) {
        p.a = s[i]; p.b = s[i + ld]; p.c = 0; p.u = p.a; p.v = p.b;
    } else {
        p.a = s[i]; p.b = 0; p.c = 0; p.u = p.a; p.v = 0;
    }
    return p;
}

template <int MAN>
__device__ __forceinline__ double basis_dist(const PointCoord &p, const BasisDev &B, int j) {
    if (man_traits<MAN>::sphere) {
        // tcrossprod(n1,n2) in reference-BLAS order ((a1 b1)+a2 b2)+a3 b3, clamp, R*acos   :1733-1738
        double t = __dadd_rn(__dadd_rn(__dmul_rn(p.a, __ldg(B.c0 + j)), __dmul_rn(p.b, __ldg(B.c1 + j))),
                             __dmul_rn(p.c, __ldg(B.c2 + j)));
        if (fabs(t) > 1.0) t = t > 0 ? 1.0 : -1.0;
        const double sd = __dmul_rn(B.radius, acos(t));
        if (MAN == FRK_SPHERE) return sd;
        // gc_dist_time: sqrt(spatdist^2 + tdist^2), tdist = distR(t1,t2) = sqrt((t1-t2)^2)   R/geometryfns.R:181-189
        const double dt = __dsub_rn(p.t, __ldg(B.c3 + j));
        const double td = __dsqrt_rn(__dmul_rn(dt, dt));
        return __dsqrt_rn(__dadd_rn(__dmul_rn(sd, sd), __dmul_rn(td, td)));
    } else if (MAN == FRK_STPLANE) {
        // Euclid_dist(dim=3L): distR over three columns, Reduce("+") left to right   R/geometryfns.R:217-218
        double dx = __dsub_rn(p.a, __ldg(B.c0 + j)), dy = __dsub_rn(p.b, __ldg(B.c1 + j)), dt = __dsub_rn(p.c, __ldg(B.c2 + j));
        return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dt, dt)));
    } else if (MAN == FRK_PLANE) {
        // sqrt(Reduce("+", (x1_d - x2_d)^2)), R/geometryfns.R:217-218
        double dx = __dsub_rn(p.a, __ldg(B.c0 + j)), dy = __dsub_rn(p.b, __ldg(B.c1 + j));
        return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    } else {
        double dx = __dsub_rn(p.a, __ldg(B.c0 + j));
        return __dsqrt_rn(__dmul_rn(dx, dx));
    }
}

template <int TYPE>
__device__ __forceinline__ double basis_phi(double y, double sc) {
    if (TYPE == FRK_BISQUARE) {            // (1-(y/R)^2)^2 * (y < R)        R/basisfns.R:1227-1228
        double t = __ddiv_rn(y, sc);
        double u = __dsub_rn(1.0, __dmul_rn(t, t));
        return y < sc ? __dmul_rn(u, u) : 0.0;
    } else if (TYPE == FRK_GAUSSIAN) {     // exp(-0.5*dist^2/std^2)          :1217-1218
        double q = __dmul_rn(y, y);
        return exp(__ddiv_rn(__dmul_rn(-0.5, q), __dmul_rn(sc, sc)));
    } else if (TYPE == FRK_EXP) {          // exp(-y/tau)                     :1238
        return exp(__ddiv_rn(-y, sc));
    } else {                               // (1+sqrt(3)y/kappa) exp(-sqrt(3)y/kappa)   :1248
        const double s3 = 1.7320508075688772;
        double a = __ddiv_rn(__dmul_rn(s3, y), sc);
        return __dmul_rn(__dadd_rn(1.0, a), exp(-a));
    }
}

template <int MAN>
__device__ __forceinline__ void cand_range(const GroupDev &G, const PointCoord &p, int &beg, int &end) {
    if (G.all) { beg = G.cell_start[0]; end = G.cell_start[1]; return; }
    double fx = floor((p.u - G.x0) * G.rhx);
    double fy = MAN == FRK_REAL_LINE ? 0.0 : floor((p.v - G.y0) * G.rhy);
    if (man_traits<MAN>::sphere) {   // closed domain: clamp the upper edges (lat = 90) into the last cell
        if (fx >= G.nx) fx = G.nx - 1;
        if (fy >= G.ny) fy = G.ny - 1;
        if (fx < 0) fx = 0;
        if (fy < 0) fy = 0;
    }
    if (!(fx >= 0 && fx < G.nx && fy >= 0 && fy < G.ny)) { beg = end = 0; return; }   // also catches NaN
    int64_t c = (int64_t)fy * G.nx + (int64_t)fx;
    beg = __ldg(G.cell_start + c);
    end = __ldg(G.cell_start + c + 1);
}

// pass 1: per-row non-zero count
template <int MAN, int TYPE>
__global__ void __launch_bounds__(256) eval_count_kernel(BasisDev B, int64_t n, const double *__restrict__ s, int64_t ld,
                                                         int *__restrict__ counts) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        PointCoord p = load_point<MAN>(s, ld, i);
        int cnt = 0;
        for (int g = 0; g < B.ngroups; g++) {
            int beg, end;
            cand_range<MAN>(B.g[g], p, beg, end);
            const int *cand = B.g[g].cand;
            for (int k = beg; k < end; k++) {
                int j = __ldg(cand + k);
                double y = basis_dist<MAN>(p, B, j);
                double v = basis_phi<TYPE>(y, __ldg(B.scale + j));
                cnt += (v != 0.0);
            }
        }
        counts[i] = cnt;
    }
}

// pass 2: fill.  FILL_T rows per block; rows staged in shared memory when the block's segment fits.
constexpr int FILL_T = 128;
constexpr int FILL_CAP = 6144;   // staged entries per block (12 B each = 72 KB)

template <int MAN, int TYPE>
__global__ void __launch_bounds__(FILL_T) eval_fill_kernel(BasisDev B, int64_t n, const double *__restrict__ s, int64_t ld,
                                                           const int *__restrict__ rowptr, int *__restrict__ col,
                                                           double *__restrict__ val, int normalise, int cols_sorted) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sval = reinterpret_cast<double *>(smem_raw);
    int *scol = reinterpret_cast<int *>(sval + FILL_CAP);
    const int64_t row0 = (int64_t)blockIdx.x * FILL_T;
    const int64_t rowN = min(n, row0 + FILL_T);
    const int seg0 = rowptr[row0];
    const int seg = rowptr[rowN] - seg0;
    const bool staged = seg <= FILL_CAP;
    const int64_t i = row0 + threadIdx.x;
    if (i < n) {
        PointCoord p = load_point<MAN>(s, ld, i);
        const int my0 = rowptr[i] - seg0;
        int w = my0;
        double ss = 0.0;
        for (int g = 0; g < B.ngroups; g++) {
            int beg, end;
            cand_range<MAN>(B.g[g], p, beg, end);
            const int *cand = B.g[g].cand;
            for (int k = beg; k < end; k++) {
                int j = __ldg(cand + k);
                double y = basis_dist<MAN>(p, B, j);
                double v = basis_phi<TYPE>(y, __ldg(B.scale + j));
                if (v != 0.0) {
                    if (staged) { scol[w] = j; sval[w] = v; }
                    else        { col[seg0 + w] = j; val[seg0 + w] = v; }
                    ss = __dadd_rn(ss, __dmul_rn(v, v));
                    w++;
                }
            }
        }
        int *cc = staged ? scol + my0 : col + seg0 + my0;
        double *vv = staged ? sval + my0 : val + seg0 + my0;
        const int k = w - my0;
        if (!cols_sorted) {   // resolutions interleaved in column space: insertion sort (rows are short)
            for (int a = 1; a < k; a++) {
                int cj = cc[a]; double vj = vv[a]; int bb = a - 1;
                while (bb >= 0 && cc[bb] > cj) { cc[bb + 1] = cc[bb]; vv[bb + 1] = vv[bb]; bb--; }
                cc[bb + 1] = cj; vv[bb + 1] = vj;
            }
            ss = 0.0;
            for (int a = 0; a < k; a++) ss = __dadd_rn(ss, __dmul_rn(vv[a], vv[a]));
        }
        if (normalise) {     // xx = sqrt(rowSums(S0*S0)); xx[xx==0] = 1; S0/xx     R/SRE.R:416-421
            double xx = __dsqrt_rn(ss);
            if (xx == 0.0) xx = 1.0;
            for (int a = 0; a < k; a++) vv[a] = __ddiv_rn(vv[a], xx);
        }
    }
    if (staged) {
        __syncthreads();
        for (int t = threadIdx.x; t < seg; t += FILL_T) {
            col[seg0 + t] = scol[t];
            val[seg0 + t] = sval[t];
        }
    }
}


// ---------------------------------------------------------------------------------
// Second-generation kernels.  The two-pass structure stays (the caller sizes col / val from the count), but
//  * bisquare functions on a Euclidean manifold ("fast" bases) are COUNTED without a square root or a division: the pattern is
//    {sqrt_rn(q) < scale} = {q < T_j} with the threshold T_j tabulated at frk_basis_create, so the count pass is ~12 instructions per
//    candidate and the expensive arithmetic (sqrt, /, the polynomial) runs once, in the fill pass, for the hits only;
//  * y / scale is formed with the tabulated correctly rounded reciprocal and two Markstein correction steps (5 FMAs, result
//    = RN(y / scale) exactly as __ddiv_rn gives it) instead of the 20-instruction division sequence;
//  * the fill pass stages at most FCAP entries per row in shared memory (6.4 KB per warp instead of 72 KB per 128 rows: 32 resident
//    warps per SM instead of 12) and flushes them, a half-warp per row, as contiguous runs of the row's CSR segment;
//  * the row normalisation re-reads the warp's own segment (still in L2) once the row sums are complete.
// ---------------------------------------------------------------------------------
constexpr int FW = 4;                 // warps per block (static shared memory: 26 KB)
constexpr int FCAP = 16, FST = FCAP + 1;
constexpr int FU = 2;                 // candidates evaluated together (fast path)
constexpr unsigned FULL = 0xffffffffu;

// RN(a / b) from rb = RN(1 / b): q0 = RN(a rb) is within 1.5 ulp, one correction makes it faithful, and for a faithful quotient
// Markstein's theorem gives RN(q + (a - b q) rb) = RN(a / b).  (No underflow / overflow for the operands met here.)
__device__ __forceinline__ double div_by_recip(double a, double b, double rb) {
    double q = __dmul_rn(a, rb);
    double r = __fma_rn(-q, b, a);
    q = __fma_rn(r, rb, q);
    r = __fma_rn(-q, b, a);
    return __fma_rn(r, rb, q);
}

__device__ __forceinline__ double4 ldg_pk(const double4 *p) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

template <int MAN>
__device__ __forceinline__ double fast_dist2(const PointCoord &p, const double4 &c) {   // the argument of distR's sqrt, same rounding
    const double dx = __dsub_rn(p.a, c.x);
    if (MAN == FRK_REAL_LINE) return __dmul_rn(dx, dx);
    const double dy = __dsub_rn(p.b, c.y);
    const double q = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    if (MAN == FRK_PLANE) return q;
    const double dt = __dsub_rn(p.c, c.z);
    return __dadd_rn(q, __dmul_rn(dt, dt));
}

template <int MAN>
__global__ void __launch_bounds__(256) eval_count_fast_kernel(BasisDev B, int64_t n, const double *__restrict__ s, int64_t ld,
                                                              int *__restrict__ counts) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const PointCoord p = load_point<MAN>(s, ld, i);
        int cnt = 0;
        for (int g = 0; g < B.ngroups; g++) {
            int beg, end;
            cand_range<MAN>(B.g[g], p, beg, end);
            const double4 *cp = B.g[g].cpk;
            for (int k = beg; k < end; k++) {
                const double4 c = ldg_pk(cp + k);
                cnt += fast_dist2<MAN>(p, c) < c.w;
            }
        }
        counts[i] = cnt;
    }
}

template <int MAN, int TYPE, bool FAST>
__global__ void __launch_bounds__(FW * 32, 8) eval_fill_warp_kernel(BasisDev B, int64_t n, const double *__restrict__ s, int64_t ld,
                                                                    const int *__restrict__ rowptr, int *__restrict__ col,
                                                                    double *__restrict__ val, int normalise) {
    __shared__ __align__(16) double sval_all[FW * 32 * FST];
    __shared__ int scol_all[FW * 32 * FST];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, half = lane >> 4, l16 = lane & 15;
    double *sval = sval_all + wid * 32 * FST;
    int *scol = scol_all + wid * 32 * FST;
    double2 *sx = reinterpret_cast<double2 *>(sval);      // (xx, 1/xx) per row; overlays the staging area once it has been flushed
    for (int64_t row0 = ((int64_t)blockIdx.x * FW + wid) * 32; row0 < n; row0 += (int64_t)gridDim.x * FW * 32) {
        const int64_t i = row0 + lane;
        const bool active = i < n;
        PointCoord p;
        p.a = p.b = p.c = p.t = p.u = p.v = 0.0;
        int base = 0, len = 0;
        if (active) {
            p = load_point<MAN>(s, ld, i);
            base = rowptr[i];
            len = rowptr[i + 1] - base;
        }
        int cnt = 0, dst = base;                       // staged entries; where the next flushed entry of this row goes
        double ss = 0.0;
        auto flush = [&]() {
            __syncwarp();
            const double *sv = sval + half * FST + l16;
            const int *sc = scol + half * FST + l16;
#pragma unroll
            for (int a = 0; a < 16; a++) {
                const int l = 2 * a + half;
                const int c = __shfl_sync(FULL, cnt, l), b = __shfl_sync(FULL, dst, l);
                if (l16 < c) {
                    col[b + l16] = sc[a * 2 * FST];
                    val[b + l16] = sv[a * 2 * FST];
                }
            }
            dst += cnt;
            cnt = 0;
            __syncwarp();
        };
        auto stage = [&](double v, int j) {
            sval[lane * FST + cnt] = v;
            scol[lane * FST + cnt] = j;
            ss = __dadd_rn(ss, __dmul_rn(v, v));
            cnt++;
        };
        for (int g = 0; g < B.ngroups; g++) {
            int beg = 0, end = 0;
            if (active) cand_range<MAN>(B.g[g], p, beg, end);
            const int nl = end - beg;
            const int nmax = __reduce_max_sync(FULL, nl);
            const int *cj = B.g[g].cand + beg;
            if (FAST) {
                const double4 *cp = B.g[g].cpk + beg;
                const double2 *cs = B.g[g].csk + beg;
                for (int t = 0; t < nmax; t += FU) {     // FU candidates between flush checks: their loads and their sqrt / division
                    double4 c[FU];                       // chains are issued together
                    double q[FU];
                    bool any = false;
#pragma unroll
                    for (int u = 0; u < FU; u++) {
                        c[u] = make_double4(0, 0, 0, -1.0);
                        if (t + u < nl) c[u] = ldg_pk(cp + t + u);
                    }
#pragma unroll
                    for (int u = 0; u < FU; u++) {
                        q[u] = fast_dist2<MAN>(p, c[u]);
                        any |= q[u] < c[u].w;
                    }
                    if (any) {
                        double v[FU];
#pragma unroll
                        for (int u = 0; u < FU; u++) {
                            double2 sc = make_double2(1.0, 1.0);
                            if (q[u] < c[u].w) sc = __ldg(cs + t + u);
                            const double y = __dsqrt_rn(q[u]);
                            const double tt = div_by_recip(y, sc.x, sc.y);
                            const double w = __dsub_rn(1.0, __dmul_rn(tt, tt));
                            v[u] = __dmul_rn(w, w);
                        }
#pragma unroll
                        for (int u = 0; u < FU; u++)
                            if (q[u] < c[u].w) stage(v[u], __ldg(cj + t + u));
                    }
                    if (__any_sync(FULL, cnt > FCAP - FU)) flush();
                }
            } else {
                for (int t = 0; t < nmax; t++) {
                    if (t < nl) {
                        const int j = __ldg(cj + t);
                        const double y = basis_dist<MAN>(p, B, j);
                        const double v = basis_phi<TYPE>(y, __ldg(B.scale + j));
                        if (v != 0.0) stage(v, j);
                    }
                    if (__any_sync(FULL, cnt > FCAP - 1)) flush();
                }
            }
        }
        flush();
        if (normalise) {     // xx = sqrt(rowSums(S0*S0)); xx[xx==0] = 1; S0/xx     R/SRE.R:416-421
            double xx = __dsqrt_rn(ss);
            if (xx == 0.0) xx = 1.0;
            sx[lane] = make_double2(xx, __ddiv_rn(1.0, xx));
            const int maxlen = __reduce_max_sync(FULL, len);
            __syncwarp();
            for (int e0 = 0; e0 < maxlen; e0 += 48) {
#pragma unroll 1
                for (int a0 = 0; a0 < 16; a0 += 4) {      // four rows per half-warp at a time: twelve loads in flight per lane
                    double vv[4][3];
                    int bb[4], LL[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int l = 2 * (a0 + u) + half;
                        LL[u] = __shfl_sync(FULL, len, l) - e0 - l16;
                        bb[u] = __shfl_sync(FULL, base, l) + e0 + l16;
#pragma unroll
                        for (int c = 0; c < 3; c++) vv[u][c] = 16 * c < LL[u] ? __ldcg(val + bb[u] + 16 * c) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const double2 X = sx[2 * (a0 + u) + half];
#pragma unroll
                        for (int c = 0; c < 3; c++)
                            if (16 * c < LL[u]) val[bb[u] + 16 * c] = div_by_recip(vv[u][c], X.x, X.y);
                    }
                }
            }
            __syncwarp();
        }
    }
}

// FRK_EVAL_LEGACY=1 selects the first-generation kernels (full evaluation in both passes, block-staged fill) for A/B timing
static bool eval_legacy() {
    static const bool v = getenv("FRK_EVAL_LEGACY") != nullptr;
    return v;
}

template <int MAN>
static int launch_count(frk_ctx *ctx, const frk_basis *b, int64_t n, const double *s, int64_t ld, int *counts) {
    unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), (int64_t)ctx->sm_count * 32);
    if (b->dev.fast && !eval_legacy()) {
        FRK_LAUNCH(ctx, (eval_count_fast_kernel<MAN>), grid, 256, 0, b->dev, n, s, ld, counts);
        return 0;
    }
    switch (b->dev.type) {
        case FRK_BISQUARE: FRK_LAUNCH(ctx, (eval_count_kernel<MAN, FRK_BISQUARE>), grid, 256, 0, b->dev, n, s, ld, counts); break;
        case FRK_GAUSSIAN: FRK_LAUNCH(ctx, (eval_count_kernel<MAN, FRK_GAUSSIAN>), grid, 256, 0, b->dev, n, s, ld, counts); break;
        case FRK_EXP:      FRK_LAUNCH(ctx, (eval_count_kernel<MAN, FRK_EXP>), grid, 256, 0, b->dev, n, s, ld, counts); break;
        default:           FRK_LAUNCH(ctx, (eval_count_kernel<MAN, FRK_MATERN32>), grid, 256, 0, b->dev, n, s, ld, counts); break;
    }
    return 0;
}

template <int MAN, int TYPE>
static int launch_fill_t(frk_ctx *ctx, const frk_basis *b, int64_t n, const double *s, int64_t ld, const int *rowptr,
                         int *col, double *val, int normalise) {
    if (b->cols_sorted && !eval_legacy()) {
        const unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, FW * 32), (int64_t)ctx->sm_count * 8 * 8);
        if (b->dev.fast && TYPE == FRK_BISQUARE && !man_traits<MAN>::sphere)
            FRK_LAUNCH(ctx, (eval_fill_warp_kernel<MAN, TYPE, true>), grid, FW * 32, 0, b->dev, n, s, ld, rowptr, col, val, normalise);
        else
            FRK_LAUNCH(ctx, (eval_fill_warp_kernel<MAN, TYPE, false>), grid, FW * 32, 0, b->dev, n, s, ld, rowptr, col, val, normalise);
        return 0;
    }
    const size_t smem = FILL_CAP * (sizeof(double) + sizeof(int));
    static bool attr_set = false;
    if (!attr_set) {
        FRK_CHECK_CUDA(cudaFuncSetAttribute(eval_fill_kernel<MAN, TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    FRK_LAUNCH(ctx, (eval_fill_kernel<MAN, TYPE>), (unsigned)cdiv(n, FILL_T), FILL_T, smem, b->dev, n, s, ld, rowptr, col, val,
               normalise, b->cols_sorted);
    return 0;
}

template <int MAN>
static int launch_fill(frk_ctx *ctx, const frk_basis *b, int64_t n, const double *s, int64_t ld, const int *rowptr,
                       int *col, double *val, int normalise) {
    switch (b->dev.type) {
        case FRK_BISQUARE: return launch_fill_t<MAN, FRK_BISQUARE>(ctx, b, n, s, ld, rowptr, col, val, normalise);
        case FRK_GAUSSIAN: return launch_fill_t<MAN, FRK_GAUSSIAN>(ctx, b, n, s, ld, rowptr, col, val, normalise);
        case FRK_EXP:      return launch_fill_t<MAN, FRK_EXP>(ctx, b, n, s, ld, rowptr, col, val, normalise);
        default:           return launch_fill_t<MAN, FRK_MATERN32>(ctx, b, n, s, ld, rowptr, col, val, normalise);
    }
}

extern "C" int frk_eval_basis_count(frk_ctx *ctx, const frk_basis *b, int64_t n, const double *d_s, int64_t ld_s,
                                    int *d_rowptr, int64_t *h_nnz) {
    FRK_REQUIRE(ctx && b && d_rowptr, "frk_eval_basis_count: NULL argument");
    FRK_REQUIRE(n >= 0 && ld_s >= n, "frk_eval_basis_count: bad n/ld");
    if (n == 0) {
        FRK_CHECK_CUDA(cudaMemsetAsync(d_rowptr, 0, sizeof(int), ctx->stream));
        if (h_nnz) *h_nnz = 0;
        return 0;
    }
    int *counts = nullptr;
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&counts, n * sizeof(int), ctx->stream));
    int rc;
    switch (b->dev.manifold) {
        case FRK_SPHERE:   rc = launch_count<FRK_SPHERE>(ctx, b, n, d_s, ld_s, counts); break;
        case FRK_PLANE:    rc = launch_count<FRK_PLANE>(ctx, b, n, d_s, ld_s, counts); break;
        case FRK_STPLANE:  rc = launch_count<FRK_STPLANE>(ctx, b, n, d_s, ld_s, counts); break;
        case FRK_STSPHERE: rc = launch_count<FRK_STSPHERE>(ctx, b, n, d_s, ld_s, counts); break;
        default:           rc = launch_count<FRK_REAL_LINE>(ctx, b, n, d_s, ld_s, counts); break;
    }
    if (!rc) rc = frk_exclusive_scan_i32(ctx, counts, n, d_rowptr, h_nnz);
    cudaFreeAsync(counts, ctx->stream);
    return rc;
}

extern "C" int frk_eval_basis_fill(frk_ctx *ctx, const frk_basis *b, int64_t n, const double *d_s, int64_t ld_s,
                                   const int *d_rowptr, int *d_col, double *d_val, int normalise) {
    FRK_REQUIRE(ctx && b && d_rowptr, "frk_eval_basis_fill: NULL argument");
    if (n == 0) return 0;
    switch (b->dev.manifold) {
        case FRK_SPHERE:   return launch_fill<FRK_SPHERE>(ctx, b, n, d_s, ld_s, d_rowptr, d_col, d_val, normalise);
        case FRK_PLANE:    return launch_fill<FRK_PLANE>(ctx, b, n, d_s, ld_s, d_rowptr, d_col, d_val, normalise);
        case FRK_STPLANE:  return launch_fill<FRK_STPLANE>(ctx, b, n, d_s, ld_s, d_rowptr, d_col, d_val, normalise);
        case FRK_STSPHERE: return launch_fill<FRK_STSPHERE>(ctx, b, n, d_s, ld_s, d_rowptr, d_col, d_val, normalise);
        default:           return launch_fill<FRK_REAL_LINE>(ctx, b, n, d_s, ld_s, d_rowptr, d_col, d_val, normalise);
    }
}


// ---------------------------------------------------------------------------------
// Dense rows on the fly (bases without compact support at scale: Gaussian / exponential / Matern on 1e7 points x r ~ 4k would be
// hundreds of GB as a stored matrix).  One CTA per point: out[i * r + j] = phi_j(s_i) for ALL j, optionally divided by the row norm
// (R/SRE.R:416-421).  The consumers (bucket.cu: gram_rhs_dense / quadform_dense) evaluate a chunk of rows, use it, and drop it.
// ---------------------------------------------------------------------------------
template <int MAN, int TYPE>
__global__ void __launch_bounds__(256) eval_dense_rows_kernel(BasisDev B, int64_t n, const double *__restrict__ s, int64_t ld, int normalise,
                                                              double *__restrict__ out) {
    __shared__ double red[8];
    __shared__ double s_xx;
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
        const PointCoord p = load_point<MAN>(s, ld, i);
        double *row = out + (size_t)i * B.r;
        double ss = 0.0;
        for (int j = threadIdx.x; j < B.r; j += blockDim.x) {
            const double v = basis_phi<TYPE>(basis_dist<MAN>(p, B, j), __ldg(B.scale + j));
            row[j] = v;
            ss += v * v;
        }
        if (normalise) {
            ss = warp_sum(ss);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
            __syncthreads();
            if (threadIdx.x == 0) {
                double t = 0.0;
                for (int w = 0; w < 8; w++) t += red[w];
                double xx = sqrt(t);
                s_xx = xx == 0.0 ? 1.0 : xx;
            }
            __syncthreads();
            const double xx = s_xx;
            for (int j = threadIdx.x; j < B.r; j += blockDim.x) row[j] = row[j] / xx;     // (each thread re-reads what it wrote)
        }
    }
}

template <int MAN>
static int launch_dense_rows(frk_ctx *ctx, const frk_basis *b, int64_t n, const double *s, int64_t ld, int normalise, double *out) {
    const unsigned grid = (unsigned)std::min<int64_t>(n, (int64_t)ctx->sm_count * 16);
    FRK_WORK(ctx, 0, (double)n * b->dev.r * 8.0);
    switch (b->dev.type) {
        case FRK_BISQUARE: FRK_LAUNCH(ctx, (eval_dense_rows_kernel<MAN, FRK_BISQUARE>), grid, 256, 0, b->dev, n, s, ld, normalise, out); break;
        case FRK_GAUSSIAN: FRK_LAUNCH(ctx, (eval_dense_rows_kernel<MAN, FRK_GAUSSIAN>), grid, 256, 0, b->dev, n, s, ld, normalise, out); break;
        case FRK_EXP:      FRK_LAUNCH(ctx, (eval_dense_rows_kernel<MAN, FRK_EXP>), grid, 256, 0, b->dev, n, s, ld, normalise, out); break;
        default:           FRK_LAUNCH(ctx, (eval_dense_rows_kernel<MAN, FRK_MATERN32>), grid, 256, 0, b->dev, n, s, ld, normalise, out); break;
    }
    return 0;
}

// rows [0, n) of eval_basis(basis, s) as a dense row-major n x r block (internal: bucket.cu's implicit matrices)
int frk_eval_basis_dense_rows(frk_ctx *ctx, const frk_basis *b, int64_t n, const double *d_s, int64_t ld_s, int normalise, double *d_out) {
    FRK_REQUIRE(ctx && b && d_s && d_out && n >= 0, "frk_eval_basis_dense_rows: bad argument");
    if (n == 0) return 0;
    switch (b->dev.manifold) {
        case FRK_SPHERE:   return launch_dense_rows<FRK_SPHERE>(ctx, b, n, d_s, ld_s, normalise, d_out);
        case FRK_PLANE:    return launch_dense_rows<FRK_PLANE>(ctx, b, n, d_s, ld_s, normalise, d_out);
        case FRK_STPLANE:  return launch_dense_rows<FRK_STPLANE>(ctx, b, n, d_s, ld_s, normalise, d_out);
        case FRK_STSPHERE: return launch_dense_rows<FRK_STSPHERE>(ctx, b, n, d_s, ld_s, normalise, d_out);
        default:           return launch_dense_rows<FRK_REAL_LINE>(ctx, b, n, d_s, ld_s, normalise, d_out);
    }
}

// ---------------------------------------------------------------------------------
// tensor-product rows and stand-alone row normalisation
// ---------------------------------------------------------------------------------
__global__ void tensor_count_kernel(int64_t n, const int *__restrict__ rp1, const int *__restrict__ rp2, int *__restrict__ cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cnt[i] = (rp1[i + 1] - rp1[i]) * (rp2[i + 1] - rp2[i]);
}

// one warp per row; S[i, j1 + r1*j2] = S1[i,j1]*S2[i,j2], time (operand 2) slowest   R/basisfns.R:821-822
__global__ void __launch_bounds__(256) tensor_fill_kernel(int64_t n, int r1, const int *__restrict__ rp1, const int *__restrict__ c1,
                                                          const double *__restrict__ v1, const int *__restrict__ rp2,
                                                          const int *__restrict__ c2, const double *__restrict__ v2,
                                                          const int *__restrict__ rp, int *__restrict__ col, double *__restrict__ val) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    int a0 = rp1[row], na = rp1[row + 1] - a0, b0 = rp2[row], nb = rp2[row + 1] - b0, o = rp[row];
    for (int t = lane; t < na * nb; t += 32) {
        int ib = t / na, ia = t - ib * na;
        col[o + t] = c1[a0 + ia] + r1 * c2[b0 + ib];
        val[o + t] = __dmul_rn(v2[b0 + ib], v1[a0 + ia]);
    }
}

extern "C" int frk_csr_tensor_count(frk_ctx *ctx, int64_t n, const int *d_rowptr1, const int *d_rowptr2, int *d_rowptr,
                                    int64_t *h_nnz) {
    if (n == 0) { FRK_CHECK_CUDA(cudaMemsetAsync(d_rowptr, 0, sizeof(int), ctx->stream)); if (h_nnz) *h_nnz = 0; return 0; }
    int *counts = nullptr;
    FRK_CHECK_CUDA(cudaMallocAsync((void **)&counts, n * sizeof(int), ctx->stream));
    FRK_LAUNCH(ctx, tensor_count_kernel, (unsigned)cdiv(n, 256), 256, 0, n, d_rowptr1, d_rowptr2, counts);
    int rc = frk_exclusive_scan_i32(ctx, counts, n, d_rowptr, h_nnz);
    cudaFreeAsync(counts, ctx->stream);
    return rc;
}

extern "C" int frk_csr_tensor_fill(frk_ctx *ctx, int64_t n, int r1, const int *d_rowptr1, const int *d_col1,
                                   const double *d_val1, const int *d_rowptr2, const int *d_col2, const double *d_val2,
                                   const int *d_rowptr, int *d_col, double *d_val) {
    if (n == 0) return 0;
    FRK_LAUNCH(ctx, tensor_fill_kernel, (unsigned)cdiv(n * 32, 256), 256, 0, n, r1, d_rowptr1, d_col1, d_val1, d_rowptr2, d_col2,
               d_val2, d_rowptr, d_col, d_val);
    return 0;
}

__global__ void row_normalise_kernel(int64_t n, const int *__restrict__ rp, double *__restrict__ val) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double ss = 0.0;
    for (int k = rp[i]; k < rp[i + 1]; k++) ss = __dadd_rn(ss, __dmul_rn(val[k], val[k]));
    double xx = __dsqrt_rn(ss);
    if (xx == 0.0) xx = 1.0;
    for (int k = rp[i]; k < rp[i + 1]; k++) val[k] = __ddiv_rn(val[k], xx);
}

extern "C" int frk_csr_row_normalise(frk_ctx *ctx, int64_t n, const int *d_rowptr, double *d_val) {
    if (n == 0) return 0;
    FRK_LAUNCH(ctx, row_normalise_kernel, (unsigned)cdiv(n, 256), 256, 0, n, d_rowptr, d_val);
    return 0;
}

// ---------------------------------------------------------------------------------
// BuildD (R/basisfns.R:1131-1153): centroid-to-centroid distance matrix of one resolution, D[a + ni*b] =
// distance(centroid idx[a], centroid idx[b]) in the reference's operation order (distR / dist_sphere).
// ---------------------------------------------------------------------------------
template <int MAN>
__global__ void basis_dist_matrix_kernel(BasisDev B, int ni, const int *__restrict__ idx, double *__restrict__ D) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)ni * ni; t += (int64_t)gridDim.x * blockDim.x) {
        const int ja = idx[t % ni], jb = idx[t / ni];
        PointCoord p;
        p.a = __ldg(B.c0 + ja); p.b = __ldg(B.c1 + ja); p.c = __ldg(B.c2 + ja); p.t = __ldg(B.c3 + ja); p.u = p.v = 0;
        D[t] = basis_dist<MAN>(p, B, jb);
    }
}

extern "C" int frk_basis_dist_matrix(frk_ctx *ctx, const frk_basis *b, int ni, const int *d_idx, double *d_D) {
    FRK_REQUIRE(ctx && b && d_idx && d_D && ni > 0, "frk_basis_dist_matrix: bad argument");
    const unsigned grid = ctx->sm_count * 8;
    switch (b->dev.manifold) {
        case FRK_SPHERE:   FRK_LAUNCH(ctx, basis_dist_matrix_kernel<FRK_SPHERE>, grid, 256, 0, b->dev, ni, d_idx, d_D); break;
        case FRK_PLANE:    FRK_LAUNCH(ctx, basis_dist_matrix_kernel<FRK_PLANE>, grid, 256, 0, b->dev, ni, d_idx, d_D); break;
        case FRK_STPLANE:  FRK_LAUNCH(ctx, basis_dist_matrix_kernel<FRK_STPLANE>, grid, 256, 0, b->dev, ni, d_idx, d_D); break;
        case FRK_STSPHERE: FRK_LAUNCH(ctx, basis_dist_matrix_kernel<FRK_STSPHERE>, grid, 256, 0, b->dev, ni, d_idx, d_D); break;
        default:           FRK_LAUNCH(ctx, basis_dist_matrix_kernel<FRK_REAL_LINE>, grid, 256, 0, b->dev, ni, d_idx, d_D); break;
    }
    return 0;
}

// ---------------------------------------------------------------------------------
// auto_basis helpers: nearest-centroid distance (scale rule) and column sums (pruning)
// ---------------------------------------------------------------------------------
template <int MAN>
__global__ void __launch_bounds__(128) nn_dist_kernel(BasisDev B, int zero_is_inf, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.r) return;
    PointCoord p;
    p.a = __ldg(B.c0 + i); p.b = __ldg(B.c1 + i); p.c = __ldg(B.c2 + i); p.t = __ldg(B.c3 + i); p.u = p.v = 0;
    double best = __longlong_as_double(0x7ff0000000000000LL);   // +Inf
    for (int j = 0; j < B.r; j++) {
        if (!zero_is_inf && j == i) continue;
        const double y = basis_dist<MAN>(p, B, j);
        if (zero_is_inf && y == 0.0) continue;
        best = fmin(best, y);
    }
    out[i] = best;
}

extern "C" int frk_basis_nn_dist(frk_ctx *ctx, const frk_basis *b, int zero_is_inf, double *d_out) {
    FRK_REQUIRE(ctx && b && d_out, "frk_basis_nn_dist: NULL argument");
    const unsigned grid = (unsigned)cdiv(b->dev.r, 128);
    switch (b->dev.manifold) {
        case FRK_SPHERE:   FRK_LAUNCH(ctx, nn_dist_kernel<FRK_SPHERE>, grid, 128, 0, b->dev, zero_is_inf, d_out); break;
        case FRK_PLANE:    FRK_LAUNCH(ctx, nn_dist_kernel<FRK_PLANE>, grid, 128, 0, b->dev, zero_is_inf, d_out); break;
        case FRK_STPLANE:  FRK_LAUNCH(ctx, nn_dist_kernel<FRK_STPLANE>, grid, 128, 0, b->dev, zero_is_inf, d_out); break;
        case FRK_STSPHERE: FRK_LAUNCH(ctx, nn_dist_kernel<FRK_STSPHERE>, grid, 128, 0, b->dev, zero_is_inf, d_out); break;
        default:           FRK_LAUNCH(ctx, nn_dist_kernel<FRK_REAL_LINE>, grid, 128, 0, b->dev, zero_is_inf, d_out); break;
    }
    return 0;
}

// one warp per row chunk; fp64 atomics (the sums feed a threshold test, not the model)
__global__ void __launch_bounds__(256) csr_colsums_kernel(int64_t nnz, const int *__restrict__ col, const double *__restrict__ val,
                                                          double *__restrict__ out) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(out + col[k], val[k]);
}

extern "C" int frk_csr_colsums(frk_ctx *ctx, int64_t n, int ncol, const int *d_rowptr, const int *d_col, const double *d_val, double *d_out) {
    FRK_REQUIRE(ctx && d_rowptr && d_out && ncol > 0 && n >= 0, "frk_csr_colsums: bad argument");
    FRK_CHECK_CUDA(cudaMemsetAsync(d_out, 0, (size_t)ncol * sizeof(double), ctx->stream));
    if (n == 0) return 0;
    int nnz = 0;
    FRK_CHECK_CUDA(cudaMemcpyAsync(&nnz, d_rowptr + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FRK_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (nnz == 0) return 0;
    const unsigned grid = (unsigned)std::min<int64_t>(cdiv(nnz, 256), (int64_t)ctx->sm_count * 16);
    FRK_LAUNCH(ctx, csr_colsums_kernel, grid, 256, 0, (int64_t)nnz, d_col, d_val, d_out);
    return 0;
}

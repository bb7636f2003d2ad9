/*
 * frkb200.h -- C ABI of libfrkb200.so: the B200 (sm_100a) implementation of the
 * Gaussian fixed-rank-kriging hot path of the R package FRK (v2.3.1).
 *
 * This is the drop-in boundary.  The reference has no native code on this path
 * (it is R on top of Matrix/sp/stats), so every entry point below replaces an R
 * function; the file:line of the reference interface each one stands in for is
 * cited with it (paths relative to the reference root).  The `.Call` shim a
 * maintainer adds on the R side is shown in INTEGRATION.md (r/src/frkb200_R.c).
 *
 * Conventions
 *   - plain C: pointers and sizes only, no C++/torch types.
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from frk_last_error() (thread-local).  No exception crosses the ABI.
 *   - matrices are column-major fp64 (R layout); indices are 0-based int32.
 *   - pointers named d_* are DEVICE pointers (cudaMalloc'd by the caller, by
 *     frk_malloc(), or torch-owned memory); h_* are HOST pointers.
 *   - all launches go to the stream given at frk_ctx_create(); calls are
 *     asynchronous unless they return a value to the host (documented per call).
 *   - there is no CPU fallback anywhere: without a CUDA device frk_ctx_create fails.
 */
#ifndef FRKB200_H
#define FRKB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRKB200_ABI_VERSION 1

typedef struct frk_ctx frk_ctx;       /* device + stream + scratch workspace        */
typedef struct frk_basis frk_basis;   /* device-resident basis + cell lists          */
typedef struct frk_polygons frk_polygons;       /* polygon BAUs (rings) + candidate cell lists  */
typedef struct frk_rowbuckets frk_rowbuckets;   /* row buckets + ELL tiles of one CSR basis matrix */
typedef struct frk_model frk_model;   /* device-resident SRE object (S, Z, X, Ve, Vfs + the r-sized slots) */

/* SRE(K_type=): R/SRE.R:322; only these two are legal for method = "EM" (R/check_args.R:184-187) */
enum { FRK_K_UNSTRUCTURED = 0, FRK_K_BLOCK_EXPONENTIAL = 1 };

/* Reduction hook for row-partitioned (multi-GPU) runs: combine buf[n] over the ranks IN PLACE and return 0.
 * on_device != 0: buf is a device pointer and the reduction must be enqueued on the ctx stream;
 * op: 0 = sum, 1 = max.  NULL = single rank.  (The plumbing -- torch.distributed / NCCL -- stays with the caller.) */
typedef int (*frk_allreduce_fn)(void *user, double *buf, int64_t n, int on_device, int op);
/* objective for the exported host optimisers: returns f(x[0..n)) */
typedef double (*frk_objective_fn)(void *user, int n, const double *x);

/* manifold(): R/geometryfns.R:19-189 */
enum { FRK_REAL_LINE = 1, FRK_PLANE = 2, FRK_SPHERE = 3, FRK_STPLANE = 4, FRK_STSPHERE = 5 };
/* local_basis(type=): R/basisfns.R:99-146 */
enum { FRK_BISQUARE = 0, FRK_GAUSSIAN = 1, FRK_EXP = 2, FRK_MATERN32 = 3 };

/* ---- context / memory ------------------------------------------------- */
int         frk_abi_version(void);
const char *frk_last_error(void);
/* stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or NULL */
int frk_ctx_create(int device, void *stream, frk_ctx **out);
int frk_ctx_destroy(frk_ctx *ctx);
int frk_sync(frk_ctx *ctx);
int frk_malloc(frk_ctx *ctx, size_t bytes, void **d_out);
int frk_free(frk_ctx *ctx, void *d_ptr);
int frk_h2d(frk_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);   /* async on ctx stream */
int frk_d2h(frk_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);   /* synchronises        */
int frk_d2d(frk_ctx *ctx, void *d_dst, const void *d_src, size_t bytes);   /* async on ctx stream */
int frk_memset0(frk_ctx *ctx, void *d_dst, size_t bytes);
/* number of kernels this library has launched on ctx since creation (bench.py: gpu_launches) */
int64_t frk_launch_count(frk_ctx *ctx);

/* Per-kernel device timing for bench.py's roofline: while enabled every launch is bracketed by CUDA events
 * on the ctx stream; the report is a JSON array of {kernel, launches, ms, flops, bytes} where flops/bytes
 * are the ALGORITHMIC work declared at the launch site.  Enabling clears earlier statistics. */
/* The factorisations replay cached CUDA graphs of their launch sequences (default on; off while profiling). */
int frk_use_graphs(frk_ctx *ctx, int enable);
/* enable != 0: run-to-run bit-reproducible results on one GPU.  The only order-dependent reduction on the EM path is the flush of
 * the per-bucket Gram tiles (fp64 atomics by default); in this mode every bucket stores its tile and a second kernel adds the
 * tiles into S'D^-1 S in ascending bucket order, one warp owning each row (frk_gram_rhs_bucketed).  Default off.             */
int frk_use_deterministic(frk_ctx *ctx, int enable);
int frk_profile(frk_ctx *ctx, int enable);
int frk_profile_report(frk_ctx *ctx, char *buf, size_t cap);
/* clock64() stamps taken at the phase boundaries of the last diag128 / panel-solve launch of frk_potrf (CTA 0): 32 values; see
 * tools/dense_phase_probe.py.  Development aid; synchronises the device. */
int frk_debug_clocks(long long *h_out32);

/* ---- multi-GPU: the context holds the NCCL communicator -------------------------------
 * One process per GPU.  Rank 0 calls frk_comm_unique_id and hands the FRK_COMM_ID_BYTES bytes to the other ranks by any means
 * the host language has (a file, MPI, a socket); every rank then calls frk_comm_init on its context.  A frk_model created on
 * such a context with allreduce == NULL sums its row-partitioned quantities over the communicator -- per EM iteration ONE fp64
 * ncclAllReduce of [packed upper triangle of S'D^-1 S | S'D^-1 rho | rho'D^-1 rho | sum log d] (r(r+1)/2 + r + 2 doubles: the row sums
 * crossprod() forms in R/SREfit.R:194,252,254) plus O(p^2)-double all-reduces of the M-step moments (:335-356).  No host
 * language is involved in the data path.  libnccl.so.2 is opened with dlopen on first use (FRKB200_NCCL_LIB overrides).       */
#define FRK_COMM_ID_BYTES 128
int frk_comm_unique_id(void *out128);
int frk_comm_init(frk_ctx *ctx, int rank, int world, const void *id128);
int frk_comm_destroy(frk_ctx *ctx);
int frk_comm_rank(const frk_ctx *ctx);
int frk_comm_world(const frk_ctx *ctx);
/* in place on the ctx stream; op: 0 = sum, 1 = max.  No-ops on a context without a communicator. */
int frk_comm_allreduce(frk_ctx *ctx, double *d_buf, int64_t n, int op);
int frk_comm_allreduce_host(frk_ctx *ctx, double *h_buf, int64_t n, int op);   /* synchronises */
int frk_comm_broadcast(frk_ctx *ctx, double *d_buf, int64_t n, int root);

/* ---- basis: local_basis()/Basis object, R/basisfns.R:99-146, R/AllClass.R:78 -----
 * h_loc: r x d column-major centroids; h_scale: r; h_res: r (1-based resolution).
 * radius: sphere radius (R/geometryfns.R:97), ignored otherwise.
 * Builds the per-resolution cell lists used by frk_eval_basis_*.                    */
int frk_basis_create(frk_ctx *ctx, int manifold, double radius, int type, int d, int r,
                     const double *h_loc, const double *h_scale, const int *h_res,
                     frk_basis **out);
int frk_basis_destroy(frk_basis *b);
/* BuildD (R/basisfns.R:1131-1153) for the ni centroids d_idx: d_D[a + ni*b] = distance(centroid idx[a], centroid idx[b]) */
int frk_basis_dist_matrix(frk_ctx *ctx, const frk_basis *b, int ni, const int *d_idx, double *d_D);
/* auto_basis scale rule (R/basisfns.R:485-511): d_out[i] = distance from centroid i to its nearest other centroid, in the manifold's
 * measure.  zero_is_inf = 0: only the centroid itself is excluded (diag(D) <- Inf, :488-490); != 0: every zero distance is
 * (D[D == 0] <- Inf, the batched branch for >= 5000 centroids, :497-509).  Inf when there is no other centroid.               */
int frk_basis_nn_dist(frk_ctx *ctx, const frk_basis *b, int zero_is_inf, double *d_out);
int frk_basis_n(const frk_basis *b);     /* number of basis functions r     */
int frk_basis_dim(const frk_basis *b);   /* manifold dimension d (1 or 2)    */

/* ---- eval_basis(Basis, matrix): R/basisfns.R:709-748 + .point_eval_fn :1157-1161 +
 *      closures :1213-1250 + distR R/geometryfns.R:202-219 / dist_sphere :1717-1740.
 * d_s: n x d column-major coordinates (leading dimension ld_s >= n).
 * Two calls: _count fills d_rowptr[n+1] (exclusive scan of the per-row non-zero count,
 * pattern = {phi != 0}) and returns nnz to the host; _fill writes ascending column indices
 * and values.  normalise != 0 fuses SRE()'s row normalisation (R/SRE.R:416-421).          */
int frk_eval_basis_count(frk_ctx *ctx, const frk_basis *b, int64_t n, const double *d_s, int64_t ld_s,
                         int *d_rowptr, int64_t *h_nnz);
int frk_eval_basis_fill(frk_ctx *ctx, const frk_basis *b, int64_t n, const double *d_s, int64_t ld_s,
                        const int *d_rowptr, int *d_col, double *d_val, int normalise);

/* eval_basis(TensorP_Basis, matrix): R/basisfns.R:812-824 + quickbind R/linalgfns.R:80-118.
 * S[i, j1 + r1*j2] = S1[i,j1]*S2[i,j2] from two CSR operands with the same rows.          */
int frk_csr_tensor_count(frk_ctx *ctx, int64_t n, const int *d_rowptr1, const int *d_rowptr2,
                         int *d_rowptr, int64_t *h_nnz);
int frk_csr_tensor_fill(frk_ctx *ctx, int64_t n, int r1,
                        const int *d_rowptr1, const int *d_col1, const double *d_val1,
                        const int *d_rowptr2, const int *d_col2, const double *d_val2,
                        const int *d_rowptr, int *d_col, double *d_val);
/* colSums(S) of a CSR matrix into d_out[ncol] (zeroed here): the influence of the data on each basis function that
 * auto_basis(prune > 0) thresholds, R/basisfns.R:536-553 */
int frk_csr_colsums(frk_ctx *ctx, int64_t n, int ncol, const int *d_rowptr, const int *d_col, const double *d_val, double *d_out);
/* SRE() row normalisation on an existing CSR: R/SRE.R:416-421 */
int frk_csr_row_normalise(frk_ctx *ctx, int64_t n, const int *d_rowptr, double *d_val);

/* ---- map_data_to_BAUs(SpatialPoints, SpatialPixels): R/geometryfns.R:1222-1320 ------
 * sp::over on a pixel grid = integer point-in-cell lookup (sp getGridIndex).
 * d_xy: n x 2 column-major; h_offset = cellcentre.offset, h_cellsize, h_dims = cells.dim.
 * d_cell2bau: map full-grid cell (row-wise from the TOP row) -> BAU row, or NULL for
 * identity.  d_bau[n] receives the 0-based BAU row, -1 for NA (outside / not a BAU).      */
int frk_grid_lookup(frk_ctx *ctx, int64_t n, const double *d_xy, int64_t ld_xy,
                    const double *h_offset, const double *h_cellsize, const int *h_dims,
                    const int *d_cell2bau, int *d_bau);
/* Same lookup for a row-partitioned run: with d_cell2bau == NULL and expand_grid_order != 0 the BAU
 * row is computed directly for BAUs stored in expand.grid order (x fastest, y ascending -- how
 * auto_BAU lays them out, R/geometryfns.R:629-637); rows outside [bau_lo, bau_hi) become NA and
 * the rest are returned relative to bau_lo (the rank's local BAU rows).                     */
int frk_grid_lookup_range(frk_ctx *ctx, int64_t n, const double *d_xy, int64_t ld_xy,
                          const double *h_offset, const double *h_cellsize, const int *h_dims,
                          const int *d_cell2bau, int expand_grid_order, int bau_lo, int bau_hi,
                          int *d_bau);
/* map_data_to_BAUs for polygon BAUs (ISEA3H hexagons): over(SpatialPoints, SpatialPolygons), R/geometryfns.R:1266,1680.
 * Polygon k is the ring h_vx/h_vy[h_ptr[k] .. h_ptr[k+1]) (no holes).  A point belongs to the first polygon (in index order)
 * whose interior, edge or vertex contains it (sp's point.in.polygon != 0; O'Rourke's crossing test); -1 = NA.
 * Rows outside [bau_lo, bau_hi) become NA and the rest are returned relative to bau_lo (row-partitioned runs).          */
int frk_polygons_create(frk_ctx *ctx, int npoly, const int *h_ptr, const double *h_vx, const double *h_vy,
                        frk_polygons **out);
int frk_polygons_destroy(frk_polygons *P);
int frk_polygon_lookup(frk_ctx *ctx, const frk_polygons *P, int64_t n, const double *d_xy, int64_t ld_xy,
                       int bau_lo, int bau_hi, int *d_bau);
/* map_data_to_BAUs(ST): R/geometryfns.R:1419-1547.  Space-time BAUs are ns spatial BAUs x nt time intervals, space fastest
 * (auto_BAUs, R/geometryfns.R:971-980); h_breaks[nt] are the interval starts time(sp_pols).  d_spatial[n] is the per-point
 * spatial BAU (from frk_grid_lookup / frk_polygon_lookup, -1 = NA), d_time[n] the observation times in the units of h_breaks.
 * Slice i takes t1_i <= time < t1_{i+1}; the last slice takes every time >= its start (:1462-1468); earlier than the first
 * start -> NA.  d_bau = spatial + ns*slice, restricted to [bau_lo, bau_hi) and relative to bau_lo (may alias d_spatial).  */
int frk_st_bau_index(frk_ctx *ctx, int64_t n, const int *d_spatial, const double *d_time, int nt, const double *h_breaks,
                     int64_t ns, int64_t bau_lo, int64_t bau_hi, int *d_bau);
/* ---- areal (polygon-support) data: BuildC(SpatialPolygons) R/geometryfns.R:1572-1616 + SRE() R/SRE.R:472-502 ----------------
 * d_group[N]: for every BAU the observation polygon that contains its centroid (frk_polygon_lookup over the BAU centroids;
 * -1 = none).  frk_areal_incidence builds Cmat as CSR over the m observations: d_rowptr[m+1], d_members (BAU rows, ascending
 * within an observation = j_idx), d_weights = BAUs$wts (d_wts, NULL = 1) divided by the row sum (normalise_wts = TRUE,
 * R/SRE.R:483-490).  *h_empty = observations that contain no BAU centroid (the reference falls back to the BAU under the
 * polygon's vertices, :1593-1598; the caller decides).  Footprints must not share BAUs (one group per BAU): Vfs stays diagonal. */
int frk_areal_incidence(frk_ctx *ctx, int64_t N, const int *d_group, int64_t m, const double *d_wts, int *d_rowptr,
                        int *d_members, double *d_weights, int64_t *h_nmem, int64_t *h_empty);
/* S = Cmat %*% S0 (R/SRE.R:502) for a general row-weighted Cmat: two-pass CSR build, columns ascending, values summed in
 * member order.  r <= ~25000 (per-observation accumulator in shared memory).                                                  */
int frk_csr_aggregate_count(frk_ctx *ctx, int64_t m, int r, const int *d_cz_rowptr, const int *d_cz_members,
                            const int *d_rowptr0, const int *d_col0, int *d_rowptr, int64_t *h_nnz);
int frk_csr_aggregate_fill(frk_ctx *ctx, int64_t m, int r, const int *d_cz_rowptr, const int *d_cz_members,
                           const double *d_cz_weights, const int *d_rowptr0, const int *d_col0, const double *d_val0,
                           const int *d_rowptr, int *d_col, double *d_val);
/* d_out[j + m*c] = sum_k w_jk^power d_in[member_k + ld_in*c]: power 1 -> X = Cmat %*% X_BAU (BAU covariates averaged over the
 * footprint, R/geometryfns.R:1352-1357); power 2 with d_in = fs -> diag(tcrossprod(Cmat %*% Diagonal(sqrt(fs)))), R/SRE.R:498 */
int frk_areal_aggregate(frk_ctx *ctx, int64_t m, const int *d_cz_rowptr, const int *d_cz_members, const double *d_cz_weights,
                        int power, int ncols, const double *d_in, int64_t ld_in, double *d_out);
/* the same lookup over all polygons, also counting the points that lie in MORE than one polygon (*h_multi): used with the
 * BAU centroids as points and the observation polygons as polygons (over(BAU_as_points, this_poly), R/geometryfns.R:1589),
 * where a BAU shared by two observations would make Vfs non-diagonal                                                          */
int frk_polygon_lookup_count(frk_ctx *ctx, const frk_polygons *P, int64_t n, const double *d_xy, int64_t ld_xy, int *d_first,
                             int64_t *h_multi);
/* General incidence for (possibly overlapping) polygons -- the prediction polygons' C_P (BuildC, R/geometryfns.R:1572-1616;
 * normalised as in R/SREpredict.R:119-128): every polygon gets ALL the points it contains.  _count: d_ptoff[n+1] = exclusive scan of
 * the number of containing polygons per point, *h_npairs the total; _fill: CSR over the polygons (d_rowptr[npoly+1], d_members and
 * d_weights[npairs], points ascending within a polygon, weights d_wts (NULL = 1) / row sum), *h_empty = polygons without a point. */
int frk_polygon_incidence_count(frk_ctx *ctx, const frk_polygons *P, int64_t n, const double *d_xy, int64_t ld_xy, int *d_ptoff,
                                int64_t *h_npairs);
int frk_polygon_incidence_fill(frk_ctx *ctx, const frk_polygons *P, int64_t n, const double *d_xy, int64_t ld_xy, const int *d_ptoff,
                               int64_t npairs, const double *d_wts, int *d_rowptr, int *d_members, double *d_weights, int64_t *h_empty);
/* group_by(BAU) %>% summarise_all(mean): R/geometryfns.R:1286-1302.  d_vals: n x ncols
 * column-major (ld n).  Outputs (capacity n rows): d_obs_bau[m] ascending BAU rows,
 * d_means m x ncols column-major with leading dimension n, d_counts[m].  Returns m.
 * Deterministic: members of a BAU are summed in ascending original point order.           */
int frk_bin_mean(frk_ctx *ctx, int64_t n, const int *d_bau, int64_t nbau, int ncols,
                 const double *d_vals, int *d_obs_bau, double *d_means, int *d_counts,
                 int64_t *h_m);
/* the same with map_data_to_BAUs(sum_variables = ...) (R/geometryfns.R:1286-1299): columns whose bit is set in sum_mask are
 * SUMMED over the observations of a BAU (.safe_sum) instead of averaged (ncols <= 32 then).                                  */
int frk_bin_mean_sum(frk_ctx *ctx, int64_t n, const int *d_bau, int64_t nbau, int ncols,
                     const double *d_vals, unsigned sum_mask, int *d_obs_bau, double *d_means, int *d_counts,
                     int64_t *h_m);

/* coordinates(BAUs) of a pixel grid in expand.grid order (x fastest), BAU rows [lo, hi): what SRE()
 * hands to eval_basis (R/SRE.R:412; grid layout R/geometryfns.R:629-637).  d_xgrid[nx], d_ygrid[ny]
 * device axes; d_out (hi-lo) x 2 column-major.                                               */
int frk_grid_centroids(frk_ctx *ctx, int64_t lo, int64_t hi, int nx, const double *d_xgrid,
                       const double *d_ygrid, double *d_out);

/* ---- S = Cmat %*% S0 for point data (row gather): R/SRE.R:502, BuildC
 *      R/geometryfns.R:1557-1568 -------------------------------------------------- */
int frk_csr_gather_count(frk_ctx *ctx, int64_t m, const int *d_rows, const int *d_rowptr_src,
                         int *d_rowptr_dst, int64_t *h_nnz);
int frk_csr_gather_fill(frk_ctx *ctx, int64_t m, const int *d_rows,
                        const int *d_rowptr_src, const int *d_col_src, const double *d_val_src,
                        const int *d_rowptr_dst, int *d_col_dst, double *d_val_dst);

/* ---- E-step / log-likelihood accumulators: R/SREfit.R:194 (crossprod(cholDinvT %*% S)),
 *      :252 (Q_eta), :254 (t(S) %*% Dinv %*% (Z - X alpha)), :198,:211 (scalars).
 * For the m local rows of S (CSR) with d_j = sigma2fs*vfs_j + ve_j, rho_j = z_j - x_j'alpha:
 *   d_acc[0 .. r*r)        += upper triangle (row<=col, column-major) of sum_j s_j s_j'/d_j
 *   d_acc[r*r .. r*r+r)    += sum_j s_j rho_j / d_j
 *   d_acc[r*r+r]           += sum_j rho_j^2 / d_j
 *   d_acc[r*r+r+1]         += sum_j log d_j
 * d_acc is ONE contiguous buffer of r*r + r + 2 doubles so that a single fp64 all-reduce per
 * EM iteration combines the ranks.  The caller zeroes it first.  X: m x p column-major
 * (ld m); h_alpha: p.                                                                     */
int frk_gram_rhs(frk_ctx *ctx, int64_t m, int r, const int *d_rowptr, const int *d_col,
                 const double *d_val, const double *d_z, const double *d_X, int p,
                 const double *h_alpha, const double *d_ve, const double *d_vfs, double sigma2fs,
                 double *d_acc);

/* ---- row buckets: group the rows of a CSR basis matrix by their last column index (the highest finest-
 * resolution basis function in the row's support).  Built once per matrix (SRE()); lets the Gram and
 * quadratic-form kernels stage the few columns a bucket touches in shared memory.
 *   d_bptr[ncols+1], d_perm[n] : rows of bucket k are d_perm[d_bptr[k] .. d_bptr[k+1])
 *   d_nU[ncols], d_ucol[ncols*64] : the columns bucket k touches (d_nU[k] > 64: not staged, direct path)
 *   d_lid[nnz] : for every stored entry, its position in its bucket's column list (one byte)            */
int frk_csr_bucket_rows(frk_ctx *ctx, int64_t n, const int *d_rowptr, const int *d_col, int ncols,
                        int *d_perm, int *d_bptr, int *d_ucol, int *d_nU, unsigned char *d_lid);
/* Owning handle: the bucket arrays above plus a bucket-ordered, entry-major copy of the matrix in tiles of 32 rows
 * ("ELL tiles": val[e][lane], lid[e][lane]) that the per-iteration kernels read with coalesced bulk copies.        */
int frk_rowbuckets_create(frk_ctx *ctx, int64_t n, int ncols, const int *d_rowptr, const int *d_col,
                          const double *d_val, frk_rowbuckets **out);
int frk_rowbuckets_destroy(frk_rowbuckets *bk);
/* frk_gram_rhs / frk_quadform_spmv with row buckets (same results up to fp64 summation order) */
int frk_gram_rhs_bucketed(frk_ctx *ctx, int64_t m, int r, const int *d_rowptr, const int *d_col,
                          const double *d_val, const frk_rowbuckets *bk, const double *d_z,
                          const double *d_X, int p, const double *h_alpha, const double *d_ve,
                          const double *d_vfs, double sigma2fs, double *d_acc);
int frk_quadform_spmv_bucketed(frk_ctx *ctx, int64_t n, int r, const int *d_rowptr, const int *d_col,
                               const double *d_val, const frk_rowbuckets *bk, const double *d_M,
                               const double *d_mu, double *d_q, double *d_t);

/* ---- gather-based sparse quadratic form + SpMV: R/SREutils.R:286-300 (.batch_compute_var,
 *      rowSums((S0 Cov) * S0)), R/SREfit.R:332-334 (Omega_diag1), S %*% mu_eta.
 * q_i = s_i' M s_i (M symmetric r x r column-major, ld r), t_i = s_i' mu.  Either output
 * may be NULL.                                                                            */
int frk_quadform_spmv(frk_ctx *ctx, int64_t n, int r, const int *d_rowptr, const int *d_col,
                      const double *d_val, const double *d_M, const double *d_mu,
                      double *d_q, double *d_t);

/* ---- M-step row sums: R/SREfit.R:335-356 (J), :391-405 (alpha), :408-417 (sigma2fs).
 * With u_j = 1/d_j, w_j = vfs_j/d_j^2 (mode 0, heteroscedastic J at sigma2fs) or u=w=1
 * (mode 1, homoscedastic closed form) writes to h_out (2p^2 + 2p + 2 doubles, synchronises):
 *   [0,p^2) sum u x x'   [p^2,p^2+p) sum u x (z-t)   [..+1) sum vfs u
 *   [..+1) sum w (Omega1 - 2 t z + z^2), Omega1 = q + t^2 (d_q = s_j' S_eta s_j)   [..+p) sum w x (t - z)   [..+p^2) sum w x x'      */
int frk_mstep_sums(frk_ctx *ctx, int64_t m, int p, const double *d_X, const double *d_z,
                   const double *d_t, const double *d_q, const double *d_ve,
                   const double *d_vfs, double sigma2fs, int mode, double *h_out);

/* h_out[4] = sum(x - shift), sum((x - shift)^2), min(x), max(x)  (synchronises).  var(Z) and
 * mean(diag Ve) of .EM_initialise (R/SRE.R:704-731); all(a == a[1]) test (R/SREfit.R:278-280).  */
int frk_vec_stats(frk_ctx *ctx, int64_t n, const double *d_x, double shift, double *h_out);

/* out = a*a  (Ve = std^2, R/SRE.R:468)  and  out = 1/a  (Ve^-1, R/SREpredict.R:286-293) */
int frk_square(frk_ctx *ctx, int64_t n, const double *d_a, double *d_out);
int frk_recip(frk_ctx *ctx, int64_t n, const double *d_a, double *d_out);
/* x *= alpha  (K_i^-1 = omega_i exp(-D_i/tau_i)^-1, R/SREfit.R:625-637) */
int frk_scale(frk_ctx *ctx, int64_t n, double alpha, double *d_x);
/* x[i] = value  (intercept column of the model matrix at the BAUs, R/SRE.R:464-465) */
int frk_fill(frk_ctx *ctx, int64_t n, double value, double *d_x);
/* dst[j, c] = src[idx[j], c] for ncols column-major columns: Vfs = fs[bau] (R/SRE.R:498) and the model
 * matrix X at the observed BAUs (R/SRE.R:464-465)                                                */
int frk_gather_rows(frk_ctx *ctx, int64_t m, int ncols, const int *d_idx, const double *d_src,
                    int64_t ld_src, double *d_dst, int64_t ld_dst);

/* ---- predict at BAUs: R/SREpredict.R:334-375 and the point-data closed form of :308-333.
 * mode 0 (obs_fs=TRUE or sigma2fs==0): mu = xa + smu, var = q.
 * mode 1 (obs_fs=FALSE, sigma2fs>0):   a = A/(A+Qf), Qf = 1/(sigma2fs*fs);
 *        var = (1-a)^2 q + 1/(A+Qf);  mu = xa + smu + (ytil - A*smu)/(A+Qf).
 * d_A / d_ytil may be NULL (treated as 0: unobserved BAUs).  sd = sqrt(var).               */
int frk_predict_finalize(frk_ctx *ctx, int64_t n, int mode, double sigma2fs, const double *d_q,
                         const double *d_smu, const double *d_xa, const double *d_A,
                         const double *d_ytil, const double *d_fs, double *d_mu, double *d_var,
                         double *d_sd);
/* scatter per-observation values to their BAU: dst[idx[j]] += src[j] (C' Ve^-1 C, C' Ve^-1 rho) */
int frk_scatter_add(frk_ctx *ctx, int64_t m, const int *d_idx, const double *d_src, double *d_dst);
/* y = z - X alpha, optionally divided by d_div (may be NULL) */
int frk_resid(frk_ctx *ctx, int64_t m, int p, const double *d_X, const double *h_alpha,
              const double *d_z, const double *d_div, double *d_out);

/* out = X alpha  (X %*% alpha at the BAUs, R/SREpredict.R:274,347) */
int frk_xalpha(frk_ctx *ctx, int64_t n, int p, const double *d_X, const double *h_alpha, double *d_out);

/* ---- dense fp64 (Matrix::chol / chol2inv / determinant / solve; R/linalgfns.R:21-36) ----
 * All r x r column-major with leading dimension r.                                         */
/* In-place lower Cholesky A = L L' (chol(), R/SREfit.R:177,195,253,273); the strict upper
 * triangle is left untouched.  *h_logdet = 2*sum(log(diag L)) (logdet, R/linalgfns.R:32).
 * Fails (status != 0) if A is not positive definite.                                       */
int frk_potrf(frk_ctx *ctx, int r, double *d_A, double *h_logdet);
/* chol2inv: given L from frk_potrf in d_L (left intact), writes the full symmetric inverse of
 * A to d_Ainv.  d_work: r*r doubles of scratch (receives L^-1, lower triangular).                             */
int frk_potri(frk_ctx *ctx, int r, const double *d_L, double *d_Ainv, double *d_work);
/* x = X b for the lower-triangular X = L^-1 that frk_potri leaves in d_work, i.e. x = L^-1 b:
 * (rDinv %*% S) %*% solve(R) in .loglik.ind (R/SREfit.R:211)                                  */
int frk_lower_mulv(frk_ctx *ctx, int r, const double *d_X, const double *d_b, double *d_x);
/* y = A x for symmetric full-storage A */
int frk_symv(frk_ctx *ctx, int r, const double *d_A, const double *d_x, double *d_y);
/* out = mirror(upper triangle of G) + B   (Q_eta = S'D^-1 S + K^-1, R/SREfit.R:252); B may be NULL */
int frk_sym_from_upper_add(frk_ctx *ctx, int r, const double *d_Gupper, const double *d_B, double *d_out);
/* A = value * I  (.EM_initialise, R/SRE.R:710-722);  diag(A) += d  (.regularise_K, R/SREfit.R:669-686) */
int frk_scaled_identity(frk_ctx *ctx, int r, double value, double *d_A);
int frk_add_diag(frk_ctx *ctx, int r, const double *d_diag, double *d_A);
/* out = A + x x'   (S_eta + tcrossprod(mu_eta), R/SREfit.R:332,689) */
int frk_add_outer(frk_ctx *ctx, int r, const double *d_A, const double *d_x, double *d_out);
/* out = A + alpha x x'   (Sherman-Morrison form of Khat_inv for the unstructured K, R/SREfit.R:273,689) */
int frk_add_scaled_outer(frk_ctx *ctx, int r, const double *d_A, const double *d_x, double alpha, double *d_out);
/* out = exp(-D / tau) * scale   (block-exponential K_i, R/SREfit.R:517,631) */
int frk_exp_kernel(frk_ctx *ctx, int64_t n, const double *d_D, double tau, double scale, double *d_out);
/* *h_out = sum(A * B) elementwise over n entries (sum(diag2(.,., symm=TRUE)), R/SREfit.R:478,524) */
int frk_dot(frk_ctx *ctx, int64_t n, const double *d_A, const double *d_B, double *h_out);
/* copy the ni x ni block A[idx, idx] (idx device int32) out of / into an r x r matrix
 * (S_eta[idx,idx], Khat[idx,idx], bdiag + reverse_permute: R/SREfit.R:459-464,625-649)      */
int frk_block_gather(frk_ctx *ctx, int r, const double *d_A, int ni, const int *d_idx, double *d_out);
int frk_block_scatter(frk_ctx *ctx, int r, double *d_A, int ni, const int *d_idx, const double *d_in);
/* out[i + ni*j] = x[idx[i]]*x[idx[j]] + A[idx[i], idx[j]]  (eta2 blocks, R/SREfit.R:459-464) */
int frk_block_gather_outer(frk_ctx *ctx, int r, const double *d_A, const double *d_x, int ni,
                           const int *d_idx, double *d_out);
/* Kronecker helpers for the block-exponential K of a tensor-product basis (time slowest, space fastest; R/SREfit.R:496-514,
 * 627-649); kronecker(At, As) is never formed.  *h_out = sum(kronecker(At, As) * E);  A[idx[a], idx[b]] = scale * kron(At, As)[a, b]. */
int frk_kron_dot(frk_ctx *ctx, int nt, int ns, const double *d_At, const double *d_As, const double *d_E, double *h_out);
int frk_kron_scatter(frk_ctx *ctx, int r, int nt, int ns, const double *d_At, const double *d_As, double scale,
                     const int *d_idx, double *d_A);
/* C(MxN, ldc) = beta*C + alpha * A(MxK, lda) * B(NxK, ldb)'  -- the fp64 tensor-core GEMM all
 * blocked factorisations are built from (exported for tests and the fp64 peak probe).       */
int frk_dgemm_nt(frk_ctx *ctx, int M, int N, int K, double alpha, const double *d_A, int lda,
                 const double *d_B, int ldb, double beta, double *d_C, int ldc);

/* general operand layouts: transa / transb != 0 -> that operand is stored with k contiguous (K x M / K x N) */
int frk_dgemm(frk_ctx *ctx, int transa, int transb, int M, int N, int K, double alpha, const double *d_A,
              int lda, const double *d_B, int ldb, double beta, double *d_C, int ldc);

/* ---- the SRE object and SRE.fit(method = "EM") / predict() as single calls ------------------------------
 * frk_model holds this rank's rows of S (m x r CSR), Z, X (m x p column-major), Ve = diag, Vfs = diag and the
 * slots mu_eta, S_eta, Q_eta, Khat, Khat_inv (device), alphahat, sigma2fshat (host): R/AllClass.R:133-175.
 * _create borrows device arrays (they must outlive the model); _create_host copies host arrays (what the R shim
 * passes: S as a dgRMatrix' p/j/x, R/SRE.R:502).                                                              */
int frk_model_create(frk_ctx *ctx, int64_t m, int r, int p, const int *d_rowptr, const int *d_col,
                     const double *d_val, const double *d_Z, const double *d_X, const double *d_Ve,
                     const double *d_Vfs, frk_allreduce_fn allreduce, void *user, frk_model **out);
int frk_model_create_host(frk_ctx *ctx, int64_t m, int r, int p, const int *h_rowptr, const int *h_col,
                          const double *h_val, const double *h_Z, const double *h_X, const double *h_Ve,
                          const double *h_Vfs, frk_model **out);
int frk_model_destroy(frk_model *mdl);
/* Position of this rank in a row-partitioned run.  With world > 1 the Nelder-Mead candidates of the block-exponential M-step (each a
 * replicated r_i x r_i factorisation) are predicted ahead and spread over the ranks; the iterate path and all results are identical to
 * the sequential run.  emulate != 0 evaluates a pretend world's candidates on this one rank (for tests).                              */
int frk_model_set_parallel(frk_model *mdl, int rank, int world, int emulate);
/* Concurrent evaluation lanes on this GPU for the block-exponential Nelder-Mead of .update_K (R/SREfit.R:618-621): each f_tau
 * evaluation is a latency-bound r_i x r_i factorisation, so up to `lanes` of the points Nelder-Mead may ask for next are evaluated
 * side by side on separate streams (combined with the ranks of a row-partitioned run: world x lanes candidates per round).
 * The sequence of iterates, hence every result, is that of the sequential algorithm.  Default 1.                                  */
int frk_model_set_lanes(frk_model *mdl, int lanes);
/* Row-partitioned runs: distribute != 0 spreads those candidates over the ranks as well (world x lanes per round; one small
 * all-reduce per round and a broadcast of the winner's inverse); 0 (default) lets every rank evaluate them on its own lanes --
 * replicated and deterministic like the rest of the r-sized work, with no exchange.  Results are identical either way.        */
int frk_model_set_nm_distribute(frk_model *mdl, int distribute);
/* SRE object whose basis-function matrix S is never stored: row j of S is eval_basis(basis, coords[j, ]) (R/basisfns.R:709-748;
 * rows divided by their norms if normalise != 0, R/SRE.R:416-421), evaluated in chunks inside every pass over the matrix and
 * dropped again.  For basis functions without compact support (Gaussian, exponential, Matern 3/2), whose S is dense: 1e7 observations
 * x r = 3608 would be 290 GB as a stored matrix (BASELINE config 5).  d_coords: the observed BAUs' coordinates, column-major with
 * leading dimension ld_coords; the other arguments as frk_model_create.  frk_rowbuckets_create_implicit is the same for S0
 * (prediction): pass the handle to frk_model_predict with NULL CSR pointers.                                                     */
int frk_model_create_implicit(frk_ctx *ctx, int64_t m, int p, const frk_basis *basis, const double *d_coords, int64_t ld_coords,
                              int normalise, const double *d_Z, const double *d_X, const double *d_Ve, const double *d_Vfs,
                              frk_allreduce_fn allreduce, void *user, frk_model **out);
int frk_rowbuckets_create_implicit(frk_ctx *ctx, int64_t n, const frk_basis *basis, const double *d_coords, int64_t ld_coords,
                                   int normalise, frk_rowbuckets **out);
/* Non-diagonal Vfs: observation footprints that share BAUs (overlapping areal supports, BuildC R/geometryfns.R:1572-1616), or
 * several point observations in one BAU with average_in_BAU = FALSE.  Vfs = tcrossprod(Cmat %*% Diagonal(sqrt(fs))) (R/SRE.R:498)
 * as the full symmetric m x m host CSR (0-based; the dgRMatrix slots of object@Vfs).  The EM then runs the reference's general
 * branches: chol(D) in .loglik.ind / .SRE.Estep.ind (R/SREfit.R:186-189,240-252) and the general J(sigma2fs) (:296-328).  D is
 * block diagonal over the connected groups of coupled observations; groups of up to 64 observations are handled, larger ones are
 * refused.  A diagonal matrix is accepted and changes nothing.  One rank only.  Call before frk_model_init.                    */
int frk_model_set_vfs_csr(frk_model *mdl, const int *h_rowptr, const int *h_col, const double *h_val);
/* number of connected groups (0: Vfs is diagonal)                                                                               */
int frk_model_coupled_groups(const frk_model *mdl);
/* BuildD() (R/basisfns.R:1131-1153) for K_type = "block-exponential": h_res[r] 1-based resolution of every basis
 * function, h_D[i] the r_i x r_i centroid distance matrix of resolution i+1 (column-major, host).              */
int frk_model_set_blocks(frk_model *mdl, int nres, const int *h_res, const double *const *h_D);
/* BuildD on the device: the per-resolution distance matrices are computed from the basis' own centroid table with the
 * arithmetic of distR / dist_sphere (no r_i x r_i host matrices, no upload).                                          */
int frk_model_set_blocks_basis(frk_model *mdl, int nres, const int *h_res, const frk_basis *basis);
/* the same for a TensorP_Basis (space x time): r = r_s * r_t, tensor column = j_s + r_s * j_t (R/basisfns.R:821-822);
 * h_res_s[r_s] spatial resolutions, h_Ds[i] the spatial distance matrices per resolution (D_basis$Basis1), h_Dt the
 * r_t x r_t temporal one (D_basis$Basis2[[1]]).  tau_i = (spatial, temporal) by 2-D Nelder-Mead.                     */
int frk_model_set_blocks_tensor(frk_model *mdl, int nres, int r_s, int r_t, const int *h_res_s,
                                const double *const *h_Ds, const double *h_Dt);
/* .EM_initialise (R/SRE.R:704-731) */
int frk_model_init(frk_model *mdl);
/* .EM_fit (R/SREfit.R:78-160): up to n_EM iterations of logLik + E-step + M-step; stops when the log-likelihood
 * changes by less than tol.  h_lambda[nlambda] ridge parameter(s) (unstructured K; R/SREfit.R:664-694), h_res as
 * above (may be NULL when nlambda <= 1).  h_llk[n_EM] receives info_fit$plot_lik$llk, *h_niter the iterations run. */
int frk_model_em_fit(frk_model *mdl, int n_EM, double tol, int K_type, int nlambda, const double *h_lambda,
                     const int *h_res, int known_sigma2fs_given, double known_sigma2fs, double *h_llk,
                     int *h_niter);
/* logLik(SRE) at the current slots (R/SREutils.R:22-32 -> .loglik.ind R/SREfit.R:174-220) */
int frk_model_loglik(frk_model *mdl, double *h_out);
/* slot access: "mu_eta" (r), "S_eta", "Q_eta", "Khat", "Khat_inv" (r x r column-major), "alphahat" (p),
 * "sigma2fshat" (1); read-only: "m_total", "homoscedastic", "tau", "omega" (per resolution, last M-step),
 * "nm_stats" (4: Nelder-Mead points requested / evaluation rounds / points evaluated / re-evaluations, last M-step),
 * "nm_prefetch_hits" (1: evaluations whose factorisation ran beside the E-step), "nm_reuse_hits" (1: evaluations served by a saved
 * inverse -- exp(-D/tau)^-1 depends on tau alone -- instead of a factorisation)  */
int frk_model_get(frk_model *mdl, const char *slot, double *h_out);
int frk_model_set(frk_model *mdl, const char *slot, const double *h_in);
int frk_model_slot_device(frk_model *mdl, const char *slot, double **d_out);
/* .SRE.predict_EM at n BAUs (R/SREpredict.R:243-375; predict_BAUs = TRUE, covariances = FALSE) with
 * .batch_compute_var (R/SREutils.R:245-305).  S0 (n x r device CSR) comes with its row buckets
 * (frk_rowbuckets_create); d_XBAU n x p column-major; d_fs[n]; d_obs_bau[m] = BAU row of every observation.
 * obs_fs != 0 or sigma2fs == 0: stored posterior; else the joint (eta, xi) solve at the final parameters.      */
int frk_model_predict(frk_model *mdl, int64_t n, const int *d_rowptr0, const int *d_col0, const double *d_val0,
                      const frk_rowbuckets *bk0, const double *d_XBAU, const double *d_fs,
                      const int *d_obs_bau, int obs_fs, double *d_mu, double *d_var, double *d_sd);
/* The same for areal data whose supports share no BAU (Cmat from frk_areal_incidence), obs_fs = FALSE and sigma2fs > 0: the joint
 * (eta, xi) solve of R/SREpredict.R:308-333 in closed form -- one more E-step, the quadratic forms over S0 and S, and one
 * cross term s_k'Sigma S_j' per BAU inside a support.  (obs_fs = TRUE / sigma2fs = 0: frk_model_predict.)                     */
int frk_model_predict_areal(frk_model *mdl, int64_t n, const int *d_rowptr0, const int *d_col0, const double *d_val0,
                            const frk_rowbuckets *bk0, const double *d_XBAU, const double *d_fs, const int *d_cz_rowptr,
                            const int *d_cz_members, const double *d_cz_weights, double *d_mu, double *d_var, double *d_sd);
/* predict(newdata = <polygons>) (R/SREpredict.R:365-419): C_P (mP x n CSR, rows normalised: frk_polygon_incidence_*) applied to the
 * BAU-level prediction.  mu = C_P mu_BAU; var = diag(C_P S0 Cov S0' C_P') for obs_fs != 0 or sigma2fs == 0; for obs_fs == 0 with
 * point-referenced data the joint-solve variance in closed form (rows of S0 weighted by c (1 - a_k), plus sum c^2 / (A_k + Q_k)). */
int frk_model_predict_polygons(frk_model *mdl, int64_t n, const int *d_rowptr0, const int *d_col0, const double *d_val0,
                               const frk_rowbuckets *bk0, const double *d_XBAU, const double *d_fs, const int *d_obs_bau, int obs_fs,
                               int64_t mP, const int *d_cp_rowptr, const int *d_cp_members, const double *d_cp_weights, int64_t cp_nnz,
                               double *d_mu, double *d_var, double *d_sd);
/* stats::optim(method = "Nelder-Mead", control = list(maxit, reltol)) and stats::uniroot (default tol) as used by
 * the M-step (R/SREfit.R:618-621, :386-388); exported so that the iterate path can be tested on its own.       */
int frk_optim_nelder_mead(frk_objective_fn fn, void *user, int n, const double *par, int maxit,
                          double reltol, double *x_out, double *f_out, int *count_out, int *fail_out);
/* The same minimisation driven the way the block-exponential M-step drives it (speculative batches of up to `width` points per
 * round, see frk_model_set_lanes): identical result; *rounds_out <= the number of sequential evaluations.                       */
int frk_optim_nelder_mead_speculative(frk_objective_fn fn, void *user, int n, const double *par, int maxit, double reltol,
                                      int width, double *x_out, double *f_out, int *requests_out, int *rounds_out, int *evals_out);
int frk_optim_zeroin(frk_objective_fn fn, void *user, double lower, double upper, double *root_out);

/* ---- host-buffer entry points (what the R .Call shim binds; H2D/D2H inside) ------------ */
/* over(SpatialPoints, SpatialPolygons) for polygon BAUs (hexagons, irregular cells) from host buffers: h_xy n x 2 column-major,
 * h_bau[n] 0-based polygon index or -1 (R/geometryfns.R:1266,1680; semantics of frk_polygon_lookup).                        */
int frk_polygon_lookup_host(frk_ctx *ctx, int npoly, const int *h_ptr, const double *h_vx, const double *h_vy, int64_t n,
                            const double *h_xy, int *h_bau);
/* eval_basis on host coordinates -> host CSR.  Call with h_col == NULL to get nnz only.     */
int frk_eval_basis_host(frk_ctx *ctx, const frk_basis *b, int64_t n, const double *h_s,
                        int normalise, int *h_rowptr, int *h_col, double *h_val, int64_t *h_nnz);
/* .SRE.predict_EM (R/SREpredict.R:243-420) from host buffers -- what the R shim binds for predict().  S0 as host CSR (the
 * dgRMatrix slots of object@S0, 0-based), h_XBAU n x p column-major, h_fs[n]; h_obs_bau[m] (0-based BAU row of every observation,
 * point data; may be NULL for obs_fs != 0) or the incidence CSR of Cmat for areal data (h_cz_*, may be NULL); mP > 0 with h_cp_*:
 * prediction over polygons (C_P as CSR, rows normalised), outputs of length mP; otherwise outputs of length n.           */
int frk_model_predict_host(frk_model *mdl, int64_t n, const int *h_rowptr0, const int *h_col0, const double *h_val0,
                           const double *h_XBAU, const double *h_fs, const int *h_obs_bau, const int *h_cz_rowptr,
                           const int *h_cz_members, const double *h_cz_weights, int obs_fs, int64_t mP,
                           const int *h_cp_rowptr, const int *h_cp_members, const double *h_cp_weights, double *h_mu,
                           double *h_var, double *h_sd);
int frk_grid_lookup_host(frk_ctx *ctx, int64_t n, const double *h_xy, const double *h_offset,
                         const double *h_cellsize, const int *h_dims, const int *h_cell2bau,
                         int64_t ncell, int *h_bau);
/* var = diag(S0 Cov S0'), smu = S0 mu for a host CSR (R/SREutils.R:245-305)                 */
int frk_quadform_host(frk_ctx *ctx, int64_t n, int r, const int *h_rowptr, const int *h_col,
                      const double *h_val, const double *h_M, const double *h_mu,
                      double *h_q, double *h_t);
/* chol2inv(chol(A)) + log-determinant for a host matrix (R/SREfit.R:178-179,253,273)        */
int frk_chol2inv_host(frk_ctx *ctx, int r, const double *h_A, double *h_Ainv, double *h_logdet);

#ifdef __cplusplus
}
#endif
#endif /* FRKB200_H */

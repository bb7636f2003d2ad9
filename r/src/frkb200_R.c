/*
 * frkb200_R.c -- the `.Call` shim between the R package FRK and libfrkb200.so.
 *
 * NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no R (no R.h / Rinternals.h / libR).  It is the
 * binding a maintainer adds to FRK's src/ (next to src/FRK-init.h, whose registration pattern it follows:
 * R_CallMethodDef + R_registerRoutines + R_useDynamicSymbols(FALSE), src/FRK-init.h:38-55).  It only unpacks
 * SEXP arguments, calls the C ABI declared in include/frkb200.h, and packs the results; every call that the
 * Python ctypes harness (frk_b200/_lib.py) makes in the test-suite maps 1:1 onto a call below.
 *
 * Build (on a machine with R and CUDA):
 *   R CMD SHLIB frkb200_R.c -I../../include -L../../frk_b200 -lfrkb200 -Wl,-rpath,<dir of libfrkb200.so>
 *
 * Conventions: R matrices are column-major doubles (as the C ABI expects); R indices are 1-based and are
 * converted here; inputs are never modified (copy-on-modify); outputs are fresh allocVector()s; errors are raised
 * with Rf_error() only after every temporary has been released; the device-resident SRE object is handed to R as
 * an external pointer with a finalizer.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdlib.h>
#include <string.h>
#include "frkb200.h"

static frk_ctx *g_ctx = NULL;
static int g_live = 0;      /* external pointers (basis / model handles) whose finalizer has not run yet: they hold g_ctx */

static frk_ctx *ctx_or_error(void) {
    if (!g_ctx && frk_ctx_create(0, NULL, &g_ctx)) Rf_error("frkb200: %s", frk_last_error());
    return g_ctx;
}

#define FRK_TRY(expr, cleanup)                                                      \
    do { if ((expr) != 0) { cleanup; Rf_error("frkb200: %s", frk_last_error()); } } while (0)

/* ---- basis handle: local_basis()/auto_basis() result -> device cell lists ------------------------------------ */
static void basis_finalizer(SEXP p) {
    if (!R_ExternalPtrAddr(p)) return;
    frk_basis_destroy((frk_basis *)R_ExternalPtrAddr(p));
    R_ClearExternalPtr(p);
    g_live--;
}

/* .Call(frkb200_basis_create, manifold_code, radius, type_code, loc (r x d), scale (r), res (r)) */
SEXP frkb200_basis_create(SEXP manifold, SEXP radius, SEXP type, SEXP loc, SEXP scale, SEXP res) {
    frk_basis *b = NULL;
    int r = Rf_nrows(loc), d = Rf_ncols(loc);
    FRK_TRY(frk_basis_create(ctx_or_error(), Rf_asInteger(manifold), Rf_asReal(radius), Rf_asInteger(type), d, r, REAL(loc),
                             REAL(scale), INTEGER(res), &b), (void)0);
    SEXP p = PROTECT(R_MakeExternalPtr(b, R_NilValue, R_NilValue));
    g_live++;
    R_RegisterCFinalizerEx(p, basis_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

/* eval_basis(Basis, matrix): returns list(p = rowptr (n+1), j = col (nnz), x = val (nnz)) -- the slots of a dgRMatrix
 * (R/basisfns.R:709-748).  The overlay wraps it with new("dgRMatrix", ...) and as(., "CsparseMatrix"). */
SEXP frkb200_eval_basis(SEXP basis, SEXP s, SEXP normalise) {
    frk_basis *b = (frk_basis *)R_ExternalPtrAddr(basis);
    int64_t n = Rf_nrows(s), nnz = 0;
    SEXP rp = PROTECT(Rf_allocVector(INTSXP, n + 1));
    FRK_TRY(frk_eval_basis_host(ctx_or_error(), b, n, REAL(s), Rf_asInteger(normalise), INTEGER(rp), NULL, NULL, &nnz), UNPROTECT(1));
    SEXP cj = PROTECT(Rf_allocVector(INTSXP, nnz)), x = PROTECT(Rf_allocVector(REALSXP, nnz));
    FRK_TRY(frk_eval_basis_host(g_ctx, b, n, REAL(s), Rf_asInteger(normalise), INTEGER(rp), INTEGER(cj), REAL(x), &nnz), UNPROTECT(3));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, rp); SET_VECTOR_ELT(out, 1, cj); SET_VECTOR_ELT(out, 2, x);
    UNPROTECT(4);
    return out;
}

/* sp::over(SpatialPoints, SpatialPixels) inside map_data_to_BAUs (R/geometryfns.R:1266): 1-based BAU row, NA outside */
SEXP frkb200_grid_lookup(SEXP xy, SEXP offset, SEXP cellsize, SEXP dims, SEXP cell2bau) {
    int64_t n = Rf_nrows(xy);
    SEXP out = PROTECT(Rf_allocVector(INTSXP, n));
    int ncell = Rf_length(cell2bau), *map = (int *)malloc(sizeof(int) * (ncell ? ncell : 1));
    for (int i = 0; i < ncell; i++) map[i] = INTEGER(cell2bau)[i] == NA_INTEGER ? -1 : INTEGER(cell2bau)[i] - 1;
    FRK_TRY(frk_grid_lookup_host(ctx_or_error(), n, REAL(xy), REAL(offset), REAL(cellsize), INTEGER(dims), ncell ? map : NULL, ncell,
                                 INTEGER(out)), (free(map), UNPROTECT(1)));
    free(map);
    for (int64_t i = 0; i < n; i++) INTEGER(out)[i] = INTEGER(out)[i] < 0 ? NA_INTEGER : INTEGER(out)[i] + 1;
    UNPROTECT(1);
    return out;
}

/* ---- the SRE object on the device ------------------------------------------------------------------------------ */
static void model_finalizer(SEXP p) {
    if (!R_ExternalPtrAddr(p)) return;
    frk_model_destroy((frk_model *)R_ExternalPtrAddr(p));
    R_ClearExternalPtr(p);
    g_live--;
}

static frk_model *model_or_error(SEXP mdl) {
    frk_model *m = model_or_error(mdl);
    if (!m) Rf_error("frkb200: the device model of this SRE object is gone (restored from disk?): re-create it with the overlay's .frkb200_model()");
    return m;
}

/* .Call(frkb200_model_create, Sp, Sj, Sx (dgRMatrix slots of object@S, 0-based), r, Z, X, diag(Ve), diag(Vfs)) */
SEXP frkb200_model_create(SEXP Sp, SEXP Sj, SEXP Sx, SEXP r, SEXP Z, SEXP X, SEXP Ve, SEXP Vfs) {
    frk_model *m = NULL;
    FRK_TRY(frk_model_create_host(ctx_or_error(), Rf_length(Z), Rf_asInteger(r), Rf_ncols(X), INTEGER(Sp), INTEGER(Sj), REAL(Sx),
                                  REAL(Z), REAL(X), REAL(Ve), REAL(Vfs), &m), (void)0);
    SEXP p = PROTECT(R_MakeExternalPtr(m, R_NilValue, R_NilValue));
    g_live++;
    R_RegisterCFinalizerEx(p, model_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

/* D_basis (list of per-resolution distance matrices, BuildD R/basisfns.R:1131-1153) + basis@df$res */
SEXP frkb200_model_set_blocks(SEXP mdl, SEXP res, SEXP Dlist) {
    int nres = Rf_length(Dlist);
    const double **D = (const double **)malloc(sizeof(double *) * nres);
    for (int i = 0; i < nres; i++) D[i] = REAL(VECTOR_ELT(Dlist, i));
    FRK_TRY(frk_model_set_blocks(model_or_error(mdl), nres, INTEGER(res), D), free(D));
    free(D);
    return R_NilValue;
}

/* sp::over(SpatialPoints, SpatialPolygons BAUs): ring k = (vx, vy)[ptr[k] .. ptr[k+1]) (outer ring of BAU k, as R/geometryfns.R:1680
 * passes them); returns the 1-based BAU row or NA */
SEXP frkb200_polygon_lookup(SEXP ptr, SEXP vx, SEXP vy, SEXP xy) {
    frk_ctx *ctx = ctx_or_error();
    R_xlen_t n = Rf_nrows(xy);
    int npoly = Rf_length(ptr) - 1;
    int *idx = (int *)R_alloc((size_t)(n > 0 ? n : 1), sizeof(int));
    FRK_TRY(frk_polygon_lookup_host(ctx, npoly, INTEGER(ptr), REAL(vx), REAL(vy), (int64_t)n, REAL(xy), idx), (void)0);
    SEXP out = PROTECT(Rf_allocVector(INTSXP, n));
    for (R_xlen_t i = 0; i < n; i++) INTEGER(out)[i] = idx[i] < 0 ? NA_INTEGER : idx[i] + 1;
    UNPROTECT(1);
    return out;
}

/* non-diagonal object@Vfs (overlapping areal supports, average_in_BAU = FALSE): the dgRMatrix slots p, j, x (R/SRE.R:498) */
SEXP frkb200_model_set_vfs(SEXP mdl, SEXP p, SEXP j, SEXP x) {
    FRK_TRY(frk_model_set_vfs_csr(model_or_error(mdl), INTEGER(p), INTEGER(j), REAL(x)), (void)0);
    return R_NilValue;
}

/* push the slots of the S4 object (so that saveRDS()/readRDS() between SRE.fit calls resumes EM) */
SEXP frkb200_model_set(SEXP mdl, SEXP name, SEXP value) {
    FRK_TRY(frk_model_set(model_or_error(mdl), CHAR(STRING_ELT(name, 0)), REAL(value)), (void)0);
    return R_NilValue;
}

SEXP frkb200_model_get(SEXP mdl, SEXP name, SEXP len) {
    SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)Rf_asReal(len)));
    FRK_TRY(frk_model_get(model_or_error(mdl), CHAR(STRING_ELT(name, 0)), REAL(out)), UNPROTECT(1));
    UNPROTECT(1);
    return out;
}

/* execution policy of optim() inside .update_K (no reference counterpart): candidate points factorised side by side */
SEXP frkb200_model_set_lanes(SEXP mdl, SEXP lanes) {
    FRK_TRY(frk_model_set_lanes(model_or_error(mdl), Rf_asInteger(lanes)), (void)0);
    return R_NilValue;
}

/* .EM_fit (R/SREfit.R:78-160) in one call: returns list(llk = numeric(niter), niter) */
SEXP frkb200_em_fit(SEXP mdl, SEXP n_EM, SEXP tol, SEXP K_type, SEXP lambda, SEXP res, SEXP known_sigma2fs) {
    int n = Rf_asInteger(n_EM), niter = 0, known = !Rf_isNull(known_sigma2fs);
    double *llk = (double *)calloc(n > 0 ? n : 1, sizeof(double));
    FRK_TRY(frk_model_em_fit(model_or_error(mdl), n, Rf_asReal(tol), Rf_asInteger(K_type), Rf_length(lambda), REAL(lambda),
                             INTEGER(res), known, known ? Rf_asReal(known_sigma2fs) : 0.0, llk, &niter), free(llk));
    SEXP v = PROTECT(Rf_allocVector(REALSXP, niter));
    memcpy(REAL(v), llk, sizeof(double) * niter);
    free(llk);
    R_CheckUserInterrupt();                 /* (may longjmp: nothing malloc'ed is live any more; v is on the protect stack) */
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, v); SET_VECTOR_ELT(out, 1, Rf_ScalarInteger(niter));
    UNPROTECT(2);
    return out;
}

/* csr = NULL or list(p, j, x) of an RsparseMatrix (0-based) */
static const int *csr_p(SEXP l) { return Rf_isNull(l) ? NULL : INTEGER(VECTOR_ELT(l, 0)); }
static const int *csr_j(SEXP l) { return Rf_isNull(l) ? NULL : INTEGER(VECTOR_ELT(l, 1)); }
static const double *csr_x(SEXP l) { return Rf_isNull(l) ? NULL : REAL(VECTOR_ELT(l, 2)); }

/* .SRE.predict_EM (R/SREpredict.R:243-420): BAU-level prediction (point or areal data) or prediction over polygons.
 * .Call(frkb200_model_predict, mdl, S0 = list(p, j, x), X_BAU (N x p), fs (N), obs_bau (0-based integer(m)) or NULL,
 *       CZ = list(p, j, x) or NULL (areal data), obs_fs, CP = list(p, j, x) or NULL)  ->  list(mu, var, sd)                   */
SEXP frkb200_model_predict(SEXP mdl, SEXP S0, SEXP X, SEXP fs, SEXP obs_bau, SEXP CZ, SEXP obs_fs, SEXP CP) {
    const int64_t n = Rf_length(VECTOR_ELT(S0, 0)) - 1;
    const int64_t mP = Rf_isNull(CP) ? 0 : Rf_length(VECTOR_ELT(CP, 0)) - 1;
    const int64_t nout = mP > 0 ? mP : n;
    SEXP mu = PROTECT(Rf_allocVector(REALSXP, nout)), var = PROTECT(Rf_allocVector(REALSXP, nout)), sd = PROTECT(Rf_allocVector(REALSXP, nout));
    FRK_TRY(frk_model_predict_host(model_or_error(mdl), n, csr_p(S0), csr_j(S0), csr_x(S0), REAL(X), REAL(fs),
                                   Rf_isNull(obs_bau) ? NULL : INTEGER(obs_bau), csr_p(CZ), csr_j(CZ), csr_x(CZ), Rf_asLogical(obs_fs), mP,
                                   csr_p(CP), csr_j(CP), csr_x(CP), REAL(mu), REAL(var), REAL(sd)), UNPROTECT(3));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, mu); SET_VECTOR_ELT(out, 1, var); SET_VECTOR_ELT(out, 2, sd);
    UNPROTECT(4);
    return out;
}

/* logLik(SRE) (R/SREutils.R:22-32) at the slots currently on the device */
SEXP frkb200_model_loglik(SEXP mdl) {
    double v = 0.0;
    FRK_TRY(frk_model_loglik(model_or_error(mdl), &v), (void)0);
    return Rf_ScalarReal(v);
}

/* multi-GPU, one R process per GPU: rank 0 calls frkb200_comm_unique_id() and hands the raw vector to the other processes
 * (a file, Rmpi, a socket); every process then calls frkb200_comm_init(rank, world, id) BEFORE creating models.  From then on
 * the row-partitioned sums of the fit go over NCCL inside libfrkb200 (include/frkb200.h, frk_comm_*). */
SEXP frkb200_comm_unique_id(void) {
    SEXP id = PROTECT(Rf_allocVector(RAWSXP, FRK_COMM_ID_BYTES));
    FRK_TRY(frk_comm_unique_id(RAW(id)), UNPROTECT(1));
    UNPROTECT(1);
    return id;
}

SEXP frkb200_comm_init(SEXP rank, SEXP world, SEXP id) {
    if (Rf_length(id) != FRK_COMM_ID_BYTES) Rf_error("frkb200: the NCCL unique id has %d bytes", FRK_COMM_ID_BYTES);
    FRK_TRY(frk_comm_init(ctx_or_error(), Rf_asInteger(rank), Rf_asInteger(world), RAW(id)), (void)0);
    return R_NilValue;
}

/* .batch_compute_var(S0, Cov) (R/SREutils.R:245-305) and S0 %*% mu for a host dgRMatrix: list(var, smu) */
SEXP frkb200_quadform(SEXP Sp, SEXP Sj, SEXP Sx, SEXP Cov, SEXP mu) {
    int64_t n = Rf_length(Sp) - 1;
    int r = Rf_nrows(Cov);
    SEXP q = PROTECT(Rf_allocVector(REALSXP, n)), t = PROTECT(Rf_allocVector(REALSXP, n));
    FRK_TRY(frk_quadform_host(ctx_or_error(), n, r, INTEGER(Sp), INTEGER(Sj), REAL(Sx), REAL(Cov), REAL(mu), REAL(q), REAL(t)), UNPROTECT(2));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, q); SET_VECTOR_ELT(out, 1, t);
    UNPROTECT(3);
    return out;
}

/* chol2inv(chol(A)) with attr "logdet" = determinant(A, logarithm = TRUE)$modulus (R/SREfit.R:178-179,253,273) */
SEXP frkb200_chol2inv(SEXP A) {
    int r = Rf_nrows(A);
    double ld = 0.0;
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, r, r));
    FRK_TRY(frk_chol2inv_host(ctx_or_error(), r, REAL(A), REAL(out), &ld), UNPROTECT(1));
    Rf_setAttrib(out, Rf_install("logdet"), Rf_ScalarReal(ld));
    UNPROTECT(1);
    return out;
}

static const R_CallMethodDef CallEntries[] = {
    {"frkb200_basis_create", (DL_FUNC)&frkb200_basis_create, 6},
    {"frkb200_eval_basis", (DL_FUNC)&frkb200_eval_basis, 3},
    {"frkb200_grid_lookup", (DL_FUNC)&frkb200_grid_lookup, 5},
    {"frkb200_model_create", (DL_FUNC)&frkb200_model_create, 8},
    {"frkb200_model_set_blocks", (DL_FUNC)&frkb200_model_set_blocks, 3},
    {"frkb200_model_set_vfs", (DL_FUNC)&frkb200_model_set_vfs, 4},
    {"frkb200_polygon_lookup", (DL_FUNC)&frkb200_polygon_lookup, 4},
    {"frkb200_model_set", (DL_FUNC)&frkb200_model_set, 3},
    {"frkb200_model_get", (DL_FUNC)&frkb200_model_get, 3},
    {"frkb200_model_set_lanes", (DL_FUNC)&frkb200_model_set_lanes, 2},
    {"frkb200_em_fit", (DL_FUNC)&frkb200_em_fit, 7},
    {"frkb200_model_predict", (DL_FUNC)&frkb200_model_predict, 8},
    {"frkb200_model_loglik", (DL_FUNC)&frkb200_model_loglik, 1},
    {"frkb200_comm_unique_id", (DL_FUNC)&frkb200_comm_unique_id, 0},
    {"frkb200_comm_init", (DL_FUNC)&frkb200_comm_init, 3},
    {"frkb200_quadform", (DL_FUNC)&frkb200_quadform, 5},
    {"frkb200_chol2inv", (DL_FUNC)&frkb200_chol2inv, 1},
    {NULL, NULL, 0}};

void R_init_frkb200(DllInfo *dll) {
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}

void R_unload_frkb200(DllInfo *dll) {
    (void)dll;
    /* handles that are still alive hold the context (their finalizers run later, onexit = TRUE, and dereference it):
     * destroy it only when none is left -- otherwise it is reclaimed with the process */
    if (g_ctx && g_live == 0) { frk_ctx_destroy(g_ctx); g_ctx = NULL; }
}

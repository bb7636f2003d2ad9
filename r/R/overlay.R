## overlay.R -- re-defines FRK's hot internal functions so that they call libfrkb200 through the .Call shim
## (r/src/frkb200_R.c).  NOT RUN IN THIS REPOSITORY (no R in the build image); it shows the R side of the drop-in:
## every exported signature in FRK's NAMESPACE stays unchanged, only internals are swapped with assignInNamespace().
##
##   library(FRK); dyn.load("frkb200_R.so"); source("overlay.R"); frkb200_enable()
##
## There is no CPU fallback: a basis or measure the device path does not implement is an error, not a silent R path.

.frkb200_basis_type <- function(fn) {
  ## the basis type is not a slot (R/AllClass.R:78): recover it from the closure environment the way
  ## R/rhipe_fns.R:65-78 does (R/basisfns.R:1213-1250: R -> bisquare, std -> Gaussian, tau -> exp, kappa -> Matern32)
  vars <- ls(environment(fn))
  if ("R" %in% vars) 0L else if ("std" %in% vars) 1L else if ("tau" %in% vars) 2L else if ("kappa" %in% vars) 3L else
    stop("frkb200: custom basis functions have no CUDA kernel (no CPU fallback)")
}

.frkb200_manifold <- function(m) {
  ## STsphere does not inherit from sphere (nor STplane from plane) in FRK's class tree: test the ST classes first
  if (is(m, "STplane")) 4L else if (is(m, "STsphere")) 5L else
  if (is(m, "real_line")) 1L else if (is(m, "plane")) 2L else if (is(m, "sphere")) 3L else
    stop("frkb200: only real_line(), plane(), sphere(), STplane() and STsphere() with their built-in measures are on the device path")
}

.frkb200_handle <- function(basis) {
  .Call("frkb200_basis_create", .frkb200_manifold(basis@manifold),
        if (is(basis@manifold, "sphere") || is(basis@manifold, "STsphere")) basis@manifold@radius else 0,
        .frkb200_basis_type(basis@fn[[1]]), as.matrix(basis@df[, seq_len(dimensions(basis@manifold)), drop = FALSE]),
        as.numeric(basis@df$scale), as.integer(basis@df$res))
}

## eval_basis(Basis, matrix), R/basisfns.R:709-748
.frkb200_eval_basis <- function(basis, s) {
  stopifnot(is.numeric(s))
  out <- .Call("frkb200_eval_basis", .frkb200_handle(basis), as.matrix(s) + 0, 0L)
  S <- new("dgRMatrix", p = out[[1]], j = out[[2]], x = out[[3]], Dim = c(nrow(s), nbasis(basis)))
  as(S, "CsparseMatrix")
}

## The device-resident SRE object for `object`: created from the S4 slots (R/AllClass.R:133-175) and filled with the current
## parameter slots, so that an object restored with readRDS() continues where it stopped.
.frkb200_model <- function(object) {
  if (object@fs_model != "ind") stop("frkb200: only fs_model = \"ind\" is on the device path")
  ## Ve is diagonal in every model SRE() builds (R/SRE.R:468); anything else is refused rather than fitted as if it were
  if (!Matrix::isDiagonal(object@Ve)) stop("frkb200: a non-diagonal Ve is not on the device path")
  S <- as(object@S, "RsparseMatrix")
  mdl <- .Call("frkb200_model_create", S@p, S@j, S@x, ncol(S), as.numeric(object@Z), as.matrix(object@X),
               as.numeric(diag(object@Ve)), as.numeric(diag(object@Vfs)))
  ## overlapping areal supports / average_in_BAU = FALSE: Vfs has off-diagonal entries and the EM takes the general branches
  ## (chol(D), J with Dinv %*% Vfs %*% Dinv, R/SREfit.R:186-189,240-252,296-328) -- the library does too, given the sparse matrix
  if (!Matrix::isDiagonal(object@Vfs)) {
    V <- as(as(object@Vfs, "generalMatrix"), "RsparseMatrix")
    .Call("frkb200_model_set_vfs", mdl, V@p, V@j, as.numeric(V@x))
  }
  if (object@K_type == "block-exponential")
    .Call("frkb200_model_set_blocks", mdl, as.integer(object@basis@df$res), lapply(object@D_basis, as.matrix))
  for (slot in c("mu_eta", "S_eta", "Q_eta", "Khat_inv", "alphahat", "Khat"))   # Khat last: it refreshes log|K|
    .Call("frkb200_model_set", mdl, slot, as.numeric(as.matrix(slot(object, slot))))
  .Call("frkb200_model_set", mdl, "sigma2fshat", as.numeric(object@sigma2fshat))
  mdl
}

## .EM_fit, R/SREfit.R:78-160: one .Call for the whole loop; the slots come back into the S4 object afterwards
.frkb200_EM_fit <- function(object, n_EM, lambda, tol, print_lik, known_sigma2fs, taper) {
  if (!is.null(taper)) stop("frkb200: covariance tapering is not on the device path")
  mdl <- .frkb200_model(object)
  .Call("frkb200_model_set_lanes", mdl, 2L)      # two Nelder-Mead candidates per round (the measured optimum on B200)
  time <- system.time(
    fit <- .Call("frkb200_em_fit", mdl, as.integer(n_EM), as.numeric(tol),
                 if (object@K_type == "unstructured") 0L else 1L, as.numeric(lambda),
                 as.integer(object@basis@df$res), known_sigma2fs))
  r <- nbasis(object@basis)
  object@mu_eta      <- Matrix(.Call("frkb200_model_get", mdl, "mu_eta", r))
  for (slot in c("S_eta", "Q_eta", "Khat", "Khat_inv"))
    slot(object, slot) <- Matrix(matrix(.Call("frkb200_model_get", mdl, slot, r * r), r, r))
  object@alphahat    <- Matrix(.Call("frkb200_model_get", mdl, "alphahat", ncol(object@X)))
  object@sigma2fshat <- .Call("frkb200_model_get", mdl, "sigma2fshat", 1)
  object@info_fit <- list(method = "EM", time = time, num_iterations = fit[[2]], converged = as.numeric(fit[[2]] != n_EM),
                          plot_lik = list(x = seq_len(fit[[2]]), llk = fit[[1]], ylab = "log likelihood", xlab = "EM iteration"),
                          sigma2fshat_equal_0 = as.numeric(object@sigma2fshat == 0),
                          frkb200_model = mdl)   # the device object stays alive with the fitted SRE object: predict() reuses it
  object@method <- "EM"
  object
}

.frkb200_csr <- function(M) { M <- as(M, "RsparseMatrix"); list(M@p, M@j, as.numeric(M@x)) }

## .SRE.predict_EM, R/SREpredict.R:243-420 -- BAU-level prediction for point and (disjoint-support) areal data in both cases of
## obs_fs, and prediction over arbitrary polygons (CP); ONE .Call.  The (r + N)-dimensional sparse factorisation + Takahashi
## inverse of the reference (:308-333) is replaced by the closed forms of DESIGN.md s2 on the device.
.frkb200_predict_EM <- function(object, obs_fs = FALSE, newdata = NULL, pred_time = NULL, covariances = FALSE, CP, predict_BAUs) {
  if (covariances) stop("frkb200: covariances = TRUE (a dense mP x mP matrix) is not on the device path")
  BAUs <- object@BAUs
  mdl <- object@info_fit$frkb200_model
  if (is.null(mdl) || !is(mdl, "externalptr") || identical(mdl, new("externalptr")))
    mdl <- .frkb200_model(object)                               # e.g. after readRDS(): rebuild from the slots
  else {                                                        # the user may have edited slots since the fit (S@sigma2fshat <- 0, ...)
    for (slot in c("mu_eta", "S_eta", "alphahat"))
      .Call("frkb200_model_set", mdl, slot, as.numeric(as.matrix(slot(object, slot))))
    .Call("frkb200_model_set", mdl, "sigma2fshat", as.numeric(object@sigma2fshat))
  }
  X <- as.matrix(FRK:::.extract_BAU_X_matrix(object@f, BAUs))
  CZ <- as(object@Cmat, "RsparseMatrix")
  point_data <- all(diff(CZ@p) == 1L) && !anyDuplicated(CZ@j)   # one BAU per observation and one observation per BAU (R/geometryfns.R:1557-1568);
                                                                # anything else (areal supports, average_in_BAU = FALSE) passes Cmat itself
  polygons <- !(predict_BAUs & !covariances)
  out <- .Call("frkb200_model_predict", mdl, .frkb200_csr(object@S0), X + 0, as.numeric(BAUs$fs),
               if (point_data) as.integer(CZ@j) else NULL,
               if (point_data) NULL else list(CZ@p, CZ@j, as.numeric(CZ@x)),
               as.logical(obs_fs), if (polygons) .frkb200_csr(CP) else NULL)
  if (!polygons) {
    BAUs[["mu"]] <- out[[1]]; BAUs[["var"]] <- out[[2]]; BAUs[["sd"]] <- out[[3]]
    if (is(newdata, "Spatial")) BAUs <- BAUs[row.names(newdata), ]
    if (!is.null(pred_time)) BAUs <- BAUs[, pred_time]
    return(BAUs)
  }
  newdata[["mu"]] <- out[[1]]; newdata[["var"]] <- out[[2]]; newdata[["sd"]] <- out[[3]]
  newdata
}

## sp::over(points, SpatialPixels BAUs) inside map_data_to_BAUs (R/geometryfns.R:1266, through .parallel_over :1675-1680): the
## integer point-in-cell lookup on the device.  Other BAU classes / fn != NULL keep the reference's call.
.frkb200_parallel_over <- function(sp1, sp2, fn = NULL, batch_size = NULL) {
  if (is.null(fn) && is(sp1, "SpatialPoints") && is(sp2, "SpatialPixelsDataFrame")) {
    g <- sp2@grid
    cell2bau <- rep(NA_integer_, prod(g@cells.dim))
    cell2bau[sp2@grid.index] <- seq_along(sp2@grid.index)
    idx <- .Call("frkb200_grid_lookup", coordinates(sp1) + 0, as.numeric(g@cellcentre.offset), as.numeric(g@cellsize),
                 as.integer(g@cells.dim), cell2bau)
    return(sp2@data[idx, , drop = FALSE])
  }
  if (is.null(fn) && is(sp1, "SpatialPoints") && is(sp2, "SpatialPolygonsDataFrame") &&
      all(vapply(sp2@polygons, function(p) length(p@Polygons) == 1L, TRUE))) {
    ## polygon BAUs with one ring each (ISEA3H hexagons, BAUs_from_points): the point-in-polygon test on the device
    rings <- lapply(sp2@polygons, function(p) p@Polygons[[1]]@coords)
    ptr <- c(0L, cumsum(vapply(rings, nrow, 1L)))
    xy <- do.call(rbind, rings)
    idx <- .Call("frkb200_polygon_lookup", as.integer(ptr), as.numeric(xy[, 1]), as.numeric(xy[, 2]), coordinates(sp1) + 0)
    return(sp2@data[idx, , drop = FALSE])
  }
  sp::over(sp1, sp2, fn = fn)
}

## .batch_compute_var, R/SREutils.R:245-305.  fs_in_process = FALSE: diag(S0 Cov S0') on the device.  fs_in_process = TRUE is
## only reached from the reference's own .SRE.predict_EM (R/SREpredict.R:328), which frkb200_enable() replaces as a whole; a direct
## caller gets the reference's own implementation (kept below as .ref_batch_compute_var), never a wrong answer or an error.
.frkb200_batch_compute_var <- function(S0, Cov, fs_in_process = FALSE) {
  if (fs_in_process) return(.ref_batch_compute_var(S0, Cov, fs_in_process = TRUE))
  S <- as(S0, "RsparseMatrix")
  r <- ncol(S)
  .Call("frkb200_quadform", S@p, S@j, S@x, as.matrix(Cov[seq_len(r), seq_len(r)]), numeric(r))[[1]]
}

frkb200_enable <- function() {
  assign(".ref_batch_compute_var", get(".batch_compute_var", envir = asNamespace("FRK")), envir = globalenv())
  assignInNamespace(".EM_fit", .frkb200_EM_fit, ns = "FRK")
  assignInNamespace(".SRE.predict_EM", .frkb200_predict_EM, ns = "FRK")
  assignInNamespace(".parallel_over", .frkb200_parallel_over, ns = "FRK")
  assignInNamespace(".batch_compute_var", .frkb200_batch_compute_var, ns = "FRK")
  setMethod("eval_basis", signature(basis = "Basis", s = "matrix"), .frkb200_eval_basis, where = asNamespace("FRK"))
  invisible(TRUE)
}

## One R process per GPU (e.g. under mpirun): rank 0 writes the NCCL unique id to `id_file`, the others read it; afterwards every model
## created in this process reduces its row-partitioned sums over NCCL inside libfrkb200 (frk_comm_*, include/frkb200.h).
frkb200_init_comm <- function(rank, world, id_file) {
  if (rank == 0) { id <- .Call("frkb200_comm_unique_id"); writeBin(id, id_file) }
  else { while (!file.exists(id_file) || file.size(id_file) < 128) Sys.sleep(0.01); id <- readBin(id_file, "raw", 128) }
  .Call("frkb200_comm_init", as.integer(rank), as.integer(world), id)
}

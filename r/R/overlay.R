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

## .EM_fit, R/SREfit.R:78-160: one .Call for the whole loop; the slots come back into the S4 object afterwards
.frkb200_EM_fit <- function(object, n_EM, lambda, tol, print_lik, known_sigma2fs, taper) {
  if (!is.null(taper)) stop("frkb200: covariance tapering is not on the device path")
  S <- as(object@S, "RsparseMatrix")
  mdl <- .Call("frkb200_model_create", S@p, S@j, S@x, ncol(S), as.numeric(object@Z), as.matrix(object@X),
               as.numeric(diag(object@Ve)), as.numeric(diag(object@Vfs)))
  if (object@K_type == "block-exponential")
    .Call("frkb200_model_set_blocks", mdl, as.integer(object@basis@df$res), lapply(object@D_basis, as.matrix))
  for (slot in c("mu_eta", "S_eta", "Q_eta", "Khat_inv", "alphahat", "Khat"))   # Khat last: it refreshes log|K|
    .Call("frkb200_model_set", mdl, slot, as.numeric(as.matrix(slot(object, slot))))
  .Call("frkb200_model_set", mdl, "sigma2fshat", as.numeric(object@sigma2fshat))
  .Call("frkb200_model_set_lanes", mdl, 2L)      # two Nelder-Mead candidates per round (the measured optimum on B200)
  time <- system.time(
    fit <- .Call("frkb200_em_fit", mdl, as.integer(n_EM), as.numeric(tol),
                 if (object@K_type == "unstructured") 0L else 1L, as.numeric(lambda),
                 as.integer(object@basis@df$res), known_sigma2fs))
  r <- ncol(S)
  object@mu_eta      <- Matrix(.Call("frkb200_model_get", mdl, "mu_eta", r))
  for (slot in c("S_eta", "Q_eta", "Khat", "Khat_inv"))
    slot(object, slot) <- Matrix(matrix(.Call("frkb200_model_get", mdl, slot, r * r), r, r))
  object@alphahat    <- Matrix(.Call("frkb200_model_get", mdl, "alphahat", ncol(object@X)))
  object@sigma2fshat <- .Call("frkb200_model_get", mdl, "sigma2fshat", 1)
  object@info_fit <- list(method = "EM", time = time, num_iterations = fit[[2]], converged = as.numeric(fit[[2]] != n_EM),
                          plot_lik = list(x = seq_len(fit[[2]]), llk = fit[[1]], ylab = "log likelihood", xlab = "EM iteration"),
                          sigma2fshat_equal_0 = as.numeric(object@sigma2fshat == 0))
  object@method <- "EM"
  object
}

## .batch_compute_var, R/SREutils.R:245-305 (fs_in_process = FALSE branch; the Case-2 closed form of SURVEY.md s8(a)
## supplies the fs_in_process = TRUE result for point data)
.frkb200_batch_compute_var <- function(S0, Cov, fs_in_process = FALSE) {
  if (fs_in_process) stop("frkb200: use the closed-form Case-2 prediction for point data (see INTEGRATION.md)")
  S <- as(S0, "RsparseMatrix")
  .Call("frkb200_quadform", S@p, S@j, S@x, as.matrix(Cov), numeric(ncol(S)))[[1]]
}

frkb200_enable <- function() {
  assignInNamespace(".EM_fit", .frkb200_EM_fit, ns = "FRK")
  assignInNamespace(".batch_compute_var", .frkb200_batch_compute_var, ns = "FRK")
  setMethod("eval_basis", signature(basis = "Basis", s = "matrix"), .frkb200_eval_basis, where = asNamespace("FRK"))
  invisible(TRUE)
}

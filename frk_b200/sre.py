"""SRE(), SRE.fit(method = "EM") and predict() -- host mirror of R/SRE.R, R/SREfit.R and
R/SREpredict.R for Gaussian data with identity link.

This module only assembles the model (formula, BAU fields, which rows this rank owns) and unpacks
arguments: the SRE object itself is the native ``frk_model`` of libfrkb200 (csrc/model.cu), and
``SRE_fit`` / ``predict`` are ONE C-ABI call each (``frk_model_em_fit`` / ``frk_model_predict``) --
exactly what the R ``.Call`` shim binds.  The EM control flow (Nelder-Mead on tau,
bracket search + uniroot on sigma2fs, p x p solves) runs on the host inside the library, as it runs
on the host in R; everything m-, N- or r^2-sized is a kernel.

With ``torch.distributed`` initialised the BAUs (and the observations that fall in them) are
row-partitioned over the ranks; the library calls back into ``Device.allreduce_hook`` for the one
large exchange per EM iteration (``[S'D^-1 S | S'D^-1 (z - X alpha) | 2 scalars]``, fp64 sum over
NCCL) and for the O(p^2) M-step moments.
"""
from __future__ import annotations

import ctypes as C
import re
import time
from typing import Optional

import numpy as np

from . import _lib
from ._lib import FRKError, call, c_i64, vp
from .basis import ImplicitBasisMatrix, BuildD, DeviceCSR, TensorP_Basis, eval_basis_device
from .geometry import GridBAUs, STBAUs, SpatialPointsDataFrame, SpatialPolygonsDataFrame, map_data_to_BAUs
from .runtime import Device, get_device, partition_rows

def _parse_formula(f: str):
    """``z ~ 1``, ``co2avgret ~ lat + 1``, ``z ~ -1 + x`` ...  (stats::model.frame as SRE() uses it,
    R/SRE.R:455-465): returns (response, covariates, intercept)."""
    lhs, rhs = f.split("~")
    resp = lhs.strip()
    terms = [t.strip() for t in re.split(r"\+", rhs.replace("-", "+-")) if t.strip()]
    intercept, covs = True, []
    for t in terms:
        if t in ("1",):
            intercept = True
        elif t in ("0", "-1", "- 1"):
            intercept = False
        else:
            covs.append(t)
    return resp, covs, intercept


class SREModel:
    """The ``SRE`` S4 object (R/AllClass.R:133-175), backed by a native ``frk_model``.  m- and N-sized slots live
    in HBM (this rank's rows only); r x r slots live in HBM and are replicated on every rank."""

    def __init__(self, dev, f, basis, BAUs, K_type, N, lo, hi, m, S0, S, obs_bau, Z, X, X_BAU, Ve, Vfs, fs_BAU, p):
        self.dev, self.f, self.basis, self.BAUs, self.K_type = dev, f, basis, BAUs, K_type
        self.r, self.p, self.N, self.lo, self.hi, self.m = basis.n, p, N, lo, hi, m
        self.S0, self.S, self.obs_bau = S0, S, obs_bau
        self.Z, self.X, self.X_BAU, self.Ve, self.Vfs, self.fs_BAU = Z, X, X_BAU, Ve, Vfs, fs_BAU     # keep the device arrays alive
        self.info_fit: dict = {}
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        # row-partitioned runs: the context's own NCCL communicator if it has one (Device.init_comm), else the Python reduction hook
        self._hook = dev.allreduce_hook() if (dev.world > 1 and not dev.has_comm) else _lib.ALLREDUCE_FN()     # must outlive the model
        h = vp()
        if isinstance(S, ImplicitBasisMatrix):
            call("frk_model_create_implicit", dev.ctx, m, p, S.basis.handle(dev), dev.ptr(S.coords), S.ld, 1 if S.normalise else 0,
                 dev.ptr(Z), dev.ptr(X), dev.ptr(Ve), dev.ptr(Vfs), self._hook, vp(None), C.byref(h))
        else:
            call("frk_model_create", dev.ctx, m, self.r, p, dev.ptr(S.rowptr), dev.ptr(S.col), dev.ptr(S.val), dev.ptr(Z), dev.ptr(X),
                 dev.ptr(Ve), dev.ptr(Vfs), self._hook, vp(None), C.byref(h))
        self.handle = h
        if dev.world > 1:
            call("frk_model_set_parallel", h, dev.rank, dev.world, 0)
        self.D_basis = None

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None:
                _lib.lib.frk_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- slots ------------------------------------------------------------------------------------
    def get(self, name: str) -> np.ndarray:
        """Host copy of a slot (what the R shim writes back into the S4 object); r x r slots as (r, r) arrays."""
        n = {"mu_eta": self.r, "alphahat": self.p, "sigma2fshat": 1, "m_total": 1, "homoscedastic": 1,
             "tau": 16, "omega": 16, "nm_stats": 4, "nm_prefetch_hits": 1, "nm_reuse_hits": 1}.get(name, self.r * self.r)
        big = name in ("S_eta", "Q_eta", "Khat", "Khat_inv")
        out = self.dev.host_buffer(n) if big else np.zeros(n)       # pinned staging for the r x r slots
        call("frk_model_get", self.handle, name.encode(), vp(out.ctypes.data))
        if big:
            self.d2h_bytes += out.nbytes
            return out.reshape(self.r, self.r)                      # symmetric: column-major == row-major
        return out

    def set(self, name: str, value) -> None:
        a = np.asarray(value, dtype=np.float64)
        a = np.ascontiguousarray(a.T).ravel() if a.ndim == 2 else np.ascontiguousarray(a).ravel()
        call("frk_model_set", self.handle, name.encode(), vp(a.ctypes.data))

    @property
    def alphahat(self) -> np.ndarray:
        return self.get("alphahat")

    @property
    def sigma2fshat(self) -> float:
        return float(self.get("sigma2fshat")[0])

    @property
    def m_total(self) -> int:
        return int(self.get("m_total")[0])

    @property
    def homoscedastic(self) -> bool:
        return bool(self.get("homoscedastic")[0])


# ---------------------------------------------------------------------------------------
# SRE()
# ---------------------------------------------------------------------------------------

def SRE(f: str, data: SpatialPointsDataFrame, basis, BAUs, est_error: bool = False, average_in_BAU: bool = True,
        K_type: str = "block-exponential", normalise_basis: bool = True, dev: Optional[Device] = None,
        implicit_basis: Optional[bool] = None, balance: str = "BAUs", obs_weight: float = 13.0) -> SREModel:
    """SRE(f, data, basis, BAUs, est_error, average_in_BAU, ..., K_type, normalise_basis), R/SRE.R:318-701
    for one point-referenced Gaussian data set.

    Steps on the device (file:line of what each replaces): BAU centroids -> eval_basis -> row
    normalisation (R/SRE.R:412-421); map_data_to_BAUs (R/geometryfns.R:1222-1320); BuildC + Cmat %*% S0
    as a CSR row gather (R/SRE.R:472-502); Ve, Vfs (:468,:498); .EM_initialise (:704-731)."""
    K_type = K_type.lower()
    if K_type not in ("block-exponential", "unstructured"):
        raise ValueError("K_type needs to be \"block-exponential\" or \"unstructured\" for method = \"EM\"")   # check_args.R:184-187
    if est_error:
        raise FRKError("est_error=TRUE needs the gstat variogram fit (host, third-party; out of the hot-path scope): "
                       "supply a 'std' column and est_error=FALSE")
    if "std" not in data.data:
        raise ValueError("data needs a field 'std' (measurement-error standard deviation) when est_error=FALSE")
    if "fs" not in BAUs.fields:
        raise ValueError("BAUs should contain a field 'fs', containing a basis vector for fine-scale variation. "
                         "Do BAUs$fs <- 1 if you don't know what this should be.")     # R/check_args.R
    if np.any(BAUs.fields["fs"] < 0):
        raise ValueError("fine-scale variation basis function needs to be nonnegative everywhere")
    resp, covs, intercept = _parse_formula(f)
    if resp not in data.data:
        raise ValueError(f"response {resp!r} not found in data")
    for c in covs:
        if c not in BAUs.fields:
            raise ValueError(f"covariate {c!r} is not a field of the BAUs (covariates must be known at every BAU)")
    if not intercept and not covs:
        raise FRKError("a model without fixed effects is not on the device path (at least an intercept is needed)")
    dev = dev or get_device()
    t0 = time.time()
    r = basis.n
    N = len(BAUs)
    areal = isinstance(data, SpatialPolygonsDataFrame)
    if areal:
        if dev.world > 1:
            raise FRKError("areal (polygon) data on a row-partitioned run: footprints may straddle ranks -- not on the device path yet")
        if isinstance(basis, TensorP_Basis) or isinstance(BAUs, STBAUs):
            raise FRKError("areal (polygon) data in space-time (BuildC for STFDF, R/geometryfns.R:1628-1672) is not on the device path yet")
    lo, hi = partition_rows(N, dev.world, dev.rank)
    if balance == "observations" and dev.world > 1 and not areal and not isinstance(BAUs, STBAUs):
        # equal WORK per rank instead of equal BAU counts: a BAU row costs 1 (prediction), an observation in it `obs_weight` (the
        # EM passes over S; 13 = ten EM iterations against one prediction pass, measured on BASELINE config 5 where the
        # observations are uniform on the sphere but the lon/lat BAU rows are not)
        from .geometry import point_BAU_index
        from .runtime import partition_rows_weighted
        b_all = point_BAU_index(data, BAUs, dev)
        cnt = np.bincount(b_all[b_all >= 0], minlength=N).astype(np.float64)
        lo, hi = partition_rows_weighted(1.0 + float(obs_weight) * cnt, dev.world, dev.rank)
    elif balance not in ("BAUs", "observations"):
        raise ValueError("balance needs to be \"BAUs\" or \"observations\"")
    nloc = hi - lo
    h2d = 0
    # Basis functions without compact support give dense rows.  Beyond ~8 GB for this rank's S0 the matrices are not stored at all:
    # rows are evaluated in chunks inside every pass (ImplicitBasisMatrix; `implicit_basis` forces the choice either way)
    dense_rows = (not isinstance(basis, TensorP_Basis)) and getattr(basis, "type", "bisquare") != "bisquare"
    if implicit_basis is None:
        implicit_basis = dense_rows and nloc * r * 8 > 8e9
    if implicit_basis and (not dense_rows or areal or not average_in_BAU):
        raise FRKError("implicit_basis needs a plain basis without compact support, point data and average_in_BAU = TRUE")

    # ---- S0 = eval_basis(basis, coordinates(BAUs)); normalise rows ------------------------
    if isinstance(BAUs, GridBAUs) and not isinstance(basis, TensorP_Basis):
        d_xg = dev.upload(BAUs.xgrid)
        d_yg = dev.upload(BAUs.ygrid)
        h2d += 8 * (BAUs.nx + BAUs.ny)
        d_s = dev.empty(max(2 * nloc, 1))
        call("frk_grid_centroids", dev.ctx, lo, hi, BAUs.nx, dev.ptr(d_xg), dev.ptr(d_yg), dev.ptr(d_s))
        if implicit_basis:
            S0 = ImplicitBasisMatrix(basis, d_s, nloc, nloc, normalise_basis)
        else:
            S0 = eval_basis_device(basis, (d_s, nloc, 2), dev, normalise=normalise_basis)
        if not areal and not implicit_basis:
            del d_s
    else:
        cc = BAUs.centroids(lo, hi)
        h2d += cc.nbytes
        if implicit_basis:
            d_s = dev.upload(cc[:, :basis.manifold.dim])
            S0 = ImplicitBasisMatrix(basis, d_s, nloc, nloc, normalise_basis)
        else:
            S0 = eval_basis_device(basis, cc, dev, normalise=normalise_basis)
        if areal:
            d_s = dev.upload(cc[:, :2])

    binned = None
    obs_bau = None
    if areal:
        # ---- BuildC(SpatialPolygons): BAU centroids in observation polygons, rows normalised; S = C_Z S0 ---------------
        m = len(data)
        hp = vp()
        call("frk_polygons_create", dev.ctx, m, vp(data.ptr.ctypes.data), vp(data.vx.ctypes.data), vp(data.vy.ctypes.data), C.byref(hp))
        d_group = dev.empty(max(N, 1), "i4")
        multi = c_i64(0)
        cz_rowptr = dev.empty(m + 1, "i4")
        nmem, empty = c_i64(0), c_i64(0)
        try:
            call("frk_polygon_lookup_count", dev.ctx, hp, N, dev.ptr(d_s), N, dev.ptr(d_group), C.byref(multi))
            if multi.value:
                # supports that share BAUs: the general incidence (every polygon a BAU centroid lies in), R/geometryfns.R:1572-1616
                ptoff = dev.empty(N + 1, "i4")
                call("frk_polygon_incidence_count", dev.ctx, hp, N, dev.ptr(d_s), N, dev.ptr(ptoff), C.byref(nmem))
                cz_mem = dev.empty(max(nmem.value, 1), "i4")
                cz_w = dev.empty(max(nmem.value, 1))
                call("frk_polygon_incidence_fill", dev.ctx, hp, N, dev.ptr(d_s), N, dev.ptr(ptoff), nmem.value, vp(None), dev.ptr(cz_rowptr),
                     dev.ptr(cz_mem), dev.ptr(cz_w), C.byref(empty))
        finally:
            _lib.lib.frk_polygons_destroy(hp)
        h2d += data.ptr.nbytes + data.vx.nbytes + data.vy.nbytes
        overlapping = bool(multi.value)
        if not overlapping:
            cz_mem = dev.empty(max(N, 1), "i4")
            cz_w = dev.empty(max(N, 1))
            call("frk_areal_incidence", dev.ctx, N, dev.ptr(d_group), m, vp(None), dev.ptr(cz_rowptr), dev.ptr(cz_mem), dev.ptr(cz_w),
                 C.byref(nmem), C.byref(empty))
        if empty.value:
            raise FRKError(f"{empty.value} observation polygons contain no BAU centroid (supports smaller than a BAU fall back to a "
                           "point lookup in the reference, R/geometryfns.R:1593-1598): use finer BAUs or point data for them")
        rowptr = dev.empty(m + 1, "i4")
        nnz = c_i64(0)
        call("frk_csr_aggregate_count", dev.ctx, m, r, dev.ptr(cz_rowptr), dev.ptr(cz_mem), dev.ptr(S0.rowptr), dev.ptr(S0.col),
             dev.ptr(rowptr), C.byref(nnz))
        col = dev.empty(max(nnz.value, 1), "i4")
        val = dev.empty(max(nnz.value, 1))
        call("frk_csr_aggregate_fill", dev.ctx, m, r, dev.ptr(cz_rowptr), dev.ptr(cz_mem), dev.ptr(cz_w), dev.ptr(S0.rowptr),
             dev.ptr(S0.col), dev.ptr(S0.val), dev.ptr(rowptr), dev.ptr(col), dev.ptr(val))
        S = DeviceCSR(m, r, rowptr, col, val, nnz.value)
        Z = dev.upload(np.ascontiguousarray(data.data[resp], dtype=np.float64))
        d_std = dev.upload(np.ascontiguousarray(data.data["std"], dtype=np.float64))
        h2d += 16 * m
        Ve = dev.empty(max(m, 1))
        call("frk_square", dev.ctx, m, dev.ptr(d_std), dev.ptr(Ve))                     # Ve = diag(std^2), R/SRE.R:468
        obs_bau = dev.zeros(1, "i4")                                                   # (no one-BAU-per-observation map for areal data)
        del d_s
    else:
        # ---- bin the observations into this rank's BAUs ------------------------------------------
        cols = [resp, "std"]
        if average_in_BAU:
            binned = map_data_to_BAUs(data, BAUs, cols, dev, row_range=(lo, hi))
            h2d += data.coords[:, :2].nbytes + 8 * len(data) * len(cols)
            m, ld = binned["m"], binned["ld"]
            obs_bau = binned["obs_bau"][:max(m, 1)]
            Z = binned["means"][0:max(m, 1)]                             # views into the binned table (column-major, ld = n)
            Ve = dev.empty(max(m, 1))
            call("frk_square", dev.ctx, m, dev.ptr(binned["means"], ld), dev.ptr(Ve))      # Ve = diag(std^2), R/SRE.R:468
        else:
            # average_in_BAU = FALSE (R/geometryfns.R:1304-1305): every observation that falls into a BAU is kept, in data order
            if dev.world > 1:
                raise FRKError("average_in_BAU=FALSE on a row-partitioned run is not on the device path (observations of one BAU are coupled)")
            from .geometry import point_BAU_index
            bau_pt = point_BAU_index(data, BAUs, dev)                    # device lookup; -1 = outside every BAU (dropped, :1271-1278)
            keep = np.nonzero(bau_pt >= 0)[0]
            h_obs_bau = np.ascontiguousarray(bau_pt[keep], dtype=np.int32)
            m = len(keep)
            obs_bau = dev.upload(h_obs_bau if m else np.zeros(1, dtype=np.int32))
            Z = dev.upload(np.ascontiguousarray(np.asarray(data.data[resp], dtype=np.float64)[keep]) if m else np.zeros(1))
            d_std = dev.upload(np.ascontiguousarray(np.asarray(data.data["std"], dtype=np.float64)[keep]) if m else np.zeros(1))
            h2d += data.coords[:, :2].nbytes + 16 * len(data)
            Ve = dev.empty(max(m, 1))
            call("frk_square", dev.ctx, m, dev.ptr(d_std), dev.ptr(Ve))

        # ---- C_Z (one 1 per row) -> S = C_Z S0, Vfs = fs[bau], X = model matrix at the observed BAUs ----
        if implicit_basis:      # S = the rows of S0 at the observed BAUs = eval_basis at THEIR centroids: gather the coordinates instead
            dcoord = basis.manifold.dim
            d_sobs = dev.empty(max(m * dcoord, 1))
            call("frk_gather_rows", dev.ctx, m, dcoord, dev.ptr(obs_bau), dev.ptr(d_s), nloc, dev.ptr(d_sobs), m)
            S = ImplicitBasisMatrix(basis, d_sobs, m, m, normalise_basis)
        else:
            rowptr = dev.empty(m + 1, "i4")
            nnz = c_i64(0)
            call("frk_csr_gather_count", dev.ctx, m, dev.ptr(obs_bau), dev.ptr(S0.rowptr), dev.ptr(rowptr), C.byref(nnz))
            col = dev.empty(max(nnz.value, 1), "i4")
            val = dev.empty(max(nnz.value, 1))
            call("frk_csr_gather_fill", dev.ctx, m, dev.ptr(obs_bau), dev.ptr(S0.rowptr), dev.ptr(S0.col), dev.ptr(S0.val),
                 dev.ptr(rowptr), dev.ptr(col), dev.ptr(val))
            S = DeviceCSR(m, r, rowptr, col, val, nnz.value)

    # BAU-level fields for this rank: fs and covariates
    fs_host = np.ascontiguousarray(BAUs.fields["fs"][lo:hi], dtype=np.float64)
    fs_BAU = dev.upload(fs_host)
    h2d += fs_host.nbytes
    p = (1 if intercept else 0) + len(covs)
    X_BAU = dev.empty(max(nloc * p, 1))                           # model matrix at this rank's BAUs, column-major (nloc x p)
    k = 0
    if intercept:                                                 # the intercept column is generated on the device
        call("frk_fill", dev.ctx, nloc, 1.0, dev.ptr(X_BAU))
        k = 1
    for c in covs:
        colv = np.ascontiguousarray(BAUs.fields[c][lo:hi], dtype=np.float64)
        if nloc:
            call("frk_h2d", dev.ctx, dev.ptr(X_BAU, k * nloc), vp(colv.ctypes.data), colv.nbytes)
            call("frk_sync", dev.ctx)
        h2d += colv.nbytes
        k += 1
    Vfs = dev.empty(max(m, 1))                                    # Vfs = diag(fs[bau]), R/SRE.R:498
    X = dev.empty(max(m * p, 1))                                  # model matrix at the observed BAUs, :464-465
    if areal:      # diag(C_Z diag(fs) C_Z') and the BAU covariates averaged over each footprint (R/geometryfns.R:1352-1357)
        call("frk_areal_aggregate", dev.ctx, m, dev.ptr(cz_rowptr), dev.ptr(cz_mem), dev.ptr(cz_w), 2, 1, dev.ptr(fs_BAU), nloc, dev.ptr(Vfs))
        call("frk_areal_aggregate", dev.ctx, m, dev.ptr(cz_rowptr), dev.ptr(cz_mem), dev.ptr(cz_w), 1, p, dev.ptr(X_BAU), nloc, dev.ptr(X))
    else:
        call("frk_gather_rows", dev.ctx, m, 1, dev.ptr(obs_bau), dev.ptr(fs_BAU), nloc, dev.ptr(Vfs), m)
        call("frk_gather_rows", dev.ctx, m, p, dev.ptr(obs_bau), dev.ptr(X_BAU), nloc, dev.ptr(X), m)

    mdl = SREModel(dev, f, basis, BAUs, K_type, N, lo, hi, m, S0, S, obs_bau, Z, X, X_BAU, Ve, Vfs, fs_BAU, p)
    mdl._keep = binned                                            # Z / obs_bau are views into these buffers
    mdl.areal = areal
    mdl.coupled = False
    Vcsr = None
    if areal and overlapping:
        # Vfs = tcrossprod(Cmat %*% Diagonal(sqrt(fs))) (R/SRE.R:498) with its off-diagonal entries: host sparse algebra on the
        # incidence matrix, as SRE() does in R; the device takes it from there (frk_model_set_vfs_csr)
        import scipy.sparse as sp
        Cm = sp.csr_matrix((dev.download(cz_w, nmem.value), dev.download(cz_mem, nmem.value), dev.download(cz_rowptr, m + 1)), shape=(m, N))
        half = sp.csr_matrix(Cm @ sp.diags(np.sqrt(fs_host)))
        Vcsr = sp.csr_matrix(half @ half.T)
    elif not areal and not average_in_BAU and m:
        import scipy.sparse as sp
        Cm = sp.csr_matrix((np.ones(m), h_obs_bau.astype(np.int64), np.arange(m + 1)), shape=(m, nloc))
        half = sp.csr_matrix(Cm @ sp.diags(np.sqrt(fs_host)))
        Vcsr = sp.csr_matrix(half @ half.T)
    if Vcsr is not None:
        Vcsr.sort_indices()
        vrp = np.ascontiguousarray(Vcsr.indptr, dtype=np.int32)
        vcl = np.ascontiguousarray(Vcsr.indices, dtype=np.int32)
        vvl = np.ascontiguousarray(Vcsr.data, dtype=np.float64)
        call("frk_model_set_vfs_csr", mdl.handle, vp(vrp.ctypes.data), vp(vcl.ctypes.data), vp(vvl.ctypes.data))
        mdl.coupled = _lib.lib.frk_model_coupled_groups(mdl.handle) > 0
        h2d += vrp.nbytes + vcl.nbytes + vvl.nbytes
    if areal:
        mdl.C_Z = (cz_rowptr, cz_mem, cz_w, int(nmem.value))      # Cmat (CSR over the observations, rows normalised)
    elif mdl.coupled:                                             # average_in_BAU = FALSE: one 1 per row, several rows per BAU
        mdl.C_Z = (dev.upload(np.arange(m + 1, dtype=np.int32)), obs_bau, dev.upload(np.ones(m)), m)
    mdl.h2d_bytes = h2d
    if K_type == "block-exponential" and isinstance(basis, TensorP_Basis):     # BuildD(basis) = list(Basis1 = ..., Basis2 = ...)
        D = BuildD(basis)
        mdl.D_basis = [np.asfortranarray(d, dtype=np.float64) for d in D["Basis1"]]
        Dt = np.asfortranarray(D["Basis2"][0], dtype=np.float64)
        res_s = np.ascontiguousarray(basis.Basis1.res, dtype=np.int32)
        ptrs = (vp * len(mdl.D_basis))(*[vp(d.ctypes.data) for d in mdl.D_basis])
        call("frk_model_set_blocks_tensor", mdl.handle, len(mdl.D_basis), basis.Basis1.n, basis.Basis2.n, vp(res_s.ctypes.data), ptrs,
             vp(Dt.ctypes.data))
        mdl.h2d_bytes += sum(d.nbytes for d in mdl.D_basis) + Dt.nbytes
        mdl._tensor = True
    elif K_type == "block-exponential":                           # BuildD(basis), R/SRE.R:424 -- on the device, from the basis handle
        res = np.ascontiguousarray(basis.res, dtype=np.int32)
        mdl.D_basis = [None] * basis.nres                         # (host copies on demand: frk_b200.BuildD(basis))
        call("frk_model_set_blocks_basis", mdl.handle, basis.nres, vp(res.ctypes.data), basis.handle(dev))
    call("frk_model_init", mdl.handle)                            # .EM_initialise, R/SRE.R:656,704-731
    mdl.info_fit["time_SRE"] = time.time() - t0
    return mdl


def loglik(mdl: SREModel) -> float:
    """logLik(SRE) at the current parameters (R/SREutils.R:22-32 -> .loglik.ind); leaves the slots untouched."""
    out = C.c_double(0.0)
    call("frk_model_loglik", mdl.handle, C.byref(out))
    return out.value


# ---------------------------------------------------------------------------------------
# SRE.fit
# ---------------------------------------------------------------------------------------

def SRE_fit(mdl: SREModel, n_EM: int = 100, tol: float = 0.01, method: str = "EM", lambda_=0.0,
            known_sigma2fs=None, print_lik: bool = False, lanes: Optional[int] = None,
            distribute_nm: bool = False) -> SREModel:
    """SRE.fit(object, n_EM, tol, method = "EM", lambda, print_lik, known_sigma2fs), R/SREfit.R:3-160: one
    ``frk_model_em_fit`` call.  Calling it again continues from the current slots, as in R.

    ``lanes`` (not in the reference): how many of the block-exponential Nelder-Mead's candidate points are factorised side by side
    on this GPU (``frk_model_set_lanes``); the iterates and results are those of the sequential algorithm for any value.  Two
    candidates per round is the measured optimum on B200 (more cost more than the rounds they save): the default is 2.
    ``distribute_nm`` (row-partitioned runs): spread the candidates over the ranks instead of evaluating them, replicated, on every
    rank's own lanes; measured slower on NVLink-connected B200s (the exchange costs more than the contention it avoids)."""
    if method != "EM":
        raise FRKError("only method = \"EM\" (Gaussian data, identity link) is on the device path")
    if lanes is None:
        lanes = 1 if (distribute_nm and mdl.dev.world > 1) else 2
    call("frk_model_set_lanes", mdl.handle, int(lanes))
    call("frk_model_set_nm_distribute", mdl.handle, 1 if distribute_nm else 0)
    t0 = time.time()
    lam = np.atleast_1d(np.asarray(lambda_, dtype=np.float64)).copy()
    res = np.ascontiguousarray(mdl.basis.res, dtype=np.int32)
    llk = np.zeros(max(n_EM, 1))
    nit = C.c_int(0)
    call("frk_model_em_fit", mdl.handle, int(n_EM), float(tol), 1 if mdl.K_type == "block-exponential" else 0, len(lam),
         vp(lam.ctypes.data), vp(res.ctypes.data), 0 if known_sigma2fs is None else 1,
         0.0 if known_sigma2fs is None else float(known_sigma2fs), vp(llk.ctypes.data), C.byref(nit))
    nit = nit.value
    if print_lik:
        for i in range(nit):
            print(f"EM iteration {i + 1}: log-likelihood {llk[i]:.10g}")
    s2 = mdl.sigma2fshat
    mdl.info_fit.update(method="EM", time=time.time() - t0, num_iterations=nit, converged=int(nit != n_EM),
                        llk=llk[:nit].copy(), sigma2fshat_equal_0=int(s2 == 0))
    if mdl.K_type == "block-exponential":
        nres = len(mdl.D_basis)
        ntau = nres * (2 if getattr(mdl, "_tensor", False) else 1)
        mdl.info_fit["tau"] = mdl.get("tau")[:ntau].tolist()
        mdl.info_fit["omega"] = mdl.get("omega")[:nres].tolist()
        mdl.info_fit["nm_stats"] = dict(zip(("requests", "rounds", "evaluated", "re_evaluated"), mdl.get("nm_stats").tolist()))   # last M-step
        mdl.info_fit["nm_stats"]["prefetched"] = float(mdl.get("nm_prefetch_hits")[0])    # evaluations whose factorisation ran beside the E-step
        mdl.info_fit["nm_stats"]["reused"] = float(mdl.get("nm_reuse_hits")[0])           # evaluations served by a saved inverse (no factorisation)
    return mdl


# ---------------------------------------------------------------------------------------
# predict
# ---------------------------------------------------------------------------------------

def _predict_polygons(mdl: SREModel, newdata, obs_fs: bool):
    """predict(newdata = SpatialPolygons), R/SREpredict.R:119-128,365-419: C_P over the BAU centroids, then frk_model_predict_polygons."""
    dev = mdl.dev
    if dev.world > 1:
        raise FRKError("prediction over polygons on a row-partitioned run: polygons may straddle ranks -- not on the device path yet")
    if isinstance(mdl.basis, TensorP_Basis) or isinstance(mdl.BAUs, STBAUs):
        raise FRKError("prediction over polygons in space-time is not on the device path yet")
    if (getattr(mdl, "areal", False) or getattr(mdl, "coupled", False)) and not obs_fs and mdl.sigma2fshat > 0:
        raise FRKError("prediction over polygons with obs_fs = FALSE for areal DATA needs the covariances between the fine-scale terms "
                       "of a support (R/SREpredict.R:392-404): not on the device path yet; use obs_fs = TRUE")
    N = mdl.N
    if isinstance(mdl.BAUs, GridBAUs):
        d_xg, d_yg = dev.upload(mdl.BAUs.xgrid), dev.upload(mdl.BAUs.ygrid)
        d_cen = dev.empty(2 * N)
        call("frk_grid_centroids", dev.ctx, 0, N, mdl.BAUs.nx, dev.ptr(d_xg), dev.ptr(d_yg), dev.ptr(d_cen))
    else:
        d_cen = dev.upload(mdl.BAUs.centroids()[:, :2])
    mP = len(newdata)
    hp = vp()
    call("frk_polygons_create", dev.ctx, mP, vp(newdata.ptr.ctypes.data), vp(newdata.vx.ctypes.data), vp(newdata.vy.ctypes.data), C.byref(hp))
    try:
        ptoff = dev.empty(N + 1, "i4")
        npairs = c_i64(0)
        call("frk_polygon_incidence_count", dev.ctx, hp, N, dev.ptr(d_cen), N, dev.ptr(ptoff), C.byref(npairs))
        cp_rowptr = dev.empty(mP + 1, "i4")
        cp_mem = dev.empty(max(npairs.value, 1), "i4")
        cp_w = dev.empty(max(npairs.value, 1))
        empty = c_i64(0)
        call("frk_polygon_incidence_fill", dev.ctx, hp, N, dev.ptr(d_cen), N, dev.ptr(ptoff), npairs.value, vp(None), dev.ptr(cp_rowptr),
             dev.ptr(cp_mem), dev.ptr(cp_w), C.byref(empty))
    finally:
        _lib.lib.frk_polygons_destroy(hp)
    if empty.value or npairs.value == 0:
        raise FRKError(f"{empty.value} prediction polygons contain no BAU centroid")
    mu, var, sd = (dev.empty(mP) for _ in range(3))
    call("frk_model_predict_polygons", mdl.handle, N, dev.ptr(mdl.S0.rowptr), dev.ptr(mdl.S0.col), dev.ptr(mdl.S0.val), mdl.S0.buckets(dev),
         dev.ptr(mdl.X_BAU), dev.ptr(mdl.fs_BAU), dev.ptr(mdl.obs_bau), 1 if obs_fs else 0, mP, dev.ptr(cp_rowptr), dev.ptr(cp_mem),
         dev.ptr(cp_w), npairs.value, dev.ptr(mu), dev.ptr(var), dev.ptr(sd))
    mdl.d2h_bytes += 3 * 8 * mP
    return dict(mu=dev.download(mu, mP), var=dev.download(var, mP), sd=dev.download(sd, mP),
                C_P=(cp_rowptr, cp_mem, cp_w, int(npairs.value)))


def predict(mdl: SREModel, obs_fs: bool = False, to_host: bool = True, newdata=None):
    """predict(SRE, obs_fs) at the BAUs (newdata = NULL, covariances = FALSE): .SRE.predict_EM,
    R/SREpredict.R:243-375 with .batch_compute_var (R/SREutils.R:245-305) -- one ``frk_model_predict`` call.

    obs_fs = TRUE or sigma2fs == 0: the stored posterior (mu_eta, S_eta of the last E-step).
    obs_fs = FALSE and sigma2fs > 0: the joint (eta, xi) solve at the final parameters; for point data its
    diagonal reduces to one more E-step plus the same sparse quadratic form.
    Returns dict(mu, var, sd) for this rank's BAU rows [lo, hi) (host arrays, or device tensors)."""
    if newdata is not None:
        if not isinstance(newdata, SpatialPolygonsDataFrame):
            raise FRKError("newdata must be a SpatialPolygonsDataFrame (prediction polygons); BAU-level prediction is newdata = None")
        return _predict_polygons(mdl, newdata, obs_fs)
    dev = mdl.dev
    nloc = mdl.hi - mdl.lo
    mu, var, sd = (dev.empty(max(nloc, 1)) for _ in range(3))
    if (getattr(mdl, "areal", False) or getattr(mdl, "coupled", False)) and not obs_fs and mdl.sigma2fshat > 0:
        cz_rowptr, cz_mem, cz_w, _ = mdl.C_Z
        call("frk_model_predict_areal", mdl.handle, nloc, dev.ptr(mdl.S0.rowptr), dev.ptr(mdl.S0.col), dev.ptr(mdl.S0.val),
             mdl.S0.buckets(dev), dev.ptr(mdl.X_BAU), dev.ptr(mdl.fs_BAU), dev.ptr(cz_rowptr), dev.ptr(cz_mem), dev.ptr(cz_w),
             dev.ptr(mu), dev.ptr(var), dev.ptr(sd))
        if not to_host:
            return dict(mu=mu, var=var, sd=sd)
        out = dict(mu=dev.download(mu, nloc), var=dev.download(var, nloc), sd=dev.download(sd, nloc))
        mdl.d2h_bytes += 3 * 8 * nloc
        return out
    call("frk_model_predict", mdl.handle, nloc, dev.ptr(mdl.S0.rowptr), dev.ptr(mdl.S0.col), dev.ptr(mdl.S0.val),
         mdl.S0.buckets(dev), dev.ptr(mdl.X_BAU), dev.ptr(mdl.fs_BAU), dev.ptr(mdl.obs_bau), 1 if obs_fs else 0,
         dev.ptr(mu), dev.ptr(var), dev.ptr(sd))
    if not to_host:
        return dict(mu=mu, var=var, sd=sd)
    out = dict(mu=dev.download(mu, nloc), var=dev.download(var, nloc), sd=dev.download(sd, nloc))
    mdl.d2h_bytes += 3 * 8 * nloc
    return out

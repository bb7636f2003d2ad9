"""Basis objects and eval_basis -- host mirror of R/basisfns.R for the hot path.

``local_basis`` / ``auto_basis`` / ``TensorP`` keep the reference's names, argument meaning and
error messages; placement (r-sized) stays on the host as it does in R, evaluation is the
libfrkb200 cell-list kernel (``eval_basis``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import scipy.sparse as sp

from ._lib import FRKError, call, c_i64, lib, vp
from .geometry import Manifold, SpatialPointsDataFrame, host_distance, plane, real_line
from .runtime import Device, get_device

BASIS_TYPES = ("bisquare", "Gaussian", "exp", "Matern32")
_TYPE_CODE = {"bisquare": 0, "Gaussian": 1, "exp": 2, "Matern32": 3}


@dataclass
class Basis:
    """``Basis`` S4 class (R/AllClass.R:78): manifold, n, pars (loc, scale), df (loc, scale, res)."""
    manifold: Manifold
    loc: np.ndarray
    scale: np.ndarray
    res: np.ndarray
    type: str = "bisquare"
    regular: int = 0
    _handle: Optional[vp] = field(default=None, repr=False, compare=False)
    _dev: Optional[Device] = field(default=None, repr=False, compare=False)

    @property
    def n(self) -> int:
        return self.loc.shape[0]

    @property
    def nres(self) -> int:
        return int(self.res.max())

    def handle(self, dev: Device):
        """Device-resident basis (centroid table + per-resolution cell lists), built once."""
        if self._handle is None or self._dev is not dev:
            loc = np.asfortranarray(self.loc, dtype=np.float64)
            scale = np.ascontiguousarray(self.scale, dtype=np.float64)
            res = np.ascontiguousarray(self.res, dtype=np.int32)
            h = vp()
            call("frk_basis_create", dev.ctx, self.manifold.code, float(self.manifold.radius), _TYPE_CODE[self.type],
                 int(self.loc.shape[1]), int(self.n), vp(loc.ctypes.data), vp(scale.ctypes.data), vp(res.ctypes.data),
                 C.byref(h))
            self._handle, self._dev = h, dev
        return self._handle

    def __del__(self):
        try:
            if self._handle is not None:
                lib.frk_basis_destroy(self._handle)
        except Exception:
            pass


@dataclass
class TensorP_Basis:
    """TensorP(Basis1, Basis2), R/basisfns.R:659-704: space x time, space index fastest."""
    Basis1: Basis
    Basis2: Basis

    @property
    def n(self) -> int:
        return self.Basis1.n * self.Basis2.n

    @property
    def res(self) -> np.ndarray:
        return np.tile(self.Basis1.res, self.Basis2.n)

    @property
    def nres(self) -> int:
        return self.Basis1.nres


def nbasis(b) -> int:
    return b.n


def local_basis(manifold: Manifold, loc, scale, type: str = "bisquare", res=1, regular=0) -> Basis:
    """local_basis(), R/basisfns.R:99-146 (same checks, same messages)."""
    loc = np.asarray(loc, dtype=np.float64)
    if loc.ndim != 2:
        raise ValueError("loc needs to be a matrix")
    if manifold.dim != loc.shape[1]:
        raise ValueError("number of columns in loc needs to be the same as the number of manifold dimensions")
    scale = np.atleast_1d(np.asarray(scale, dtype=np.float64))
    if scale.shape[0] != loc.shape[0]:
        raise ValueError("need to have as many scale parameters as centroids")
    if type not in BASIS_TYPES:
        raise ValueError(f"'arg' should be one of {BASIS_TYPES}")
    if np.any(~(scale > 0)):
        raise ValueError("The scale parameter needs to be greater than zero")
    res = np.broadcast_to(np.asarray(res, dtype=np.int32), (loc.shape[0],)).copy()
    return Basis(manifold, loc.copy(), scale.copy(), res, type, int(regular))


def TensorP(Basis1: Basis, Basis2: Basis) -> TensorP_Basis:
    if Basis2.nres != 1:
        raise ValueError("Only one basis can be multiresolution (Basis 1)")     # R/basisfns.R:667
    return TensorP_Basis(Basis1, Basis2)


# R/basisfns.R:268-283: seconds per temporal unit (months = 7 days x 52/12, years = 12 of those)
_TUNIT_SECS = {"secs": 1.0, "mins": 60.0, "hours": 60.0 * 60, "days": 60.0 * 60 * 24, "weeks": 60.0 * 60 * 24 * 7,
               "months": 60.0 * 60 * 24 * 7 * (52 / 12), "years": 60.0 * 60 * 24 * 7 * (52 / 12) * 12}


def _seq_len(a, b, n):
    if n == 1:
        return np.array([a], dtype=np.float64)
    return a + np.arange(n, dtype=np.float64) * ((b - a) / (n - 1))


def _nn_scale_device(dev: Device, manifold: Manifold, this: np.ndarray, zero_is_inf: bool) -> np.ndarray:
    """distance to the nearest other centroid on the device (frk_basis_nn_dist), R/basisfns.R:485-511."""
    k = this.shape[0]
    tmp = Basis(manifold, this, np.ones(k), np.ones(k, dtype=np.int32), "Gaussian", 0)    # non-compact type: no cell lists are built
    out = dev.empty(k)
    call("frk_basis_nn_dist", dev.ctx, tmp.handle(dev), 1 if zero_is_inf else 0, dev.ptr(out))
    return dev.download(out, k)


def _colsums_device(dev: Device, basis, coords: np.ndarray) -> np.ndarray:
    """colSums(eval_basis(basis, coords)) on the device, R/basisfns.R:543."""
    S = eval_basis_device(basis, coords, dev)
    out = dev.empty(basis.n)
    call("frk_csr_colsums", dev.ctx, S.n, S.ncol, dev.ptr(S.rowptr), dev.ptr(S.col), dev.ptr(S.val), dev.ptr(out))
    return dev.download(out, basis.n)


def auto_basis(manifold: Manifold, data, regular: int = 1, nres: int = 3, prune: float = 0, type: str = "bisquare",
               isea3h_lo: int = 2, scale_aperture: Optional[float] = None, isea3h: Optional[dict] = None,
               subsamp: Optional[int] = None, dev: Optional[Device] = None, tunit: Optional[str] = None,
               end_time=None):
    """auto_basis() regular placement, R/basisfns.R:218-585.

    The placement uses all data coordinates (the reference subsamples 10,000 points with R's RNG when n > subsamp, which no
    other implementation can reproduce -- pass the coordinates you want the ranges taken from).  ``prune > 0`` (deprecated in
    the reference, still used by its vignettes): basis functions with colSums(eval_basis(., data)) < prune are removed and the
    scales re-derived (:536-553) -- eval_basis and the column sums run on the device, as does the nearest-centroid search for
    resolutions with >= 5000 centroids (:497-509).  Sphere placement needs the ISEA3H centroids (``isea3h`` = {resolution:
    (n x 2) lon/lat}).
    """
    if manifold.kind in ("STplane", "STsphere"):
        # R/basisfns.R:243-289: TensorP(auto_basis(plane(), <spatial part>), ten Gaussian functions on the time axis).  The
        # spatial call is made with manifold = plane() for BOTH space-time manifolds in the reference (:254); followed here.
        if tunit is None:
            raise ValueError("Please specify tunit (the temporal unit) if you want to construct a spatio-temporal basis. "
                             "Note that this should be the same as the BAUs.")
        if tunit not in _TUNIT_SECS:
            raise ValueError("tunit should be a one of: secs, mins, hours, days, weeeks, months, years")
        coords = data.coords if isinstance(data, SpatialPointsDataFrame) else np.asarray(data, dtype=np.float64)
        if end_time is None:
            if coords.shape[1] < 3:
                raise ValueError("space-time auto_basis needs the end times (end_time=, seconds) or a third coordinate column")
            end_time = coords[:, 2]
        end_time = np.asarray(end_time, dtype=np.float64)
        B_spatial = auto_basis(plane(), coords[:, :2], regular=regular, nres=nres, prune=prune, type=type, isea3h_lo=isea3h_lo,
                               scale_aperture=1.25 if scale_aperture is None else scale_aperture,   # the default is evaluated on the ST manifold (:228)
                               isea3h=isea3h, subsamp=subsamp, dev=dev)
        end_loc = float(np.ceil((end_time.max() - end_time.min()) / _TUNIT_SECS[tunit]) + 1.0)      # "+ 1 as a buffer"
        B_temporal = local_basis(real_line(), _seq_len(0.0, end_loc, 10)[:, None], np.full(10, end_loc / 10.0), "Gaussian",
                                 regular=1)
        return TensorP(B_spatial, B_temporal)
    if prune < 0:
        raise ValueError("prune needs to be greater than zero")
    if not regular and manifold.kind == "plane":
        raise FRKError("irregular placement needs fmesher (host geometry, out of scope); use regular >= 1")
    if type not in BASIS_TYPES:
        raise ValueError(f"'arg' should be one of {BASIS_TYPES}")
    coords = data.coords if isinstance(data, SpatialPointsDataFrame) else np.asarray(data, dtype=np.float64)
    if coords.ndim == 1:
        coords = coords[:, None]
    if scale_aperture is None:
        scale_aperture = 1.0 if manifold.kind == "sphere" else 1.25
    mult = 1.5 if type == "bisquare" else 0.7
    xr = (coords[:, 0].min(), coords[:, 0].max())
    yr = (coords[:, 1].min(), coords[:, 1].max()) if coords.shape[1] > 1 else (0.0, 0.0)
    if manifold.kind == "plane":
        asp = (yr[1] - yr[0]) / (xr[1] - xr[0])
        if asp < 1:
            ny = regular
            nx = ny / asp
        else:
            nx = regular
            ny = nx * asp
    locs, scales, ress = [], [], []
    for i in range(1, nres + 1):
        xg = yg = None
        if manifold.kind == "plane":
            xg = _seq_len(xr[0], xr[1], int(np.rint(nx * 3 ** i)))
            yg = _seq_len(yr[0], yr[1], int(np.rint(ny * 3 ** i)))
            gx, gy = np.meshgrid(xg, yg, indexing="xy")
            this = np.column_stack([gx.ravel(), gy.ravel()])
        elif manifold.kind == "real_line":
            this = _seq_len(xr[0], xr[1], i * regular)[:, None]
        else:
            if isea3h is None:
                raise FRKError("auto_basis on the sphere needs the ISEA3H centroids (isea3h=...)")
            this = np.asarray(isea3h[i + isea3h_lo - 1], dtype=np.float64)
        base = np.zeros(0)
        for j in (1, 2):                                   # twice when pruning: the scales are re-derived from the survivors
            k = this.shape[0]
            if k == 0:
                base = np.zeros(0)
                break
            if k == 1:
                base = np.array([max(xr[1] - xr[0], yr[1] - yr[0]) / 2.0])
            elif k < 5000:
                D = host_distance(manifold, this, this)
                np.fill_diagonal(D, np.inf)
                base = D.min(axis=1) * scale_aperture
            elif not prune and regular and manifold.kind == "plane":
                base = np.full(k, min(abs(xg[1] - xg[0]), abs(yg[1] - yg[0]))) * scale_aperture
            else:                                          # :497-509, batched in the reference; one kernel here
                base = _nn_scale_device(dev or get_device(), manifold, this, zero_is_inf=True) * scale_aperture
            if not prune:
                break
            if j == 1:
                cand = Basis(manifold, this, mult * base, np.full(k, i, dtype=np.int32), type, int(regular))
                keep = ~(_colsums_device(dev or get_device(), cand, coords[:, :manifold.dim]) < prune)
                if not keep.any():
                    import warnings
                    warnings.warn("prune is too large -- all functions at a resolution removed. "
                                  "Consider also removing number of resolutions.")
                if keep.all():
                    break                                  # nothing removed: the second pass would reproduce the same scales
                this = this[keep]
        if this.shape[0] == 0:
            continue
        locs.append(this)
        scales.append(mult * base)
        ress.append(np.full(this.shape[0], i, dtype=np.int32))
    if not locs:
        raise FRKError("auto_basis: prune removed every basis function")
    return Basis(manifold, np.vstack(locs), np.concatenate(scales), np.concatenate(ress), type, int(regular))


def BuildD(basis):
    """BuildD(), R/basisfns.R:1131-1153: per-resolution centroid distance matrices (r_i x r_i,
    one-off, host -- as in the reference)."""
    if isinstance(basis, TensorP_Basis):
        return {"Basis1": BuildD(basis.Basis1), "Basis2": BuildD(basis.Basis2)}
    return [host_distance(basis.manifold, basis.loc[basis.res == i], basis.loc[basis.res == i])
            for i in range(1, basis.nres + 1)]


# ---------------------------------------------------------------------------------------
# eval_basis
# ---------------------------------------------------------------------------------------

class DeviceCSR:
    """CSR matrix resident in HBM: rowptr (n+1 int32), col (nnz int32), val (nnz fp64)."""

    def __init__(self, n, ncol, rowptr, col, val, nnz):
        self.n, self.ncol, self.rowptr, self.col, self.val, self.nnz = int(n), int(ncol), rowptr, col, val, int(nnz)

    _buckets = None      # frk_rowbuckets handle (row buckets + ELL tiles), built on first use

    def buckets(self, dev: Device):
        if self._buckets is None:
            h = vp()
            call("frk_rowbuckets_create", dev.ctx, self.n, self.ncol, dev.ptr(self.rowptr), dev.ptr(self.col), dev.ptr(self.val),
                 C.byref(h))
            self._buckets = h
        return self._buckets

    def __del__(self):
        try:
            if self._buckets is not None:
                lib.frk_rowbuckets_destroy(self._buckets)
                self._buckets = None
        except Exception:
            pass

    def to_scipy(self, dev: Device) -> sp.csr_matrix:
        rp = dev.download(self.rowptr, self.n + 1)
        c = dev.download(self.col, self.nnz)
        v = dev.download(self.val, self.nnz)
        return sp.csr_matrix((v, c, rp), shape=(self.n, self.ncol))


class ImplicitBasisMatrix:
    """An n x r basis-function matrix that is never stored: row i = eval_basis(basis, coords[i, ]) (normalised if asked), evaluated in
    chunks inside every pass over it (frk_rowbuckets_create_implicit / frk_model_create_implicit).  For bases without compact
    support at scale -- Gaussian / exponential / Matern 3/2 give DENSE rows: 1e7 x 3608 doubles = 290 GB.  Same duck type as
    DeviceCSR where the device path needs it (n, ncol, nnz, NULL CSR pointers, buckets())."""
    rowptr = col = val = None

    def __init__(self, basis: Basis, d_coords, n: int, ld: int, normalise: bool):
        self.basis, self.coords, self.n, self.ld, self.normalise = basis, d_coords, int(n), int(ld), bool(normalise)
        self.ncol = basis.n
        self.nnz = self.n * self.ncol
        self._buckets = None

    def buckets(self, dev: Device):
        if self._buckets is None:
            h = vp()
            call("frk_rowbuckets_create_implicit", dev.ctx, self.n, self.basis.handle(dev), dev.ptr(self.coords), self.ld,
                 1 if self.normalise else 0, C.byref(h))
            self._buckets = h
        return self._buckets

    def __del__(self):
        try:
            if self._buckets is not None:
                lib.frk_rowbuckets_destroy(self._buckets)
                self._buckets = None
        except Exception:
            pass


def _eval_single(dev: Device, basis: Basis, d_s, n: int, normalise: bool) -> DeviceCSR:
    h = basis.handle(dev)
    rowptr = dev.empty(n + 1, "i4")
    nnz = c_i64(0)
    call("frk_eval_basis_count", dev.ctx, h, n, dev.ptr(d_s), n, dev.ptr(rowptr), C.byref(nnz))
    col = dev.empty(max(nnz.value, 1), "i4")
    val = dev.empty(max(nnz.value, 1), "f8")
    call("frk_eval_basis_fill", dev.ctx, h, n, dev.ptr(d_s), n, dev.ptr(rowptr), dev.ptr(col), dev.ptr(val),
         1 if normalise else 0)
    return DeviceCSR(n, basis.n, rowptr, col, val, nnz.value)


def eval_basis_device(basis, s, dev: Optional[Device] = None, normalise: bool = False) -> DeviceCSR:
    """eval_basis with the result left in HBM (what SRE() uses).  ``s``: host (n x d) array or a
    device tensor holding the column-major coordinates."""
    dev = dev or get_device()
    if isinstance(s, np.ndarray):
        s = np.asarray(s, dtype=np.float64)
        if s.ndim == 1:
            s = s[:, None]
        n, d = s.shape
        d_s = dev.upload(s)
    else:
        d_s, n, d = s
    if isinstance(basis, TensorP_Basis):
        n1 = basis.Basis1.manifold.dim
        if d != n1 + basis.Basis2.manifold.dim:
            raise ValueError("incorrect number of coordinate columns for the tensor basis")
        S1 = _eval_single(dev, basis.Basis1, d_s[: n * n1], n, False)
        S2 = _eval_single(dev, basis.Basis2, d_s[n * n1:], n, False)
        rowptr = dev.empty(n + 1, "i4")
        nnz = c_i64(0)
        call("frk_csr_tensor_count", dev.ctx, n, dev.ptr(S1.rowptr), dev.ptr(S2.rowptr), dev.ptr(rowptr), C.byref(nnz))
        col = dev.empty(max(nnz.value, 1), "i4")
        val = dev.empty(max(nnz.value, 1), "f8")
        call("frk_csr_tensor_fill", dev.ctx, n, basis.Basis1.n, dev.ptr(S1.rowptr), dev.ptr(S1.col), dev.ptr(S1.val),
             dev.ptr(S2.rowptr), dev.ptr(S2.col), dev.ptr(S2.val), dev.ptr(rowptr), dev.ptr(col), dev.ptr(val))
        out = DeviceCSR(n, basis.n, rowptr, col, val, nnz.value)
        if normalise:
            call("frk_csr_row_normalise", dev.ctx, n, dev.ptr(rowptr), dev.ptr(val))
        return out
    if d < basis.manifold.dim:
        raise ValueError("ncol(s) needs to be at least the number of manifold dimensions")
    return _eval_single(dev, basis, d_s, n, normalise)


def eval_basis(basis, s, dev: Optional[Device] = None) -> sp.csr_matrix:
    """eval_basis(basis, s), R/basisfns.R:709-857 (matrix / SpatialPointsDataFrame methods).

    Returns the N x r sparse matrix on the host (the reference returns a ``dgCMatrix``; this is
    the same matrix in CSR with ascending column indices).  Plain ``Basis`` goes through the
    host-buffer C-ABI entry point ``frk_eval_basis_host`` -- the call the R shim binds."""
    dev = dev or get_device()
    if isinstance(s, SpatialPointsDataFrame):
        s = s.coords
    s = np.asarray(s, dtype=np.float64)
    if s.ndim == 1:
        s = s[:, None]
    if isinstance(basis, TensorP_Basis):
        return eval_basis_device(basis, s, dev).to_scipy(dev)
    n = s.shape[0]
    d = basis.manifold.dim
    if s.shape[1] < d:
        raise ValueError("ncol(s) needs to be at least the number of manifold dimensions")
    sf = np.asfortranarray(s[:, :d])
    h = basis.handle(dev)
    rowptr = np.empty(n + 1, dtype=np.int32)
    nnz = c_i64(0)
    call("frk_eval_basis_host", dev.ctx, h, n, vp(sf.ctypes.data), 0, vp(rowptr.ctypes.data), vp(None), vp(None), C.byref(nnz))
    col = np.empty(nnz.value, dtype=np.int32)
    val = np.empty(nnz.value, dtype=np.float64)
    call("frk_eval_basis_host", dev.ctx, h, n, vp(sf.ctypes.data), 0, vp(rowptr.ctypes.data), vp(col.ctypes.data),
         vp(val.ctypes.data), C.byref(nnz))
    return sp.csr_matrix((val, col, rowptr), shape=(n, basis.n))

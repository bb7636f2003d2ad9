"""Manifolds, BAU containers and the point->BAU mapping -- the host-side mirror of the reference's
R/geometryfns.R interface for the hot path (same names, argument meaning and error behaviour).

Only r-sized or O(grid-axis) work happens here in NumPy (as it does in R in the reference); every
N- or m-sized computation is a libfrkb200 kernel.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from ._lib import FRKError, call, c_i64, vp
from .runtime import Device, get_device

FRK_REAL_LINE, FRK_PLANE, FRK_SPHERE, FRK_STPLANE, FRK_STSPHERE = 1, 2, 3, 4, 5


@dataclass(frozen=True)
class Manifold:
    """real_line() / plane() / sphere(radius) / STplane() / STsphere(radius)  (R/geometryfns.R:19-189).

    Only the built-in measures exist on the device; a custom ``measure@dist`` has no CUDA
    kernel and is refused (north_star: no CPU fallback)."""
    kind: str
    radius: float = 6371.0

    @property
    def dim(self) -> int:
        return {"real_line": 1, "plane": 2, "sphere": 2, "STplane": 3, "STsphere": 3}[self.kind]

    @property
    def code(self) -> int:
        return {"real_line": FRK_REAL_LINE, "plane": FRK_PLANE, "sphere": FRK_SPHERE,
                "STplane": FRK_STPLANE, "STsphere": FRK_STSPHERE}[self.kind]


def real_line() -> Manifold:
    return Manifold("real_line")


def plane() -> Manifold:
    return Manifold("plane")


def sphere(radius: float = 6371.0) -> Manifold:
    """sphere(radius=6371), R/geometryfns.R:97-103."""
    if not radius > 0:
        raise ValueError("radius needs to be positive")
    return Manifold("sphere", float(radius))


def STplane() -> Manifold:
    """STplane(): plane x time with Euclid_dist(dim=3L), R/geometryfns.R:60-80."""
    return Manifold("STplane")


def STsphere(radius: float = 6371.0) -> Manifold:
    """STsphere(radius=6371): measure gc_dist_time, R/geometryfns.R:120-134,181-189."""
    if not radius > 0:
        raise ValueError("radius needs to be positive")
    return Manifold("STsphere", float(radius))


def host_distance(m: Manifold, x1: np.ndarray, x2: np.ndarray) -> np.ndarray:
    """distance(manifold, x1, x2) for r-sized inputs (basis placement, BuildD); the reference keeps
    these on the host as well (R/basisfns.R:485-490,1131-1153).  Same operation order as
    distR / dist_sphere (R/geometryfns.R:202-219,1717-1740)."""
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    if m.kind == "STsphere":   # gc_dist_time: sqrt(spatdist^2 + tdist^2), R/geometryfns.R:181-189
        sd = host_distance(Manifold("sphere", m.radius), x1[:, :2], x2[:, :2])
        dt = x1[:, 2][:, None] - x2[:, 2][None, :]
        td = np.sqrt(dt * dt)
        return np.sqrt(sd * sd + td * td)
    if m.kind == "sphere":
        def normals(x):
            xr = (x * np.pi) / 180.0
            cl = np.cos(xr[:, 1])
            return cl * np.cos(xr[:, 0]), cl * np.sin(xr[:, 0]), np.sin(xr[:, 1])
        a1, a2, a3 = normals(x1)
        b1, b2, b3 = normals(x2)
        dl = (a1[:, None] * b1[None, :] + a2[:, None] * b2[None, :]) + a3[:, None] * b3[None, :]
        dl = np.where(np.abs(dl) > 1.0, np.sign(dl), dl)
        return m.radius * np.arccos(dl)
    acc = None
    for k in range(x1.shape[1]):
        df = x1[:, k][:, None] - x2[:, k][None, :]
        acc = df * df if acc is None else acc + df * df
    return np.sqrt(acc)


# ---------------------------------------------------------------------------------------
# data and BAU containers
# ---------------------------------------------------------------------------------------

@dataclass
class SpatialPointsDataFrame:
    """coords (n x d) + named numeric columns -- mirror of sp's class as FRK uses it."""
    coords: np.ndarray
    data: Dict[str, np.ndarray]

    def __post_init__(self):
        self.coords = np.asarray(self.coords, dtype=np.float64)
        if self.coords.ndim == 1:
            self.coords = self.coords[:, None]
        n = self.coords.shape[0]
        for k, v in self.data.items():
            v = np.asarray(v, dtype=np.float64)
            if v.shape[0] != n:
                raise ValueError(f"column {k!r} has {v.shape[0]} rows, expected {n}")
            self.data[k] = v

    def __len__(self):
        return self.coords.shape[0]


@dataclass
class SpatialPolygonsDataFrame:
    """Areal observations (sp's SpatialPolygonsDataFrame as FRK uses it): polygon k is the single ring
    ``vx/vy[ptr[k]:ptr[k+1]]`` (no holes) and ``data`` holds one value per polygon (the response, ``std``, ...).
    An observation is the average of the process over the BAUs whose centroid lies in its polygon (BuildC,
    R/geometryfns.R:1572-1616 with normalise_wts = TRUE, R/SRE.R:483-490)."""
    ptr: np.ndarray
    vx: np.ndarray
    vy: np.ndarray
    data: Dict[str, np.ndarray]

    def __post_init__(self):
        self.ptr = np.ascontiguousarray(self.ptr, dtype=np.int32)
        self.vx = np.ascontiguousarray(self.vx, dtype=np.float64)
        self.vy = np.ascontiguousarray(self.vy, dtype=np.float64)
        n = len(self.ptr) - 1
        if n < 1 or self.ptr[0] != 0 or np.any(np.diff(self.ptr) < 3) or self.ptr[-1] != len(self.vx) or len(self.vx) != len(self.vy):
            raise ValueError("polygons need ptr[0] = 0, at least 3 vertices each and ptr[-1] = len(vx) = len(vy)")
        for k, v in self.data.items():
            v = np.asarray(v, dtype=np.float64)
            if v.shape[0] != n:
                raise ValueError(f"column {k!r} has {v.shape[0]} rows, expected {n}")
            self.data[k] = v

    def __len__(self):
        return len(self.ptr) - 1


@dataclass
class GridBAUs:
    """SpatialPixelsDataFrame of grid BAUs on the plane (auto_BAU, R/geometryfns.R:553-697).

    BAU row i has centroid (xgrid[i % nx], ygrid[i // nx]) -- expand.grid order.  ``fields`` holds
    BAU-level columns; ``fs`` (fine-scale variance field) must be present before SRE()."""
    xgrid: np.ndarray
    ygrid: np.ndarray
    cellsize: np.ndarray
    fields: Dict[str, np.ndarray] = field(default_factory=dict)

    @property
    def nx(self):
        return len(self.xgrid)

    @property
    def ny(self):
        return len(self.ygrid)

    def __len__(self):
        return self.nx * self.ny

    @property
    def offset(self):
        return np.array([self.xgrid[0], self.ygrid[0]], dtype=np.float64)

    @property
    def dims(self):
        return np.array([self.nx, self.ny], dtype=np.int32)

    def centroids(self, lo: int = 0, hi: Optional[int] = None) -> np.ndarray:
        hi = len(self) if hi is None else hi
        idx = np.arange(lo, hi)
        return np.column_stack([self.xgrid[idx % self.nx], self.ygrid[idx // self.nx]])

    def cell2bau(self) -> np.ndarray:
        """sp full-grid cell (rows counted from the top) -> BAU row."""
        cell = np.arange(len(self), dtype=np.int64)
        return ((self.ny - 1 - cell // self.nx) * self.nx + cell % self.nx).astype(np.int32)

    def __setitem__(self, k, v):
        v = np.broadcast_to(np.asarray(v, dtype=np.float64), (len(self),)).copy()
        self.fields[k] = v

    def __getitem__(self, k):
        return self.fields[k]


@dataclass
class PointBAUs:
    """BAUs given by their centroids only (e.g. ISEA3H hexagons on the sphere).  The point-in-
    polygon lookup for such BAUs is on the host in the reference (sp::over) and is not part of the
    device path yet (SURVEY.md s8 'next' #2): data must carry a precomputed ``bau`` column."""
    coords: np.ndarray
    fields: Dict[str, np.ndarray] = field(default_factory=dict)

    def __len__(self):
        return self.coords.shape[0]

    def centroids(self, lo: int = 0, hi: Optional[int] = None) -> np.ndarray:
        return np.asarray(self.coords[lo:hi], dtype=np.float64)

    def __setitem__(self, k, v):
        self.fields[k] = np.broadcast_to(np.asarray(v, dtype=np.float64), (len(self),)).copy()

    def __getitem__(self, k):
        return self.fields[k]


@dataclass
class PolygonBAUs(PointBAUs):
    """SpatialPolygonsDataFrame BAUs (e.g. ISEA3H hexagons from auto_BAUs(sphere(), type = "hex"), R/geometryfns.R:765-800): polygon k
    is the ring ``vx/vy[ptr[k]:ptr[k+1]]``; ``coords`` are the centroids that SRE() evaluates the basis at (R/SRE.R:412).  The
    point-in-polygon lookup of map_data_to_BAUs runs on the device (frk_polygon_lookup).  BAU *construction* (DGGRID, dateline
    splitting with sf) stays outside: the rings are an input."""
    ptr: np.ndarray = None
    vx: np.ndarray = None
    vy: np.ndarray = None
    _handle: object = field(default=None, repr=False, compare=False)

    def handle(self, dev):
        if self._handle is None:
            ptr = np.ascontiguousarray(self.ptr, dtype=np.int32)
            vx = np.ascontiguousarray(self.vx, dtype=np.float64)
            vy = np.ascontiguousarray(self.vy, dtype=np.float64)
            h = vp()
            call("frk_polygons_create", dev.ctx, len(ptr) - 1, vp(ptr.ctypes.data), vp(vx.ctypes.data), vp(vy.ctypes.data), C.byref(h))
            self._handle = h
        return self._handle

    def __del__(self):
        try:
            if self._handle is not None:
                from ._lib import lib
                lib.frk_polygons_destroy(self._handle)
        except Exception:
            pass


@dataclass
class STBAUs:
    """STFDF BAUs: ``spatial`` BAUs (grid, polygon or point) x ``times`` (ascending interval starts, numeric), space running
    fastest -- auto_BAUs(STplane()/STsphere()), R/geometryfns.R:871-991.  BAU row i is spatial BAU ``i % ns`` in time interval
    ``i // ns``; its time coordinate for eval_basis is the reference's ``t = rep(1:nt, each = ns)`` (:974-975)."""
    spatial: object
    times: np.ndarray
    fields: Dict[str, np.ndarray] = field(default_factory=dict)

    def __post_init__(self):
        self.times = np.ascontiguousarray(self.times, dtype=np.float64)
        if self.times.ndim != 1 or len(self.times) == 0 or np.any(np.diff(self.times) <= 0):
            raise ValueError("times needs to be a non-empty increasing vector of interval starts")

    @property
    def ns(self):
        return len(self.spatial)

    @property
    def nt(self):
        return len(self.times)

    def __len__(self):
        return self.ns * self.nt

    def centroids(self, lo: int = 0, hi: Optional[int] = None) -> np.ndarray:
        hi = len(self) if hi is None else hi
        idx = np.arange(lo, hi)
        sp_i = idx % self.ns
        cs = self.spatial.centroids()
        return np.column_stack([cs[sp_i], (idx // self.ns + 1).astype(np.float64)])

    def __setitem__(self, k, v):
        self.fields[k] = np.broadcast_to(np.asarray(v, dtype=np.float64), (len(self),)).copy()

    def __getitem__(self, k):
        return self.fields[k]


def _seq_by(a: float, b: float, by: float) -> np.ndarray:
    n = int(math.floor((b - a) / by + 1e-10))        # seq.default(from, to, by)
    return a + np.arange(n + 1, dtype=np.float64) * by


def auto_BAUs(manifold: Manifold, cellsize=None, data: Optional[SpatialPointsDataFrame] = None,
              xlims=None, ylims=None, type: str = "grid", nonconvex_hull: bool = False) -> GridBAUs:
    """auto_BAUs(manifold = plane(), type = "grid", ...), R/geometryfns.R:340,553-697.

    The grid is the reference's (20 % buffer unless limits are given, seq(..., by = cellsize),
    expand.grid order).  Hull pruning of the grid (fmesher / sp::over on a polygon, :665-682) is
    BAU construction, out of the hot-path scope: the full rectangle is returned."""
    if manifold.kind != "plane" or type != "grid":
        raise FRKError("auto_BAUs on the device path builds grid BAUs on the plane only; "
                       "pass PointBAUs for other BAU types")
    if nonconvex_hull:
        raise FRKError("nonconvex_hull=TRUE needs fmesher (host geometry, out of scope); use nonconvex_hull=False")
    if data is None and (xlims is None or ylims is None):
        raise ValueError("Data must be supplied when generating BAUs on a plane")
    if data is not None:
        coords = data.coords
        xr = (coords[:, 0].min(), coords[:, 0].max()) if xlims is None else tuple(xlims)
        yr = (coords[:, 1].min(), coords[:, 1].max()) if ylims is None else tuple(ylims)
    else:
        xr, yr = tuple(xlims), tuple(ylims)
    if cellsize is None:
        cellsize = max(xr[1] - xr[0], yr[1] - yr[0]) / 100.0     # .choose_cellsize default, :1963-1979
    cellsize = np.broadcast_to(np.asarray(cellsize, dtype=np.float64), (2,)).copy()
    dx, dy = xr[1] - xr[0], yr[1] - yr[0]
    bx = 0.2 if xlims is None else 0.0
    by = 0.2 if ylims is None else 0.0
    xg = _seq_by(xr[0] - dx * bx, xr[1] + dx * bx, cellsize[0])
    yg = _seq_by(yr[0] - dy * by, yr[1] + dy * by, cellsize[1])
    b = GridBAUs(xg, yg, cellsize)
    return b


def map_data_to_BAUs(data: SpatialPointsDataFrame, BAUs, columns, dev: Optional[Device] = None,
                     row_range=None, sum_variables=None):
    """map_data_to_BAUs(SpatialPoints, BAUs, average_in_BAU = TRUE), R/geometryfns.R:1222-1320.

    Device pipeline: integer point-in-cell lookup (sp::over on SpatialPixels) -> drop NAs ->
    group by BAU -> mean of every requested column.  Returns device tensors
    (obs_bau[m] ascending BAU rows, means (ncols x m, i.e. m x ncols column-major), counts[m]) and
    the per-point BAU index.  ``row_range=(lo,hi)`` keeps only points whose BAU is owned by this
    rank (multi-GPU row partition) and returns BAU rows relative to ``lo``.  ``sum_variables``: columns that are summed over the
    observations of a BAU rather than averaged (R/geometryfns.R:1286-1299)."""
    dev = dev or get_device()
    n = len(data)
    ncols = len(columns)
    nbau = len(BAUs)
    lo, hi = (0, nbau) if row_range is None else row_range
    if isinstance(BAUs, STBAUs):
        # every time slice binned in space (R/geometryfns.R:1419-1547): spatial lookup over all spatial BAUs, then the combined index
        if data.coords.shape[1] < 3:
            raise ValueError("space-time BAUs need data coordinates (x, y, time)")
        d_bau = _spatial_lookup(dev, data, BAUs.spatial, n, 0, BAUs.ns)
        d_t = dev.upload(data.coords[:, 2])
        call("frk_st_bau_index", dev.ctx, n, dev.ptr(d_bau), dev.ptr(d_t), BAUs.nt, vp(BAUs.times.ctypes.data), BAUs.ns,
             int(lo), int(hi), dev.ptr(d_bau))
    else:
        d_bau = _spatial_lookup(dev, data, BAUs, n, lo, hi)
    return _bin(dev, data, columns, d_bau, n, ncols, lo, hi, sum_variables)


def point_BAU_index(data: SpatialPointsDataFrame, BAUs, dev: Optional[Device] = None) -> np.ndarray:
    """BAU row of every observation (0-based; -1 = in no BAU): the sp::over step of map_data_to_BAUs (R/geometryfns.R:1266) without
    the averaging, as average_in_BAU = FALSE keeps it (:1304-1305)."""
    dev = dev or get_device()
    n = len(data)
    if isinstance(BAUs, STBAUs):
        raise FRKError("average_in_BAU=FALSE with space-time BAUs is not on the device path")
    d_bau = _spatial_lookup(dev, data, BAUs, n, 0, len(BAUs))
    return dev.download(d_bau, n).astype(np.int64) if n else np.zeros(0, dtype=np.int64)


def _spatial_lookup(dev, data, BAUs, n, lo, hi):
    """per-point BAU row in [lo, hi) relative to lo, -1 = NA (sp::over, R/geometryfns.R:1266)."""
    if isinstance(BAUs, GridBAUs):
        d_xy = dev.upload(data.coords[:, :2])
        d_bau = dev.empty(max(n, 1), "i4")
        off, cs, dm = BAUs.offset, np.asarray(BAUs.cellsize, dtype=np.float64), BAUs.dims
        # BAUs are in expand.grid order, so sp's top-down cell index maps to the BAU row by formula
        call("frk_grid_lookup_range", dev.ctx, n, dev.ptr(d_xy), n, vp(off.ctypes.data), vp(cs.ctypes.data),
             vp(dm.ctypes.data), vp(None), 1, int(lo), int(hi), dev.ptr(d_bau))
    elif isinstance(BAUs, PolygonBAUs):
        d_xy = dev.upload(data.coords[:, :2])
        d_bau = dev.empty(max(n, 1), "i4")
        call("frk_polygon_lookup", dev.ctx, BAUs.handle(dev), n, dev.ptr(d_xy), n, int(lo), int(hi), dev.ptr(d_bau))
    else:
        if "bau" not in data.data:
            raise FRKError("point-in-polygon lookup for non-grid BAUs is not on the device path; "
                           "supply a precomputed 'bau' column (0-based BAU row, -1 = NA)")
        b = np.asarray(data.data["bau"]).astype(np.int64)
        b = np.where((b >= lo) & (b < hi), b - lo, -1)     # host: the lookup itself was done on the host
        d_bau = dev.upload(b.astype(np.int32))
    return d_bau


def _bin(dev, data, columns, d_bau, n, ncols, lo, hi, sum_variables=None):
    """group_by(BAU) %>% summarise_all(mean), R/geometryfns.R:1286-1302."""
    d_vals = dev.empty(max(n * ncols, 1))                          # n x ncols column-major, one H2D per column (no host-side stacking)
    for k, c in enumerate(columns):
        colv = np.ascontiguousarray(data.data[c], dtype=np.float64)
        if n:
            call("frk_h2d", dev.ctx, dev.ptr(d_vals, k * n), vp(colv.ctypes.data), colv.nbytes)
    call("frk_sync", dev.ctx)
    d_obs = dev.empty(max(n, 1), "i4")
    d_means = dev.empty(max(n * ncols, 1), "f8")
    d_cnt = dev.empty(max(n, 1), "i4")
    m = c_i64(0)
    mask = 0
    for v in (sum_variables or []):
        if v not in columns:
            raise ValueError(f"sum_variables: {v!r} is not one of the binned columns")
        mask |= 1 << list(columns).index(v)
    call("frk_bin_mean_sum", dev.ctx, n, dev.ptr(d_bau), hi - lo, ncols, dev.ptr(d_vals), mask, dev.ptr(d_obs),
         dev.ptr(d_means), dev.ptr(d_cnt), C.byref(m))
    return dict(m=int(m.value), obs_bau=d_obs, means=d_means, ld=n, counts=d_cnt, bau=d_bau)

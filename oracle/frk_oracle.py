"""CPU oracle for the Gaussian-FRK hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This file is a NumPy/SciPy fp64 *restatement* of the algorithm in the reference
R package FRK v2.3.1 (paths below are relative to /root/reference).  It exists so
that the CUDA path in ``frk_b200`` can be checked for parity.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it; the product path (``frk_b200``) never does.

Provenance / pinning.  The reference is pure R on top of Matrix / sparseinv /
sp / stats (all *unvendored, unpinned* third-party packages) and there is no R
interpreter in this image or on the GPU box, so the reference itself can never
be executed here.  The oracle is pinned by:
  * every known-answer test the reference's own tests hold for this path
    (tests/testthat/test_basis.R:23-24,51-52,62-63; test_domains.R:25-36;
    test_other.R:114-120; test_linalg.R:9-25; test_SRE.R:46-52) --
    see tests/test_oracle_kat.py;
  * independent algebra (direct dense Gaussian conditioning vs. the Woodbury
    forms; general (r+N) joint solve vs. the closed form used for prediction).
EM estimates, the log-likelihood trace and predictions are *parity-unpinned* by
the reference's own tests (it holds no golden vectors for them); for those the
oracle is a line-by-line restatement of the cited reference lines only.

Third-party arithmetic restated from the published algorithms:
  * stats::optim(method="Nelder-Mead")  -> nmmin()   (R src/appl/optim.c)
  * stats::uniroot                      -> zeroin()  (R src/library/stats/src/zeroin.c)
  * sp::over(SpatialPoints, SpatialPixels) -> grid_index() (sp getGridIndex)
  * Matrix::chol / chol2inv / determinant -> numpy.linalg / scipy.linalg (LAPACK)

Arithmetic order follows SURVEY.md Appendix A.  Order-sensitive arithmetic
(distance sums, the 3-term sphere dot product, basis functions) is written as
explicit elementwise NumPy ops -- never ``@``/``dot`` -- because the bundled
OpenBLAS uses FMA.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

EARTH_RADIUS = 6371.0          # sphere() default, R/geometryfns.R:97
DEFAULT_GC_RADIUS = 6378.137   # dist_sphere() when R=NULL, R/geometryfns.R:1720

# --------------------------------------------------------------------------
# A.1 distances
# --------------------------------------------------------------------------

def distR(x1: np.ndarray, x2: Optional[np.ndarray] = None) -> np.ndarray:
    """Euclidean distance matrix, R/geometryfns.R:202-219.

    ``sqrt(Reduce("+", lapply(1:d, function(i) outer(x1[,i],x2[,i],'-')^2)))``:
    squared displacements added in dimension order, every op rounded (no FMA).
    """
    x1 = np.asarray(x1, dtype=np.float64)
    if x1.ndim == 1:
        x1 = x1[:, None]
    x2 = x1 if x2 is None else np.asarray(x2, dtype=np.float64)
    if x2.ndim == 1:
        x2 = x2[:, None]
    if x1.shape[1] != x2.shape[1]:
        raise ValueError("x1 and x2 have to have same number of columns")
    acc = None
    for i in range(x1.shape[1]):
        d = x1[:, i][:, None] - x2[:, i][None, :]
        sq = d * d
        acc = sq if acc is None else acc + sq
    return np.sqrt(acc)


def _unit_normals(x: np.ndarray):
    """lon/lat (deg) -> (cos(lat)cos(lon), cos(lat)sin(lon), sin(lat)); R/geometryfns.R:1726-1731."""
    xr = (x * np.pi) / 180.0                      # x * pi/180 is left-associative in R
    lon, lat = xr[:, 0], xr[:, 1]
    cl = np.cos(lat)
    return cl * np.cos(lon), cl * np.sin(lon), np.sin(lat)


def dist_sphere(x1: np.ndarray, x2: Optional[np.ndarray] = None, R: Optional[float] = None) -> np.ndarray:
    """Great-circle distance, R/geometryfns.R:1717-1740.

    ``tcrossprod(n1,n2)`` is a k=3 dgemm: restated in reference-BLAS order
    ((a1*b1)+a2*b2)+a3*b3 with no FMA.  Clamp |delta|>1 to sign(delta).
    """
    if R is None:
        R = DEFAULT_GC_RADIUS
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = x1 if x2 is None else np.asarray(x2, dtype=np.float64)
    a1, a2, a3 = _unit_normals(x1)
    b1, b2, b3 = _unit_normals(x2)
    delta = (a1[:, None] * b1[None, :] + a2[:, None] * b2[None, :]) + a3[:, None] * b3[None, :]
    delta = np.where(np.abs(delta) > 1.0, np.sign(delta), delta)
    return R * np.arccos(delta)


def gc_dist_time(x1: np.ndarray, x2: np.ndarray, R: Optional[float] = None) -> np.ndarray:
    """sqrt(gc^2 + dt^2), R/geometryfns.R:181-189."""
    sd = dist_sphere(x1[:, :2], x2[:, :2], R=R)
    td = distR(x1[:, 2], x2[:, 2])
    return np.sqrt(sd * sd + td * td)


@dataclass
class Manifold:
    """real_line / plane / sphere / STplane / STsphere (R/geometryfns.R:19-189).  ``radius`` only on spheres."""
    kind: str                      # "real_line" | "plane" | "sphere" | "STplane" | "STsphere"
    radius: float = EARTH_RADIUS

    @property
    def dim(self) -> int:
        return {"real_line": 1, "plane": 2, "sphere": 2, "STplane": 3, "STsphere": 3}[self.kind]

    def distance(self, x1, x2=None):
        if self.kind == "sphere":
            return dist_sphere(x1, x2, R=self.radius)
        if self.kind == "STsphere":
            x1 = np.asarray(x1, dtype=np.float64)
            return gc_dist_time(x1, x1 if x2 is None else np.asarray(x2, dtype=np.float64), R=self.radius)
        return distR(x1, x2)


def real_line():
    return Manifold("real_line")


def plane():
    return Manifold("plane")


def sphere(radius: float = EARTH_RADIUS):
    return Manifold("sphere", radius)


def STplane():
    """plane x time, Euclid_dist(dim=3L) (R/geometryfns.R:60-80)."""
    return Manifold("STplane")


def STsphere(radius: float = EARTH_RADIUS):
    """sphere x time, gc_dist_time (R/geometryfns.R:120-134)."""
    return Manifold("STsphere", radius)

# --------------------------------------------------------------------------
# A.2 basis functions
# --------------------------------------------------------------------------

BASIS_TYPES = ("bisquare", "Gaussian", "exp", "Matern32")


def basis_fn(kind: str, y: np.ndarray, scale) -> np.ndarray:
    """The four closures, R/basisfns.R:1213-1250 (y = distance, one scale per function)."""
    if kind == "bisquare":
        t = y / scale
        u = 1.0 - t * t
        return (u * u) * (y < scale)
    if kind == "Gaussian":
        q = y * y
        return np.exp((-0.5 * q) / (scale * scale))
    if kind == "exp":
        return np.exp(-y / scale)
    if kind == "Matern32":
        a = (math.sqrt(3.0) * y) / scale
        return (1.0 + a) * np.exp(-a)
    raise ValueError(kind)


@dataclass
class Basis:
    """Restates the ``Basis`` S4 class (R/AllClass.R:78) as plain arrays."""
    manifold: Manifold
    loc: np.ndarray            # r x d
    scale: np.ndarray          # r
    res: np.ndarray            # r (1-based resolution)
    type: str = "bisquare"
    regular: int = 0

    @property
    def n(self) -> int:
        return self.loc.shape[0]

    @property
    def nres(self) -> int:
        return int(self.res.max())


@dataclass
class TensorP_Basis:
    """R/basisfns.R:659-704: space x time, column index = i_space + r_s*(i_time-1)."""
    Basis1: Basis
    Basis2: Basis

    @property
    def n(self) -> int:
        return self.Basis1.n * self.Basis2.n

    @property
    def res(self) -> np.ndarray:
        # df of the tensor basis: space runs fastest; res = spatial res (Basis2 has one res)
        return np.tile(self.Basis1.res, self.Basis2.n)


def local_basis(manifold: Manifold, loc, scale, type: str = "bisquare", res=1, regular=0) -> Basis:
    """R/basisfns.R:99-146."""
    loc = np.asarray(loc, dtype=np.float64)
    if loc.ndim != 2:
        raise ValueError("loc needs to be a matrix")
    if manifold.dim != loc.shape[1]:
        raise ValueError("number of columns in loc needs to be the same as the number of manifold dimensions")
    scale = np.atleast_1d(np.asarray(scale, dtype=np.float64))
    if scale.shape[0] != loc.shape[0]:
        raise ValueError("need to have as many scale parameters as centroids")
    if np.any(~(scale > 0)):
        raise ValueError("The scale parameter needs to be greater than zero")
    if type not in BASIS_TYPES:
        raise ValueError(type)
    res = np.broadcast_to(np.asarray(res, dtype=np.int32), (loc.shape[0],)).copy()
    return Basis(manifold, loc, scale, res, type, int(regular))


def _seq_len(a: float, b: float, n: int) -> np.ndarray:
    """R's seq(a, b, length.out=n): ``from + (0:(n-1))*by`` with by=(b-a)/(n-1) (seq.default)."""
    if n == 1:
        return np.array([a], dtype=np.float64)
    by = (b - a) / (n - 1)
    out = a + np.arange(n, dtype=np.float64) * by
    return out


def _r_round(x: float) -> int:
    """R >= 4.0 round(): half-to-even."""
    return int(np.rint(x))


def auto_basis(manifold: Manifold, coords: np.ndarray, regular: int = 1, nres: int = 3,
               type: str = "bisquare", isea3h_lo: int = 2, scale_aperture: Optional[float] = None,
               isea3h: Optional[dict] = None, prune: float = 0) -> Basis:
    """Regular placement, R/basisfns.R:358-585 (subsamp>=n, no buffer), including the deprecated ``prune`` refinement (:479-553).

    ``isea3h``: {res: (n_res x 2) lon/lat centroids} from data/isea3h.rda for the sphere.
    """
    coords = np.asarray(coords, dtype=np.float64)
    if scale_aperture is None:
        scale_aperture = 1.0 if manifold.kind == "sphere" else 1.25       # :228
    xrange_ = (coords[:, 0].min(), coords[:, 0].max())                     # :396
    yrange_ = (coords[:, 1].min(), coords[:, 1].max()) if coords.shape[1] > 1 else (0.0, 0.0)
    if manifold.kind == "plane":
        asp = (yrange_[1] - yrange_[0]) / (xrange_[1] - xrange_[0])      # :414-423
        if asp < 1:
            ny = regular
            nx = ny / asp
        else:
            nx = regular
            ny = nx * asp
    mult = 1.5 if type == "bisquare" else 0.7                             # :532
    locs, scales, ress = [], [], []
    for i in range(1, nres + 1):
        xgrid = ygrid = None
        if manifold.kind == "plane":
            xgrid = _seq_len(xrange_[0], xrange_[1], _r_round(nx * 3 ** i))  # :460-461
            ygrid = _seq_len(yrange_[0], yrange_[1], _r_round(ny * 3 ** i))
            gx, gy = np.meshgrid(xgrid, ygrid, indexing="xy")              # expand.grid: x fastest
            this = np.column_stack([gx.ravel(), gy.ravel()])
        elif manifold.kind == "real_line":
            this = _seq_len(xrange_[0], xrange_[1], i * regular)[:, None]  # :468
        else:
            this = np.asarray(isea3h[i + isea3h_lo - 1], dtype=np.float64)  # :472-474
        base = np.zeros(0)
        for j in (1, 2):                                                   # :479 "go over this twice for refinement when pruning"
            nloc = this.shape[0]
            if nloc == 0:
                base = np.zeros(0)
                break
            if nloc == 1:                                                  # :482-483
                base = np.array([max(xrange_[1] - xrange_[0], yrange_[1] - yrange_[0]) / 2.0])
            elif nloc < 5000:                                              # :485-490
                D = manifold.distance(this, this)
                np.fill_diagonal(D, np.inf)
                base = D.min(axis=1) * scale_aperture
            elif not prune and regular:                                    # :492-495
                diffx = abs(xgrid[1] - xgrid[0])
                diffy = abs(ygrid[1] - ygrid[0])
                base = np.full(nloc, min(diffx, diffy)) * scale_aperture
            else:                                                          # :497-509, batches of 5000 rows against all centroids
                parts = []
                for i0 in range(0, nloc, 5000):
                    D = manifold.distance(this[i0:i0 + 5000], this)
                    D[D == 0] = np.inf
                    parts.append(D.min(axis=1) * scale_aperture)
                base = np.concatenate(parts)
            if not (prune > 0):                                            # :551-553
                break
            if j == 1:                                                     # :536-549
                cand = Basis(manifold, this, mult * base, np.full(nloc, i, dtype=np.int32), type, int(regular))
                cs = np.zeros(nloc)
                for r0 in range(0, coords.shape[0], 2000):
                    cs += eval_basis_dense(cand, coords[r0:r0 + 2000]).sum(axis=0)
                rm = cs < prune
                if rm.any():
                    this = this[~rm]
        if this.shape[0] == 0:
            continue
        locs.append(this)
        scales.append(mult * base)
        ress.append(np.full(this.shape[0], i, dtype=np.int32))
    return Basis(manifold, np.vstack(locs), np.concatenate(scales), np.concatenate(ress), type, int(regular))


def eval_basis_dense(basis: Basis, s: np.ndarray) -> np.ndarray:
    """.point_eval_fn (R/basisfns.R:1157-1161): cbind over the r closures, dense n x r."""
    s = np.asarray(s, dtype=np.float64)
    if s.ndim == 1:
        s = s[:, None]
    d = basis.manifold.dim
    s = s[:, :d]
    n, r = s.shape[0], basis.n
    out = np.empty((n, r), dtype=np.float64)
    if basis.manifold.kind in ("sphere", "STsphere"):
        a1, a2, a3 = _unit_normals(s)
        b1, b2, b3 = _unit_normals(basis.loc)
        CH = 256
        for j0 in range(0, r, CH):
            j1 = min(r, j0 + CH)
            delta = (a1[:, None] * b1[None, j0:j1] + a2[:, None] * b2[None, j0:j1]) + a3[:, None] * b3[None, j0:j1]
            delta = np.where(np.abs(delta) > 1.0, np.sign(delta), delta)
            y = basis.manifold.radius * np.arccos(delta)
            if basis.manifold.kind == "STsphere":     # gc_dist_time, R/geometryfns.R:181-189
                dt = s[:, 2][:, None] - basis.loc[None, j0:j1, 2]
                td = np.sqrt(dt * dt)
                y = np.sqrt(y * y + td * td)
            out[:, j0:j1] = basis_fn(basis.type, y, basis.scale[None, j0:j1])
        return out
    CH = 256
    for j0 in range(0, r, CH):
        j1 = min(r, j0 + CH)
        acc = None
        for k in range(d):
            df = s[:, k][:, None] - basis.loc[None, j0:j1, k]
            sq = df * df
            acc = sq if acc is None else acc + sq
        y = np.sqrt(acc)
        out[:, j0:j1] = basis_fn(basis.type, y, basis.scale[None, j0:j1])
    return out


def eval_basis(basis, s: np.ndarray, batch: int = 10000) -> sp.csr_matrix:
    """eval_basis(Basis|TensorP_Basis, matrix), R/basisfns.R:709-748, 812-824.

    ``as(x,"Matrix")`` drops exact zeros, so the pattern is {phi != 0}.  Returned as CSR
    with ascending column indices (the reference returns the same matrix as CSC).
    """
    s = np.asarray(s, dtype=np.float64)
    if s.ndim == 1:
        s = s[:, None]
    if isinstance(basis, TensorP_Basis):
        n1 = basis.Basis1.manifold.dim
        S1 = eval_basis(basis.Basis1, s[:, :n1], batch)
        S2 = eval_basis(basis.Basis2, s[:, n1:], batch)
        return tensor_rows(S1, S2)
    blocks = []
    for i0 in range(0, s.shape[0], batch):
        blocks.append(sp.csr_matrix(eval_basis_dense(basis, s[i0:i0 + batch])))
    if not blocks:
        return sp.csr_matrix((0, basis.n))
    out = sp.vstack(blocks, format="csr")
    out.sort_indices()
    return out


def eval_basis_fast(basis: Basis, s: np.ndarray) -> sp.csr_matrix:
    """Same arithmetic as eval_basis() for a *bisquare* basis, but candidates come from a
    KD-tree superset instead of all N*r pairs (used for larger parity cases; checked equal
    to eval_basis() in the CPU tests)."""
    from scipy.spatial import cKDTree
    assert basis.type == "bisquare"
    s = np.asarray(s, dtype=np.float64)
    if s.ndim == 1:
        s = s[:, None]
    d = basis.manifold.dim
    s = s[:, :d]
    if basis.manifold.kind == "sphere":
        a = np.column_stack(_unit_normals(s))
        b = np.column_stack(_unit_normals(basis.loc))
        ang = basis.scale / basis.manifold.radius
        rad = 2.0 * np.sin(np.minimum(ang, np.pi) / 2.0) * (1 + 1e-9) + 1e-12   # chord length
        pts, cen = a, b
    else:
        pts, cen, rad = s, basis.loc[:, :d], basis.scale * (1 + 1e-12)
    rows, cols = [], []
    tree = cKDTree(pts)
    for rr in np.unique(rad):
        idx = np.nonzero(rad == rr)[0]
        lists = tree.query_ball_point(cen[idx], rr)
        for j, l in zip(idx, lists):
            if len(l):
                rows.append(np.asarray(l, dtype=np.int64))
                cols.append(np.full(len(l), j, dtype=np.int64))
    if not rows:
        return sp.csr_matrix((s.shape[0], basis.n))
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    if basis.manifold.kind == "sphere":
        a1, a2, a3 = _unit_normals(s)
        b1, b2, b3 = _unit_normals(basis.loc)
        delta = (a1[rows] * b1[cols] + a2[rows] * b2[cols]) + a3[rows] * b3[cols]
        delta = np.where(np.abs(delta) > 1.0, np.sign(delta), delta)
        y = basis.manifold.radius * np.arccos(delta)
    else:
        acc = None
        for k in range(d):
            df = s[rows, k] - basis.loc[cols, k]
            sq = df * df
            acc = sq if acc is None else acc + sq
        y = np.sqrt(acc)
    v = basis_fn("bisquare", y, basis.scale[cols])
    keep = v != 0
    out = sp.csr_matrix((v[keep], (rows[keep], cols[keep])), shape=(s.shape[0], basis.n))
    out.sort_indices()
    return out


def tensor_rows(S1: sp.csr_matrix, S2: sp.csr_matrix) -> sp.csr_matrix:
    """quickcbind(lapply(1:ncol(S2), function(i) S2[,i]*S1)), R/basisfns.R:821-822 +
    R/linalgfns.R:80-118: S[i, j_s + r_s*j_t] = S1[i,j_s]*S2[i,j_t] (0-based), space fastest."""
    S1 = S1.tocsr()
    S2 = S2.tocsr()
    n, rs = S1.shape
    rt = S2.shape[1]
    n1 = np.diff(S1.indptr)
    n2 = np.diff(S2.indptr)
    cnt = n1 * n2
    indptr = np.concatenate([[0], np.cumsum(cnt)])
    nnz = int(indptr[-1])
    indices = np.empty(nnz, dtype=np.int64)
    data = np.empty(nnz, dtype=np.float64)
    for i in range(n):
        c1 = S1.indices[S1.indptr[i]:S1.indptr[i + 1]]
        v1 = S1.data[S1.indptr[i]:S1.indptr[i + 1]]
        c2 = S2.indices[S2.indptr[i]:S2.indptr[i + 1]]
        v2 = S2.data[S2.indptr[i]:S2.indptr[i + 1]]
        if len(c1) and len(c2):
            indices[indptr[i]:indptr[i + 1]] = (c1[None, :] + rs * c2[:, None]).ravel()
            data[indptr[i]:indptr[i + 1]] = (v2[:, None] * v1[None, :]).ravel()
    out = sp.csr_matrix((data, indices, indptr), shape=(n, rs * rt))
    return out


def normalise_rows(S0: sp.csr_matrix) -> sp.csr_matrix:
    """R/SRE.R:416-421: xx = sqrt(rowSums(S0*S0)); xx[xx==0] = 1; S0 = S0/xx."""
    S0 = S0.tocsr().copy()
    sq = S0.data * S0.data
    xx = np.zeros(S0.shape[0])
    # sequential left-to-right row sums (rowSums accumulates in column order)
    rows = np.repeat(np.arange(S0.shape[0]), np.diff(S0.indptr))
    np.add.at(xx, rows, sq)
    xx = np.sqrt(xx)
    xx = xx + 1.0 * (xx == 0)
    S0.data = S0.data / xx[rows]
    return S0


def BuildD(basis):
    """R/basisfns.R:1131-1153: per-resolution centroid distance matrices."""
    if isinstance(basis, TensorP_Basis):
        return {"Basis1": BuildD(basis.Basis1), "Basis2": BuildD(basis.Basis2)}
    out = []
    for i in range(1, basis.nres + 1):
        x = basis.loc[basis.res == i]
        out.append(basis.manifold.distance(x, x))
    return out

# --------------------------------------------------------------------------
# A.3 BAUs, point -> BAU lookup, binning, C_Z
# --------------------------------------------------------------------------

@dataclass
class GridBAUs:
    """A SpatialPixels-like regular grid of BAUs on the plane (R/geometryfns.R:629-697).

    BAU i (0-based) has centroid (xgrid[i % nx], ygrid[i // nx]) -- ``expand.grid`` order,
    x fastest, y ascending.  ``grid_index`` is sp's full-grid index of every BAU (sp numbers
    grid cells row-wise from the TOP row).  ``fs`` is the fine-scale variance field.
    """
    xgrid: np.ndarray
    ygrid: np.ndarray
    cellsize: np.ndarray
    fs: Optional[np.ndarray] = None

    def __post_init__(self):
        if self.fs is None:
            self.fs = np.ones(self.n)

    @property
    def nx(self):
        return len(self.xgrid)

    @property
    def ny(self):
        return len(self.ygrid)

    @property
    def n(self):
        return self.nx * self.ny

    @property
    def offset(self):
        """cellcentre.offset = centre of the lower-left cell."""
        return np.array([self.xgrid[0], self.ygrid[0]])

    @property
    def dims(self):
        return np.array([self.nx, self.ny], dtype=np.int64)

    def centroids(self) -> np.ndarray:
        gx, gy = np.meshgrid(self.xgrid, self.ygrid, indexing="xy")
        return np.column_stack([gx.ravel(), gy.ravel()])

    def cell2bau(self) -> np.ndarray:
        """Map sp full-grid cell index (0-based, top row first) -> BAU row (0-based)."""
        cell = np.arange(self.n)
        col = cell % self.nx
        row_from_top = cell // self.nx
        return ((self.ny - 1 - row_from_top) * self.nx + col).astype(np.int32)


def _seq_by(a: float, b: float, by: float) -> np.ndarray:
    """R's seq(from, to, by): n = floor((to-from)/by + 1e-10); from + (0:n)*by (seq.default)."""
    n = int(math.floor((b - a) / by + 1e-10))
    return a + np.arange(n + 1, dtype=np.float64) * by


def auto_BAUs_plane(coords: np.ndarray, cellsize, xlims=None, ylims=None) -> GridBAUs:
    """Grid BAUs on the plane, R/geometryfns.R:583-637 (grid construction only).

    Hull pruning (:665-682; fmesher / sp::over on a polygon) is BAU *construction*, which
    SURVEY.md s8 marks out of scope: the full rectangular grid is used and the identical BAU
    table is fed to the oracle and to the GPU path.
    """
    coords = np.asarray(coords, dtype=np.float64)
    cellsize = np.broadcast_to(np.asarray(cellsize, dtype=np.float64), (2,))
    xr = (coords[:, 0].min(), coords[:, 0].max()) if xlims is None else xlims
    yr = (coords[:, 1].min(), coords[:, 1].max()) if ylims is None else ylims
    dx, dy = xr[1] - xr[0], yr[1] - yr[0]
    bx = 0.2 if xlims is None else 0.0
    by = 0.2 if ylims is None else 0.0
    xgrid = _seq_by(xr[0] - dx * bx, xr[1] + dx * bx, cellsize[0])
    ygrid = _seq_by(yr[0] - dy * by, yr[1] + dy * by, cellsize[1])
    return GridBAUs(xgrid, ygrid, cellsize.copy())


def grid_index(xy: np.ndarray, offset, cellsize, dims) -> np.ndarray:
    """sp::getGridIndex(cc, grid, all.inside=FALSE), 0-based, -1 for NA.

    Restated from the published sp algorithm (sp is not vendored; unpinned):
      idx_d = round((cc[,d] - cellcentre.offset[d]) / cellsize[d])   (R round = half-even)
      d == 2: idx_2 = cells.dim[2] - (idx_2 + 1)                      (rows from the top)
      outside if idx_d >= cells.dim[d] or idx_d < 0  -> NA
      idx = 1 + idx_1 + idx_2*cells.dim[1]
    """
    xy = np.asarray(xy, dtype=np.float64)
    ix = np.rint((xy[:, 0] - offset[0]) / cellsize[0])
    iy = np.rint((xy[:, 1] - offset[1]) / cellsize[1])
    iy = dims[1] - (iy + 1)
    outside = (ix >= dims[0]) | (ix < 0) | (iy >= dims[1]) | (iy < 0) | ~np.isfinite(ix) | ~np.isfinite(iy)
    cell = np.where(outside, -1, ix + iy * dims[0])
    return cell.astype(np.int64)


def points_to_BAUs(xy: np.ndarray, baus: GridBAUs, cell2bau: Optional[np.ndarray] = None) -> np.ndarray:
    """over(SpatialPoints, SpatialPixels): full-grid cell -> row of the pixel set (R/geometryfns.R:1266)."""
    cell = grid_index(xy, baus.offset, baus.cellsize, baus.dims)
    if cell2bau is None:
        cell2bau = baus.cell2bau()
    out = np.where(cell >= 0, cell2bau[np.maximum(cell, 0)], -1)
    return out.astype(np.int32)


def in_poly(px: float, py: float, vx: np.ndarray, vy: np.ndarray) -> int:
    """sp::point.in.polygon for one point and one ring: 0 outside, 1 inside, 2 on an edge, 3 on a vertex.

    sp is not vendored in the reference (unpinned CRAN dependency); its pip.c ``InPoly`` is O'Rourke's crossing test
    ("Computational Geometry in C", 2nd ed., code 7.13), restated here: shift the ring so that the point is the origin,
    count the edges that cross the positive / negative x axis, using x = (x_i y_{i-1} - x_{i-1} y_i) / (y_{i-1} - y_i).
    Every operation is a single IEEE operation, in this order (the CUDA kernel uses the same order without FMA)."""
    n = len(vx)
    x = vx - px
    y = vy - py
    rcross = lcross = 0
    for i in range(n):
        if x[i] == 0.0 and y[i] == 0.0:
            return 3
        i1 = (i + n - 1) % n
        rstrad = (y[i] > 0) != (y[i1] > 0)
        lstrad = (y[i] < 0) != (y[i1] < 0)
        if rstrad or lstrad:
            xc = (x[i] * y[i1] - x[i1] * y[i]) / (y[i1] - y[i])
            if rstrad and xc > 0:
                rcross += 1
            if lstrad and xc < 0:
                lcross += 1
    if (rcross % 2) != (lcross % 2):
        return 2
    return 1 if (rcross % 2) == 1 else 0


def points_to_polygon_BAUs(xy: np.ndarray, ptr: np.ndarray, vx: np.ndarray, vy: np.ndarray) -> np.ndarray:
    """over(SpatialPoints, SpatialPolygons) as map_data_to_BAUs uses it (R/geometryfns.R:1266,1680): the index of the polygon that
    contains the point (interior, edge or vertex all count: point.in.polygon != 0), -1 (NA) if none.  A point on a shared edge
    belongs to the FIRST polygon in index order here -- sp's tie rule is flagged "to verify against real sp" (SURVEY.md s8(c));
    polygons are single rings without holes (ISEA3H cells).  Bounding boxes are a pre-filter only."""
    xy = np.asarray(xy, dtype=np.float64)
    npoly = len(ptr) - 1
    xmin = np.array([vx[ptr[k]:ptr[k + 1]].min() for k in range(npoly)]); xmax = np.array([vx[ptr[k]:ptr[k + 1]].max() for k in range(npoly)])
    ymin = np.array([vy[ptr[k]:ptr[k + 1]].min() for k in range(npoly)]); ymax = np.array([vy[ptr[k]:ptr[k + 1]].max() for k in range(npoly)])
    out = np.full(len(xy), -1, dtype=np.int32)
    for i, (px, py) in enumerate(xy):
        if not (np.isfinite(px) and np.isfinite(py)):
            continue
        cand = np.nonzero((px >= xmin) & (px <= xmax) & (py >= ymin) & (py <= ymax))[0]
        for k in cand:
            if in_poly(px, py, vx[ptr[k]:ptr[k + 1]], vy[ptr[k]:ptr[k + 1]]) != 0:
                out[i] = k
                break
    return out


def st_bau_index(spatial: np.ndarray, time: np.ndarray, breaks: np.ndarray, ns: int) -> np.ndarray:
    """map_data_to_BAUs(ST), R/geometryfns.R:1419-1547: for BAU time point i the data with
    ``time >= t1 & time < t2`` (t2 = next BAU time; for the last one t2 = last(endTime) + 1, i.e. everything from t1 on,
    :1455-1470) are binned in space; ST BAUs are numbered space-fastest (:971-980).  Plain loop over the time points, as in
    the reference's lapply."""
    spatial = np.asarray(spatial, dtype=np.int64)
    time = np.asarray(time, dtype=np.float64)
    out = np.full(time.shape[0], -1, dtype=np.int64)
    nt = len(breaks)
    for i in range(nt):
        t1 = breaks[i]
        sel = (time >= t1) & (time < breaks[i + 1]) if i < nt - 1 else (time >= t1)
        sel &= spatial >= 0
        out[sel] = spatial[sel] + ns * i
    return out.astype(np.int32)


def bin_mean(bau: np.ndarray, cols: np.ndarray):
    """map_data_to_BAUs: drop NA, group by BAU, mean of every numeric column
    (R/geometryfns.R:1277-1302).  Rows are returned in ascending BAU index (the reference
    orders by the BAU-name string; parity is defined given BAU names whose sort order is
    the index order).  Means use R's two-pass mean (sum/n, then + mean of residuals)."""
    cols = np.asarray(cols, dtype=np.float64)
    if cols.ndim == 1:
        cols = cols[:, None]
    keep = bau >= 0
    b = bau[keep].astype(np.int64)
    v = cols[keep]
    order = np.argsort(b, kind="stable")
    b = b[order]
    v = v[order]
    uniq, start, cnt = np.unique(b, return_index=True, return_counts=True)
    sums = np.add.reduceat(v, start, axis=0) if len(b) else np.zeros((0, cols.shape[1]))
    mean1 = sums / cnt[:, None]
    grp = np.repeat(np.arange(len(uniq)), cnt)
    resid = v - mean1[grp]
    rs = np.add.reduceat(resid, start, axis=0) if len(b) else resid
    means = mean1 + rs / cnt[:, None]
    return uniq.astype(np.int32), means, cnt.astype(np.int32)


def BuildC(obs_bau: np.ndarray, nbau: int) -> sp.csr_matrix:
    """BuildC for points + sparseMatrix(i,j,x=1), R/geometryfns.R:1557-1568, R/SRE.R:476-480."""
    m = len(obs_bau)
    return sp.csr_matrix((np.ones(m), (np.arange(m), obs_bau)), shape=(m, nbau))

# --------------------------------------------------------------------------
# stats::optim Nelder-Mead and stats::uniroot, restated
# --------------------------------------------------------------------------

def nmmin(fn, par, maxit=100, reltol=1e-4, abstol=-np.inf, alpha=1.0, bet=0.5, gamm=2.0):
    """R's optim(method="Nelder-Mead") core (src/appl/optim.c:nmmin), as called at
    R/SREfit.R:618-621 with control=list(maxit=100, reltol=1e-4)."""
    big = 1.0e35
    B = np.array(par, dtype=np.float64).copy()
    n = len(B)
    if maxit <= 0:
        return B, fn(B), 0, 0
    P = np.zeros((n + 1, n + 2))
    f = fn(B)
    if not np.isfinite(f):
        raise RuntimeError("function cannot be evaluated at initial parameters")
    funcount = 1
    convtol = reltol * (abs(f) + reltol)
    n1 = n + 1
    C = n + 2
    P[n1 - 1, 0] = f
    P[:n, 0] = B
    L = 1
    size = 0.0
    step = 0.0
    for i in range(n):
        if 0.1 * abs(B[i]) > step:
            step = 0.1 * abs(B[i])
    if step == 0.0:
        step = 0.1
    for j in range(2, n1 + 1):
        P[:n, j - 1] = B
        trystep = step
        while P[j - 2, j - 1] == B[j - 2]:
            P[j - 2, j - 1] = B[j - 2] + trystep
            trystep *= 10
        size += trystep
    oldsize = size
    calcvert = True
    fail = 0
    while True:
        if calcvert:
            for j in range(n1):
                if j + 1 != L:
                    B = P[:n, j].copy()
                    f = fn(B)
                    if not np.isfinite(f):
                        f = big
                    funcount += 1
                    P[n1 - 1, j] = f
            calcvert = False
        VL = P[n1 - 1, L - 1]
        VH = VL
        H = L
        for j in range(1, n1 + 1):
            if j != L:
                f = P[n1 - 1, j - 1]
                if f < VL:
                    L = j
                    VL = f
                if f > VH:
                    H = j
                    VH = f
        if VH <= VL + convtol or VL <= abstol:
            break
        for i in range(n):
            temp = -P[i, H - 1]
            for j in range(n1):
                temp += P[i, j]
            P[i, C - 1] = temp / n
        B = (1.0 + alpha) * P[:n, C - 1] - alpha * P[:n, H - 1]
        f = fn(B)
        if not np.isfinite(f):
            f = big
        funcount += 1
        VR = f
        if VR < VL:
            P[n1 - 1, C - 1] = f
            for i in range(n):
                f = gamm * B[i] + (1 - gamm) * P[i, C - 1]
                P[i, C - 1] = B[i]
                B[i] = f
            f = fn(B)
            if not np.isfinite(f):
                f = big
            funcount += 1
            if f < VR:
                P[:n, H - 1] = B
                P[n1 - 1, H - 1] = f
            else:
                P[:n, H - 1] = P[:n, C - 1]
                P[n1 - 1, H - 1] = VR
        else:
            if VR < VH:
                P[:n, H - 1] = B
                P[n1 - 1, H - 1] = VR
            B = (1 - bet) * P[:n, H - 1] + bet * P[:n, C - 1]
            f = fn(B)
            if not np.isfinite(f):
                f = big
            funcount += 1
            if f < P[n1 - 1, H - 1]:
                P[:n, H - 1] = B
                P[n1 - 1, H - 1] = f
            elif VR >= VH:
                calcvert = True
                size = 0.0
                for j in range(n1):
                    if j + 1 != L:
                        for i in range(n):
                            P[i, j] = bet * (P[i, j] - P[i, L - 1]) + P[i, L - 1]
                            size += abs(P[i, j] - P[i, L - 1])
                if size < oldsize:
                    oldsize = size
                else:
                    fail = 10
                    break
        if not (funcount <= maxit):
            break
    fmin = P[n1 - 1, L - 1]
    X = P[:n, L - 1].copy()
    if funcount > maxit:
        fail = 1
    return X, fmin, funcount, fail


def zeroin(f, lower, upper, tol=np.finfo(np.float64).eps ** 0.25, maxiter=1000):
    """stats::uniroot -> R_zeroin2 (Brent), default tol=.Machine$double.eps^0.25, as called
    at R/SREfit.R:386-388.  Returns the root."""
    EPS = np.finfo(np.float64).eps
    fa = f(lower)
    fb = f(upper)
    xmax = np.finfo(np.float64).max
    fa = min(max(fa, -xmax), xmax)
    fb = min(max(fb, -xmax), xmax)
    if fa * fb > 0:
        raise RuntimeError("f() values at end points not of opposite sign")
    a, b = lower, upper
    c, fc = a, fa
    if fa == 0.0:
        return a
    if fb == 0.0:
        return b
    maxit = maxiter + 1
    while maxit > 0:
        maxit -= 1
        prev_step = b - a
        if abs(fc) < abs(fb):
            a, b, c = b, c, b
            fa, fb, fc = fb, fc, fb
        tol_act = 2 * EPS * abs(b) + tol / 2
        new_step = (c - b) / 2
        if abs(new_step) <= tol_act or fb == 0.0:
            return b
        if abs(prev_step) >= tol_act and abs(fa) > abs(fb):
            cb = c - b
            if a == c:
                t1 = fb / fa
                p = cb * t1
                q = 1.0 - t1
            else:
                q = fa / fc
                t1 = fb / fc
                t2 = fb / fa
                p = t2 * (cb * q * (q - t1) - (b - a) * (t1 - 1.0))
                q = (q - 1.0) * (t1 - 1.0) * (t2 - 1.0)
            if p > 0:
                q = -q
            else:
                p = -p
            if p < (0.75 * cb * q - abs(tol_act * q) / 2) and p < abs(prev_step * q / 2):
                new_step = p / q
        if abs(new_step) < tol_act:
            new_step = tol_act if new_step > 0 else -tol_act
        a, fa = b, fb
        b += new_step
        fb = f(b)
        if (fb > 0 and fc > 0) or (fb < 0 and fc < 0):
            c, fc = a, fa
    return b

# --------------------------------------------------------------------------
# linalg helpers (R/linalgfns.R:21-36) and Matrix::chol / chol2inv
# --------------------------------------------------------------------------

def tr(X):
    return float(np.sum(np.diag(X)))


def diag2(X, Y, symm=False):
    return np.sum(X * Y, axis=1) if symm else np.sum(X * Y.T, axis=1)


def chol_upper(A):
    """Matrix::chol -> upper R with A = R'R."""
    return sla.cholesky(np.asarray(A), lower=False)


def logdet(L):
    return 2.0 * float(np.sum(np.log(np.diag(L))))


def chol2inv(R):
    """chol2inv(R) = (R'R)^{-1} (LAPACK dpotri)."""
    inv, info = sla.lapack.dpotri(np.asarray(R), lower=0)
    if info != 0:
        raise np.linalg.LinAlgError("dpotri failed")
    inv = np.triu(inv)
    return inv + np.triu(inv, 1).T

# --------------------------------------------------------------------------
# The SRE model (A.3-A.7)
# --------------------------------------------------------------------------

@dataclass
class SREModel:
    """Plain-array restatement of the ``SRE`` S4 object (R/AllClass.R:133-175), point data only."""
    basis: object
    S0: sp.csr_matrix          # N x r (normalised)
    S: sp.csr_matrix           # m x r
    obs_bau: np.ndarray        # m (0-based BAU row of each binned observation)
    Z: np.ndarray              # m
    X: np.ndarray              # m x p
    X_BAU: np.ndarray          # N x p
    Ve: np.ndarray             # m (diag)
    Vfs: np.ndarray            # m (diag)
    fs_BAU: np.ndarray         # N
    D_basis: object
    K_type: str
    mu_eta: np.ndarray = None
    S_eta: np.ndarray = None
    Q_eta: np.ndarray = None
    Khat: np.ndarray = None
    Khat_inv: np.ndarray = None
    alphahat: np.ndarray = None
    sigma2fshat: float = None
    info_fit: dict = field(default_factory=dict)
    Vmat: object = None        # dense m x m Vfs when it is NOT diagonal (footprints sharing BAUs); Vfs then holds its diagonal

    @property
    def r(self):
        return self.S.shape[1]


def EM_initialise(Z, X, Ve, r):
    """R/SRE.R:704-731."""
    varZ = float(np.var(Z, ddof=1))
    K = np.eye(r) * (1.0 / (1.0 / varZ))
    Kinv = np.eye(r) * (1.0 / (1.0 / (1.0 / varZ)))      # solve(Diagonal(x)) = 1/x
    XtX = X.T @ X
    alpha = np.linalg.solve(XtX, X.T @ Z)
    return dict(mu_eta=np.zeros(r), S_eta=np.eye(r), Q_eta=np.eye(r), Khat=K, Khat_inv=Kinv,
                alphahat=alpha, sigma2fshat=float(np.mean(Ve)) / 4.0)


def BuildC_polygons(ptr: np.ndarray, vx: np.ndarray, vy: np.ndarray, bau_xy: np.ndarray, wts: Optional[np.ndarray] = None):
    """BuildC(data = "SpatialPolygons"), R/geometryfns.R:1572-1616: for every observation polygon i (in order) the BAUs whose
    centroid it contains (over(BAU_as_points, this_poly) == 1: interior, edge and vertex all count), weight BAUs$wts.
    Returns the triplets (i_idx, j_idx, x_idx), 0-based.  (The fallback for polygons that contain no centroid, :1593-1598, is not
    restated: such inputs are refused on the device path.)"""
    bau_xy = np.asarray(bau_xy, dtype=np.float64)
    i_idx, j_idx, x_idx = [], [], []
    for i in range(len(ptr) - 1):
        px, py = vx[ptr[i]:ptr[i + 1]], vy[ptr[i]:ptr[i + 1]]
        cand = np.nonzero((bau_xy[:, 0] >= px.min()) & (bau_xy[:, 0] <= px.max()) &
                          (bau_xy[:, 1] >= py.min()) & (bau_xy[:, 1] <= py.max()))[0]        # bounding box: pre-filter only
        overlap = [k for k in cand if in_poly(bau_xy[k, 0], bau_xy[k, 1], px, py) != 0]
        i_idx += [i] * len(overlap)
        j_idx += overlap
        x_idx += [1.0 if wts is None else float(wts[k]) for k in overlap]
    return np.asarray(i_idx, dtype=np.int64), np.asarray(j_idx, dtype=np.int64), np.asarray(x_idx, dtype=np.float64)


def SRE(basis, baus_coords: np.ndarray, fs_BAU: np.ndarray, obs_bau: np.ndarray, Z: np.ndarray,
        std: np.ndarray, X: Optional[np.ndarray] = None, X_BAU: Optional[np.ndarray] = None,
        K_type: str = "block-exponential", normalise_basis: bool = True, fast_eval: bool = False,
        C: Optional[sp.spmatrix] = None, allow_overlap: bool = False) -> SREModel:
    """SRE() for binned point data, R/SRE.R:395-701.

    Inputs are the *binned* observations (map_data_to_BAUs output): BAU row, mean z, mean std.
    X defaults to the intercept-only model matrix (f = z ~ 1).

    ``C`` (m x N incidence matrix from BuildC_polygons, un-normalised): areal data.  Rows are divided by their sums
    (normalise_wts = TRUE, :483-490), S = Cmat %*% S0 (:502), Vfs = tcrossprod(Cmat %*% Diagonal(sqrt(fs))) (:498), which must be
    diagonal here (supports that share no BAU), and X = Cmat %*% X_BAU (BAU covariates averaged over the support).
    ``obs_bau`` is ignored then.
    """
    if C is not None:
        return _SRE_areal(basis, baus_coords, fs_BAU, sp.csr_matrix(C), Z, std, X_BAU, K_type, normalise_basis, allow_overlap)
    N = baus_coords.shape[0]
    if fast_eval and not isinstance(basis, TensorP_Basis) and basis.type == "bisquare":
        S0 = eval_basis_fast(basis, baus_coords)
    else:
        S0 = eval_basis(basis, baus_coords)                       # :412
    if normalise_basis:
        S0 = normalise_rows(S0)                                   # :416-421
    D_basis = BuildD(basis)                                       # :424
    m = len(obs_bau)
    Z = np.asarray(Z, dtype=np.float64)
    Ve = np.asarray(std, dtype=np.float64) ** 2                   # :468 (x^2 == x*x)
    sq = np.sqrt(fs_BAU[obs_bau])
    Vfs = sq * sq                                                 # :498 tcrossprod(C diag(sqrt(fs)))
    S = S0[obs_bau]                                               # :502 Cmat %*% S0
    if X is None:
        X = np.ones((m, 1))
    if X_BAU is None:
        X_BAU = np.ones((N, 1))
    mdl = SREModel(basis=basis, S0=S0, S=S.tocsr(), obs_bau=np.asarray(obs_bau), Z=Z, X=X, X_BAU=X_BAU,
                   Ve=Ve, Vfs=Vfs, fs_BAU=np.asarray(fs_BAU, dtype=np.float64), D_basis=D_basis, K_type=K_type)
    init = EM_initialise(Z, X, Ve, S.shape[1])
    for k, v in init.items():
        setattr(mdl, k, v)
    return mdl


def _SRE_areal(basis, baus_coords, fs_BAU, C, Z, std, X_BAU, K_type, normalise_basis, allow_overlap=False) -> SREModel:
    N = baus_coords.shape[0]
    S0 = eval_basis(basis, baus_coords)
    if normalise_basis:
        S0 = normalise_rows(S0)
    D_basis = BuildD(basis)
    m = C.shape[0]
    Cmat = sp.diags(1.0 / np.asarray(C.sum(axis=1)).ravel()) @ C          # Cmat / rowSums(Cmat), :487-489
    Cmat = sp.csr_matrix(Cmat)
    Z = np.asarray(Z, dtype=np.float64)
    Ve = np.asarray(std, dtype=np.float64) ** 2
    fs_BAU = np.asarray(fs_BAU, dtype=np.float64)
    half = sp.csr_matrix(Cmat @ sp.diags(np.sqrt(fs_BAU)))
    V = sp.csr_matrix(half @ half.T)                                      # tcrossprod(Cmat %*% Diagonal(sqrt(fs))), :498
    off = V - sp.diags(V.diagonal())
    Vmat = None
    if off.nnz and abs(off).max() > 0:
        if not allow_overlap:
            raise ValueError("supports share BAUs: Vfs is not diagonal")
        Vmat = np.asarray(V.todense())                                    # the general branches of the EM (R/SREfit.R:186-189,240-252,296-328)
    Vfs = V.diagonal()
    S = sp.csr_matrix(Cmat @ S0)
    S.sort_indices()
    if X_BAU is None:
        X_BAU = np.ones((N, 1))
    X = np.asarray(Cmat @ X_BAU)
    mdl = SREModel(basis=basis, S0=S0, S=S, obs_bau=np.zeros(0, dtype=np.int64), Z=Z, X=X, X_BAU=X_BAU,
                   Ve=Ve, Vfs=Vfs, fs_BAU=fs_BAU, D_basis=D_basis, K_type=K_type)
    mdl.Cmat = Cmat
    mdl.Vmat = Vmat
    init = EM_initialise(Z, X, Ve, S.shape[1])
    for k, v in init.items():
        setattr(mdl, k, v)
    return mdl


def _gram(S: sp.csr_matrix, w: np.ndarray) -> np.ndarray:
    """crossprod(diag(sqrt(w)) S) = S' diag(w) S, dense r x r."""
    Sw = sp.diags(np.sqrt(w)) @ S
    return np.asarray((Sw.T @ Sw).todense())


def loglik(M: SREModel) -> float:
    """.loglik.ind, R/SREfit.R:174-220."""
    S, K = M.S, M.Khat
    chol_K = chol_upper(K)
    Kinv = chol2inv(chol_K)
    resid = M.Z - M.X @ M.alphahat
    N = len(M.Z)
    if M.Vmat is not None:
        return _loglik_general(M, K, Kinv, resid)
    D = M.sigma2fshat * M.Vfs + M.Ve
    cholD = np.sqrt(D)
    cholDinv = 1.0 / cholD
    S_Dinv_S = _gram(S, cholDinv * cholDinv)                      # crossprod(cholDinvT %*% S)
    R = chol_upper(Kinv + S_Dinv_S)
    sign, logdetK = np.linalg.slogdet(K)                          # determinant(K, logarithm=TRUE)
    log_det_SigmaZ = logdet(R) + logdetK + 2.0 * float(np.sum(np.log(cholD)))
    rDinv = (cholDinv * resid) * cholDinv
    v = S.T @ rDinv                                               # (rDinv %*% S)
    x = sla.solve_triangular(R, v, trans="T", lower=False)        # v' R^{-1}
    quad = float(np.sum((cholDinv * resid) ** 2)) - float(x @ x)
    return -0.5 * N * math.log(2 * math.pi) - 0.5 * log_det_SigmaZ - 0.5 * quad


def _loglik_general(M: SREModel, K, Kinv, resid) -> float:
    """.loglik.ind with a non-diagonal D: cholD <- chol(D); cholDinvT <- t(solve(cholD)), R/SREfit.R:186-189 and the lines that follow."""
    N = len(M.Z)
    D = M.sigma2fshat * M.Vmat + np.diag(M.Ve)
    cholD = chol_upper(D)                                             # upper factor, D = cholD' cholD
    cholDinvT = sla.solve_triangular(cholD, np.eye(N), lower=False).T
    A = cholDinvT @ np.asarray(M.S.todense())
    S_Dinv_S = A.T @ A                                                # crossprod(cholDinvT %*% S)
    R = chol_upper(Kinv + S_Dinv_S)
    sign, logdetK = np.linalg.slogdet(K)
    log_det_SigmaZ = logdet(R) + logdetK + logdet(cholD)
    wr = cholDinvT @ resid
    rDinv = wr @ cholDinvT                                            # crossprod(cholDinvT %*% resid, cholDinvT)
    v = M.S.T @ rDinv
    x = sla.solve_triangular(R, v, trans="T", lower=False)
    quad = float(wr @ wr) - float(x @ x)
    return -0.5 * N * math.log(2 * math.pi) - 0.5 * log_det_SigmaZ - 0.5 * quad


def Estep(M: SREModel) -> None:
    """.SRE.Estep.ind, R/SREfit.R:232-260."""
    if M.Vmat is not None:                                            # :244-247
        N = len(M.Z)
        D = M.sigma2fshat * M.Vmat + np.diag(M.Ve)
        cholD = chol_upper(D)
        cholDinv = sla.solve_triangular(cholD, np.eye(N), lower=False)
        Dinv = chol2inv(cholD)
        A = cholDinv.T @ np.asarray(M.S.todense())
        Q_eta = A.T @ A + M.Khat_inv
        S_eta = chol2inv(chol_upper(Q_eta))
        mu_eta = S_eta @ (M.S.T @ (Dinv @ (M.Z - M.X @ M.alphahat)))
        M.mu_eta, M.S_eta, M.Q_eta = mu_eta, S_eta, Q_eta
        return
    D = M.sigma2fshat * M.Vfs + M.Ve
    Dinv = 1.0 / D
    cholDinv = 1.0 / np.sqrt(D)
    Q_eta = _gram(M.S, cholDinv * cholDinv) + M.Khat_inv
    S_eta = chol2inv(chol_upper(Q_eta))
    mu_eta = S_eta @ (M.S.T @ (Dinv * (M.Z - M.X @ M.alphahat)))
    M.mu_eta, M.S_eta, M.Q_eta = mu_eta, S_eta, Q_eta


def _res_index_sets(basis):
    res = basis.res
    return [np.nonzero(res == i)[0] for i in range(1, int(res.max()) + 1)]


def update_K(M: SREModel, lam=0.0) -> np.ndarray:
    """.update_K / .regularise_K, R/SREfit.R:441-694."""
    S_eta, mu_eta = M.S_eta, M.mu_eta
    E = S_eta + np.outer(mu_eta, mu_eta)
    if M.K_type == "unstructured":
        lam_arr = np.atleast_1d(np.asarray(lam, dtype=np.float64))
        if np.any(lam_arr > 0):                                   # :669-686
            if lam_arr.size == 1:
                reg = np.eye(E.shape[0]) * lam_arr[0]
            else:
                reg = np.diag(lam_arr[np.asarray(M.basis.res) - 1])
            Q = chol2inv(chol_upper(E)) + reg
            return chol2inv(chol_upper(Q))
        return E                                                  # :689
    if M.K_type != "block-exponential":
        raise ValueError(M.K_type)
    tensor = isinstance(M.basis, TensorP_Basis)
    idxs = _res_index_sets(M.basis)
    nres = len(idxs)
    eta2 = [E[np.ix_(ix, ix)] for ix in idxs]                     # :459-464
    omega = []
    for i, ix in enumerate(idxs):                                 # :467-482
        Ki = M.Khat[np.ix_(ix, ix)] / M.Khat[ix[0], ix[0]]
        Ki_inv = chol2inv(chol_upper(Ki))
        omega.append(len(ix) / float(np.sum(diag2(Ki_inv, eta2[i]))))
    if tensor:
        Ds, Dt = M.D_basis["Basis1"], M.D_basis["Basis2"][0]
        nsp = [int(np.sum(M.basis.Basis1.res == i + 1)) for i in range(nres)]
    else:
        Ds = M.D_basis

    def f_tau(tau, i):                                            # :485-529
        tau = np.atleast_1d(tau)
        if np.any(tau <= 1e-10):
            return np.inf
        if tensor:
            Ki1 = np.exp(-Dt / tau[1])
            Ki2 = np.exp(-Ds[i] / tau[0])
            Qi1 = chol2inv(chol_upper(Ki1))
            Qi2 = chol2inv(chol_upper(Ki2))
            Ki_inv = np.kron(Qi1, Qi2)
            R1 = chol_upper(Qi1)
            R2 = chol_upper(Qi2)
            det_part = 0.5 * (R2.shape[0] * logdet(R1) + R1.shape[0] * logdet(R2))
            return -float(det_part - omega[i] / 2 * np.sum(diag2(Ki_inv, eta2[i], symm=True)))
        Ki = np.exp(-Ds[i] / tau[0])
        Ki_inv = chol2inv(chol_upper(Ki))
        sign, ld = np.linalg.slogdet(Ki_inv)
        return -float(0.5 * ld - omega[i] / 2 * np.sum(diag2(Ki_inv, eta2[i], symm=True)))

    max_l = float(np.max(Ds[0]))                                  # :584
    taus = []
    for i, ix in enumerate(idxs):                                 # :588-622
        Ki = M.Khat[np.ix_(ix, ix)] / M.Khat[ix[0], ix[0]]
        with np.errstate(divide="ignore", invalid="ignore"):
            if tensor:
                p0 = max(-Ds[i][0, 1] / np.log(Ki[0, 1]), 1e-9) if nsp[i] > 1 else 1e-9
                p1 = max(-Dt[0, 1] / np.log(Ki[0, nsp[i]]), 1e-9)
                if p1 == 1e-9:
                    p1 = 1.0
                par = [p0, p1]
            else:
                p0 = max(-Ds[i][0, 1] / np.log(Ki[0, 1]), 1e-9) if len(ix) > 1 else 1e-9
                par = [p0]
        if par[0] == 1e-9:
            par[0] = max_l / 10
        x, _, _, _ = nmmin(lambda t, i=i: f_tau(t, i), par, maxit=100, reltol=1e-4)
        taus.append(x)
    K = np.zeros_like(M.Khat)
    for i, ix in enumerate(idxs):                                 # :625-649 (bdiag + reverse_permute)
        if tensor:
            Ki = np.kron(np.exp(-Dt / taus[i][1]), np.exp(-Ds[i] / taus[i][0])) / omega[i]
        else:
            Ki = np.exp(-Ds[i] / taus[i][0]) / omega[i]
        K[np.ix_(ix, ix)] = Ki
    M.info_fit["tau"] = [t.copy() for t in taus]
    M.info_fit["omega"] = list(omega)
    return K


def Mstep(M: SREModel, lam=0.0, known_sigma2fs=None) -> None:
    """.SRE.Mstep.ind for diagonal Ve/Vfs (point data), R/SREfit.R:263-438."""
    mu_eta, S_eta = M.mu_eta, M.S_eta
    sigma2fs = M.sigma2fshat
    K = update_K(M, lam)
    Khat_inv = chol2inv(chol_upper(K))
    a, b = M.Ve, M.Vfs
    if M.Vmat is not None:
        return _Mstep_general(M, K, Khat_inv, known_sigma2fs)
    homoscedastic = bool(np.all(a == a[0]) and np.all(b == b[0]))
    est = known_sigma2fs is None
    sigma2fs_new = sigma2fs if est else known_sigma2fs
    S, X, Z = M.S, M.X, M.Z
    t = S @ mu_eta
    alpha = M.alphahat
    if not np.all(b == 0):
        R_eta = chol_upper(S_eta + np.outer(mu_eta, mu_eta))      # :332
        S_R = np.asarray((S @ R_eta.T))                           # :333 (dense m x r in the reference)
        Omega_diag1 = np.sum(S_R * S_R, axis=1)                   # :334

        def gls(Dinv):
            XtD = X.T * Dinv
            return np.linalg.solve(XtD @ X, XtD @ (Z - t))

        def J(s2):                                                # :335-356
            if s2 < 0:
                return np.inf
            D = s2 * b + a
            Dinv = 1.0 / D
            DinvV = Dinv * b
            al = gls(Dinv)
            resid = Z - X @ al
            Om = Omega_diag1 - 2 * (t * resid) + resid * resid
            return -(-0.5 * np.sum(DinvV) + 0.5 * np.sum(DinvV * Dinv * Om))

        if not homoscedastic:
            if est:
                amp = 10.0
                OK = False
                while not OK:                                     # :368-380
                    amp *= 10
                    if not (np.sign(J(sigma2fs / amp)) == np.sign(J(sigma2fs * amp))):
                        OK = True
                    if amp > 1e9:
                        OK = True
                if amp > 1e9:
                    sigma2fs_new = 0.0
                else:
                    sigma2fs_new = zeroin(J, sigma2fs / amp, sigma2fs * amp)
            alpha = gls(1.0 / (sigma2fs_new * b + a))             # :391-400
        else:
            alpha = np.linalg.solve(X.T @ X, X.T @ (Z - t))       # :404-405
            resid = Z - X @ alpha
            if est:
                Om = Omega_diag1 - 2 * (t * resid) + resid * resid
                sigma2fs_new = 1.0 / b[0] * (np.sum(Om) / len(Z) - a[0])
                if sigma2fs_new < 0:
                    sigma2fs_new = 0.0
    else:                                                         # :424-428
        XtD = X.T * (1.0 / a)
        alpha = np.linalg.solve(XtD @ X, XtD @ (Z - t))
        if est:
            sigma2fs_new = 0.0
    M.Khat, M.Khat_inv, M.alphahat, M.sigma2fshat = K, Khat_inv, alpha, float(sigma2fs_new)


def _Mstep_general(M: SREModel, K, Khat_inv, known_sigma2fs) -> None:
    """.SRE.Mstep.ind when Vfs is not diagonal (diagonal_mats FALSE, hence not homoscedastic): the general J, R/SREfit.R:296-328,
    the bracket search :366-389 and the GLS step :391-400, dense algebra (toy sizes)."""
    mu_eta, S_eta = M.mu_eta, M.S_eta
    sigma2fs = M.sigma2fshat
    est = known_sigma2fs is None
    sigma2fs_new = sigma2fs if est else known_sigma2fs
    S, X, Z, V = np.asarray(M.S.todense()), M.X, M.Z, M.Vmat
    Ve = np.diag(M.Ve)
    E = S_eta + np.outer(mu_eta, mu_eta)

    def gls(Dinv):
        return np.linalg.solve(X.T @ Dinv @ X, X.T @ Dinv @ (Z - S @ mu_eta))

    def J(s2):
        if s2 < 0:
            return np.inf
        D = s2 * V + Ve
        Dinv = chol2inv(chol_upper(D))
        DinvV = Dinv @ V
        DinvVDinv = Dinv @ V @ Dinv
        al = gls(Dinv)
        resid = Z - X @ al
        Dinvr = DinvVDinv @ resid
        DinvS = DinvVDinv @ S
        tr1 = np.trace(DinvV)
        tr2 = (np.sum((DinvS @ E) * S) - 2.0 * np.sum((DinvS @ mu_eta) * resid) + np.sum(Dinvr * resid))
        return -(-0.5 * tr1 + 0.5 * tr2)

    alpha = M.alphahat
    if not np.all(np.diag(V) == 0):
        if est:
            amp = 10.0
            OK = False
            while not OK:
                amp *= 10
                if not (np.sign(J(sigma2fs / amp)) == np.sign(J(sigma2fs * amp))):
                    OK = True
                if amp > 1e9:
                    OK = True
            sigma2fs_new = 0.0 if amp > 1e9 else zeroin(J, sigma2fs / amp, sigma2fs * amp)
        alpha = gls(chol2inv(chol_upper(sigma2fs_new * V + Ve)))
    else:
        alpha = gls(np.diag(1.0 / M.Ve))
        if est:
            sigma2fs_new = 0.0
    M.Khat, M.Khat_inv, M.alphahat, M.sigma2fshat = K, Khat_inv, alpha, float(sigma2fs_new)


def SRE_fit(M: SREModel, n_EM: int = 100, tol: float = 0.01, lam=0.0, known_sigma2fs=None) -> SREModel:
    """.EM_fit, R/SREfit.R:78-160."""
    if known_sigma2fs is not None:
        M.sigma2fshat = known_sigma2fs                            # R/SREfit.R:19
    llk = np.zeros(n_EM)
    i = 0
    for i in range(n_EM):
        llk[i] = loglik(M)
        Estep(M)
        Mstep(M, lam, known_sigma2fs)
        if i > 0 and abs(llk[i] - llk[i - 1]) < tol:
            break
    nit = i + 1
    M.info_fit.update(method="EM", num_iterations=nit, converged=int(nit != n_EM),
                      llk=llk[:nit].copy(), sigma2fshat_equal_0=int(M.sigma2fshat == 0))
    return M


def batch_compute_var(S0: sp.csr_matrix, Cov: np.ndarray) -> np.ndarray:
    """.batch_compute_var(fs_in_process=FALSE), R/SREutils.R:245-305: rowSums((S0 Cov) * S0)."""
    out = np.empty(S0.shape[0])
    for i0 in range(0, S0.shape[0], 10000):
        blk = S0[i0:i0 + 10000]
        out[i0:i0 + 10000] = np.asarray(blk.multiply(blk @ Cov).sum(axis=1)).ravel()
    return out


def predict(M: SREModel, obs_fs: bool = False):
    """.SRE.predict_EM at the BAUs (predict_BAUs=TRUE, covariances=FALSE), R/SREpredict.R:243-375.

    obs_fs=FALSE and sigma2fs>0 uses the closed form of the joint (eta, xi) solve that is
    valid for point data; predict_general() below is the literal joint
    solve and the two are checked equal in the CPU tests.
    """
    S0, alpha, s2 = M.S0, M.alphahat, M.sigma2fshat
    N = S0.shape[0]
    Xa = M.X_BAU @ alpha
    if obs_fs or s2 == 0:
        mu = Xa + S0 @ M.mu_eta
        var = batch_compute_var(S0, M.S_eta)
        return mu, var, np.sqrt(var)
    if getattr(M, "Cmat", None) is not None:                      # areal data: no point-data shortcut -- the literal joint solve
        return predict_general(M)
    A = np.zeros(N)
    np.add.at(A, M.obs_bau, 1.0 / M.Ve)                           # C' Ve^-1 C (diagonal)
    Q = 1.0 / (s2 * M.fs_BAU)
    ytil = np.zeros(N)
    np.add.at(ytil, M.obs_bau, (M.Z - M.X @ alpha) / M.Ve)
    # eta | Z at the FINAL parameters (xi integrated out)
    Dinv = 1.0 / (s2 * M.Vfs + M.Ve)
    Qeta = _gram(M.S, Dinv) + M.Khat_inv
    Sig = chol2inv(chol_upper(Qeta))
    mu_t = Sig @ (M.S.T @ (Dinv * (M.Z - M.X @ alpha)))
    q = batch_compute_var(S0, Sig)
    a = A / (A + Q)
    var = (1 - a) ** 2 * q + 1.0 / (A + Q)
    smu = S0 @ mu_t
    xi = (ytil - A * smu) / (A + Q)
    mu = Xa + smu + xi
    return mu, var, np.sqrt(var)


def predict_general(M: SREModel):
    """Literal Case-2 joint solve (R/SREpredict.R:308-333) with dense algebra; toy sizes only."""
    S0 = np.asarray(M.S0.todense())
    N, r = S0.shape
    s2 = M.sigma2fshat
    Cm = getattr(M, "Cmat", None)                                 # areal data: the row-normalised incidence matrix
    C = np.asarray(BuildC(M.obs_bau, N).todense()) if Cm is None else np.asarray(Cm.todense())
    LAMBDAinv = np.zeros((r + N, r + N))
    LAMBDAinv[:r, :r] = M.Khat_inv
    LAMBDAinv[r:, r:] = np.diag(1.0 / (s2 * M.fs_BAU))
    PI = np.hstack([S0, np.eye(N)])
    tCVC = C.T @ np.diag(1.0 / M.Ve) @ C
    Qx = PI.T @ tCVC @ PI + LAMBDAinv
    ybar = PI.T @ C.T @ ((M.Z - M.X @ M.alphahat) / M.Ve)
    Cov = np.linalg.inv(Qx)
    x_mean = Cov @ ybar
    var = (np.sum((S0 @ Cov[:r, :r]) * S0, axis=1) + np.diag(Cov)[r:] + 2 * np.sum(Cov[r:, :r] * S0, axis=1))
    mu = M.X_BAU @ M.alphahat + PI @ x_mean
    return mu, var, np.sqrt(var)


def predict_polygons(M: SREModel, CP: sp.spmatrix, obs_fs: bool = False):
    """predict(newdata = <polygons>), R/SREpredict.R:365-419: CP = row-normalised incidence of the prediction polygons over the BAUs.
    mu = CP %*% mu_BAU; obs_fs | sigma2fs == 0: var = diag2(CPM %*% Cov, t(CPM)) with CPM = CP %*% S0, Cov = S_eta (:385-391);
    otherwise CPM = CP %*% [S0 I] against the full joint covariance Qx^-1 (:392-404) -- literal dense solve, toy sizes only."""
    CP = sp.csr_matrix(CP)
    s2 = M.sigma2fshat
    if obs_fs or s2 == 0:
        mu_b = M.X_BAU @ M.alphahat + M.S0 @ M.mu_eta
        CPM = sp.csr_matrix(CP @ M.S0)
        var = np.asarray(CPM.multiply(CPM @ M.S_eta).sum(axis=1)).ravel()
        return CP @ mu_b, var, np.sqrt(var)
    S0 = np.asarray(M.S0.todense())
    N, r = S0.shape
    Cm = getattr(M, "Cmat", None)
    C = np.asarray(BuildC(M.obs_bau, N).todense()) if Cm is None else np.asarray(Cm.todense())
    LAMBDAinv = np.zeros((r + N, r + N))
    LAMBDAinv[:r, :r] = M.Khat_inv
    LAMBDAinv[r:, r:] = np.diag(1.0 / (s2 * M.fs_BAU))
    PI = np.hstack([S0, np.eye(N)])
    Qx = PI.T @ (C.T @ np.diag(1.0 / M.Ve) @ C) @ PI + LAMBDAinv
    ybar = PI.T @ C.T @ ((M.Z - M.X @ M.alphahat) / M.Ve)
    Cov = np.linalg.inv(Qx)
    mu_b = M.X_BAU @ M.alphahat + PI @ (Cov @ ybar)
    CPM = np.asarray(CP.todense()) @ PI
    var = np.sum((CPM @ Cov) * CPM, axis=1)
    return CP @ mu_b, var, np.sqrt(var)

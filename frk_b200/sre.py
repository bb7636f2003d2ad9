"""SRE(), SRE.fit(method = "EM") and predict() -- host mirror of R/SRE.R, R/SREfit.R and
R/SREpredict.R for Gaussian data with identity link.

As in the reference the host keeps the control flow (the EM loop, Nelder-Mead on tau, the bracket
search + uniroot on sigma2fs, p x p solves for the fixed effects) and r-sized bookkeeping; every
computation that touches observations, BAUs or an r x r matrix is a libfrkb200 kernel reached
through the C ABI.  With ``torch.distributed`` initialised the BAUs (and the observations that fall
in them) are row-partitioned over the ranks; the only large exchange is one fp64 all-reduce of
``[S'D^-1 S | S'D^-1 (z - X alpha) | 2 scalars]`` per EM iteration.

Arithmetic that the reference computes twice per iteration with identical inputs is computed once:
``.loglik.ind`` (R/SREfit.R:174-220) and ``.SRE.Estep.ind`` (:232-260) both form S'D^-1 S and factor
K^-1 + S'D^-1 S (the reference's two K^-1 are the same ``chol2inv(chol(Khat))``), and the M-step's
``Khat_inv`` factorisation (:273) supplies the next iteration's ``chol(K)`` / ``log|K|``.
"""
from __future__ import annotations

import ctypes as C
import math
import re
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from ._lib import FRKError, call, c_i64, vp
from .basis import Basis, BuildD, DeviceCSR, TensorP_Basis, eval_basis_device
from .geometry import GridBAUs, PointBAUs, SpatialPointsDataFrame, map_data_to_BAUs
from .optim import nmmin, zeroin
from .runtime import Device, get_device, partition_rows

_DBL4 = C.c_double * 4


# ---------------------------------------------------------------------------------------
# small helpers around the C ABI
# ---------------------------------------------------------------------------------------

def _alpha_ptr(alpha: np.ndarray):
    a = np.ascontiguousarray(alpha, dtype=np.float64)
    return a, vp(a.ctypes.data)


def _vec_stats(dev: Device, t, n: int, shift: float = 0.0):
    out = _DBL4()
    call("frk_vec_stats", dev.ctx, int(n), dev.ptr(t), float(shift), out)
    return np.array(out[:], dtype=np.float64)


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


@dataclass
class SREModel:
    """The ``SRE`` S4 object (R/AllClass.R:133-175).  m- and N-sized slots live in HBM (this rank's
    rows only); r x r slots live in HBM and are replicated on every rank; alphahat / sigma2fshat /
    info_fit are host values."""
    dev: Device
    f: str
    basis: object
    BAUs: object
    K_type: str
    r: int
    p: int
    N: int                      # BAUs (all ranks)
    lo: int                     # this rank owns BAU rows [lo, hi)
    hi: int
    m: int                      # binned observations on this rank
    m_total: int
    S0: DeviceCSR               # (hi-lo) x r
    S: DeviceCSR                # m x r
    obs_bau: object             # int32[m], relative to lo
    Z: object
    X: object                   # m x p column-major
    X_BAU: object               # (hi-lo) x p column-major
    Ve: object
    Vfs: object
    fs_BAU: object
    D_basis: object
    homoscedastic: bool = False
    Ve0: float = 0.0
    Vfs0: float = 0.0
    all_vfs_zero: bool = False
    # r-sized state
    mu_eta: object = None
    S_eta: object = None
    Q_eta: object = None
    Khat: object = None
    Khat_inv: object = None
    alphahat: np.ndarray = None
    sigma2fshat: float = None
    info_fit: dict = field(default_factory=dict)
    # scratch (device)
    _acc: object = None
    _w1: object = None
    _w2: object = None
    _logdetK: float = None
    _t: object = None
    _q: object = None
    _vec: object = None
    _blocks: dict = field(default_factory=dict)
    h2d_bytes: int = 0
    d2h_bytes: int = 0

    # ---- host copies of the r-sized slots (what the R shim writes back into the S4 object) ----
    def get(self, name: str) -> np.ndarray:
        t = getattr(self, name)
        a = self.dev.download(t)
        self.d2h_bytes += a.nbytes
        if a.size == self.r * self.r:
            return a.reshape(self.r, self.r).T.copy()
        return a


# ---------------------------------------------------------------------------------------
# SRE()
# ---------------------------------------------------------------------------------------

def SRE(f: str, data: SpatialPointsDataFrame, basis, BAUs, est_error: bool = False, average_in_BAU: bool = True,
        K_type: str = "block-exponential", normalise_basis: bool = True, dev: Optional[Device] = None) -> SREModel:
    """SRE(f, data, basis, BAUs, est_error, average_in_BAU, ..., K_type, normalise_basis), R/SRE.R:318-701
    for one point-referenced Gaussian data set.

    Steps on the device (file:line of what each replaces): BAU centroids -> eval_basis -> row
    normalisation (R/SRE.R:412-421); map_data_to_BAUs (R/geometryfns.R:1222-1320); BuildC + Cmat %*% S0
    as a CSR row gather (R/SRE.R:472-502); Ve, Vfs (:468,:498); .EM_initialise (:704-731)."""
    dev = dev or get_device()
    K_type = K_type.lower()
    if K_type not in ("block-exponential", "unstructured"):
        raise ValueError("K_type needs to be \"block-exponential\" or \"unstructured\" for method = \"EM\"")   # check_args.R:184-187
    if est_error:
        raise FRKError("est_error=TRUE needs the gstat variogram fit (host, third-party; out of the hot-path scope): "
                       "supply a 'std' column and est_error=FALSE")
    if not average_in_BAU:
        raise FRKError("average_in_BAU=FALSE gives several observations per BAU (non-diagonal C'C): not on the device path yet")
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
    t0 = time.time()
    r = basis.n
    N = len(BAUs)
    lo, hi = partition_rows(N, dev.world, dev.rank)
    nloc = hi - lo
    h2d = 0

    # ---- S0 = eval_basis(basis, coordinates(BAUs)); normalise rows ------------------------
    if isinstance(BAUs, GridBAUs) and not isinstance(basis, TensorP_Basis):
        d_xg = dev.upload(BAUs.xgrid)
        d_yg = dev.upload(BAUs.ygrid)
        h2d += 8 * (BAUs.nx + BAUs.ny)
        d_s = dev.empty(max(2 * nloc, 1))
        call("frk_grid_centroids", dev.ctx, lo, hi, BAUs.nx, dev.ptr(d_xg), dev.ptr(d_yg), dev.ptr(d_s))
        S0 = eval_basis_device(basis, (d_s, nloc, 2), dev, normalise=normalise_basis)
        del d_s
    else:
        cc = BAUs.centroids(lo, hi)
        h2d += cc.nbytes
        S0 = eval_basis_device(basis, cc, dev, normalise=normalise_basis)

    # ---- bin the observations into this rank's BAUs ------------------------------------------
    cols = [resp, "std"]
    binned = map_data_to_BAUs(data, BAUs, cols, dev, row_range=(lo, hi))
    h2d += data.coords[:, :2].nbytes + 8 * len(data) * len(cols)
    m, ld = binned["m"], binned["ld"]
    obs_bau = binned["obs_bau"][:max(m, 1)]
    Z = binned["means"][0:max(m, 1)]                             # views into the binned table (column-major, ld = n)
    Ve = dev.empty(max(m, 1))
    call("frk_square", dev.ctx, m, dev.ptr(binned["means"], ld), dev.ptr(Ve))      # Ve = diag(std^2), R/SRE.R:468
    m_total = int(dev.allreduce_host(np.array([float(m)]))[0])
    if m_total == 0:
        raise FRKError("no observation falls inside a BAU")

    # ---- C_Z (one 1 per row) -> S = C_Z S0, Vfs = fs[bau], X = model matrix at the observed BAUs ----
    rowptr = dev.empty(m + 1, "i4")
    nnz = c_i64(0)
    call("frk_csr_gather_count", dev.ctx, m, dev.ptr(obs_bau), dev.ptr(S0.rowptr), dev.ptr(rowptr), C.byref(nnz))
    col = dev.empty(max(nnz.value, 1), "i4")
    val = dev.empty(max(nnz.value, 1))
    call("frk_csr_gather_fill", dev.ctx, m, dev.ptr(obs_bau), dev.ptr(S0.rowptr), dev.ptr(S0.col), dev.ptr(S0.val),
         dev.ptr(rowptr), dev.ptr(col), dev.ptr(val))
    S = DeviceCSR(m, r, rowptr, col, val, nnz.value)

    # BAU-level fields for this rank: fs and covariates (coordinates count as BAU fields)
    fs_host = np.ascontiguousarray(BAUs.fields["fs"][lo:hi], dtype=np.float64)
    fs_BAU = dev.upload(fs_host)
    h2d += fs_host.nbytes
    xcols = []
    if intercept:
        xcols.append(np.ones(nloc))
    for c in covs:
        if c in BAUs.fields:
            xcols.append(np.asarray(BAUs.fields[c][lo:hi], dtype=np.float64))
        else:
            raise ValueError(f"covariate {c!r} is not a field of the BAUs (covariates must be known at every BAU)")
    p = len(xcols)
    if p == 0:
        raise FRKError("a model without fixed effects is not on the device path (at least an intercept is needed)")
    X_BAU_host = np.column_stack(xcols) if nloc else np.zeros((0, p))
    X_BAU = dev.upload(X_BAU_host)
    h2d += X_BAU_host.nbytes
    Vfs = dev.empty(max(m, 1))                                    # Vfs = diag(fs[bau]), R/SRE.R:498
    call("frk_gather_rows", dev.ctx, m, 1, dev.ptr(obs_bau), dev.ptr(fs_BAU), nloc, dev.ptr(Vfs), m)
    X = dev.empty(max(m * p, 1))                                  # model matrix at the observed BAUs, :464-465
    call("frk_gather_rows", dev.ctx, m, p, dev.ptr(obs_bau), dev.ptr(X_BAU), nloc, dev.ptr(X), m)

    mdl = SREModel(dev=dev, f=f, basis=basis, BAUs=BAUs, K_type=K_type, r=r, p=p, N=N, lo=lo, hi=hi, m=m, m_total=m_total,
                   S0=S0, S=S, obs_bau=obs_bau, Z=Z, X=X, X_BAU=X_BAU, Ve=Ve, Vfs=Vfs, fs_BAU=fs_BAU,
                   D_basis=BuildD(basis))
    mdl.h2d_bytes = h2d
    _em_initialise(mdl)
    mdl.info_fit["time_SRE"] = time.time() - t0
    return mdl


def _global_stats(mdl: SREModel, t, shift=0.0):
    """sum, sumsq, min, max over ALL ranks' rows."""
    dev = mdl.dev
    st = _vec_stats(dev, t, mdl.m, shift)
    if dev.world > 1:
        s = dev.allreduce_host(np.array([st[0], st[1]]))
        mn = -dev.allreduce_host_max(np.array([-st[2]]))[0]
        mx = dev.allreduce_host_max(np.array([st[3]]))[0]
        st = np.array([s[0], s[1], mn, mx])
    return st


def _em_initialise(mdl: SREModel) -> None:
    """.EM_initialise, R/SRE.R:704-731: mu_eta = 0, S_eta = Q_eta = I, Khat = var(Z) I, Khat_inv = I / var(Z),
    alphahat = OLS, sigma2fshat = mean(diag(Ve)) / 4."""
    dev, r, m, p = mdl.dev, mdl.r, mdl.m, mdl.p
    mt = mdl.m_total
    zs = _global_stats(mdl, mdl.Z)
    zbar = zs[0] / mt
    zs2 = _global_stats(mdl, mdl.Z, zbar)
    varZ = zs2[1] / (mt - 1) if mt > 1 else float("nan")
    ves = _global_stats(mdl, mdl.Ve)
    vfs = _global_stats(mdl, mdl.Vfs)
    mdl.homoscedastic = bool(ves[2] == ves[3] and vfs[2] == vfs[3])          # all(a == a[1]) & all(b == b[1]), R/SREfit.R:278-280
    mdl.Ve0, mdl.Vfs0 = float(ves[2]), float(vfs[2])
    mdl.all_vfs_zero = bool(vfs[2] == 0.0 and vfs[3] == 0.0)
    mdl.sigma2fshat = float(ves[0] / mt) / 4.0
    # r x r slots
    mdl.S_eta, mdl.Q_eta, mdl.Khat, mdl.Khat_inv = (dev.empty(r * r) for _ in range(4))
    call("frk_scaled_identity", dev.ctx, r, 1.0, dev.ptr(mdl.S_eta))
    call("frk_scaled_identity", dev.ctx, r, 1.0, dev.ptr(mdl.Q_eta))
    call("frk_scaled_identity", dev.ctx, r, 1.0 / (1.0 / varZ), dev.ptr(mdl.Khat))
    call("frk_scaled_identity", dev.ctx, r, 1.0 / (1.0 / (1.0 / varZ)), dev.ptr(mdl.Khat_inv))
    mdl.mu_eta = dev.zeros(r)
    mdl._acc = dev.empty(r * r + r + 2)
    mdl._w1 = dev.empty(r * r)
    mdl._w2 = dev.empty(r * r)
    mdl._vec = dev.empty(r)
    mdl._t = dev.empty(max(m, 1))
    mdl._q = dev.zeros(max(m, 1))
    mdl._logdetK = r * math.log(varZ)
    # alpha = (X'X)^-1 X'Z : moments with t = 0, Omega1 = 0 in mode 1
    call("frk_memset0", dev.ctx, dev.ptr(mdl._t), 8 * max(m, 1))
    sums = _mstep_sums(mdl, 0.0, 1)
    A, c = sums["uxx"], sums["uxz"]
    mdl.alphahat = np.linalg.solve(A, c)


def _mstep_sums(mdl: SREModel, sigma2fs: float, mode: int) -> Dict[str, np.ndarray]:
    dev, p = mdl.dev, mdl.p
    ns = 2 * p * p + 2 * p + 2
    out = np.zeros(ns)
    call("frk_mstep_sums", dev.ctx, mdl.m, p, dev.ptr(mdl.X), dev.ptr(mdl.Z), dev.ptr(mdl._t), dev.ptr(mdl._q),
         dev.ptr(mdl.Ve), dev.ptr(mdl.Vfs), float(sigma2fs), int(mode), vp(out.ctypes.data))
    out = dev.allreduce_host(out)                                  # O(p^2) doubles
    o = 0
    uxx = out[o:o + p * p].reshape(p, p).T; o += p * p
    uxz = out[o:o + p]; o += p
    sv = out[o]; o += 1
    w0 = out[o]; o += 1
    wx = out[o:o + p]; o += p
    wxx = out[o:o + p * p].reshape(p, p).T
    return dict(uxx=uxx, uxz=uxz, sv=sv, w0=w0, wx=wx, wxx=wxx)


# ---------------------------------------------------------------------------------------
# E-step + log-likelihood
# ---------------------------------------------------------------------------------------

def _gram(mdl: SREModel, alpha: np.ndarray, sigma2fs: float):
    """acc = [upper(S'D^-1 S) | S'D^-1 rho | rho'D^-1 rho | sum log d], summed over ranks."""
    dev, r = mdl.dev, mdl.r
    call("frk_memset0", dev.ctx, dev.ptr(mdl._acc), (r * r + r + 2) * 8)
    a, ap = _alpha_ptr(alpha)
    call("frk_gram_rhs_bucketed", dev.ctx, mdl.m, r, dev.ptr(mdl.S.rowptr), dev.ptr(mdl.S.col), dev.ptr(mdl.S.val),
         *mdl.S.buckets(dev), dev.ptr(mdl.Z), dev.ptr(mdl.X), mdl.p, ap, dev.ptr(mdl.Ve), dev.ptr(mdl.Vfs),
         float(sigma2fs), dev.ptr(mdl._acc))
    dev.allreduce_(mdl._acc)                                       # the ONE large exchange per EM iteration
    return mdl._acc


def _posterior(mdl: SREModel, alpha, sigma2fs, Q_out, Sig_out, mu_out):
    """Q = S'D^-1 S + Khat_inv; Sigma = chol2inv(chol(Q)); mu = Sigma S'D^-1 rho   (R/SREfit.R:252-254).
    Returns (logdet(chol Q) as R's logdet(R) = 2 sum log diag, rho'D^-1 rho, sum log d, ||L^-1 b||^2)."""
    dev, r = mdl.dev, mdl.r
    acc = _gram(mdl, alpha, sigma2fs)
    b = acc[r * r:r * r + r]
    call("frk_sym_from_upper_add", dev.ctx, r, dev.ptr(acc), dev.ptr(mdl.Khat_inv), dev.ptr(Q_out))
    call("frk_d2d", dev.ctx, dev.ptr(mdl._w1), dev.ptr(Q_out), 8 * r * r)
    logdetQ = C.c_double(0.0)
    call("frk_potrf", dev.ctx, r, dev.ptr(mdl._w1), C.byref(logdetQ))
    call("frk_potri", dev.ctx, r, dev.ptr(mdl._w1), dev.ptr(Sig_out), dev.ptr(mdl._w2))
    call("frk_symv", dev.ctx, r, dev.ptr(Sig_out), dev.ptr(b), dev.ptr(mu_out))
    # || (rDinv S) solve(R) ||^2 with solve(R)' = L^-1 left in _w2 by potri   (R/SREfit.R:211)
    call("frk_lower_mulv", dev.ctx, r, dev.ptr(mdl._w2), dev.ptr(b), dev.ptr(mdl._vec))
    nrm = C.c_double(0.0)
    call("frk_dot", dev.ctx, r, dev.ptr(mdl._vec), dev.ptr(mdl._vec), C.byref(nrm))
    sc = dev.download(acc[r * r + r:r * r + r + 2])
    return logdetQ.value, float(sc[0]), float(sc[1]), nrm.value


def _estep_loglik(mdl: SREModel) -> float:
    """One pass producing llk (.loglik.ind, R/SREfit.R:174-220) and the E-step slots
    mu_eta, S_eta, Q_eta (.SRE.Estep.ind, :232-260)."""
    logdetQ, rDr, slogd, nrm = _posterior(mdl, mdl.alphahat, mdl.sigma2fshat, mdl.Q_eta, mdl.S_eta, mdl.mu_eta)
    log_det_SigmaZ = logdetQ + mdl._logdetK + slogd               # logdet(R) + log|K| + logdet(cholD)
    quad = rDr - nrm
    return -0.5 * mdl.m_total * math.log(2 * math.pi) - 0.5 * log_det_SigmaZ - 0.5 * quad


def loglik(mdl: SREModel) -> float:
    """logLik(SRE) at the current parameters (R/SREutils.R:22-32 -> .loglik.ind); leaves the slots untouched."""
    dev = mdl.dev
    Q = dev.empty(mdl.r * mdl.r)
    Sg = dev.empty(mdl.r * mdl.r)
    mu = dev.empty(mdl.r)
    logdetQ, rDr, slogd, nrm = _posterior(mdl, mdl.alphahat, mdl.sigma2fshat, Q, Sg, mu)
    return -0.5 * mdl.m_total * math.log(2 * math.pi) - 0.5 * (logdetQ + mdl._logdetK + slogd) - 0.5 * (rDr - nrm)


# ---------------------------------------------------------------------------------------
# M-step
# ---------------------------------------------------------------------------------------

def _res_index_sets(basis) -> List[np.ndarray]:
    res = np.asarray(basis.res)
    return [np.nonzero(res == i)[0].astype(np.int32) for i in range(1, int(res.max()) + 1)]


def _block_state(mdl: SREModel):
    """Per-resolution device buffers for the block-exponential K (index sets, D_i, scratch)."""
    if mdl._blocks:
        return mdl._blocks
    dev = mdl.dev
    if isinstance(mdl.basis, TensorP_Basis):
        raise FRKError("K_type='block-exponential' with a tensor-product basis is not on the device path yet; "
                       "use K_type='unstructured'")
    idxs = _res_index_sets(mdl.basis)
    blocks = []
    for i, ix in enumerate(idxs):
        ni = len(ix)
        D = np.asarray(mdl.D_basis[i], dtype=np.float64)
        blocks.append(dict(ix=ix, ni=ni, d_ix=dev.upload(ix), D=dev.upload(D), D01=float(D[0, 1]) if ni > 1 else 0.0,
                           eta2=dev.empty(ni * ni), Ki=dev.empty(ni * ni), Kinv=dev.empty(ni * ni), work=dev.empty(ni * ni)))
        mdl.h2d_bytes += D.nbytes + ix.nbytes
    mdl._blocks = dict(blocks=blocks, max_l=float(np.max(mdl.D_basis[0])))
    return mdl._blocks


def _chol2inv_logdet(dev: Device, n: int, A, Ainv, work):
    """A <- chol(A) in place; Ainv <- chol2inv; returns log|A|."""
    ld = C.c_double(0.0)
    call("frk_potrf", dev.ctx, n, dev.ptr(A), C.byref(ld))
    call("frk_potri", dev.ctx, n, dev.ptr(A), dev.ptr(Ainv), dev.ptr(work))
    return ld.value


def _update_K(mdl: SREModel, lam=0.0) -> None:
    """.update_K / .regularise_K, R/SREfit.R:441-694.  Writes the new Khat into mdl.Khat."""
    dev, r = mdl.dev, mdl.r
    if mdl.K_type == "unstructured":
        lam_arr = np.atleast_1d(np.asarray(lam, dtype=np.float64))
        call("frk_add_outer", dev.ctx, r, dev.ptr(mdl.S_eta), dev.ptr(mdl.mu_eta), dev.ptr(mdl.Khat))    # :689
        if np.any(lam_arr > 0):                                                                     # :669-686
            reg = lam_arr[0] * np.ones(r) if lam_arr.size == 1 else lam_arr[np.asarray(mdl.basis.res) - 1]
            Q = dev.empty(r * r)
            _chol2inv_logdet(dev, r, mdl.Khat, Q, mdl._w2)
            call("frk_add_diag", dev.ctx, r, dev.ptr(dev.upload(np.asarray(reg, dtype=np.float64))), dev.ptr(Q))
            _chol2inv_logdet(dev, r, Q, mdl.Khat, mdl._w2)
        return
    st = _block_state(mdl)
    Khat_old = mdl.Khat
    Knew = dev.zeros(r * r)
    taus, omegas = [], []
    tiny = np.zeros(2)
    for i, B in enumerate(st["blocks"]):
        ni, ix0 = B["ni"], int(B["ix"][0])
        # eta2_i = (S_eta + mu mu')[idx, idx]                                                      :459-464
        call("frk_block_gather_outer", dev.ctx, r, dev.ptr(mdl.S_eta), dev.ptr(mdl.mu_eta), ni, dev.ptr(B["d_ix"]),
             dev.ptr(B["eta2"]))
        # omega_i = n_i / tr(C_i^-1 eta2_i),  C_i = Khat[idx,idx] / Khat[idx1,idx1]                 :466-482
        call("frk_block_gather", dev.ctx, r, dev.ptr(Khat_old), ni, dev.ptr(B["d_ix"]), dev.ptr(B["Ki"]))
        call("frk_d2h", dev.ctx, vp(tiny.ctypes.data), dev.ptr(B["Ki"]), 8 * min(2, ni * ni))
        k00 = float(tiny[0])
        k01 = float(tiny[1]) if ni > 1 else 0.0
        _chol2inv_logdet(dev, ni, B["Ki"], B["Kinv"], B["work"])
        trc = C.c_double(0.0)
        call("frk_dot", dev.ctx, ni * ni, dev.ptr(B["Kinv"]), dev.ptr(B["eta2"]), C.byref(trc))
        omega = ni / (k00 * trc.value)                             # (K/k00)^-1 = k00 K^-1
        omegas.append(omega)

        def f_tau(tau, B=B, ni=ni, omega=omega):                   # :485-529
            tau = float(tau[0])
            if tau <= 1e-10:
                return np.inf
            call("frk_exp_kernel", dev.ctx, ni * ni, dev.ptr(B["D"]), tau, 1.0, dev.ptr(B["Ki"]))
            logdetKi = _chol2inv_logdet(dev, ni, B["Ki"], B["Kinv"], B["work"])
            t = C.c_double(0.0)
            call("frk_dot", dev.ctx, ni * ni, dev.ptr(B["Kinv"]), dev.ptr(B["eta2"]), C.byref(t))
            return -(0.5 * (-logdetKi) - omega / 2.0 * t.value)

        with np.errstate(divide="ignore", invalid="ignore"):       # :608-614
            c01 = k01 / k00
            par0 = max(-B["D01"] / np.log(c01), 1e-9) if ni > 1 else 1e-9
        if not np.isfinite(par0) or par0 == 1e-9:
            par0 = st["max_l"] / 10.0
        x, _, _, _ = nmmin(f_tau, [par0], maxit=100, reltol=1e-4)  # optim(): Nelder-Mead, :618-621
        taus.append(float(x[0]))
        call("frk_exp_kernel", dev.ctx, ni * ni, dev.ptr(B["D"]), float(x[0]), 1.0 / omega, dev.ptr(B["Ki"]))   # :625-637
        call("frk_block_scatter", dev.ctx, r, dev.ptr(Knew), ni, dev.ptr(B["d_ix"]), dev.ptr(B["Ki"]))
    mdl.Khat = Knew
    mdl.info_fit["tau"] = taus
    mdl.info_fit["omega"] = omegas


def _mstep(mdl: SREModel, lam=0.0, known_sigma2fs=None) -> None:
    """.SRE.Mstep.ind for diagonal Ve / Vfs, R/SREfit.R:263-438."""
    dev, r, p = mdl.dev, mdl.r, mdl.p
    sigma2fs = mdl.sigma2fshat
    _update_K(mdl, lam)                                            # :270-272
    call("frk_d2d", dev.ctx, dev.ptr(mdl._w1), dev.ptr(mdl.Khat), 8 * r * r)
    mdl._logdetK = _chol2inv_logdet(dev, r, mdl._w1, mdl.Khat_inv, mdl._w2)   # Khat_inv = chol2inv(chol(K)), :273
    est = known_sigma2fs is None
    sigma2fs_new = sigma2fs if est else float(known_sigma2fs)
    # Omega_diag1 = rowSums((S R_eta')^2) = s_j'(S_eta + mu mu')s_j = q_j + t_j^2,  t = S mu_eta     :332-334
    call("frk_quadform_spmv_bucketed", dev.ctx, mdl.m, r, dev.ptr(mdl.S.rowptr), dev.ptr(mdl.S.col), dev.ptr(mdl.S.val),
         *mdl.S.buckets(dev), dev.ptr(mdl.S_eta), dev.ptr(mdl.mu_eta), dev.ptr(mdl._q), dev.ptr(mdl._t))

    def gls_and_J(s2, mode=0):
        s = _mstep_sums(mdl, s2, mode)
        al = np.linalg.solve(s["uxx"], s["uxz"])
        sumwOm = s["w0"] + 2.0 * float(al @ s["wx"]) + float(al @ s["wxx"] @ al)
        return al, s, sumwOm

    def J(s2):                                                     # :335-356
        if s2 < 0:
            return np.inf
        al, s, swo = gls_and_J(s2)
        return -(-0.5 * s["sv"] + 0.5 * swo)

    if not mdl.all_vfs_zero:
        if not mdl.homoscedastic:
            if est:
                amp, OK = 10.0, False
                while not OK:                                      # :368-380
                    amp *= 10
                    if not (np.sign(J(sigma2fs / amp)) == np.sign(J(sigma2fs * amp))):
                        OK = True
                    if amp > 1e9:
                        OK = True
                if amp > 1e9:
                    sigma2fs_new = 0.0
                else:
                    sigma2fs_new = zeroin(J, sigma2fs / amp, sigma2fs * amp)     # uniroot, :386-388
            alpha, _, _ = gls_and_J(sigma2fs_new)                   # :391-400
        else:
            alpha, s, sumOm = gls_and_J(0.0, mode=1)                # :404-405
            if est:
                sigma2fs_new = 1.0 / mdl.Vfs0 * (sumOm / mdl.m_total - mdl.Ve0)   # :408-417
                if sigma2fs_new < 0:
                    sigma2fs_new = 0.0
    else:                                                          # :424-428
        alpha, _, _ = gls_and_J(0.0)
        if est:
            sigma2fs_new = 0.0
    mdl.alphahat, mdl.sigma2fshat = np.asarray(alpha, dtype=np.float64), float(sigma2fs_new)


# ---------------------------------------------------------------------------------------
# SRE.fit
# ---------------------------------------------------------------------------------------

def SRE_fit(mdl: SREModel, n_EM: int = 100, tol: float = 0.01, method: str = "EM", lambda_=0.0,
            known_sigma2fs=None, print_lik: bool = False) -> SREModel:
    """SRE.fit(object, n_EM, tol, method = "EM", lambda, print_lik, known_sigma2fs), R/SREfit.R:3-160."""
    if method != "EM":
        raise FRKError("only method = \"EM\" (Gaussian data, identity link) is on the device path")
    if known_sigma2fs is not None:
        mdl.sigma2fshat = float(known_sigma2fs)                     # R/SREfit.R:19
    t0 = time.time()
    llk = np.zeros(n_EM)
    i = 0
    for i in range(n_EM):
        llk[i] = _estep_loglik(mdl)                                # logLik + E-step, :105-106
        _mstep(mdl, lambda_, known_sigma2fs)                       # :107
        if print_lik:
            print(f"EM iteration {i + 1}: log-likelihood {llk[i]:.10g}")
        if i > 0 and abs(llk[i] - llk[i - 1]) < tol:               # :110-114
            break
    nit = i + 1
    mdl.dev.sync()
    mdl.info_fit.update(method="EM", time=time.time() - t0, num_iterations=nit, converged=int(nit != n_EM),
                        llk=llk[:nit].copy(), sigma2fshat_equal_0=int(mdl.sigma2fshat == 0))
    return mdl


# ---------------------------------------------------------------------------------------
# predict
# ---------------------------------------------------------------------------------------

def predict(mdl: SREModel, obs_fs: bool = False, to_host: bool = True):
    """predict(SRE, obs_fs) at the BAUs (newdata = NULL, covariances = FALSE): .SRE.predict_EM,
    R/SREpredict.R:243-375 with .batch_compute_var (R/SREutils.R:245-305).

    obs_fs = TRUE or sigma2fs == 0: the stored posterior (mu_eta, S_eta of the last E-step).
    obs_fs = FALSE and sigma2fs > 0: the joint (eta, xi) solve at the final parameters; for point data its
    diagonal reduces to one more E-step plus the same sparse quadratic form.
    Returns dict(mu, var, sd) for this rank's BAU rows [lo, hi) (host arrays, or device tensors)."""
    dev, r = mdl.dev, mdl.r
    nloc = mdl.hi - mdl.lo
    s2 = mdl.sigma2fshat
    xa = dev.empty(max(nloc, 1))
    a, ap = _alpha_ptr(mdl.alphahat)
    call("frk_xalpha", dev.ctx, nloc, mdl.p, dev.ptr(mdl.X_BAU), ap, dev.ptr(xa))
    q = dev.empty(max(nloc, 1))
    smu = dev.empty(max(nloc, 1))
    mu = dev.empty(max(nloc, 1))
    var = dev.empty(max(nloc, 1))
    sd = dev.empty(max(nloc, 1))
    if obs_fs or s2 == 0:
        Sig, mue = mdl.S_eta, mdl.mu_eta
        call("frk_quadform_spmv_bucketed", dev.ctx, nloc, r, dev.ptr(mdl.S0.rowptr), dev.ptr(mdl.S0.col), dev.ptr(mdl.S0.val),
             *mdl.S0.buckets(dev), dev.ptr(Sig), dev.ptr(mue), dev.ptr(q), dev.ptr(smu))
        call("frk_predict_finalize", dev.ctx, nloc, 0, s2, dev.ptr(q), dev.ptr(smu), dev.ptr(xa), vp(None), vp(None),
             vp(None), dev.ptr(mu), dev.ptr(var), dev.ptr(sd))
    else:
        Q = dev.empty(r * r)
        Sig = dev.empty(r * r)
        mue = dev.empty(r)
        _posterior(mdl, mdl.alphahat, s2, Q, Sig, mue)
        del Q
        call("frk_quadform_spmv_bucketed", dev.ctx, nloc, r, dev.ptr(mdl.S0.rowptr), dev.ptr(mdl.S0.col), dev.ptr(mdl.S0.val),
             *mdl.S0.buckets(dev), dev.ptr(Sig), dev.ptr(mue), dev.ptr(q), dev.ptr(smu))
        # A = diag(C' Ve^-1 C), ytil = C' Ve^-1 (Z - X alpha)
        A = dev.zeros(max(nloc, 1))
        ytil = dev.zeros(max(nloc, 1))
        tmp = dev.empty(max(mdl.m, 1))
        inv_ve = dev.empty(max(mdl.m, 1))
        call("frk_resid", dev.ctx, mdl.m, mdl.p, dev.ptr(mdl.X), ap, dev.ptr(mdl.Z), dev.ptr(mdl.Ve), dev.ptr(tmp))
        call("frk_scatter_add", dev.ctx, mdl.m, dev.ptr(mdl.obs_bau), dev.ptr(tmp), dev.ptr(ytil))
        call("frk_recip", dev.ctx, mdl.m, dev.ptr(mdl.Ve), dev.ptr(inv_ve))
        call("frk_scatter_add", dev.ctx, mdl.m, dev.ptr(mdl.obs_bau), dev.ptr(inv_ve), dev.ptr(A))
        call("frk_predict_finalize", dev.ctx, nloc, 1, s2, dev.ptr(q), dev.ptr(smu), dev.ptr(xa), dev.ptr(A), dev.ptr(ytil),
             dev.ptr(mdl.fs_BAU), dev.ptr(mu), dev.ptr(var), dev.ptr(sd))
    if not to_host:
        return dict(mu=mu, var=var, sd=sd)
    out = dict(mu=dev.download(mu, nloc), var=dev.download(var, nloc), sd=dev.download(sd, nloc))
    mdl.d2h_bytes += 3 * 8 * nloc
    return out

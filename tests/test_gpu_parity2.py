"""GPU parity, second file (round 2): the configurations at their STATED size, the M-step corner branches, and the dense kernels at the
benchmarked r.  Same provenance note as test_gpu_parity.py: the oracle is a CPU restatement of the cited reference lines (no R in this
image); EM numerics are parity-unpinned by reference-held vectors."""
import ctypes as C
import os

import numpy as np
import pytest

from cases import readme_points
from oracle import frk_oracle as O
from test_gpu_parity import GOLD, RTOL, RTOL_UNIROOT, _isea3h, _mk_basis, _same_pattern, close

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------------------
# dense kernels at the benchmarked size (C3: r = 3276, finest block 2916) and beyond
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("r", [2916, 3276, 4100])
def test_chol2inv_at_benchmark_size(dev, r):
    from frk_b200._lib import call, vp
    rng = np.random.default_rng(r)
    A = rng.standard_normal((r, r + 5))
    A = A @ A.T + r * np.eye(r)
    Ainv = np.empty_like(A)
    ld = C.c_double(0)
    Af = np.asfortranarray(A)
    call("frk_chol2inv_host", dev.ctx, r, vp(Af.ctypes.data), vp(Ainv.ctypes.data), C.byref(ld))
    ref = O.chol2inv(O.chol_upper(A))
    close(Ainv.T, ref, 1e-11)
    np.testing.assert_allclose(ld.value, np.linalg.slogdet(A)[1], rtol=1e-12)
    assert np.array_equal(Ainv, Ainv.T)


def test_chol2inv_exponential_covariance_2916(dev):
    """The matrix the block-exponential M-step factorises at C3: exp(-D / tau) for the 54 x 54 centroid grid of resolution 3
    (R/SREfit.R:503,517-519) -- far worse conditioned than a random Wishart matrix."""
    from frk_b200._lib import call, vp
    g = (np.arange(54) + 0.5) / 54
    X, Y = np.meshgrid(g, g)
    P = np.column_stack([X.ravel(), Y.ravel()])
    D = np.sqrt(((P[:, None, :] - P[None, :, :]) ** 2).sum(-1))
    for tau in (0.02, 0.15):
        K = np.exp(-D / tau)
        Kinv = np.empty_like(K)
        ld = C.c_double(0)
        Kf = np.asfortranarray(K)                       # (keep the array alive across the call)
        call("frk_chol2inv_host", dev.ctx, len(K), vp(Kf.ctypes.data), vp(Kinv.ctypes.data), C.byref(ld))
        ref = O.chol2inv(O.chol_upper(K))
        close(Kinv.T, ref, 1e-9)
        np.testing.assert_allclose(ld.value, np.linalg.slogdet(K)[1], rtol=1e-10)
        # what f_tau consumes: tr(K^-1 E) for a PSD E
        E = np.exp(-D / 0.05) * 0.7
        np.testing.assert_allclose(np.sum(Kinv.T * E), np.sum(ref * E), rtol=1e-9)


def test_block_exponential_iteration_at_r3276(dev):
    """One-and-a-half EM iterations with the BENCHMARKED basis (auto_basis(plane, nres = 3, regular = 2): r = 3276 = 36 + 324 + 2916)
    on a problem small enough in m for the oracle: log-likelihood trace, tau, mu_eta, Sigma_eta, K and predictive variances."""
    import frk_b200 as F
    xy, z = readme_points(30000, seed=21)
    std = np.full(len(z), 0.5)
    ob = O.auto_basis(O.plane(), np.array([[0.0, 0.0], [1.0, 1.0]]), regular=2, nres=3)
    assert ob.n == 3276
    ba = O.auto_BAUs_plane(xy, cellsize=1.0 / 299, xlims=(0, 1), ylims=(0, 1))
    uo, means, _ = O.bin_mean(O.points_to_BAUs(xy, ba), np.column_stack([z, std]))
    Mo = O.SRE(ob, ba.centroids(), np.ones(ba.n), uo, means[:, 0], means[:, 1], K_type="block-exponential")
    gb = F.auto_BAUs(F.plane(), cellsize=1.0 / 299, xlims=(0, 1), ylims=(0, 1))
    gb["fs"] = 1.0
    Mg = F.SRE("z ~ 1", F.SpatialPointsDataFrame(xy, {"z": z, "std": std}), _mk_basis(F, ob), gb, est_error=False, K_type="block-exponential")
    assert np.array_equal(dev.download(Mg.obs_bau, Mg.m), Mo.obs_bau) and _same_pattern(Mo.S, Mg.S.to_scipy(dev))
    O.SRE_fit(Mo, n_EM=2, tol=0.0)
    F.SRE_fit(Mg, n_EM=2, tol=0.0)
    np.testing.assert_allclose(Mg.info_fit["llk"], Mo.info_fit["llk"], rtol=RTOL)
    np.testing.assert_allclose(Mg.info_fit["tau"], np.concatenate([np.atleast_1d(t) for t in Mo.info_fit["tau"]]), rtol=RTOL)
    np.testing.assert_allclose(Mg.sigma2fshat, Mo.sigma2fshat, rtol=RTOL)
    np.testing.assert_allclose(Mg.alphahat, Mo.alphahat, rtol=RTOL)
    close(Mg.get("mu_eta"), Mo.mu_eta)
    close(Mg.get("S_eta"), Mo.S_eta)
    close(Mg.get("Khat"), Mo.Khat, RTOL)
    for obs_fs in (True, False):
        po = O.predict(Mo, obs_fs=obs_fs)
        pg = F.predict(Mg, obs_fs=obs_fs)
        close(pg["mu"], po[0])
        np.testing.assert_allclose(pg["var"], po[1], rtol=RTOL)


# ---------------------------------------------------------------------------------------------------------
# BASELINE config 2 at its stated size
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K_type", ["block-exponential", "unstructured"])
def test_config2_airs_full_size_hexagon_BAUs(dev, K_type):
    """BASELINE config 2 as stated: AIRS_05_2003 days 1-3 (all 43,059 soundings; std = co2std -> heteroscedastic / uniroot branch),
    sphere(), ISEA3H bisquare basis nres = 3 (r = 1176), f = co2avgret ~ lat + 1, BAUs = the reference's own ISEA3H resolution-6
    hexagons (tests/golden/isea3h_res6_polygons.npz: 7204 of the 7292 cells, the 88 dateline-straddling ones -- split with sf in the
    reference, BAU construction -- left out).  Point -> hexagon lookup on the device, bit-exact against the restated sp::over."""
    import frk_b200 as F
    g = np.load(os.path.join(GOLD, "isea3h_res6_polygons.npz"))
    a = np.load(os.path.join(GOLD, "airs_days123.npz"))
    assert len(a["lon"]) == 43059 and len(g["ptr"]) - 1 == 7204
    ll = np.column_stack([a["lon"], a["lat"]])
    cen = np.column_stack([g["cx"], g["cy"]])
    isea = _isea3h()
    ob = O.auto_basis(O.sphere(), ll, nres=3, isea3h=isea, isea3h_lo=2)
    assert ob.n == 1176
    ref = O.points_to_polygon_BAUs(ll, g["ptr"], g["vx"], g["vy"])
    uo, means, _ = O.bin_mean(ref, np.column_stack([a["co2avgret"], a["co2std"]]))
    X_BAU = np.column_stack([np.ones(len(cen)), cen[:, 1]])
    Mo = O.SRE(ob, cen, np.ones(len(cen)), uo, means[:, 0], means[:, 1], X=X_BAU[uo], X_BAU=X_BAU, K_type=K_type)
    baus = F.PolygonBAUs(cen.copy(), ptr=g["ptr"], vx=g["vx"], vy=g["vy"])
    baus["fs"] = 1.0
    baus["lat"] = cen[:, 1]
    data = F.SpatialPointsDataFrame(ll, {"co2avgret": a["co2avgret"], "std": a["co2std"]})
    out = F.map_data_to_BAUs(data, baus, ["co2avgret", "std"], dev)
    assert np.array_equal(dev.download(out["bau"], len(ll)), ref)                      # C_Z: bit-exact
    Mg = F.SRE("co2avgret ~ lat + 1", data, _mk_basis(F, ob), baus, est_error=False, K_type=K_type)
    assert np.array_equal(dev.download(Mg.obs_bau, Mg.m), uo) and not Mg.homoscedastic
    close(Mg.S.to_scipy(dev).toarray(), Mo.S.toarray(), 1e-9)
    n_EM = 4
    O.SRE_fit(Mo, n_EM=n_EM, tol=0.0)
    F.SRE_fit(Mg, n_EM=n_EM, tol=0.0)
    np.testing.assert_allclose(Mg.info_fit["llk"], Mo.info_fit["llk"], rtol=RTOL)
    np.testing.assert_allclose(Mg.alphahat, Mo.alphahat, rtol=RTOL_UNIROOT)
    np.testing.assert_allclose(Mg.sigma2fshat, Mo.sigma2fshat, rtol=RTOL_UNIROOT)
    close(Mg.get("mu_eta"), Mo.mu_eta, RTOL_UNIROOT)
    for obs_fs in (True, False):
        po = O.predict(Mo, obs_fs=obs_fs)
        pg = F.predict(Mg, obs_fs=obs_fs)
        np.testing.assert_allclose(pg["mu"], po[0], rtol=RTOL_UNIROOT)
        np.testing.assert_allclose(pg["var"], po[1], rtol=RTOL_UNIROOT)


# ---------------------------------------------------------------------------------------------------------
# M-step corner branches (R/SREfit.R:377-384, :424-428) and prediction after sigma2fs = 0
# ---------------------------------------------------------------------------------------------------------
def _plane_case(F, m, seed, std, fs=1.0, K_type="unstructured", nres=2, cellsize=0.02):
    xy, z = readme_points(m, seed)
    ob = O.auto_basis(O.plane(), xy, regular=1, nres=nres)
    obaus = O.auto_BAUs_plane(xy, cellsize=cellsize)
    uo, means, _ = O.bin_mean(O.points_to_BAUs(xy, obaus), np.column_stack([z, std]))
    fsv = np.full(obaus.n, fs)
    Mo = O.SRE(ob, obaus.centroids(), fsv, uo, means[:, 0], means[:, 1], K_type=K_type)
    gb = F.auto_BAUs(F.plane(), cellsize=cellsize, data=F.SpatialPointsDataFrame(xy, {}))
    gb["fs"] = fs
    Mg = F.SRE("z ~ 1", F.SpatialPointsDataFrame(xy, {"z": z, "std": std}), _mk_basis(F, ob), gb, est_error=False, K_type=K_type)
    return Mo, Mg


@pytest.mark.parametrize("K_type", ["unstructured", "block-exponential"])
def test_mstep_no_fine_scale_branch(dev, K_type):
    """all(Vfs == 0) (BAUs$fs = 0): alpha by GLS with D = Ve, sigma2fs <- 0 (R/SREfit.R:424-428); predict then takes the
    sigma2fs == 0 path (R/SREpredict.R:334-361) for both values of obs_fs."""
    import frk_b200 as F
    std = 0.3 + 0.4 * np.random.default_rng(7).uniform(size=1200)          # heteroscedastic Ve: the GLS weights matter
    Mo, Mg = _plane_case(F, 1200, 31, std, fs=0.0, K_type=K_type)
    O.SRE_fit(Mo, n_EM=4, tol=0.0)
    F.SRE_fit(Mg, n_EM=4, tol=0.0)
    assert Mg.sigma2fshat == 0.0 == Mo.sigma2fshat and Mg.info_fit["sigma2fshat_equal_0"] == 1
    np.testing.assert_allclose(Mg.info_fit["llk"], Mo.info_fit["llk"], rtol=RTOL)
    np.testing.assert_allclose(Mg.alphahat, Mo.alphahat, rtol=RTOL)
    close(Mg.get("mu_eta"), Mo.mu_eta)
    close(Mg.get("S_eta"), Mo.S_eta)
    for obs_fs in (True, False):
        po = O.predict(Mo, obs_fs=obs_fs)
        pg = F.predict(Mg, obs_fs=obs_fs)
        close(pg["mu"], po[0])
        np.testing.assert_allclose(pg["var"], po[1], rtol=RTOL)


def test_mstep_bracket_give_up_sets_sigma2fs_to_zero(dev):
    """Heteroscedastic branch where J(sigma2fs) never changes sign (the stated measurement error already exceeds the residual
    variance): the bracket search runs to amp > 1e9 and gives up with sigma2fs <- 0 (R/SREfit.R:368-384); the fit continues and
    predict(obs_fs = FALSE) runs with sigma2fshat = 0 (tests/testthat/test_sre_polygon_predict.R:51-55 does the same by assignment)."""
    import frk_b200 as F
    std = 1.5 + 1.0 * np.random.default_rng(9).uniform(size=1500)           # noise in z is 0.5
    Mo, Mg = _plane_case(F, 1500, 33, std, K_type="unstructured")
    assert not Mg.homoscedastic
    O.SRE_fit(Mo, n_EM=4, tol=0.0)
    F.SRE_fit(Mg, n_EM=4, tol=0.0)
    assert Mo.sigma2fshat == 0.0, "the case is meant to hit the give-up branch"
    assert Mg.sigma2fshat == 0.0 and Mg.info_fit["sigma2fshat_equal_0"] == 1
    np.testing.assert_allclose(Mg.info_fit["llk"], Mo.info_fit["llk"], rtol=RTOL)
    np.testing.assert_allclose(Mg.alphahat, Mo.alphahat, rtol=RTOL)
    close(Mg.get("mu_eta"), Mo.mu_eta)
    for obs_fs in (True, False):
        po = O.predict(Mo, obs_fs=obs_fs)
        pg = F.predict(Mg, obs_fs=obs_fs)
        close(pg["mu"], po[0])
        np.testing.assert_allclose(pg["var"], po[1], rtol=RTOL)


def test_predict_after_assigning_sigma2fs_zero(dev):
    """tests/testthat/test_sre_polygon_predict.R:51-55: S@sigma2fshat <- 0 after the fit, then predict."""
    import frk_b200 as F
    Mo, Mg = _plane_case(F, 1000, 1, np.full(1000, 0.5), K_type="block-exponential")
    O.SRE_fit(Mo, n_EM=3, tol=0.0)
    F.SRE_fit(Mg, n_EM=3, tol=0.0)
    Mo.sigma2fshat = 0.0
    Mg.set("sigma2fshat", [0.0])
    for obs_fs in (True, False):
        po = O.predict(Mo, obs_fs=obs_fs)
        pg = F.predict(Mg, obs_fs=obs_fs)
        close(pg["mu"], po[0])
        np.testing.assert_allclose(pg["var"], po[1], rtol=RTOL)


# ---------------------------------------------------------------------------------------------------------
# the host-buffer predict entry point the R shim binds (frkb200_model_predict -> frk_model_predict_host)
# ---------------------------------------------------------------------------------------------------------
def test_predict_host_entry_point_matches_device_predict(dev):
    """.SRE.predict_EM through the entry point of the R overlay: S0 / X_BAU / fs / obs_bau as HOST arrays (what object@S0, the BAU
    data frame and object@Cmat hold in R) give the same mu / var / sd as the device-resident predict and the oracle."""
    import frk_b200 as F
    from frk_b200._lib import call, vp
    Mo, Mg = _plane_case(F, 1500, 41, 0.3 + 0.4 * np.random.default_rng(5).uniform(size=1500), K_type="block-exponential")
    O.SRE_fit(Mo, n_EM=3, tol=0.0)
    F.SRE_fit(Mg, n_EM=3, tol=0.0)
    S0 = Mg.S0.to_scipy(dev).tocsr()
    S0.sort_indices()
    n = S0.shape[0]
    rp, cj, x = S0.indptr.astype(np.int32), S0.indices.astype(np.int32), S0.data.astype(np.float64)
    X = np.ones(n)
    fs = np.ones(n)
    ob = np.ascontiguousarray(dev.download(Mg.obs_bau, Mg.m), dtype=np.int32)
    for obs_fs in (0, 1):
        mu, var, sd = np.empty(n), np.empty(n), np.empty(n)
        call("frk_model_predict_host", Mg.handle, n, vp(rp.ctypes.data), vp(cj.ctypes.data), vp(x.ctypes.data), vp(X.ctypes.data),
             vp(fs.ctypes.data), vp(ob.ctypes.data), vp(None), vp(None), vp(None), obs_fs, 0, vp(None), vp(None), vp(None),
             vp(mu.ctypes.data), vp(var.ctypes.data), vp(sd.ctypes.data))
        pg = F.predict(Mg, obs_fs=bool(obs_fs))
        np.testing.assert_allclose(mu, pg["mu"], rtol=1e-12, atol=1e-13)     # (not bit-identical: the row buckets are rebuilt, atomics order)
        np.testing.assert_allclose(var, pg["var"], rtol=1e-12)
        po = O.predict(Mo, obs_fs=bool(obs_fs))
        close(mu, po[0], RTOL_UNIROOT)
        close(var, po[1], RTOL_UNIROOT)
        close(sd, np.sqrt(po[1]), RTOL_UNIROOT)


# ---------------------------------------------------------------------------------------------------------
# eval_basis, second-generation kernels: threshold-table count, reciprocal division, multi-flush staging, in-kernel normalisation
# ---------------------------------------------------------------------------------------------------------
def test_eval_basis_long_rows_and_normalisation_bit_exact(dev):
    """C3's basis shape (regular = 2, nres = 3: rows of up to ~39 entries, several staging flushes per row) on a regular BAU-like grid
    whose points coincide with centroids and support edges; with and without the fused row normalisation (R/SRE.R:416-421)."""
    import frk_b200 as F
    from frk_b200.basis import eval_basis_device
    xy, _ = readme_points(2000, seed=5)
    ob = O.auto_basis(O.plane(), xy, regular=2, nres=3)
    g = np.linspace(0.0, 1.0, 257)
    s = np.column_stack([np.tile(g, len(g)), np.repeat(g, len(g))])
    s = np.vstack([s, ob.loc, ob.loc + np.array([ob.scale, np.zeros(ob.n)]).T])      # on centroids; exactly one scale to the right
    S_o = O.eval_basis_fast(ob, s).tocsr()
    S_o.sort_indices()
    assert np.diff(S_o.indptr).max() > 32
    gb = _mk_basis(F, ob)
    S_g = F.eval_basis(gb, s, dev)
    assert _same_pattern(S_o, S_g)
    assert np.array_equal(S_o.data, S_g.data)
    S_n = eval_basis_device(gb, s, dev, normalise=True).to_scipy(dev)
    S_on = O.normalise_rows(S_o).tocsr()
    assert _same_pattern(S_on, S_n)
    assert np.array_equal(S_on.data, S_n.data)


@pytest.mark.parametrize("man", ["plane", "real_line"])
def test_eval_basis_support_edge_ties_one_ulp(dev, man):
    """Points whose distance to a centroid is the scale itself, or one / two ulps either side of it: the threshold table
    (q < T  <=>  sqrt_rn(q) < scale) must reproduce the reference's  y < R  exactly."""
    import frk_b200 as F
    rng = np.random.default_rng(11)
    nb = 40
    scale = rng.uniform(0.05, 3.0, nb)
    if man == "plane":
        loc = rng.uniform(-50, 50, (nb, 2))
        ob = O.local_basis(O.plane(), loc, scale, "bisquare")
    else:
        loc = rng.uniform(-50, 50, (nb, 1))
        ob = O.local_basis(O.real_line(), loc, scale, "bisquare")
    pts = []
    for j in range(nb):
        for k in range(-3, 4):
            d = scale[j]
            for _ in range(abs(k)):
                d = np.nextafter(d, np.inf if k > 0 else 0.0)
            if man == "plane":
                for (a, b) in [(1.0, 0.0), (0.0, 1.0), (0.6, 0.8), (-0.8, 0.6)]:
                    pts.append(loc[j] + d * np.array([a, b]))
            else:
                pts.append(loc[j] + d)
                pts.append(loc[j] - d)
    s = np.asarray(pts).reshape(len(pts), -1)
    S_o = O.eval_basis(ob, s).tocsr()
    S_o.sort_indices()
    S_g = F.eval_basis(_mk_basis(F, ob), s, dev)
    assert _same_pattern(S_o, S_g)
    assert np.array_equal(S_o.data, S_g.data)


# ---------------------------------------------------------------------------------------------------------
# Non-diagonal Vfs: observations coupled through shared BAUs (R/SREfit.R:186-189, 240-252, 296-328)
# ---------------------------------------------------------------------------------------------------------
def _overlapping_supports(F, K_type, seed=3):
    """Square supports on the unit square, laid out in short chains of two or three overlapping squares (connected groups of 2-3
    observations) plus isolated ones; a covariate known at every BAU."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    gb = F.auto_BAUs(F.plane(), cellsize=0.025, data=F.SpatialPointsDataFrame(np.array([[0.0, 0.0], [1.0, 1.0]]), {}))
    cen = gb.centroids()
    gb["fs"] = 1.0 + 0.5 * (cen[:, 1] > 0.5)
    gb["x1"] = cen[:, 0] - cen[:, 1]
    ptr, vx, vy, z = [0], [], [], []
    s = 0.11
    for iy in range(6):
        for ix in range(6):
            x0, y0 = 0.004 + 0.16 * ix, 0.004 + 0.16 * iy
            chain = (ix + iy) % 3                                     # 0: one square; 1: two overlapping; 2: three in an L
            offs = [(0.0, 0.0)] + ([(0.04, 0.03)] if chain >= 1 else []) + ([(-0.0, 0.05)] if chain == 2 else [])
            for dx, dy in offs:
                a, b = x0 + dx, y0 + dy
                vx += [a, a + s, a + s, a, a]; vy += [b, b, b + s, b + s, b]; ptr.append(len(vx))
                xm, ym = a + s / 2, b + s / 2
                z.append(np.sin(5 * xm) + np.cos(4 * ym) + 0.3 * (xm - ym) + 0.6 * rng.standard_normal())
    m = len(z)
    std = 0.1 + 0.05 * rng.uniform(size=m)
    data = F.SpatialPolygonsDataFrame(np.array(ptr), np.array(vx), np.array(vy), {"z": np.array(z), "std": std})
    ob = O.auto_basis(O.plane(), cen, regular=1, nres=2)
    i_idx, j_idx, x_idx = O.BuildC_polygons(data.ptr, data.vx, data.vy, cen)
    Cm = sp.csr_matrix((x_idx, (i_idx, j_idx)), shape=(m, len(cen)))
    X_BAU = np.column_stack([np.ones(len(cen)), gb["x1"]])
    Mo = O.SRE(ob, cen, gb["fs"], None, np.array(z), std, X_BAU=X_BAU, K_type=K_type, C=Cm, allow_overlap=True)
    Mg = F.SRE("z ~ x1", data, _mk_basis(F, ob), gb, est_error=False, K_type=K_type)
    return Mo, Mg, Cm


@pytest.mark.parametrize("K_type", ["unstructured", "block-exponential"])
def test_overlapping_areal_supports(dev, K_type):
    """SURVEY.md section 8(f) #1, the remainder: supports that share BAUs.  BuildC with shared BAUs (incidence exact), S = Cmat S0,
    the full Vfs = tcrossprod(Cmat sqrt(fs)) with its off-diagonal blocks, and the EM through the reference's GENERAL branches --
    chol(D) in the log-likelihood and the E-step, J(sigma2fs) with Dinv %*% Vfs %*% Dinv -- against their dense restatement."""
    import scipy.sparse as sp
    import frk_b200 as F
    Mo, Mg, Cm = _overlapping_supports(F, K_type)
    assert Mo.Vmat is not None and Mg.coupled
    rp, mem, w, nmem = Mg.C_Z
    Cn = sp.csr_matrix(Mo.Cmat); Cn.sort_indices()
    assert nmem == Cn.nnz and np.array_equal(dev.download(rp, Mg.m + 1), Cn.indptr)
    assert np.array_equal(dev.download(mem, nmem), Cn.indices)
    np.testing.assert_allclose(dev.download(w, nmem), Cn.data, rtol=1e-15)
    Sg = Mg.S.to_scipy(dev)
    assert _same_pattern(Mo.S, Sg)
    np.testing.assert_allclose(Sg.data, Mo.S.data, rtol=1e-12)
    np.testing.assert_allclose(dev.download(Mg.Vfs, Mg.m), Mo.Vfs, rtol=1e-13)
    O.SRE_fit(Mo, n_EM=5, tol=0.0)
    F.SRE_fit(Mg, n_EM=5, tol=0.0)
    assert Mo.sigma2fshat > 0
    np.testing.assert_allclose(Mg.info_fit["llk"], Mo.info_fit["llk"], rtol=RTOL)
    np.testing.assert_allclose(Mg.sigma2fshat, Mo.sigma2fshat, rtol=RTOL_UNIROOT, atol=1e-12)
    np.testing.assert_allclose(Mg.alphahat, Mo.alphahat, rtol=RTOL_UNIROOT)
    close(Mg.get("mu_eta"), Mo.mu_eta, RTOL_UNIROOT)
    po = O.predict(Mo, obs_fs=True)
    pg = F.predict(Mg, obs_fs=True)
    close(pg["mu"], po[0], RTOL_UNIROOT)
    close(pg["var"], po[1], RTOL_UNIROOT)
    assert abs(F.loglik(Mg) - O.loglik(Mo)) <= RTOL * abs(O.loglik(Mo))
    # Case 2 (obs_fs = FALSE, sigma2fs > 0): the joint (eta, xi) solve, R/SREpredict.R:308-333, against its literal dense restatement
    po2 = O.predict(Mo, obs_fs=False)
    pg2 = F.predict(Mg, obs_fs=False)
    close(pg2["mu"], po2[0], RTOL)
    close(pg2["var"], po2[1], RTOL)
    inside = np.asarray(Cm.sum(axis=0)).ravel() > 0
    assert inside.any() and (~inside).any()


def test_average_in_BAU_false(dev):
    """SRE(average_in_BAU = FALSE), R/geometryfns.R:1304-1305: every observation is kept, several of them share a BAU, so
    Vfs = C diag(fs) C' couples them (blocks of fs_b) and the EM takes the general branches.  Known sigma2fs as well."""
    import scipy.sparse as sp
    import frk_b200 as F
    xy, z = readme_points(400, seed=21)
    std = 0.2 + 0.1 * np.random.default_rng(2).uniform(size=len(z))
    ob = O.auto_basis(O.plane(), xy, regular=1, nres=2)
    obaus = O.auto_BAUs_plane(xy, cellsize=0.06)
    b = O.points_to_BAUs(xy, obaus)
    keep = b >= 0
    m = int(keep.sum())
    assert len(np.unique(b[keep])) < m                                 # several observations per BAU
    Cm = sp.csr_matrix((np.ones(m), b[keep], np.arange(m + 1)), shape=(m, obaus.n))
    fs = 0.5 + obaus.centroids()[:, 0]
    Mo = O.SRE(ob, obaus.centroids(), fs, None, z[keep], std[keep], K_type="block-exponential", C=Cm, allow_overlap=True)
    BAUs = F.auto_BAUs(F.plane(), cellsize=0.06, data=F.SpatialPointsDataFrame(xy, {}))
    BAUs["fs"] = fs
    Mg = F.SRE("z ~ 1", F.SpatialPointsDataFrame(xy, {"z": z, "std": std}), _mk_basis(F, ob), BAUs, est_error=False,
               average_in_BAU=False)
    assert Mg.m == m and Mg.coupled
    assert np.array_equal(dev.download(Mg.obs_bau, m), b[keep])
    O.SRE_fit(Mo, n_EM=4, tol=0.0)
    F.SRE_fit(Mg, n_EM=4, tol=0.0)
    np.testing.assert_allclose(Mg.info_fit["llk"], Mo.info_fit["llk"], rtol=RTOL)
    np.testing.assert_allclose(Mg.sigma2fshat, Mo.sigma2fshat, rtol=RTOL_UNIROOT, atol=1e-12)
    np.testing.assert_allclose(Mg.alphahat, Mo.alphahat, rtol=RTOL_UNIROOT)
    pg, po = F.predict(Mg, obs_fs=True), O.predict(Mo, obs_fs=True)
    close(pg["mu"], po[0], RTOL_UNIROOT)
    close(pg["var"], po[1], RTOL_UNIROOT)
    assert Mo.sigma2fshat > 0
    pg2, po2 = F.predict(Mg, obs_fs=False), O.predict(Mo, obs_fs=False)      # literal joint solve in the oracle (predict_general)
    close(pg2["mu"], po2[0], RTOL)
    close(pg2["var"], po2[1], RTOL)
    # known sigma2fs: no root search, so the full 1e-9 bar applies
    Mo2 = O.SRE(ob, obaus.centroids(), fs, None, z[keep], std[keep], K_type="unstructured", C=Cm, allow_overlap=True)
    Mg2 = F.SRE("z ~ 1", F.SpatialPointsDataFrame(xy, {"z": z, "std": std}), _mk_basis(F, ob), BAUs, est_error=False,
                average_in_BAU=False, K_type="unstructured")
    O.SRE_fit(Mo2, n_EM=3, tol=0.0, known_sigma2fs=0.05)
    F.SRE_fit(Mg2, n_EM=3, tol=0.0, known_sigma2fs=0.05)
    np.testing.assert_allclose(Mg2.info_fit["llk"], Mo2.info_fit["llk"], rtol=RTOL)
    np.testing.assert_allclose(Mg2.alphahat, Mo2.alphahat, rtol=RTOL)
    close(Mg2.get("mu_eta"), Mo2.mu_eta)
    close(Mg2.get("S_eta"), Mo2.S_eta)


# ---------------------------------------------------------------------------------------------------------
# BASELINE config 5 at scale: basis functions without compact support, S and S0 never stored (evaluate-and-accumulate)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K_type", ["unstructured", "block-exponential"])
def test_implicit_basis_matrix_matches_stored(dev, K_type):
    """SRE(implicit_basis = TRUE): the rows of S / S0 (Matern-3/2 on the sphere: dense) are evaluated in chunks inside the Gram and
    quadratic-form passes and dropped again (frk_model_create_implicit / frk_rowbuckets_create_implicit).  Same numbers as the stored
    dense path and as the oracle: log-likelihood trace, estimates, predictions under both obs_fs."""
    import frk_b200 as F
    from cases import pseudo_isea3h, sphere_points
    ll, z = sphere_points(4000, seed=29)
    iso = pseudo_isea3h()
    ob = O.auto_basis(O.sphere(), ll, nres=3, type="Matern32", isea3h=iso, isea3h_lo=1)       # r = 32 + 92 + 272 = 396
    obaus = O.GridBAUs(np.arange(-177.5, 180.0, 5.0), np.arange(-87.5, 90.0, 5.0), np.array([5.0, 5.0]))
    bau = O.points_to_BAUs(ll, obaus)
    std = np.full(len(z), 0.3)
    uo, means, _ = O.bin_mean(bau, np.column_stack([z, std]))
    Mo = O.SRE(ob, obaus.centroids(), np.ones(obaus.n), uo, means[:, 0], means[:, 1], K_type=K_type)
    gb = F.GridBAUs(obaus.xgrid.copy(), obaus.ygrid.copy(), np.array([5.0, 5.0]))
    gb["fs"] = 1.0
    data = F.SpatialPointsDataFrame(ll, {"z": z, "std": std})
    Mi = F.SRE("z ~ 1", data, _mk_basis(F, ob), gb, est_error=False, K_type=K_type, implicit_basis=True)
    Ms = F.SRE("z ~ 1", data, _mk_basis(F, ob), gb, est_error=False, K_type=K_type, implicit_basis=False)
    assert Mi.S.val is None and Mi.S0.val is None and Mi.S.nnz == Mi.m * Mi.r and Ms.S.val is not None
    O.SRE_fit(Mo, n_EM=4, tol=0.0)
    F.SRE_fit(Mi, n_EM=4, tol=0.0)
    F.SRE_fit(Ms, n_EM=4, tol=0.0)
    np.testing.assert_allclose(Mi.info_fit["llk"], Mo.info_fit["llk"], rtol=RTOL)
    np.testing.assert_allclose(Mi.info_fit["llk"], Ms.info_fit["llk"], rtol=RTOL)      # (row norms are summed in a different order)
    np.testing.assert_allclose(Mi.sigma2fshat, Mo.sigma2fshat, rtol=RTOL_UNIROOT, atol=1e-12)
    close(Mi.get("mu_eta"), Ms.get("mu_eta"), 1e-8)
    for obs_fs in (True, False):
        po = O.predict(Mo, obs_fs=obs_fs)
        pi, ps = F.predict(Mi, obs_fs=obs_fs), F.predict(Ms, obs_fs=obs_fs)
        close(pi["mu"], po[0], RTOL)
        close(pi["var"], po[1], RTOL)
        close(pi["var"], ps["var"], 1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("n,nres", [(40000, 2), (60000, 3), (3000, 3)])
def test_quadform_warp_specialised_kernel_matches_v2_and_scipy(dev, n, nres, monkeypatch):
    """rowSums((S M) * S) and S mu (.batch_compute_var, R/SREutils.R:286-300; Omega_diag1, R/SREfit.R:332-334) through the
    warp-specialised kernel (bulk-copy producers / DMMA consumers, FRK_QUAD_MODE=ws): q bit-identical to the second-generation
    kernel (same arithmetic, same order) and equal to SciPy.  Long buckets (several tiles per producer/consumer pair, buffers and
    mbarrier phases wrap), many short buckets per CTA (phases carried across buckets), and buckets with ragged last tiles."""
    import scipy.sparse as sp
    import frk_b200 as F
    from frk_b200._lib import call
    from test_gpu_parity import _dense_case
    rng = np.random.default_rng(n + nres)
    xy, _ = readme_points(n, seed=nres)
    ob = O.auto_basis(O.plane(), xy, regular=1, nres=nres)
    So = O.normalise_rows(O.eval_basis_fast(ob, xy)).tocsr()
    Sg = F.eval_basis_device(_mk_basis(F, ob), xy, dev, normalise=True)
    m, r = So.shape
    Mm = _dense_case(r, 3)
    mu = rng.standard_normal(r)
    q_ref = np.asarray(So.multiply(So @ Mm).sum(axis=1)).ravel()
    t_ref = So @ mu
    dM, dmu = dev.upload(Mm), dev.upload(mu)
    bk = Sg.buckets(dev)
    out = {}
    for mode in ("v2", "ws"):
        monkeypatch.setenv("FRK_QUAD_MODE", mode)
        for with_mu in (True, False):
            q, t = dev.zeros(m), dev.zeros(m)
            call("frk_quadform_spmv_bucketed", dev.ctx, m, r, dev.ptr(Sg.rowptr), dev.ptr(Sg.col), dev.ptr(Sg.val), bk, dev.ptr(dM),
                 dev.ptr(dmu) if with_mu else None, dev.ptr(q), dev.ptr(t) if with_mu else None)
            out[mode, with_mu] = (dev.download(q), dev.download(t))
    for with_mu in (True, False):
        assert np.array_equal(out["ws", with_mu][0], out["v2", with_mu][0])
        # t = S mu comes out of the tensor product in ws mode (a spare column of the staged matrix holds mu): other summation order
        np.testing.assert_allclose(out["ws", with_mu][1], out["v2", with_mu][1], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(out["ws", True][0], q_ref, rtol=1e-11)
    np.testing.assert_allclose(out["ws", True][1], t_ref, rtol=1e-10, atol=1e-13)


@pytest.mark.gpu
def test_quadform_warp_specialised_kernel_long_rows_and_empty_rows(dev, monkeypatch):
    """The side paths of the warp-specialised quadratic form: tiles holding a row of more than 40 non-zeros (evaluated directly by the
    producer warp while the consumers skip the tile), rows without any non-zero, ragged last tiles, buckets of every width up to 60
    columns (block counts 1..8 with the spare column) -- on a synthetic matrix, against SciPy and the second-generation kernel."""
    import scipy.sparse as sp
    import frk_b200 as F
    from frk_b200._lib import call
    from frk_b200.basis import DeviceCSR
    from test_gpu_parity import _dense_case
    rng = np.random.default_rng(77)
    n, r = 30011, 60
    rows, cols, vals = [], [], []
    for i in range(n):
        u = rng.random()
        hi = int(rng.integers(1, r + 1))                       # the row lives on columns < hi: the bucket's union is at most hi wide
        k = 0 if u < 0.01 else (min(hi, int(rng.integers(41, 56))) if u < 0.06 else min(hi, int(rng.integers(1, 31))))
        c = np.sort(rng.choice(hi, size=k, replace=False))
        rows += [i] * k; cols += c.tolist(); vals += rng.standard_normal(k).tolist()
    So = sp.csr_matrix((vals, (rows, cols)), shape=(n, r))
    So.sort_indices()
    Sg = DeviceCSR(n, r, dev.upload(So.indptr.astype(np.int32)), dev.upload(So.indices.astype(np.int32)), dev.upload(So.data), So.nnz)
    Mm = _dense_case(r, 5)
    mu = rng.standard_normal(r)
    q_ref = np.asarray(So.multiply(So @ Mm).sum(axis=1)).ravel()
    t_ref = So @ mu
    dM, dmu = dev.upload(Mm), dev.upload(mu)
    bk = Sg.buckets(dev)
    out = {}
    for mode in ("v2", "ws"):
        monkeypatch.setenv("FRK_QUAD_MODE", mode)
        q, t = dev.zeros(n) + 7.0, dev.zeros(n) + 7.0          # (every entry must be written, empty rows included)
        call("frk_quadform_spmv_bucketed", dev.ctx, n, r, dev.ptr(Sg.rowptr), dev.ptr(Sg.col), dev.ptr(Sg.val), bk, dev.ptr(dM),
             dev.ptr(dmu), dev.ptr(q), dev.ptr(t))
        out[mode] = (dev.download(q), dev.download(t))
    scale = abs(q_ref).max()
    np.testing.assert_allclose(out["ws"][0], q_ref, rtol=1e-11, atol=1e-13 * scale)
    np.testing.assert_allclose(out["ws"][1], t_ref, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(out["ws"][0], out["v2"][0], rtol=1e-13, atol=1e-14 * scale)


@pytest.mark.gpu
@pytest.mark.parametrize("lanes", [1, 2])
def test_block_exponential_inverse_memo_changes_nothing(dev, lanes, monkeypatch):
    """The memo of Nelder-Mead inverses (model.cu f_tau_lookup: exp(-D/tau)^-1 and its log-determinant are functions of tau alone; the
    next M-step's start value is the previous winner read back from K, R/SREfit.R:596-614) against the literal evaluation of every
    point (FRK_NO_KINV_REUSE=1, what the reference does): same number of Nelder-Mead requests, same tau / omega bit for bit or to
    the last digits, same log-likelihood trace and predictions; and the oracle agrees with both.  The memo must actually be used."""
    import frk_b200 as F
    from test_gpu_parity import _setup_plane
    dev.set_deterministic(True)                                 # (atomics' last-bit noise would hide what is compared here)
    try:
        fits = {}
        for off in (True, False):
            if off:
                monkeypatch.setenv("FRK_NO_KINV_REUSE", "1")
            else:
                monkeypatch.delenv("FRK_NO_KINV_REUSE", raising=False)
            Mo, Mg = _setup_plane(F, m=3000, seed=11, nres=3, K_type="block-exponential")
            reused = 0.0
            llk, taus, reqs = [], [], []
            for _ in range(12):                                 # one iteration per call: the statistics are those of the last M-step
                F.SRE_fit(Mg, n_EM=1, tol=0.0, lanes=lanes)
                st = Mg.info_fit["nm_stats"]
                reused += st["reused"]; reqs.append(st["requests"])
                llk.append(float(Mg.info_fit["llk"][-1])); taus.append(list(Mg.info_fit["tau"]))
            pr = F.predict(Mg, obs_fs=False)
            fits[off] = dict(llk=np.array(llk), tau=np.array(taus), reqs=reqs, reused=reused, mu=pr["mu"], var=pr["var"], Mo=Mo)
    finally:
        dev.set_deterministic(False)
    a, b = fits[True], fits[False]
    assert a["reused"] == 0 and b["reused"] >= 10, (a["reused"], b["reused"])
    assert a["reqs"] == b["reqs"]
    np.testing.assert_allclose(b["tau"], a["tau"], rtol=1e-12)
    np.testing.assert_allclose(b["llk"], a["llk"], rtol=1e-12)
    close(b["mu"], a["mu"], 1e-11)
    close(b["var"], a["var"], 1e-11)
    Mo = b["Mo"]
    O.SRE_fit(Mo, n_EM=12, tol=0.0)
    np.testing.assert_allclose(b["llk"][-1], Mo.info_fit["llk"][-1], rtol=RTOL)

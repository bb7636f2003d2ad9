"""Pins the CPU oracle to every known-answer test the reference's own tests hold for the hot path
 and to independent algebra.  Citations are tests/testthat/<file>:<line> in the reference."""
import json
import math
import os

import numpy as np
import pytest
import scipy.sparse as sp

from cases import readme_points
from oracle import frk_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_kat1_gaussian_basis_at_distance_one():
    """test_basis.R:23-24: G2(s=(0,1)) with mu=(1,1), sigma=1 is identical to dnorm(1)*sqrt(2*pi) = exp(-0.5); same on the real line."""
    G2 = O.local_basis(O.plane(), np.array([[1.0, 1.0]]), [1.0], "Gaussian")
    assert O.eval_basis_dense(G2, np.array([[0.0, 1.0]]))[0, 0] == math.exp(-0.5)
    G1 = O.local_basis(O.real_line(), np.array([[1.0]]), [1.0], "Gaussian")
    assert O.eval_basis_dense(G1, np.array([[0.0]]))[0, 0] == math.exp(-0.5)


def test_kat2_value_one_at_centroids():
    """test_basis.R:51-52,62-63: diag(eval_basis(G, centroids)) == 1 and the result is sparse."""
    xy, _ = readme_points(300, 2)
    for t in O.BASIS_TYPES:
        G = O.auto_basis(O.plane(), xy, nres=2, type=t)
        S = O.eval_basis(G, G.loc)
        assert sp.issparse(S)
        assert np.all(S.diagonal() == 1.0)


def test_kat3_sphere_distances():
    """test_domains.R:25-36: unit-sphere distance (-90,0)<->(90,0) identical to pi; space-time identical to sqrt(pi^2 + 2^2)."""
    d = O.dist_sphere(np.array([[-90.0, 0.0]]), np.array([[90.0, 0.0]]), R=1.0)
    assert d[0, 0] == math.pi
    dt = O.gc_dist_time(np.array([[-90.0, 0.0, 0.0]]), np.array([[90.0, 0.0, 2.0]]), R=1.0)
    assert dt[0, 0] == math.sqrt(math.pi ** 2 + 2 ** 2)


def test_kat4_distR_vs_pairwise():
    """test_other.R:114-120: distR(x, x) agrees with stats::dist (sum of differences < 1e-12)."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((40, 3))
    D = O.distR(x, x)
    ref = np.sqrt(((x[:, None, :] - x[None, :, :]) ** 2).sum(-1))
    assert abs(D - ref).sum() < 1e-12


def test_kat6_linalg_identities():
    """test_linalg.R:4-25: tr, diag2(A,A) == diag(A %*% A), chol2inv / cholesky helpers == solve(A) on a 2x2."""
    A = np.array([[2.0, 0.5], [0.5, 1.0]])
    assert O.tr(A) == 3.0
    assert np.allclose(O.diag2(A, A), np.diag(A @ A))
    R = O.chol_upper(A)
    assert np.allclose(R.T @ R, A)
    assert np.allclose(O.chol2inv(R), np.linalg.inv(A))
    assert np.isclose(O.logdet(R), np.log(np.linalg.det(A)))


def test_kat7_isea3h_counts():
    """data/isea3h.rda: centroids per resolution 0..5 are 12/32/92/272/812/2432; nres=3 from isea3h_lo=2 gives 1176 basis
    functions (R/basisfns.R:191)."""
    z = np.load(os.path.join(GOLD, "isea3h_centroids.npz"))
    assert np.bincount(z["res"]).tolist() == [12, 32, 92, 272, 812, 2432]
    isea = {r: np.column_stack([z["lon"][z["res"] == r], z["lat"][z["res"] == r]]) for r in range(6)}
    G = O.auto_basis(O.sphere(), np.array([[0.0, 0.0], [1.0, 1.0]]), nres=3, isea3h=isea, isea3h_lo=2)
    assert G.n == 1176 == 92 + 272 + 812


def test_kat8_tensor_basis_count_and_order():
    """test_basis.R:119: 10 x 3 tensor basis has 30 functions; column order is space-fastest (R/basisfns.R:821-822)."""
    G1 = O.local_basis(O.real_line(), np.arange(10.0)[:, None], np.full(10, 2.0), "bisquare")
    G2 = O.local_basis(O.real_line(), np.arange(3.0)[:, None], np.full(3, 2.0), "bisquare")
    T = O.TensorP_Basis(G1, G2)
    assert T.n == 30
    s = np.array([[3.2, 1.1], [7.0, 0.3]])
    S = O.eval_basis(T, s).toarray()
    S1 = O.eval_basis_dense(G1, s[:, :1]); S2 = O.eval_basis_dense(G2, s[:, 1:])
    for jt in range(3):
        assert np.array_equal(S[:, jt * 10:(jt + 1) * 10], S1 * S2[:, [jt]])


def test_structural_counts_from_reference_tests():
    """test_other.R:11-22 analogue (seq/by grid construction): 1 x 1 BAUs over [0,10]^2 with no buffer give 11 x 11 centroids;
    auto_basis(plane, nres=3) has 9+81+729 = 819 functions, regular=2 gives 3276 (R/basisfns.R:458-464)."""
    b = O.auto_BAUs_plane(np.array([[0.0, 0.0], [10.0, 10.0]]), cellsize=1.0, xlims=(0, 10), ylims=(0, 10))
    assert b.nx == 11 and b.ny == 11 and np.array_equal(b.xgrid, np.arange(11.0))
    u = np.array([[0.0, 0.0], [1.0, 1.0]])
    assert O.auto_basis(O.plane(), u, nres=3).n == 819
    assert O.auto_basis(O.plane(), u, nres=3, regular=2).n == 3276


def test_fast_eval_matches_literal_eval():
    xy, _ = readme_points(500, 3)
    G = O.auto_basis(O.plane(), xy, nres=3)
    A = O.eval_basis(G, xy).tocsr(); B = O.eval_basis_fast(G, xy).tocsr()
    A.sort_indices(); B.sort_indices()
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices) and np.array_equal(A.data, B.data)


def _c1(K_type="unstructured", hetero=False, m=300, nres=1, cellsize=0.1):
    xy, z = readme_points(m, 1)
    std = np.full(m, 0.5) if not hetero else 0.3 + 0.4 * np.random.default_rng(101).uniform(size=m)
    ob = O.auto_basis(O.plane(), xy, regular=1, nres=nres)
    ba = O.auto_BAUs_plane(xy, cellsize=cellsize)
    uo, means, _ = O.bin_mean(O.points_to_BAUs(xy, ba), np.column_stack([z, std]))
    return O.SRE(ob, ba.centroids(), np.ones(ba.n), uo, means[:, 0], means[:, 1], K_type=K_type)


def test_kat5_predict_var_equals_diag_of_joint_cov():
    """test_SRE.R:46-52: predict var == diag(Cov) for both obs_fs modes.  Here: the closed form used for point data equals the
    literal (r+N)-dimensional joint solve of R/SREpredict.R:308-333, and obs_fs=TRUE equals direct conditioning."""
    M = _c1(hetero=True)
    O.SRE_fit(M, n_EM=4, tol=0.0)
    assert M.sigma2fshat > 0
    mu_c, var_c, _ = O.predict(M, obs_fs=False)
    mu_g, var_g, _ = O.predict_general(M)
    np.testing.assert_allclose(var_c, var_g, rtol=1e-9)
    np.testing.assert_allclose(mu_c, mu_g, rtol=1e-9, atol=1e-11)


def test_loglik_equals_dense_gaussian_density():
    """Independent algebra: the Woodbury log-likelihood (.loglik.ind) equals the dense N(X alpha, S K S' + D) log-density."""
    M = _c1(hetero=True)
    O.SRE_fit(M, n_EM=2, tol=0.0)
    S = M.S.toarray()
    Sig = S @ M.Khat @ S.T + np.diag(M.sigma2fshat * M.Vfs + M.Ve)
    res = M.Z - M.X @ M.alphahat
    _, ld = np.linalg.slogdet(Sig)
    dense = -0.5 * len(res) * math.log(2 * math.pi) - 0.5 * ld - 0.5 * res @ np.linalg.solve(Sig, res)
    np.testing.assert_allclose(O.loglik(M), dense, rtol=1e-10)


def test_em_monotone_and_golden_trace():
    """EM never decreases the log-likelihood for the unstructured K; traces match the committed oracle fixture
    (tests/golden/c1_oracle.json, written by tools/make_golden.py -- pins the oracle against drift)."""
    gold = json.load(open(os.path.join(GOLD, "c1_oracle.json")))
    xy, z = readme_points(1000, 1)
    for key, g in gold.items():
        K, het = key.split("|")
        het = het.endswith("True")
        std = np.full(1000, 0.5) if not het else 0.3 + 0.4 * np.random.default_rng(101).uniform(size=1000)
        ob = O.auto_basis(O.plane(), xy, regular=1, nres=2)
        ba = O.auto_BAUs_plane(xy, cellsize=0.02)
        uo, means, _ = O.bin_mean(O.points_to_BAUs(xy, ba), np.column_stack([z, std]))
        M = O.SRE(ob, ba.centroids(), np.ones(ba.n), uo, means[:, 0], means[:, 1], K_type=K)
        O.SRE_fit(M, n_EM=6, tol=0.0)
        np.testing.assert_allclose(M.info_fit["llk"], g["llk"], rtol=1e-9)
        np.testing.assert_allclose(M.sigma2fshat, g["sigma2fs"], rtol=1e-7, atol=1e-12)
        if K == "unstructured":
            assert np.all(np.diff(M.info_fit["llk"]) > -1e-8)


def test_grid_index_matches_nearest_centre():
    """sp::over on SpatialPixels: a point belongs to the cell whose centre is nearest; outside -> NA."""
    b = O.auto_BAUs_plane(np.zeros((1, 2)), cellsize=0.25, xlims=(0, 2), ylims=(0, 1))
    rng = np.random.default_rng(1)
    xy = np.column_stack([rng.uniform(-0.1, 2.1, 2000), rng.uniform(-0.1, 1.1, 2000)])
    got = O.points_to_BAUs(xy, b)
    cen = b.centroids()
    d = abs(xy[:, None, :] - cen[None, :, :]).max(-1)
    inside = d.min(1) < 0.125
    assert np.array_equal(got[inside], d.argmin(1)[inside])
    assert np.all(got[d.min(1) > 0.125 + 1e-12] == -1)


def test_bin_mean_groups_and_means():
    bau = np.array([3, 1, 3, -1, 1, 7, 3], dtype=np.int32)
    cols = np.arange(14.0).reshape(7, 2)
    u, mth, c = O.bin_mean(bau, cols)
    assert u.tolist() == [1, 3, 7] and c.tolist() == [2, 3, 1]
    np.testing.assert_allclose(mth, [[(2 + 8) / 2, (3 + 9) / 2], [(0 + 4 + 12) / 3, (1 + 5 + 13) / 3], [10, 11]])


def test_point_in_polygon_classification_and_hexagon_tiling():
    """point.in.polygon semantics (0 outside, 1 inside, 2 edge, 3 vertex) on a square, and test_BAUs.R:121-139's property on the
    reference's own ISEA3H cells: every point (away from the dateline cells dropped from the fixture) falls in exactly one hexagon."""
    vx = np.array([0.0, 1.0, 1.0, 0.0, 0.0]); vy = np.array([0.0, 0.0, 1.0, 1.0, 0.0])
    assert O.in_poly(0.5, 0.5, vx, vy) == 1 and O.in_poly(1.5, 0.5, vx, vy) == 0
    assert O.in_poly(1.0, 0.5, vx, vy) == 2 and O.in_poly(1.0, 1.0, vx, vy) == 3 and O.in_poly(0.5, 0.0, vx, vy) == 2
    g = np.load(os.path.join(GOLD, "isea3h_res4_polygons.npz"))
    rng = np.random.default_rng(0)
    pts = np.column_stack([rng.uniform(-150, 150, 400), np.degrees(np.arcsin(rng.uniform(-0.9, 0.9, 400)))])[:120]
    idx = O.points_to_polygon_BAUs(pts, g["ptr"], g["vx"], g["vy"])
    assert (idx >= 0).all()
    for i in range(0, 120, 17):          # exactly one polygon contains the point
        hits = [k for k in range(len(g["ptr"]) - 1)
                if O.in_poly(pts[i, 0], pts[i, 1], g["vx"][g["ptr"][k]:g["ptr"][k + 1]], g["vy"][g["ptr"][k]:g["ptr"][k + 1]]) != 0]
        assert hits == [idx[i]]
    # the cell's own centroid is inside it
    cidx = O.points_to_polygon_BAUs(np.column_stack([g["cx"], g["cy"]])[::13], g["ptr"], g["vx"], g["vy"])
    assert np.array_equal(cidx, np.arange(len(g["cx"]))[::13])


def test_st_bau_index_time_slices():
    """map_data_to_BAUs(ST), R/geometryfns.R:1455-1470: slice i takes t1_i <= time < t1_{i+1}; the LAST slice takes everything from
    its start on; before the first start -> NA; a spatial NA stays NA; BAU numbering is space-fastest."""
    brk = np.array([0.0, 1.0, 2.5])
    t = np.array([-1.0, 0.0, 0.5, 1.0, 2.5, 9.0, np.nan, 2.4999])
    spi = np.array([0, 1, 2, 3, -1, 5, 1, 4])
    assert O.st_bau_index(spi, t, brk, 10).tolist() == [-1, 1, 2, 13, -1, 25, -1, 14]


def test_buildc_polygons_and_areal_sre_by_hand():
    """BuildC(SpatialPolygons), R/geometryfns.R:1572-1616, on a 4 x 4 grid of BAU centroids (0.5 .. 3.5): polygon A = [0,2] x [0,2]
    holds centroids (0.5|1.5, 0.5|1.5) = BAUs 0,1,4,5; polygon B = [2,4] x [0,1] holds BAUs 2,3; a centroid ON an edge counts.
    SRE() then normalises the rows (R/SRE.R:483-490): weights 1/4 and 1/2, Vfs = sum c^2 fs, X = BAU covariate averages."""
    gx, gy = np.meshgrid(np.arange(0.5, 4.0), np.arange(0.5, 4.0), indexing="xy")
    cen = np.column_stack([gx.ravel(), gy.ravel()])
    ptr = np.array([0, 5, 10, 15])
    vx = np.array([0, 2, 2, 0, 0, 2, 4, 4, 2, 2, 0.5, 1.5, 1.5, 0.5, 0.5], dtype=float)
    vy = np.array([0, 0, 2, 2, 0, 0, 0, 1, 1, 0, 2.5, 2.5, 3.5, 3.5, 2.5], dtype=float)        # third: edge through centroids 8,9,12,13
    i, j, x = O.BuildC_polygons(ptr, vx, vy, cen)
    assert i.tolist() == [0] * 4 + [1] * 2 + [2] * 4 and j.tolist() == [0, 1, 4, 5, 2, 3, 8, 9, 12, 13] and np.all(x == 1.0)
    import scipy.sparse as sp
    C = sp.csr_matrix((x, (i, j)), shape=(3, 16))
    basis = O.local_basis(O.plane(), np.array([[1.0, 1.0], [3.0, 3.0]]), [3.0, 3.0], "bisquare")
    fs = np.arange(1.0, 17.0)
    M = O.SRE(basis, cen, fs, None, np.array([1.0, 2.0, 3.0]), np.array([0.1, 0.2, 0.3]),
              X_BAU=np.column_stack([np.ones(16), cen[:, 0]]), K_type="unstructured", C=C)
    np.testing.assert_allclose(M.Vfs, [(1 + 2 + 5 + 6) / 16.0, (3 + 4) / 4.0, (9 + 10 + 13 + 14) / 16.0], rtol=1e-15)
    np.testing.assert_allclose(M.X[:, 1], [1.0, 3.0, 1.0], rtol=1e-15)
    np.testing.assert_allclose(M.S.toarray()[1], M.S0.toarray()[[2, 3]].mean(axis=0), rtol=1e-15)
    np.testing.assert_allclose(M.Ve, [0.01, 0.04, 0.09], rtol=1e-15)


def test_areal_case2_closed_form_equals_joint_solve():
    """The closed form of predict(obs_fs = FALSE) that the device implements for areal data with disjoint supports
    (frk_model_predict_areal) against the literal joint (eta, xi) solve of R/SREpredict.R:308-333."""
    import scipy.sparse as sp
    baus = O.auto_BAUs_plane(np.array([[0.0, 0.0], [1.0, 1.0]]), cellsize=0.1)
    cen = baus.centroids()
    N = len(cen)
    ptr, vx, vy = [0], [], []
    for iy in range(3):
        for ix in range(3):
            if (ix, iy) == (1, 1):
                continue
            x0, y0 = 0.3 * ix + 0.013, 0.3 * iy + 0.013
            vx += [x0, x0 + 0.3, x0 + 0.3, x0, x0]; vy += [y0, y0, y0 + 0.3, y0 + 0.3, y0]; ptr.append(len(vx))
    i, j, x = O.BuildC_polygons(np.array(ptr), np.array(vx), np.array(vy), cen)
    m = len(ptr) - 1
    C = sp.csr_matrix((x, (i, j)), shape=(m, N))
    rng = np.random.default_rng(3)
    fs = 1.0 + 0.5 * (cen[:, 0] > 0.5)
    M = O.SRE(O.auto_basis(O.plane(), cen, regular=1, nres=1), cen, fs, None, rng.standard_normal(m), np.full(m, 0.1),
              X_BAU=np.column_stack([np.ones(N), cen.sum(axis=1)]), K_type="unstructured", C=C)
    O.SRE_fit(M, n_EM=3, tol=0.0)
    s2, a = M.sigma2fshat, M.alphahat
    assert s2 > 0
    mu, var, _ = O.predict(M, obs_fs=False)                       # literal (r + N) x (r + N) solve
    d = s2 * M.Vfs + M.Ve
    Sig = O.chol2inv(O.chol_upper(O._gram(M.S, 1.0 / d) + M.Khat_inv))
    mut = Sig @ (M.S.T @ ((M.Z - M.X @ a) / d))
    q, QS = O.batch_compute_var(M.S0, Sig), O.batch_compute_var(M.S, Sig)
    res = M.Z - M.X @ a - M.S @ mut
    mu2, var2 = M.X_BAU @ a + M.S0 @ mut, q + s2 * fs
    V = Sig @ M.S.T.toarray()
    Cm = sp.csr_matrix(M.Cmat)
    S0d = M.S0.toarray()
    for jj in range(m):
        for kk in range(Cm.indptr[jj], Cm.indptr[jj + 1]):
            k, c = Cm.indices[kk], Cm.data[kk]
            g = s2 * fs[k] * c / d[jj]
            mu2[k] += g * res[jj]
            var2[k] = q[k] - 2 * g * (S0d[k] @ V[:, jj]) + g * g * QS[jj] + s2 * fs[k] - g * g * d[jj]
    np.testing.assert_allclose(mu2, mu, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(var2, var, rtol=1e-9)


def test_auto_basis_prune_rule():
    """auto_basis(prune > 0), R/basisfns.R:536-553: a resolution's function survives iff the column sum of the FIRST-pass basis over
    the data is >= prune; scales of the survivors are then re-derived (nearest surviving centroid), centroids keep their order."""
    rng = np.random.default_rng(5)
    xy = np.vstack([rng.uniform(0, 1, (800, 2)) ** 2, [[0.0, 0.0], [1.0, 1.0]]])
    full = O.auto_basis(O.plane(), xy, regular=1, nres=2)
    pr = O.auto_basis(O.plane(), xy, regular=1, nres=2, prune=12.0)
    cs = np.asarray(O.eval_basis(full, xy).sum(axis=0)).ravel()
    keep = cs >= 12.0
    assert 0 < keep.sum() < full.n and pr.n == keep.sum()
    assert np.array_equal(pr.loc, full.loc[keep]) and np.array_equal(pr.res, full.res[keep])
    # re-derived scales: 1.5 * 1.25 * distance to the nearest surviving centroid of the same resolution
    for rr in (1, 2):
        L = pr.loc[pr.res == rr]
        D = O.distR(L, L); np.fill_diagonal(D, np.inf)
        np.testing.assert_array_equal(pr.scale[pr.res == rr], 1.5 * (D.min(axis=1) * 1.25))


def test_polygon_prediction_closed_form_equals_joint_solve():
    """predict(newdata = <polygons>, obs_fs = FALSE) for point data: the closed form the device implements
    (frk_model_predict_polygons: rows of S0 weighted by c (1 - a_k), plus sum c^2 / (A_k + Q_k)) against the literal
    C_P [S0 I] Qx^-1 [S0 I]' C_P' of R/SREpredict.R:392-404; overlapping polygons."""
    import scipy.sparse as sp
    from cases import readme_points
    xy, z = readme_points(300, 1)
    std = np.full(300, 0.5)
    ob = O.auto_basis(O.plane(), xy, regular=1, nres=1)
    baus = O.auto_BAUs_plane(xy, cellsize=0.1)
    uo, means, _ = O.bin_mean(O.points_to_BAUs(xy, baus), np.column_stack([z, std]))
    cen = baus.centroids()
    M = O.SRE(ob, cen, np.ones(baus.n), uo, means[:, 0], means[:, 1], K_type="unstructured")
    O.SRE_fit(M, n_EM=3, tol=0.0)
    s2 = M.sigma2fshat
    assert s2 > 0
    ptr, vx, vy = [0], [], []
    for x0, y0, w, h in [(0.03, 0.03, 0.5, 0.5), (0.33, 0.23, 0.6, 0.4), (0.0, 0.0, 1.0, 1.0)]:
        vx += [x0, x0 + w, x0 + w, x0, x0]; vy += [y0, y0, y0 + h, y0 + h, y0]; ptr.append(len(vx))
    i, j, x = O.BuildC_polygons(np.array(ptr), np.array(vx), np.array(vy), cen)
    CP = sp.csr_matrix((x, (i, j)), shape=(3, len(cen)))
    CP = sp.csr_matrix(sp.diags(1.0 / np.asarray(CP.sum(axis=1)).ravel()) @ CP)
    mu, var, _ = O.predict_polygons(M, CP, obs_fs=False)
    N = len(cen)
    A = np.zeros(N); np.add.at(A, M.obs_bau, 1.0 / M.Ve)
    Q = 1.0 / (s2 * M.fs_BAU)
    Sig = O.chol2inv(O.chol_upper(O._gram(M.S, 1.0 / (s2 * M.Vfs + M.Ve)) + M.Khat_inv))
    CPM = sp.csr_matrix(sp.csr_matrix(CP.multiply(1.0 - A / (A + Q))) @ M.S0)
    var2 = np.asarray(CPM.multiply(CPM @ Sig).sum(axis=1)).ravel() + np.asarray(CP.multiply(CP) @ (1.0 / (A + Q))).ravel()
    np.testing.assert_allclose(var2, var, rtol=1e-10)
    np.testing.assert_allclose(CP @ O.predict(M, obs_fs=False)[0], mu, rtol=1e-10, atol=1e-12)


def test_widening_golden_fixture():
    """tests/golden/widening_oracle.json (tools/make_golden_widening.py): the oracle's outputs for areal data, prediction polygons,
    auto_basis pruning and space-time slicing -- pins the ORACLE against drift (these are not reference outputs)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_widening", os.path.join(os.path.dirname(GOLD), "..", "tools", "make_golden_widening.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    now = mod.build()
    gold = json.load(open(os.path.join(GOLD, "widening_oracle.json")))
    assert now.keys() == gold.keys()
    for sect in gold:
        for k, v in gold[sect].items():
            if isinstance(v, (int,)) or (isinstance(v, list) and v and isinstance(v[0], int)):
                assert now[sect][k] == v, (sect, k)
            else:
                np.testing.assert_allclose(now[sect][k], v, rtol=1e-8, err_msg=f"{sect}.{k}")


def test_general_branches_nondiagonal_Vfs_against_brute_force():
    """The oracle's restatement of the reference's GENERAL branches (non-diagonal Vfs: footprints sharing BAUs; R/SREfit.R:186-189,
    240-252, 296-328) pinned by independent algebra: the log-likelihood equals the dense Gaussian log-density of
    Z ~ N(X alpha, S K S' + D); the E-step equals Gaussian conditioning; J(sigma2fs) is minus the derivative of the profiled
    expected complete-data log-likelihood (alpha at its GLS value), checked by central differences."""
    import scipy.sparse as sp
    rng = np.random.default_rng(5)
    xy = rng.uniform(size=(60, 2))
    z = np.sin(4 * xy[:, 0]) + 0.3 * rng.standard_normal(60)
    std = 0.2 + 0.1 * rng.uniform(size=60)
    ob = O.auto_basis(O.plane(), xy, regular=1, nres=1)
    baus = O.auto_BAUs_plane(xy, cellsize=0.15)
    b = O.points_to_BAUs(xy, baus)
    keep = b >= 0
    m = int(keep.sum())
    Cm = sp.csr_matrix((np.ones(m), b[keep], np.arange(m + 1)), shape=(m, baus.n))
    fs = 0.5 + baus.centroids()[:, 1]
    M = O.SRE(ob, baus.centroids(), fs, None, z[keep], std[keep], K_type="unstructured", C=Cm, allow_overlap=True)
    assert M.Vmat is not None and np.count_nonzero(M.Vmat - np.diag(np.diag(M.Vmat))) > 0
    O.SRE_fit(M, n_EM=2, tol=0.0)
    S = np.asarray(M.S.todense())
    D = M.sigma2fshat * M.Vmat + np.diag(M.Ve)
    Sig = S @ M.Khat @ S.T + D
    res = M.Z - M.X @ M.alphahat
    ll = -0.5 * (m * np.log(2 * np.pi) + np.linalg.slogdet(Sig)[1] + res @ np.linalg.solve(Sig, res))
    assert abs(O.loglik(M) - ll) <= 1e-10 * abs(ll)
    O.Estep(M)
    mu = M.Khat @ S.T @ np.linalg.solve(Sig, res)
    Se = M.Khat - M.Khat @ S.T @ np.linalg.solve(Sig, S @ M.Khat)
    np.testing.assert_allclose(M.mu_eta, mu, rtol=1e-8, atol=1e-10 * abs(mu).max())
    np.testing.assert_allclose(M.S_eta, Se, rtol=1e-7, atol=1e-9 * abs(Se).max())

    E = M.S_eta + np.outer(M.mu_eta, M.mu_eta)

    def Qprof(s2):
        Dm = s2 * M.Vmat + np.diag(M.Ve)
        Di = np.linalg.inv(Dm)
        al = np.linalg.solve(M.X.T @ Di @ M.X, M.X.T @ Di @ (M.Z - S @ M.mu_eta))
        r = M.Z - M.X @ al
        Om = S @ E @ S.T - np.outer(S @ M.mu_eta, r) - np.outer(r, S @ M.mu_eta) + np.outer(r, r)
        return -0.5 * np.linalg.slogdet(Dm)[1] - 0.5 * np.trace(Di @ Om)

    # J as _Mstep_general evaluates it (re-stated here through the oracle's own pieces)
    def J(s2):
        Dm = s2 * M.Vmat + np.diag(M.Ve)
        Di = O.chol2inv(O.chol_upper(Dm))
        W = Di @ M.Vmat @ Di
        al = np.linalg.solve(M.X.T @ Di @ M.X, M.X.T @ Di @ (M.Z - S @ M.mu_eta))
        r = M.Z - M.X @ al
        tr2 = np.sum(((W @ S) @ E) * S) - 2.0 * np.sum(((W @ S) @ M.mu_eta) * r) + np.sum((W @ r) * r)
        return -(-0.5 * np.trace(Di @ M.Vmat) + 0.5 * tr2)

    for s2 in (0.02, 0.1, 0.5):
        h = 1e-6 * s2
        num = (Qprof(s2 + h) - Qprof(s2 - h)) / (2 * h)
        assert abs(-J(s2) - num) <= 1e-5 * max(abs(num), 1e-3)
    # and the M-step's root brackets a zero of that J within uniroot's tolerance (eps^0.25 = 1.2e-4 on sigma2fs)
    O.Mstep(M)
    if M.sigma2fshat > 0:
        lo, hi = max(M.sigma2fshat - 2.5e-4, 1e-12), M.sigma2fshat + 2.5e-4
        assert np.sign(J(lo)) != np.sign(J(hi))

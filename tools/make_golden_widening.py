"""Writes tests/golden/widening_oracle.json: ORACLE outputs for the paths added after the first cut (areal data, prediction polygons,
auto_basis pruning, space-time slicing) -- pins the oracle against drift.  NOT reference outputs (R cannot run here); needs no
/root/reference.  python tools/make_golden_widening.py"""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import areal_tiles  # noqa: E402
from oracle import frk_oracle as O  # noqa: E402


def build():
    out = {}
    # areal data, supports that share no BAU
    baus = O.auto_BAUs_plane(np.array([[0.0, 0.0], [1.0, 1.0]]), cellsize=0.04)
    cen = baus.centroids()
    ptr, vx, vy, z, std = areal_tiles(tile=0.2, noise=1.0)
    i, j, x = O.BuildC_polygons(ptr, vx, vy, cen)
    C = sp.csr_matrix((x, (i, j)), shape=(len(z), len(cen)))
    fs = 1.0 + 0.5 * (cen[:, 0] > 0.5)
    M = O.SRE(O.auto_basis(O.plane(), cen, regular=1, nres=2), cen, fs, None, z, std,
              X_BAU=np.column_stack([np.ones(len(cen)), cen[:, 0] + cen[:, 1]]), C=C)
    O.SRE_fit(M, n_EM=4, tol=0.0)
    p1, p2 = O.predict(M, obs_fs=True), O.predict(M, obs_fs=False)
    out["areal"] = dict(m=len(z), N=len(cen), nnzC=int(C.nnz), llk=M.info_fit["llk"].tolist(), sigma2fs=float(M.sigma2fshat),
                        alpha=M.alphahat.tolist(), var_sum_obs_fs=float(p1[1].sum()), var_sum_case2=float(p2[1].sum()),
                        mu_sum_case2=float(p2[0].sum()))
    # prediction polygons (overlapping) on the same model
    pp, px, py = [0], [], []
    for x0, y0, w, h in [(0.03, 0.03, 0.5, 0.5), (0.33, 0.23, 0.6, 0.4), (0.0, 0.0, 1.0, 1.0)]:
        px += [x0, x0 + w, x0 + w, x0, x0]; py += [y0, y0, y0 + h, y0 + h, y0]; pp.append(len(px))
    i, j, x = O.BuildC_polygons(np.array(pp), np.array(px), np.array(py), cen)
    CP = sp.csr_matrix((x, (i, j)), shape=(3, len(cen)))
    CP = sp.csr_matrix(sp.diags(1.0 / np.asarray(CP.sum(axis=1)).ravel()) @ CP)
    q = O.predict_polygons(M, CP, obs_fs=True)
    out["polygons"] = dict(members=[int(v) for v in np.diff(CP.indptr)], mu=q[0].tolist(), var=q[1].tolist())
    # auto_basis pruning
    rng = np.random.default_rng(41)
    xy = np.vstack([rng.uniform(0, 1, (3000, 2)) ** 2, [[0.0, 0.0], [1.0, 1.0]]])
    pb = O.auto_basis(O.plane(), xy, regular=1, nres=3, prune=2.0)
    out["prune"] = dict(n=int(pb.n), per_res=np.bincount(pb.res).tolist(), scale_sum=float(pb.scale.sum()), loc_sum=float(pb.loc.sum()))
    # space-time slicing
    t = rng.uniform(-1.0, 15.0, 2000)
    spi = rng.integers(-1, 50, 2000)
    idx = O.st_bau_index(spi, t, np.array([0.0, 1.0, 2.5, 2.75, 7.0, 11.0]), 50)
    out["st_index"] = dict(n_na=int((idx < 0).sum()), checksum=int((idx.astype(np.int64) * (np.arange(2000) % 97 + 1)).sum()))
    return out


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "widening_oracle.json")
    json.dump(build(), open(path, "w"), indent=1)
    print("written", path)

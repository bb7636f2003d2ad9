"""BASELINE config 5 at reduced size: points on the sphere, Matern-3/2 basis with the ISEA3H resolution counts 92+272+812+2432 = 3608
(Fibonacci lattices stand in for the DGGRID centroids), lon/lat grid BAUs.  Every row of S0 is dense -> the GEMM formulation of the
Gram and of the quadratic form (bucket.cu: gram_rhs_dense / quadform_dense).  python tools/c5_scaled.py [n_obs] [cell_deg]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import frk_b200 as F  # noqa: E402

n_obs = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
cell = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
dev = F.get_device()
rng = np.random.default_rng(5)
lon = rng.uniform(-180, 180, n_obs)
lat = np.degrees(np.arcsin(rng.uniform(-1, 1, n_obs)))
z = 380 + 3 * np.sin(np.radians(lat) * 2) + np.cos(np.radians(lon)) + 0.5 * rng.standard_normal(n_obs)
iso = {}
for res, k in zip((2, 3, 4, 5), (92, 272, 812, 2432)):
    i = np.arange(k) + 0.5
    iso[res] = np.column_stack([(np.degrees(i * np.pi * (3 - np.sqrt(5))) + 180.0) % 360.0 - 180.0, np.degrees(np.arcsin(1 - 2 * i / k))])
ll = np.column_stack([lon, lat])
t0 = time.time()
basis = F.auto_basis(F.sphere(), ll, nres=4, type="Matern32", isea3h=iso, isea3h_lo=2)
gb = F.GridBAUs(np.arange(-180 + cell / 2, 180, cell), np.arange(-90 + cell / 2, 90, cell), np.array([cell, cell]))
gb["fs"] = 1.0
data = F.SpatialPointsDataFrame(ll, {"z": z, "std": np.full(n_obs, 0.3)})
mdl = F.SRE("z ~ 1", data, basis, gb, est_error=False, K_type="unstructured")
dev.sync()
t1 = time.time()
F.SRE_fit(mdl, n_EM=2, tol=0.0)
dev.sync()
t2 = time.time()
F.SRE_fit(mdl, n_EM=5, tol=0.0)
dev.sync()
t3 = time.time()
p = F.predict(mdl, obs_fs=False)
t4 = time.time()
print(f"r={basis.n} N_BAU={len(gb)} m={mdl.m} S0 nnz={mdl.S0.nnz:.3e}: SRE {t1 - t0:.2f} s, EM {1e3 * (t3 - t2) / 5:.1f} ms/iteration "
      f"(first two {1e3 * (t2 - t1) / 2:.1f}), predict {1e3 * (t4 - t3):.1f} ms, llk {mdl.info_fit['llk'][-1]:.6f}, var[0] {p['var'][0]:.6g}")

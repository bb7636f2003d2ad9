"""Generates the committed fixtures under tests/golden/ (run in the build container, where /root/reference exists).

  isea3h_centroids.npz  ISEA3H centroids, resolutions 0..5, from the reference's data/isea3h.rda (R/datadoc.R:56-68)
  isea3h_res4_polygons.npz  the 780 resolution-4 cells that do not straddle the dateline, as polygon rings (hexagon BAUs)
  airs_day1.npz         AIRS_05_2003 day 1: lon, lat, co2avgret, co2std   (data/AIRS_05_2003.rda, R/datadoc.R:5-20)
  noaa_jul1993.npz      NOAA_df_1990, Tmax, July 1993: lon, lat, day, z    (data/NOAA_df_1990.rda; vignettes/FRK_intro.Rnw:690-700)
  c1_oracle.json        oracle outputs on BASELINE config 1 (log-likelihood traces etc.) -- pins the oracle against drift.
                        NOT reference outputs: R cannot run here; EM numerics are parity-unpinned by the reference's own tests.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from rda import data_frame, read_rda  # noqa: E402

REF = "/root/reference/data"
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

d = data_frame(read_rda(f"{REF}/isea3h.rda"))
keep = (d["centroid"] == 1) & (d["res"] <= 5)
np.savez_compressed(f"{OUT}/isea3h_centroids.npz", lon=d["lon"][keep], lat=d["lat"][keep], res=d["res"][keep].astype(np.int32))
print("isea3h centroids per res:", np.bincount(d["res"][d["centroid"] == 1]))

# ISEA3H resolution-4 cells as polygon rings (hexagon BAUs; auto_BAUs(sphere, type = "hex"), R/geometryfns.R:765-800), without the
# dateline-straddling cells that the reference splits with sf (R/geometryfns.R:1816-1935, out of scope)
m4 = d["res"] == 4
ids = d["id"][m4].astype(int); lon4 = d["lon"][m4]; lat4 = d["lat"][m4]; cen4 = d["centroid"][m4]
ptr, vx, vy, cx, cy, kept_ids = [0], [], [], [], [], []
for k_ in np.unique(ids):
    sel = (ids == k_) & (cen4 == 0)
    if lon4[sel].max() - lon4[sel].min() > 180:
        continue
    vx += lon4[sel].tolist(); vy += lat4[sel].tolist(); ptr.append(len(vx)); kept_ids.append(k_)
    c_ = (ids == k_) & (cen4 == 1); cx.append(lon4[c_][0]); cy.append(lat4[c_][0])
np.savez_compressed(f"{OUT}/isea3h_res4_polygons.npz", ptr=np.array(ptr, dtype=np.int32), vx=np.array(vx), vy=np.array(vy),
                    cx=np.array(cx), cy=np.array(cy), id=np.array(kept_ids, dtype=np.int32))
print("ISEA3H res-4 polygons kept:", len(kept_ids))

a = data_frame(read_rda(f"{REF}/AIRS_05_2003.rda"))
k = a["day"] == 1
np.savez_compressed(f"{OUT}/airs_day1.npz", lon=a["lon"][k], lat=a["lat"][k], co2avgret=a["co2avgret"][k], co2std=a["co2std"][k])
print("AIRS day 1 rows:", int(k.sum()), "days 1-3:", int((a["day"] <= 3).sum()))

n = data_frame(read_rda(f"{REF}/NOAA_df_1990.rda"))
proc = np.array(n["proc"])
k = (proc == "Tmax") & (n["month"] == 7) & (n["year"] == 1993)
np.savez_compressed(f"{OUT}/noaa_jul1993.npz", lon=n["lon"][k], lat=n["lat"][k], day=n["day"][k], z=n["z"][k], id=n["id"][k])
print("NOAA Tmax Jul 1993 rows:", int(k.sum()), "stations:", len(np.unique(n["id"][k])))

from cases import readme_points  # noqa: E402
from oracle import frk_oracle as O  # noqa: E402
out = {}
xy, z = readme_points(1000, 1)
for K in ("unstructured", "block-exponential"):
    for het in (False, True):
        std = np.full(1000, 0.5) if not het else 0.3 + 0.4 * np.random.default_rng(101).uniform(size=1000)
        ob = O.auto_basis(O.plane(), xy, regular=1, nres=2)
        ba = O.auto_BAUs_plane(xy, cellsize=0.02)
        uo, means, _ = O.bin_mean(O.points_to_BAUs(xy, ba), np.column_stack([z, std]))
        M = O.SRE(ob, ba.centroids(), np.ones(ba.n), uo, means[:, 0], means[:, 1], K_type=K)
        O.SRE_fit(M, n_EM=6, tol=0.0)
        mu, var, sd = O.predict(M, obs_fs=False)
        out[f"{K}|hetero={het}"] = dict(r=ob.n, N=ba.n, m=len(uo), llk=M.info_fit["llk"].tolist(), sigma2fs=M.sigma2fshat,
                                        alpha=M.alphahat.tolist(), mu_sum=float(mu.sum()), var_sum=float(var.sum()),
                                        mu_eta_norm=float(np.linalg.norm(M.mu_eta)))
json.dump(out, open(f"{OUT}/c1_oracle.json", "w"), indent=1)
print("c1_oracle.json written")

"""Fixtures for BASELINE config 2 at its stated size (run in the build container, where /root/reference exists):

  isea3h_res6_polygons.npz  the ISEA3H resolution-6 cells (7292 in data/isea3h.rda, R/datadoc.R:56-68) that do not straddle the
                            dateline, as polygon rings -- the BAUs of auto_BAUs(sphere(), type = "hex", isea3h_res = 6)
                            (R/geometryfns.R:765-800; the dateline split with sf, :1816-1935, is BAU construction: out of scope)
  airs_days123.npz          AIRS_05_2003 days 1-3: all 43,059 soundings (lon, lat, co2avgret, co2std, day); float32 would lose bits,
                            so the columns are stored as float64 (zip-compressed)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from rda import data_frame, read_rda  # noqa: E402

REF = "/root/reference/data"
OUT = os.path.join(ROOT, "tests", "golden")

d = data_frame(read_rda(f"{REF}/isea3h.rda"))
m6 = d["res"] == 6
ids = d["id"][m6].astype(np.int64); lon = d["lon"][m6]; lat = d["lat"][m6]; cen = d["centroid"][m6]
order = np.argsort(ids, kind="stable")
ids, lon, lat, cen = ids[order], lon[order], lat[order], cen[order]
bounds = np.flatnonzero(np.diff(ids)) + 1
ptr, vx, vy, cx, cy, kept = [0], [], [], [], [], []
for a, b in zip(np.r_[0, bounds], np.r_[bounds, len(ids)]):
    v = cen[a:b] == 0
    lo, la = lon[a:b][v], lat[a:b][v]
    if lo.max() - lo.min() > 180:
        continue
    vx += lo.tolist(); vy += la.tolist(); ptr.append(len(vx)); kept.append(int(ids[a]))
    c = cen[a:b] == 1
    cx.append(float(lon[a:b][c][0])); cy.append(float(lat[a:b][c][0]))
np.savez_compressed(f"{OUT}/isea3h_res6_polygons.npz", ptr=np.array(ptr, dtype=np.int32), vx=np.array(vx), vy=np.array(vy),
                    cx=np.array(cx), cy=np.array(cy), id=np.array(kept, dtype=np.int32))
print("ISEA3H res-6 cells:", len(np.unique(ids)), "kept (not straddling the dateline):", len(kept))

a = data_frame(read_rda(f"{REF}/AIRS_05_2003.rda"))
k = a["day"] <= 3
np.savez_compressed(f"{OUT}/airs_days123.npz", lon=a["lon"][k], lat=a["lat"][k], co2avgret=a["co2avgret"][k], co2std=a["co2std"][k],
                    day=a["day"][k].astype(np.int32))
print("AIRS days 1-3 rows:", int(k.sum()))

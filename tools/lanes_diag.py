"""Small block-exponential fit with concurrent Nelder-Mead lanes; used to exercise the host-thread path under ncu /
compute-sanitizer.  usage: python tools/lanes_diag.py <lanes> <graphs 0|1> [n_EM]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import frk_b200 as F  # noqa: E402
from frk_b200._lib import call  # noqa: E402

lanes, graphs = int(sys.argv[1]), int(sys.argv[2])
n_em = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = F.get_device()
call("frk_use_graphs", dev.ctx, graphs)
rng = np.random.default_rng(1)
m = 2000
xy = rng.uniform(0, 1, (m, 2))
z = np.sin(8 * xy[:, 0]) + np.cos(8 * xy[:, 1]) + 0.5 * rng.standard_normal(m)
data = F.SpatialPointsDataFrame(xy, {"z": z, "std": np.full(m, 0.5)})
baus = F.auto_BAUs(F.plane(), cellsize=0.02, data=data)
baus["fs"] = 1.0
basis = F.auto_basis(F.plane(), xy, regular=1, nres=2)
mdl = F.SRE("z ~ 1", data, basis, baus, est_error=False, K_type="block-exponential")
F.SRE_fit(mdl, n_EM=n_em, tol=0.0, lanes=lanes)
print("lanes", lanes, "graphs", graphs, "llk", mdl.info_fit["llk"][-1], mdl.info_fit["nm_stats"], flush=True)

"""Wall time of every EM iteration from .EM_initialise at the bench workload, with the Nelder-Mead statistics of its M-step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import frk_b200 as F
import bench
args = bench.parse()
dev = F.get_device(0)
xy, z, std = bench.workload(args)
data = F.SpatialPointsDataFrame(xy, {"z": z, "std": std})
basis = F.auto_basis(F.plane(), np.array([[0.0, 0.0], [1.0, 1.0]]), regular=args.regular, nres=args.nres)
BAUs = F.auto_BAUs(F.plane(), cellsize=1.0 / (args.grid - 1), xlims=(0, 1), ylims=(0, 1)); BAUs["fs"] = 1.0
for rep in range(2):
    mdl = F.SRE("z ~ 1", data, basis, BAUs, est_error=False, K_type=args.k_type, dev=dev)
    dev.sync()
    out = []
    for it in range(args.steps):
        t0 = time.perf_counter(); F.SRE_fit(mdl, n_EM=1, tol=0.0, lanes=args.lanes); dev.sync(); dt = time.perf_counter() - t0
        st = mdl.info_fit.get("nm_stats", {})
        out.append((1e3 * dt, st.get("requests"), st.get("rounds"), st.get("evaluated")))
    print("pass", rep, " ".join(f"{a:.1f}ms/{int(b or 0)}req/{int(c or 0)}rnd/{int(d or 0)}ev" for a, b, c, d in out))
    del mdl

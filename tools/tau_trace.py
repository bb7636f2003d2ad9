"""tau of every resolution after each EM iteration at the bench workload (is the start value of the next Nelder-Mead run stationary?)."""
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
mdl = F.SRE("z ~ 1", data, basis, BAUs, est_error=False, K_type=args.k_type, dev=dev)
prev = None
for it in range(args.steps):
    t0 = time.perf_counter(); F.SRE_fit(mdl, n_EM=1, tol=0.0, lanes=args.lanes); dev.sync(); dt = time.perf_counter() - t0
    tau = np.array([float(np.atleast_1d(t)[0]) for t in mdl.info_fit["tau"]])
    rel = "" if prev is None else " rel.change " + " ".join(f"{(a - b) / b:+.2e}" for a, b in zip(tau, prev))
    print(f"it {it + 1:2d} {1e3 * dt:6.2f} ms tau " + " ".join(f"{t:.17g}" for t in tau) + rel, flush=True)
    prev = tau

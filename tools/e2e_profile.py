"""Wall-clock breakdown of the end-to-end path (SRE() from host arrays -> SRE_fit -> slots to host -> predict to host)."""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import frk_b200 as F
import bench
args = bench.parse()
dev = F.get_device(0)
xy, z, std = bench.workload(args)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
xy_p = pin(np.asfortranarray(xy).T.copy()).T
data = F.SpatialPointsDataFrame(xy_p, {"z": pin(z), "std": pin(std)})
basis = F.auto_basis(F.plane(), np.array([[0.0, 0.0], [1.0, 1.0]]), regular=args.regular, nres=args.nres)
BAUs = F.auto_BAUs(F.plane(), cellsize=1.0 / (args.grid - 1), xlims=(0, 1), ylims=(0, 1)); BAUs.fields["fs"] = pin(np.ones(len(BAUs)))
def run():
    t = [time.perf_counter()]
    m = F.SRE("z ~ 1", data, basis, BAUs, est_error=False, K_type=args.k_type, dev=dev); dev.sync(); t.append(time.perf_counter())
    F.SRE_fit(m, n_EM=args.steps, tol=0.0); t.append(time.perf_counter())
    s = {k: m.get(k) for k in ("mu_eta", "S_eta", "Khat", "Khat_inv")}; t.append(time.perf_counter())
    p = F.predict(m, obs_fs=False, to_host=True); t.append(time.perf_counter())
    return np.diff(t)
for i in range(3):
    d = run()
    print("SRE %.3f  fit %.3f  slots->host %.3f  predict->host %.3f  s" % tuple(d))
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)

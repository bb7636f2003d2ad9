"""Per-kernel device time of the whole C3 pipeline (SRE() build, EM iterations, predict) via the in-library profiler."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import frk_b200 as F
from frk_b200._lib import call
import bench
args = bench.parse()
dev = F.get_device(0)
xy, z, std = bench.workload(args)
data = F.SpatialPointsDataFrame(xy, {"z": z, "std": std})
basis = F.auto_basis(F.plane(), np.array([[0.0, 0.0], [1.0, 1.0]]), regular=args.regular, nres=args.nres)
BAUs = F.auto_BAUs(F.plane(), cellsize=1.0 / (args.grid - 1), xlims=(0, 1), ylims=(0, 1)); BAUs["fs"] = 1.0
def report(tag):
    buf = C.create_string_buffer(1 << 16)
    call("frk_profile_report", dev.ctx, buf, len(buf))
    rep = sorted(json.loads(buf.value.decode()), key=lambda k: -k["ms"])
    print(f"== {tag}: {sum(k['ms'] for k in rep):.2f} ms in {sum(k['launches'] for k in rep)} launches")
    for k in rep[:14] + [k for k in rep[14:] if "gram" in k["kernel"] or "quadform" in k["kernel"] or "eval_" in k["kernel"]]:
        print(f"   {k['kernel'][:52]:52s} n={k['launches']:5d} {k['ms']:9.3f} ms")
for rep_i in range(2):
    call("frk_profile", dev.ctx, 1)
    t0 = time.time(); mdl = F.SRE("z ~ 1", data, basis, BAUs, est_error=False, K_type=args.k_type, dev=dev); dev.sync(); t1 = time.time()
    mdl.S.buckets(dev); mdl.S0.buckets(dev); dev.sync(); t2 = time.time()
    report(f"SRE() wall {t1-t0:.3f}s + buckets {t2-t1:.3f}s")
    call("frk_profile", dev.ctx, 0); call("frk_profile", dev.ctx, 1)
    t0 = time.time(); F.SRE_fit(mdl, n_EM=args.steps, tol=0.0); t1 = time.time()
    report(f"SRE_fit({args.steps}) wall {t1-t0:.3f}s")
    call("frk_profile", dev.ctx, 0); call("frk_profile", dev.ctx, 1)
    t0 = time.time(); F.predict(mdl, obs_fs=False, to_host=False); dev.sync(); t1 = time.time()
    report(f"predict wall {t1-t0:.3f}s")
    call("frk_profile", dev.ctx, 0)
    del mdl

"""A/B timing of the bucketed quadratic form at C3 (S0: 1.0e7 BAU rows, S: 9.5e5 observation rows): the second-generation kernel
(FRK_QUAD_MODE=v2) against the warp-specialised one (ws).  CUDA events on the library's stream, 5 repetitions after 2 warm-ups."""
import os, sys
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
mdl = F.SRE("z ~ 1", data, basis, BAUs, est_error=False, K_type="unstructured", dev=dev)
os.environ["FRK_QUAD_MODE"] = "v2"
F.SRE_fit(mdl, n_EM=1, tol=0.0)
r = mdl.r
rng = np.random.default_rng(1)
A = rng.standard_normal((r, r + 5)); Mm = A @ A.T + r * np.eye(r)
dM, dmu = dev.upload(Mm), dev.upload(rng.standard_normal(r))
res = {}
for name, Mx in (("S0", mdl.S0), ("S", mdl.S)):
    bk = Mx.buckets(dev)
    n = Mx.n
    for mode in (("ws",) if os.environ.get("FRK_QUAD_NCU") else ("v2", "ws", "v2", "ws")):
        os.environ["FRK_QUAD_MODE"] = mode
        q, t = dev.zeros(n), dev.zeros(n)
        ts = []
        for rep in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dev.sync(); e0.record(dev.stream)
            call("frk_quadform_spmv_bucketed", dev.ctx, n, r, dev.ptr(Mx.rowptr), dev.ptr(Mx.col), dev.ptr(Mx.val), bk, dev.ptr(dM),
                 dev.ptr(dmu), dev.ptr(q), dev.ptr(t))
            e1.record(dev.stream); dev.sync()
            ts.append(e0.elapsed_time(e1))
        qh = dev.download(q)
        key = (name, mode)
        if key in res:
            assert np.array_equal(res[key][1], qh)
        res[key] = (min(ts[2:]), qh)
        print(f"{name:3s} n={n:9d} mode={mode}: min {min(ts[2:]):.3f} ms  median {sorted(ts[2:])[2]:.3f} ms", flush=True)
    if not os.environ.get("FRK_QUAD_NCU"):
        print(f"{name}: ws == v2 bit-identical: {np.array_equal(res[name, 'ws'][1], res[name, 'v2'][1])}", flush=True)

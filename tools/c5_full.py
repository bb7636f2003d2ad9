"""BASELINE config 5 at its stated size (or a share of it): n_obs points uniform on the sphere, Matern-3/2 basis with the ISEA3H
resolution counts 92 + 272 + 812 + 2432 = 3608 (Fibonacci lattices stand in for the DGGRID centroids), lon/lat grid BAUs, EM +
prediction at every BAU.  The basis has no compact support, so S and S0 are dense: they are never stored (implicit_basis: rows are
evaluated in chunks inside every Gram / quadratic-form pass).  BAU rows are block-partitioned over the ranks.

  python tools/c5_full.py --n-obs 1250000 --nx 2236 --ny 1118                       # one GPU's share of the 8-GPU run
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/c5_full.py --n-obs 10000000 --nx 6325 --ny 3163

Prints one JSON line (rank 0): sizes, seconds per EM iteration, predicted BAUs/s, the log-likelihood trace."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--n-obs", type=int, default=1250000)
ap.add_argument("--nx", type=int, default=2236)
ap.add_argument("--ny", type=int, default=1118)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--k-type", default="unstructured")
ap.add_argument("--balance", default="observations", choices=["BAUs", "observations"])
args = ap.parse_args()

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import frk_b200 as F  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = F.get_device(local)
if world > 1:
    dev.init_comm()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


rng = np.random.default_rng(5)
n_obs = args.n_obs
lon = rng.uniform(-180, 180, n_obs)
lat = np.degrees(np.arcsin(rng.uniform(-1, 1, n_obs)))
z = 380 + 3 * np.sin(np.radians(lat) * 2) + np.cos(np.radians(lon)) + 0.5 * rng.standard_normal(n_obs)
iso = {}
for res, k in zip((2, 3, 4, 5), (92, 272, 812, 2432)):
    i = np.arange(k) + 0.5
    iso[res] = np.column_stack([(np.degrees(i * np.pi * (3 - np.sqrt(5))) + 180.0) % 360.0 - 180.0, np.degrees(np.arcsin(1 - 2 * i / k))])
ll = np.column_stack([lon, lat])
cx, cy = 360.0 / args.nx, 180.0 / args.ny
basis = F.auto_basis(F.sphere(), ll[:200000], nres=4, type="Matern32", isea3h=iso, isea3h_lo=2)
gb = F.GridBAUs(-180 + cx / 2 + cx * np.arange(args.nx), -90 + cy / 2 + cy * np.arange(args.ny), np.array([cx, cy]))
gb["fs"] = 1.0
data = F.SpatialPointsDataFrame(ll, {"z": z, "std": np.full(n_obs, 0.3)})
barrier()
t0 = time.time()
mdl = F.SRE("z ~ 1", data, basis, gb, est_error=False, K_type=args.k_type, implicit_basis=True, balance=args.balance)
barrier()
t1 = time.time()
F.SRE_fit(mdl, n_EM=1, tol=0.0)              # warm-up iteration (allocator pools, CUDA graphs)
barrier()
t2 = time.time()
F.SRE_fit(mdl, n_EM=args.steps, tol=0.0)
barrier()
t3 = time.time()
p = F.predict(mdl, obs_fs=False, to_host=False)
barrier()
t4 = time.time()
finite = bool(torch.isfinite(p["var"]).all().item() and (p["var"] > 0).all().item())
mem = torch.cuda.max_memory_allocated() / 2**30
free, total = torch.cuda.mem_get_info()
if rank == 0:
    print(json.dumps({"config": "C5", "n_gpus": world, "n_obs": n_obs, "r": basis.n, "n_bau": len(gb), "m_rank0": mdl.m,
                      "K_type": args.k_type, "implicit_basis": True, "balance": args.balance, "rows_rank0": [mdl.lo, mdl.hi], "sre_s": round(t1 - t0, 3), "first_iteration_s": round(t2 - t1, 3),
                      "em_s_per_iteration": round((t3 - t2) / args.steps, 4), "em_iterations": args.steps,
                      "predict_s": round(t4 - t3, 3), "predict_baus_per_s": len(gb) / (t4 - t3),
                      "llk": [float(v) for v in mdl.info_fit["llk"]], "sigma2fshat": mdl.sigma2fshat,
                      "predict_var_finite_positive": finite, "gpu_mem_used_gib_rank0": round((total - free) / 2**30, 2)}))
if world > 1:
    dist.destroy_process_group()

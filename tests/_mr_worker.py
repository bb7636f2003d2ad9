"""Worker of tests/test_gpu_multirank.py: one rank of a row-partitioned SRE -> SRE.fit(EM) -> predict run THROUGH THE LIBRARY.
Launched by torch.distributed.run.  --comm hook : reductions through the library's frk_allreduce_fn hook (gloo when the ranks share one
GPU, NCCL otherwise);  --comm nccl : the library's own communicator (frk_comm_init; torch.distributed only carries the unique id)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def problem():
    from cases import readme_points
    xy, z = readme_points(6000, seed=11)
    std = 0.3 + 0.4 * np.random.default_rng(12).uniform(size=len(z))
    return xy, z, std


def fit(F, dev, K_type, n_EM, distribute_nm=False, balance="BAUs"):
    xy, z, std = problem()
    if balance == "observations":            # make the load uneven over the BAU index: drop most of the points with y > 0.6
        keep = (xy[:, 1] < 0.6) | (np.arange(len(z)) % 5 == 0)
        xy, z, std = xy[keep], z[keep], std[keep]
    basis = F.auto_basis(F.plane(), np.array([[0.0, 0.0], [1.0, 1.0]]), regular=1, nres=3)
    BAUs = F.auto_BAUs(F.plane(), cellsize=0.01, xlims=(0, 1), ylims=(0, 1))
    BAUs["fs"] = 1.0
    mdl = F.SRE("z ~ 1", F.SpatialPointsDataFrame(xy, {"z": z, "std": std}), basis, BAUs, est_error=False, K_type=K_type, dev=dev,
                balance=balance)
    F.SRE_fit(mdl, n_EM=n_EM, tol=0.0, distribute_nm=distribute_nm)
    out = {"llk": np.asarray(mdl.info_fit["llk"]), "alpha": mdl.alphahat, "sigma2fs": np.array([mdl.sigma2fshat]),
           "mu_eta": mdl.get("mu_eta"), "S_eta_diag": np.diag(mdl.get("S_eta")).copy(), "lo": np.array([mdl.lo]), "hi": np.array([mdl.hi])}
    for obs_fs in (False, True):
        pr = F.predict(mdl, obs_fs=obs_fs)
        out[f"mu{int(obs_fs)}"] = np.asarray(pr["mu"]); out[f"var{int(obs_fs)}"] = np.asarray(pr["var"])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--comm", default="hook", choices=["hook", "nccl"])
    ap.add_argument("--same-gpu", action="store_true")
    ap.add_argument("--k-type", default="block-exponential")
    ap.add_argument("--n-em", type=int, default=4)
    ap.add_argument("--distribute-nm", action="store_true")
    ap.add_argument("--balance", default="BAUs")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    index = 0 if a.same_gpu else local
    torch.cuda.set_device(index)
    if a.comm == "hook" and not a.same_gpu:
        dist.init_process_group("nccl", device_id=torch.device("cuda", index))
    else:
        dist.init_process_group("gloo")
    import frk_b200 as F
    dev = F.get_device(index)
    if a.comm == "nccl":
        dev.init_comm()
    res = fit(F, dev, a.k_type, a.n_em, a.distribute_nm, a.balance)
    np.savez(f"{a.out}.rank{rank}.npz", **res)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

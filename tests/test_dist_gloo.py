"""world_size = 2 on CPU (gloo): the row-partitioned accumulation that the multi-GPU path performs -- each rank owns a
contiguous range of BAU rows and the observations in them, accumulates [S'D^-1 S | S'D^-1 rho | 2 scalars] and the M-step
moments locally, and ONE sum all-reduce combines them.  The per-rank arithmetic here is the
oracle's (the kernels need a GPU); what is tested is frk_b200's partitioning and reduction plumbing."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cases import readme_points
    from frk_b200.runtime import allreduce_host, partition_rows
    from oracle import frk_oracle as O
    xy, z = readme_points(2000, 5)
    std = 0.3 + 0.4 * np.random.default_rng(1).uniform(size=len(z))
    ob = O.auto_basis(O.plane(), xy, nres=2)
    ba = O.auto_BAUs_plane(xy, cellsize=0.05)
    bau = O.points_to_BAUs(xy, ba)
    lo, hi = partition_rows(ba.n, world, rank)
    mine = (bau >= lo) & (bau < hi)                                     # observations live with their BAU
    uo, means, _ = O.bin_mean(np.where(mine, bau, -1), np.column_stack([z, std]))
    S0 = O.normalise_rows(O.eval_basis(ob, ba.centroids()[lo:hi]))
    S = S0[uo - lo]
    Z, Ve = means[:, 0], means[:, 1] ** 2
    d = 0.2 * np.ones(len(Z)) + Ve
    alpha = np.array([0.3])
    rho = Z - alpha[0]
    r = ob.n
    acc = np.concatenate([O._gram(S.tocsr(), 1 / d).ravel(), S.T @ (rho / d), [np.sum(rho * rho / d), np.sum(np.log(d))], [len(Z)]])
    tot = allreduce_host(acc, "sum")
    mx = allreduce_host(np.array([Ve.max()]), "max")
    if rank == 0:
        np.save(out, np.concatenate([tot, mx]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partitioned_accumulation_equals_single_rank(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cases import readme_points
    from oracle import frk_oracle as O
    out = str(tmp_path / "acc.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    xy, z = readme_points(2000, 5)
    std = 0.3 + 0.4 * np.random.default_rng(1).uniform(size=len(z))
    ob = O.auto_basis(O.plane(), xy, nres=2)
    ba = O.auto_BAUs_plane(xy, cellsize=0.05)
    uo, means, _ = O.bin_mean(O.points_to_BAUs(xy, ba), np.column_stack([z, std]))
    S = O.normalise_rows(O.eval_basis(ob, ba.centroids()))[uo].tocsr()
    Z, Ve = means[:, 0], means[:, 1] ** 2
    d = 0.2 + Ve
    rho = Z - 0.3
    ref = np.concatenate([O._gram(S, 1 / d).ravel(), S.T @ (rho / d), [np.sum(rho * rho / d), np.sum(np.log(d))], [len(Z)], [Ve.max()]])
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)

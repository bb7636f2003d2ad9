"""Row-partitioned (multi-rank) runs through the library, checked against the single-rank run and the oracle.

* two ranks SHARING cuda:0 (works on a one-GPU box): the library's reduction hook over gloo -- partition_rows, the packed
  upper-triangle exchange of S'D^-1 S, the M-step moment reductions, rank-local prediction;
* with >= 2 GPUs: the library's own NCCL communicator (frk_comm_init), no torch in the data path; also with the Nelder-Mead
  candidates distributed over the ranks (ncclBroadcast of the winning K_i^-1)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
pytestmark = pytest.mark.gpu


def _launch(tmp_path, world, extra):
    out = str(tmp_path / "mr")
    port = 29600 + os.getpid() % 1500
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_mr_worker.py"), "--out", out] + extra
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return [dict(np.load(f"{out}.rank{k}.npz")) for k in range(world)]


def _check(ranks, single, oracle):
    def close(a, b, rtol):
        np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * np.abs(b).max())
    r0 = ranks[0]
    for rk in ranks[1:]:                                   # replicated r-sized state: identical on every rank
        for k in ("llk", "alpha", "sigma2fs", "mu_eta", "S_eta_diag"):
            np.testing.assert_array_equal(rk[k], r0[k])
    # N ranks vs one rank: same arithmetic up to the order of the partial Gram sums
    close(r0["llk"], single["llk"], 1e-12)
    close(r0["mu_eta"], single["mu_eta"], 1e-9)
    close(r0["S_eta_diag"], single["S_eta_diag"], 1e-9)
    close(r0["alpha"], single["alpha"], 1e-8)
    close(r0["sigma2fs"], single["sigma2fs"], 1e-7)        # uniroot (tol eps^0.25) in the heteroscedastic branch
    for key in ("mu0", "var0", "mu1", "var1"):
        full = np.concatenate([rk[key] for rk in ranks])
        assert [int(rk["lo"][0]) for rk in ranks] == sorted(int(rk["lo"][0]) for rk in ranks)
        close(full, single[key], 1e-8)
    # and against the oracle
    close(r0["llk"], oracle["llk"], 1e-9)
    close(np.concatenate([rk["var1"] for rk in ranks]), oracle["var1"], 1e-7)


@pytest.fixture(scope="module")
def baseline(dev):
    import frk_b200 as F
    import _mr_worker as W
    from oracle import frk_oracle as O
    out = {}
    for K_type in ("block-exponential", "unstructured"):
        single = W.fit(F, dev, K_type, 4)
        xy, z, std = W.problem()
        ob = O.auto_basis(O.plane(), np.array([[0.0, 0.0], [1.0, 1.0]]), regular=1, nres=3)
        ba = O.auto_BAUs_plane(xy, cellsize=0.01, xlims=(0, 1), ylims=(0, 1))
        uo, means, _ = O.bin_mean(O.points_to_BAUs(xy, ba), np.column_stack([z, std]))
        Mo = O.SRE(ob, ba.centroids(), np.ones(ba.n), uo, means[:, 0], means[:, 1], K_type=K_type)
        O.SRE_fit(Mo, n_EM=4, tol=0.0)
        out[K_type] = (single, {"llk": np.asarray(Mo.info_fit["llk"]), "var1": O.predict(Mo, obs_fs=True)[1]})
    return out


@pytest.mark.parametrize("K_type", ["block-exponential", "unstructured"])
def test_two_ranks_sharing_one_gpu_match_single_rank(baseline, tmp_path, K_type):
    ranks = _launch(tmp_path, 2, ["--comm", "hook", "--same-gpu", "--k-type", K_type])
    assert ranks[0]["hi"][0] == ranks[1]["lo"][0]
    _check(ranks, *baseline[K_type])


def test_work_balanced_partition_matches_single_rank(dev, tmp_path):
    """SRE(balance = "observations"): the row partition equalises work (observations weigh 13 BAU rows), so with most observations in
    the lower part of the domain rank 0 owns FEWER BAU rows; results equal the single-rank run of the same data."""
    import frk_b200 as F
    import _mr_worker as W
    single = W.fit(F, dev, "unstructured", 3, balance="observations")
    ranks = _launch(tmp_path, 2, ["--comm", "hook", "--same-gpu", "--k-type", "unstructured", "--n-em", "3", "--balance", "observations"])
    cut = int(ranks[0]["hi"][0])
    assert cut == int(ranks[1]["lo"][0]) and cut < (int(ranks[1]["hi"][0]) // 2) * 0.9
    np.testing.assert_allclose(ranks[0]["llk"], single["llk"], rtol=1e-12)
    for key in ("mu0", "var0", "mu1", "var1"):
        full = np.concatenate([rk[key] for rk in ranks])
        np.testing.assert_allclose(full, single[key], rtol=1e-8, atol=1e-8 * np.abs(single[key]).max())


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("extra", [[], ["--distribute-nm"]])
def test_library_nccl_communicator_two_gpus(baseline, tmp_path, extra):
    if _ngpu() < 2:
        pytest.skip("needs two GPUs (the in-library NCCL communicator cannot put two ranks on one device)")
    ranks = _launch(tmp_path, 2, ["--comm", "nccl", "--k-type", "block-exponential"] + extra)
    _check(ranks, *baseline["block-exponential"])

"""Host-side logic of the frk_b200 mirror (no GPU needed): optimisers, formula parsing, grid construction, basis placement,
row partitioning -- checked against the oracle's independent copies and against SciPy."""
import numpy as np
import pytest
import scipy.optimize as so

import frk_b200 as F
from frk_b200.optim import nmmin, nmmin_speculative, zeroin
from frk_b200.sre import _parse_formula
from oracle import frk_oracle as O


def test_nmmin_matches_oracle_port_and_finds_minimum():
    f = lambda x: (x[0] - 1.3) ** 2 + 0.5 * (x[1] + 0.7) ** 4 + 0.1 * x[0] * x[1]
    a = nmmin(f, [0.2, 0.1], maxit=500, reltol=1e-10)
    b = O.nmmin(f, [0.2, 0.1], maxit=500, reltol=1e-10)
    assert np.array_equal(a[0], b[0]) and a[1] == b[1] and a[2] == b[2]
    ref = so.minimize(f, [0.2, 0.1], method="Nelder-Mead", options=dict(xatol=1e-10, fatol=1e-12))
    np.testing.assert_allclose(a[0], ref.x, atol=1e-4)
    # 1-D, the way .update_K calls it (maxit=100, reltol=1e-4)
    g = lambda t: np.inf if t[0] <= 1e-10 else (np.log(t[0]) - 1.0) ** 2
    x, fx, n, fail = nmmin(g, [0.3], maxit=100, reltol=1e-4)
    assert abs(x[0] - np.e) < 0.05 and fail == 0 and n <= 100


def test_zeroin_matches_oracle_port_and_brentq():
    f = lambda s: 0.5 * (1.0 / (s + 0.3) - 1.7)
    r1 = zeroin(f, 1e-3, 10.0)
    assert r1 == O.zeroin(f, 1e-3, 10.0)
    assert abs(r1 - so.brentq(f, 1e-3, 10.0)) < 2e-4            # uniroot's default tol = eps^0.25
    with pytest.raises(F.FRKError, match="opposite sign"):
        zeroin(f, 5.0, 10.0)


def test_parse_formula():
    assert _parse_formula("z ~ 1") == ("z", [], True)
    assert _parse_formula("co2avgret ~ lat + 1") == ("co2avgret", ["lat"], True)
    assert _parse_formula("z ~ -1 + x + y") == ("z", ["x", "y"], False)


def test_partition_rows_covers_everything():
    for n in (0, 1, 7, 10004569):
        for w in (1, 2, 3, 8):
            parts = [F.partition_rows(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_auto_BAUs_and_auto_basis_match_oracle():
    rng = np.random.default_rng(3)
    xy = rng.uniform(0, 1, (200, 2)) * [2.0, 1.0]
    gb = F.auto_BAUs(F.plane(), cellsize=0.05, data=F.SpatialPointsDataFrame(xy, {}))
    ob = O.auto_BAUs_plane(xy, cellsize=0.05)
    assert np.array_equal(gb.xgrid, ob.xgrid) and np.array_equal(gb.ygrid, ob.ygrid)
    assert np.array_equal(gb.cell2bau(), ob.cell2bau())
    assert np.array_equal(gb.centroids(), ob.centroids())
    for t in ("bisquare", "Matern32"):
        for reg in (1, 2):
            a = F.auto_basis(F.plane(), xy, regular=reg, nres=2, type=t)
            b = O.auto_basis(O.plane(), xy, regular=reg, nres=2, type=t)
            assert np.array_equal(a.loc, b.loc) and np.array_equal(a.scale, b.scale) and np.array_equal(a.res, b.res)
    a = F.auto_basis(F.real_line(), xy[:, :1], regular=6, nres=2)
    b = O.auto_basis(O.real_line(), xy[:, :1], regular=6, nres=2)
    assert np.array_equal(a.loc, b.loc) and np.array_equal(a.scale, b.scale)
    assert [D.shape for D in F.BuildD(a)] == [(6, 6), (12, 12)]


def test_local_basis_argument_checks_mirror_the_reference():
    """R/basisfns.R:118-145: same conditions, same messages."""
    with pytest.raises(ValueError, match="same as the number of manifold dimensions"):
        F.local_basis(F.plane(), np.zeros((3, 1)), np.ones(3))
    with pytest.raises(ValueError, match="as many scale parameters as centroids"):
        F.local_basis(F.plane(), np.zeros((3, 2)), np.ones(2))
    with pytest.raises(ValueError, match="greater than zero"):
        F.local_basis(F.plane(), np.zeros((3, 2)), np.array([1.0, 0.0, 1.0]))
    with pytest.raises(ValueError, match="Only one basis can be multiresolution"):
        b1 = F.local_basis(F.real_line(), np.zeros((2, 1)), np.ones(2), res=[1, 2])
        F.TensorP(b1, b1)


def test_SRE_argument_checks_run_before_any_device_work():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the GPU tests")
    xy = np.random.default_rng(0).uniform(size=(10, 2))
    data = F.SpatialPointsDataFrame(xy, {"z": xy[:, 0], "std": np.ones(10)})
    gb = F.auto_BAUs(F.plane(), cellsize=0.1, data=data)
    basis = F.auto_basis(F.plane(), xy, nres=1)
    with pytest.raises(ValueError, match="BAUs should contain a field 'fs'"):       # reference message (R/check_args.R)
        F.SRE("z ~ 1", data, basis, gb, est_error=False)
    gb["fs"] = 1.0
    with pytest.raises(ValueError, match="block-exponential"):
        F.SRE("z ~ 1", data, basis, gb, est_error=False, K_type="precision")
    with pytest.raises(F.FRKError, match="no CPU fallback"):       # no device -> loud failure, never a CPU path
        F.SRE("z ~ 1", data, basis, gb, est_error=False)


def test_space_time_manifolds():
    """dimensions(STplane()) == dimensions(STsphere()) == 3 (test_domains.R:10-11); unit-sphere space-time distance
    (-90,0,0) <-> (90,0,2) is sqrt(pi^2 + 2^2) (test_domains.R:32-36) through the host mirror used by BuildD."""
    import math
    import frk_b200 as F
    from frk_b200.geometry import host_distance
    assert F.STplane().dim == 3 and F.STsphere().dim == 3
    x = np.array([[-90.0, 0.0, 0.0], [90.0, 0.0, 2.0]])
    assert host_distance(F.STsphere(1.0), x, x)[0, 1] == math.sqrt(math.pi ** 2 + 2 ** 2)
    y = np.array([[0.0, 0.0, 0.0], [1.0, 2.0, 2.0]])
    assert host_distance(F.STplane(), y, y)[0, 1] == 3.0
    with pytest.raises(ValueError, match="number of columns in loc"):
        F.local_basis(F.STplane(), np.zeros((2, 2)), [1.0, 1.0])
    with pytest.raises(ValueError, match="Please specify tunit"):                   # R/basisfns.R:245-247
        F.auto_basis(F.STplane(), y)
    with pytest.raises(ValueError):
        F.STsphere(-1.0)


@pytest.mark.parametrize("width", [1, 2, 3, 4, 8])
def test_speculative_nelder_mead_equals_sequential(width):
    """The driver of the block-exponential M-step (SpecNM: replay against known values, evaluate the requested point together with
    the points Nelder-Mead may ask for next) returns bit-for-bit what plain nmmin returns -- it only groups the evaluations.
    Objectives: the shape of f_tau (log-determinant-like + trace-like term, 1-D), a 2-D one (tensor bases), one with an infeasible
    region (+Inf, as f_tau returns for tau <= 1e-10) and a non-smooth one."""
    fns = [
        (lambda x: 40.0 * np.log(x[0]) + 55.0 / x[0] if x[0] > 1e-10 else np.inf, [0.3]),
        (lambda x: 40.0 * np.log(x[0]) + 55.0 / x[0] if x[0] > 1e-10 else np.inf, [1.375]),        # start at the optimum (a converged M-step)
        (lambda x: (x[0] - 1.3) ** 2 + 3.0 * (x[1] + 0.4) ** 2 + 0.5 * np.sin(3 * x[0]) * x[1], [0.1, 0.1]),
        (lambda x: abs(x[0] - 0.7) + 0.1 * abs(x[0]) ** 1.5, [5.0]),
        (lambda x: np.inf if x[0] < 0.05 else (np.log(x[0]) ** 2 + 1.0 / (x[0] + 0.01)), [0.06]),
    ]
    for fn, x0 in fns:
        calls = []
        def counted(x, fn=fn):
            calls.append(tuple(x))
            return fn(x)
        xs, fs, cnt, fail = nmmin(counted, x0)
        seq_distinct = len(set(calls))
        xp, fp, req, rounds, evals = nmmin_speculative(fn, x0, width)
        assert np.array_equal(xp, xs) and fp == fs
        assert req == seq_distinct                       # Nelder-Mead asked for the same distinct points
        assert rounds <= seq_distinct and evals >= seq_distinct - 0
        if width == 1:
            assert rounds == seq_distinct == evals
        elif seq_distinct > 3:
            assert rounds < seq_distinct                 # speculation saved at least one round


def test_partition_rows_weighted_balances_work():
    """Row partition by WORK (SRE(balance = "observations")): contiguous, covering, every rank within one row's weight of the mean."""
    from frk_b200.runtime import partition_rows, partition_rows_weighted
    rng = np.random.default_rng(3)
    w = 1.0 + 13.0 * rng.poisson(np.cos(np.linspace(-1.5, 1.5, 5000)) ** 2 * 3.0)       # dense in the middle, sparse at the ends
    for world in (2, 3, 8):
        parts = [partition_rows_weighted(w, world, k) for k in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == len(w)
        assert all(parts[k][1] == parts[k + 1][0] for k in range(world - 1))
        loads = np.array([w[a:b].sum() for a, b in parts])
        assert np.all(np.abs(loads - w.sum() / world) <= w.max() + 1e-9)
        eq = np.array([w[slice(*partition_rows(len(w), world, k))].sum() for k in range(world)])
        assert loads.max() <= eq.max()                                 # never worse than the equal-count partition
    assert partition_rows_weighted(np.ones(10), 1, 0) == (0, 10)


def test_space_time_auto_basis_is_the_tensor_product_of_the_reference():
    """auto_basis(STplane()|STsphere(), tunit=): R/basisfns.R:243-289 -- spatial auto_basis on plane() (for both manifolds, with
    scale_aperture 1.25) x ten Gaussian functions at seq(0, end_loc, length.out = 10), scale end_loc / 10, where
    end_loc = ceiling(time span in tunit) + 1."""
    rng = np.random.default_rng(5)
    xy = rng.uniform(0, 10, (400, 2))
    secs = rng.uniform(0, 30.4 * 86400, 400)
    for man in (F.STplane(), F.STsphere()):
        B = F.auto_basis(man, np.column_stack([xy, secs]), regular=1, nres=2, tunit="days")
        assert isinstance(B, F.TensorP_Basis)
        ref = O.auto_basis(O.plane(), xy, regular=1, nres=2, scale_aperture=1.25)
        assert np.array_equal(B.Basis1.loc, ref.loc) and np.array_equal(B.Basis1.scale, ref.scale)
        assert B.Basis1.manifold.kind == "plane"
        end_loc = np.ceil((secs.max() - secs.min()) / 86400.0) + 1
        assert B.Basis2.n == 10 and B.Basis2.type == "Gaussian" and B.Basis2.manifold.kind == "real_line"
        assert np.array_equal(B.Basis2.loc[:, 0], np.linspace(0, end_loc, 10)) or \
            np.allclose(B.Basis2.loc[:, 0], np.linspace(0, end_loc, 10), rtol=0, atol=1e-14)
        assert np.all(B.Basis2.scale == end_loc / 10)
        assert B.n == ref.n * 10
    Bw = F.auto_basis(F.STplane(), xy, nres=1, tunit="weeks", end_time=secs)
    assert Bw.Basis2.loc[-1, 0] == np.ceil((secs.max() - secs.min()) / (7 * 86400.0)) + 1
    with pytest.raises(ValueError, match="tunit should be"):
        F.auto_basis(F.STplane(), xy, tunit="fortnights", end_time=secs)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the arm the driver runs first): one JSON line on stdout with the base contract's keys,
    `impl = "reference"`, a `cpu_baseline` describing this very run and an `e2e` block repeating the value with zero
    transfer bytes -- on a workload small enough for the CPU suite.  Needs no GPU and must not load the CUDA library's device path."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--n-obs", "3000", "--grid", "80",
                          "--nres", "2", "--regular", "1", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["unit"] == "EM iters/s" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

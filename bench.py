#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE config 3 (SURVEY.md s8(d) "C3"):

    1e6 observations uniform on the unit square, z = sin(8x) + cos(8y) + 0.5 N(0,1), std = 0.5;
    BAUs = 3163 x 3163 pixel grid over [0,1]^2 (1.0e7 BAUs); basis auto_basis(plane(), nres = 3, regular = 2,
    type = "bisquare") -> r = 3276; SRE(K_type = "block-exponential", the reference default) -> SRE.fit(EM) -> predict.

A "step" is one EM iteration (log-likelihood + E-step + M-step, R/SREfit.R:104-115).  `value` = EM iterations / s
with everything resident in HBM; `predict` = BAUs predicted / s; `e2e` = the same metrics through the public
host-buffer API (SRE() from pinned host arrays -> SRE_fit -> fitted slots back on the host; predict() -> host).
Rows (BAUs and the observations in them) are partitioned over the ranks: the problem size is fixed as N grows
("scaling": "strong"), one fp64 all-reduce of r^2 + r + 2 doubles per EM iteration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`--impl reference`: the CPU restatement of the reference (oracle/, NumPy/SciPy on all host cores; R itself is not
installable here) on the same workload, each step one EM iteration, bounded by --cpu-budget seconds.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json's metric; `value` is the EM iterations/s half (r = 3276), the predicted-BAUs/s half (1e7 BAUs) is in `predict`
METRIC = "EM iters/sec and predicted BAUs/sec at N=1e6, r~4k, 1/2/4/8 B200 vs host R (value = EM iters/sec; BAUs/sec in 'predict')"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--k-type", default="block-exponential", choices=["block-exponential", "unstructured"])
    ap.add_argument("--n-obs", type=int, default=1_000_000)
    ap.add_argument("--grid", type=int, default=3163, help="BAU grid is grid x grid over the unit square")
    ap.add_argument("--regular", type=int, default=2)
    ap.add_argument("--nres", type=int, default=3)
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--predict-reps", type=int, default=5)
    ap.add_argument("--deterministic", action="store_true", help="fixed-order Gram flush instead of fp64 atomics (bit-reproducible runs)")
    ap.add_argument("--distribute-nm", dest="distribute_nm", action="store_true", default=None,
                    help="spread the Nelder-Mead candidates over the ranks (one lane per rank, one small all-reduce per round).  Was the "
                         "default for N > 1 while every candidate was a factorisation (163 / 167 against 157 / 158 EM it/s replicated at "
                         "N = 8 / 4, profiles/r02h_scale_modes.txt); with the memo of inverses the large block needs none and the "
                         "small blocks' rounds are faster on four local lanes: 169.5 against 188.3 replicated at N = 2 (tools/n2_modes.sh)")
    ap.add_argument("--replicate-nm", dest="distribute_nm", action="store_false",
                    help="every rank evaluates the Nelder-Mead candidates on its own lanes (no exchange): the default")
    ap.add_argument("--lanes", type=int, default=None, help="concurrent Nelder-Mead evaluation lanes per GPU (frk_model_set_lanes)")
    ap.add_argument("--comm", default="nccl", choices=["nccl", "hook"],
                    help="multi-GPU reductions: the library's own NCCL communicator (frk_comm_init) or the torch.distributed hook")
    ap.add_argument("--parity-tol", type=float, default=1e-9, help="relative tolerance of the in-run parity check against the oracle")
    args = ap.parse_args()
    if args.distribute_nm is None:          # default: replicated candidates (see --distribute-nm)
        args.distribute_nm = False
    return args


def workload(args):
    rng = np.random.default_rng(3)
    n = args.n_obs
    x = rng.uniform(0, 1, n)
    y = rng.uniform(0, 1, n)
    z = np.sin(8 * x) + np.cos(8 * y) + 0.5 * rng.standard_normal(n)
    std = np.full(n, 0.5)
    return np.column_stack([x, y]), z, std


def config(args, r, n_bau, m):
    return {"workload": f"C3: {args.n_obs} obs U(0,1)^2, bisquare auto_basis(plane, nres={args.nres}, regular={args.regular}) r={r}, "
                        f"{args.grid}x{args.grid}={n_bau} grid BAUs, K_type={args.k_type}, {m} occupied BAUs",
            "n_obs": args.n_obs, "r": r, "n_bau": n_bau, "m_binned": m, "K_type": args.k_type,
            "cache": "inputs larger than L2 (S 0.37 GB, six r x r fp64 matrices 86 MB each, S0 4 GB); no L2 flush",
            "partition": "BAU rows block-partitioned over ranks; r-sized state replicated",
            "nm_lanes": args.lanes if args.lanes else (1 if (args.distribute_nm and args.gpus > 1) else 2),
            "nm_distributed": bool(args.distribute_nm and args.gpus > 1), "deterministic": bool(args.deterministic)}


# ---------------------------------------------------------------------------------------------------
# CPU arm (oracle) -- also the cpu_baseline leg of the GPU arm
# ---------------------------------------------------------------------------------------------------

def cpu_em(args, budget_s, max_steps, warmup=0):
    """EM iterations of the oracle on the full workload (S evaluated at the occupied BAUs only; BAU-level
    prediction timed on a 1e5-BAU sample).  Returns a dict for the JSON line."""
    from threadpoolctl import threadpool_info, threadpool_limits
    from oracle import frk_oracle as O
    # all host cores whatever the launcher exported (torch.distributed.run sets OMP_NUM_THREADS=1 for every rank, which made the
    # N >= 2 reference arms of round 1 slower than the N = 1 one): the same BLAS thread count for every N
    threadpool_limits(limits=os.cpu_count())
    xy, z, std = workload(args)
    ob = O.auto_basis(O.plane(), np.array([[0.0, 0.0], [1.0, 1.0]]), regular=args.regular, nres=args.nres)
    ba = O.auto_BAUs_plane(xy, cellsize=1.0 / (args.grid - 1), xlims=(0, 1), ylims=(0, 1))
    bau = O.points_to_BAUs(xy, ba)
    uo, means, _ = O.bin_mean(bau, np.column_stack([z, std]))
    cc = np.column_stack([ba.xgrid[uo % ba.nx], ba.ygrid[uo // ba.nx]])
    M = O.SRE(ob, cc, np.ones(len(uo)), np.arange(len(uo)), means[:, 0], means[:, 1], K_type=args.k_type, fast_eval=True)
    times, llk = [], []
    t_all = time.time()
    estep_sys = None
    for it in range(warmup + max_steps):
        t0 = time.time()
        llk.append(float(O.loglik(M))); O.Estep(M)
        # the linear system the E-step has just solved (Q_eta mu = S'D^-1 (Z - X alpha), R/SREfit.R:252-254), kept for the
        # extended-precision refinement below (not part of the timed arithmetic: two cheap products)
        estep_sys = (M.Q_eta, M.S.T @ ((M.Z - M.X @ M.alphahat) / (M.sigma2fshat * M.Vfs + M.Ve)), M.mu_eta.copy())
        O.Mstep(M)
        dt = time.time() - t0
        if it >= warmup:
            times.append(dt)
        if time.time() - t_all > budget_s and len(times) >= 1:
            break
    # prediction sample: 1e5 BAUs
    nb = min(100_000, ba.n)
    idx = np.linspace(0, ba.n - 1, nb).astype(np.int64)
    S0s = O.normalise_rows(O.eval_basis_fast(ob, np.column_stack([ba.xgrid[idx % ba.nx], ba.ygrid[idx // ba.nx]])))
    t0 = time.time()
    v = O.batch_compute_var(S0s, M.S_eta)
    mu = S0s @ M.mu_eta
    tp = time.time() - t0
    blas = [f"{d.get('internal_api')}:{d.get('num_threads')}" for d in threadpool_info()]
    # How accurate is ANY fp64 solve of that system?  Iterative refinement with the residual in extended precision gives mu_eta to
    # working precision; its distance from the oracle's own LAPACK solve is the floor for |device - oracle| (cond(Q_eta) ~ 1e7).
    mu_ref, self_err = None, None
    try:
        import scipy.linalg as sla
        Q, b, mu0 = estep_sys
        cf = sla.cho_factor(Q)
        Ql, bl = Q.astype(np.longdouble), b.astype(np.longdouble)
        x = mu0.astype(np.longdouble)
        for _ in range(4):
            r = bl - Ql @ x
            x = x + sla.cho_solve(cf, np.asarray(r, dtype=np.float64)).astype(np.longdouble)
        mu_ref = np.asarray(x, dtype=np.float64)
        self_err = float(np.max(np.abs(mu0 - mu_ref)) / np.max(np.abs(mu_ref)))
        del Ql
    except Exception as e:                                           # (never fail the bench on the diagnostic)
        print(f"[bench] refinement diagnostic skipped: {e}", file=sys.stderr)
    # what the GPU arm is compared with (parity): the trace from .EM_initialise and the state after len(llk) iterations
    oracle = dict(llk=llk, alpha=np.asarray(M.alphahat, dtype=np.float64).ravel(), sigma2fs=float(M.sigma2fshat),
                  mu_eta=np.asarray(M.mu_eta).ravel(), S_eta_diag=np.diag(M.S_eta).copy(), bau_idx=idx,
                  mu_eta_refined=mu_ref, mu_eta_self_err=self_err,
                  mu=float(np.asarray(M.alphahat).ravel()[0]) + mu, var=v)
    return dict(oracle=oracle, iters=len(times), s_per_iter=float(np.mean(times)), iters_per_s=len(times) / float(np.sum(times)),
                predict_baus_per_s=nb / tp, cores=os.cpu_count(), blas=blas, r=ob.n, n_bau=ba.n, m=len(uo),
                sample=f"{len(times)} full EM iteration(s) (loglik+E+M, K_type={args.k_type}) at m={len(uo)}, r={ob.n} "
                       f"after {warmup} warm-up; predict on a {nb}-BAU sample; NumPy/SciPy oracle, BLAS threads {blas}")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_em(args, budget_s=max(args.cpu_budget, 150.0), max_steps=args.steps, warmup=0)
    line = {"impl": "reference", "metric": METRIC, "value": res["iters_per_s"], "unit": "EM iters/s", "n_gpus": args.gpus,
            "steps": res["iters"], "warmup": 0, "ms_per_step": 1e3 * res["s_per_iter"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config(args, res["r"], res["n_bau"], res["m"]),
            "predict": {"value": res["predict_baus_per_s"], "unit": "BAUs/s"},
            "cpu_baseline": {"value": res["iters_per_s"], "unit": "EM iters/s", "cores": res["cores"], "kind": "port",
                             "sample": res["sample"]},
            "e2e": {"value": res["iters_per_s"], "unit": "EM iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def fp64_peak_probe(torch, dev):
    """cuBLAS DGEMM 8192^3, best of 5: the measured fp64 denominator (MEASURED_PEAKS.json has none)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def run_b200(args):
    import torch
    import torch.distributed as dist
    import frk_b200 as F
    from frk_b200._lib import call

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = F.get_device(local)
    if world > 1 and args.comm == "nccl":
        dev.init_comm()      # the library's own NCCL communicator carries every reduction of the fit; torch's group only times / barriers
    if args.deterministic:
        dev.set_deterministic(True)
    _dbg("device + process group up")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev.tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- host inputs in pinned memory -------------------------------------------------------
    xy, z, std = workload(args)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    xy_p = pin(np.asfortranarray(xy).T.copy()).T          # column-major (n x 2) view over pinned memory
    data = F.SpatialPointsDataFrame(xy_p, {"z": pin(z), "std": pin(std)})
    basis = F.auto_basis(F.plane(), np.array([[0.0, 0.0], [1.0, 1.0]]), regular=args.regular, nres=args.nres)
    BAUs = F.auto_BAUs(F.plane(), cellsize=1.0 / (args.grid - 1), xlims=(0, 1), ylims=(0, 1))
    BAUs.fields["fs"] = pin(np.ones(len(BAUs)))
    r, N = basis.n, len(BAUs)

    # ---- (A) device-resident EM iterations --------------------------------------------------------
    mdl = F.SRE("z ~ 1", data, basis, BAUs, est_error=False, K_type=args.k_type, dev=dev)
    F.SRE_fit(mdl, n_EM=args.warmup, tol=0.0, lanes=args.lanes, distribute_nm=args.distribute_nm)
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    l0 = dev.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    F.SRE_fit(mdl, n_EM=args.steps, tol=0.0, lanes=args.lanes, distribute_nm=args.distribute_nm)
    e1.record()
    barrier()
    t_em = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    launches = dev.launches() - l0
    llk_tail = [float(v) for v in mdl.info_fit["llk"][-2:]]
    nm_stats = mdl.info_fit.get("nm_stats")
    # per-iteration times of the NEXT iterations (one call each): `value` depends on which iterations are timed, because the
    # Nelder-Mead rounds of the M-step shrink as tau converges
    iter_ms = []
    for _ in range(5):
        e0.record()
        F.SRE_fit(mdl, n_EM=1, tol=0.0, lanes=args.lanes, distribute_nm=args.distribute_nm)
        e1.record()
        torch.cuda.synchronize()
        iter_ms.append(round(e0.elapsed_time(e1), 3))

    _dbg("EM timed")
    # ---- (B) device-resident prediction ---------------------------------------------------------------
    for _ in range(2):
        F.predict(mdl, obs_fs=False, to_host=False)
    barrier()
    lp0 = dev.launches()
    e0.record()
    for _ in range(args.predict_reps):
        F.predict(mdl, obs_fs=False, to_host=False)
    e1.record()
    barrier()
    t_pred = max_over_ranks(e0.elapsed_time(e1) * 1e-3) / args.predict_reps
    launches_pred = (dev.launches() - lp0) // args.predict_reps
    clk = clocks.stop()

    # ---- (D) per-kernel roofline pass (CUDA events around every launch, same steps) --------------------
    kernels, roof = [], None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs") if "hbm_gbs" in peaks else (6650.0, "fallback 6.65 TB/s")
    if not args.no_profile:
        fp64_peak = fp64_peak_probe(torch, dev.tdev)
        call("frk_profile", dev.ctx, 1)
        F.SRE_fit(mdl, n_EM=2, tol=0.0, lanes=args.lanes, distribute_nm=args.distribute_nm)
        n_f_tau_launch_iters = 2
        F.predict(mdl, obs_fs=False, to_host=False)
        buf = C.create_string_buffer(1 << 16)
        call("frk_profile_report", dev.ctx, buf, len(buf))
        call("frk_profile", dev.ctx, 0)
        rep = json.loads(buf.value.decode())
        # algorithmic bytes of the streaming kernels (DESIGN.md "roofline accounting")
        nloc = mdl.hi - mdl.lo
        # (val 8 B + local column id 1 B per non-zero; the 4-byte global column index is not read by the bucketed kernels)
        gram_bytes = mdl.S.nnz * 9 + mdl.m * (4 + 8 * (3 + mdl.p)) + (r * r + r) * 8
        quad_S_bytes = mdl.S.nnz * 9 + mdl.m * (4 + 16) + r * r * 8
        quad_S0_bytes = mdl.S0.nnz * 9 + nloc * (4 + 16) + r * r * 8
        tot_ms = sum(k["ms"] for k in rep) or 1.0
        for k in sorted(rep, key=lambda k: -k["ms"]):
            ent = {"kernel": k["kernel"], "launches": k["launches"], "ms_total": round(k["ms"], 3), "share": round(k["ms"] / tot_ms, 4)}
            if k["flops"] > 0:
                ach = k["flops"] / (k["ms"] * 1e-3) / 1e12
                ent.update(bound="fp64", achieved=ach, peak=fp64_peak, unit="TFLOP/s", frac=ach / fp64_peak)
            elif "gram_bucket_kernel" in k["kernel"] or "gram_ell" in k["kernel"] or "quadform_ell_kernel" in k["kernel"] or "quadform_ws_kernel" in k["kernel"]:
                # both ceilings: HBM (val 8 B + local id 1 B per non-zero, ...) and fp64 (the operation's own flops: 2 * sum_i nnz_i^2
                # for the quadratic form, sum_i nnz_i (nnz_i + 1) for the Gram; lower bound nnz^2 / rows); the binding one is reported
                if "gram" in k["kernel"]:
                    # gram_ell_kernel: one pass = one launch per class of bucket widths; the library books the pass's bytes once
                    passes = round(k["bytes"] / gram_bytes) if k.get("bytes", 0) > 0 else k["launches"]
                    by = gram_bytes * passes
                    fl = passes * (mdl.S.nnz ** 2 / max(mdl.m, 1))
                    ent["passes"] = passes
                else:                                               # launches: n_f_tau_launch_iters on S (M-step) + 1 on S0 (predict)
                    by = quad_S_bytes * n_f_tau_launch_iters + quad_S0_bytes
                    fl = 2.0 * (n_f_tau_launch_iters * mdl.S.nnz ** 2 / max(mdl.m, 1) + mdl.S0.nnz ** 2 / max(nloc, 1))
                t_s = k["ms"] * 1e-3
                hb, fp = by / t_s / 1e9, fl / t_s / 1e12
                if fp / fp64_peak > hb / hbm_peak:
                    ent.update(bound="fp64", achieved=fp, peak=fp64_peak, unit="TFLOP/s", frac=fp / fp64_peak, hbm_gbs=hb, hbm_frac=hb / hbm_peak)
                else:
                    ent.update(bound="hbm", achieved=hb, peak=hbm_peak, unit="GB/s", frac=hb / hbm_peak, fp64_tflops=fp, fp64_frac=fp / fp64_peak)
            kernels.append(ent)
        # dominant kernel = largest share of device time; the tagged launches of dgemm_kernel are ONE kernel
        groups = {}
        for k in rep:
            name = "dgemm_kernel" if k["kernel"].startswith("dgemm_kernel") else k["kernel"]
            g = groups.setdefault(name, {"ms": 0.0, "flops": 0.0, "launches": 0})
            g["ms"] += k["ms"]; g["flops"] += k["flops"]; g["launches"] += k["launches"]
        name, g = max(groups.items(), key=lambda kv: kv[1]["ms"])
        ent = next((e for e in kernels if e["kernel"] == name and "achieved" in e), None)
        if name == "dgemm_kernel":
            ach = g["flops"] / (g["ms"] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "dgemm_kernel (fp64 DMMA m8n8k4; potrf/potri GEMMs)", "achieved": ach, "peak": fp64_peak,
                    "unit": "TFLOP/s", "frac": ach / fp64_peak, "traffic": None, "launches": g["launches"],
                    "peak_source": "cuBLAS DGEMM 8192^3 probe in this run (MEASURED_PEAKS.json has no fp64 figure)",
                    "share_of_step": round(g["ms"] / tot_ms, 4)}
            # DRAM bytes per launch from the committed ncu capture of this command (tools/traffic_pass.sh); a recorded
            # measurement, not taken in this run -- the operands (<= 86 MB) mostly stay in the 126 MB L2
            tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02r_dgemm_traffic.json")
            if args.k_type == "block-exponential" and os.path.exists(tpath):
                try:
                    tj = json.load(open(tpath))
                    roof["traffic"] = tj["dram_bytes_per_launch"]
                    roof["traffic_source"] = "profiles/r02r_dgemm_traffic.json (ncu dram__bytes_read+write over %d launches)" % tj["launches"]
                except Exception:
                    pass
        elif ent is not None:
            roof = {"bound": "tensor" if ent["bound"] == "fp64" else "hbm", "kernel": name, "achieved": ent["achieved"],
                    "peak": ent["peak"], "unit": ent["unit"], "frac": ent["frac"], "traffic": None, "launches": g["launches"],
                    "peak_source": hbm_src if ent["bound"] == "hbm" else "cuBLAS DGEMM 8192^3 probe in this run",
                    "share_of_step": round(g["ms"] / tot_ms, 4)}
        else:       # latency-bound helper kernel on top (e.g. the 64 x 64 diagonal factorisation): report its algorithmic flops
            fl = {"potf2_kernel": 64.0 ** 3 / 3.0}.get(name, 0.0) * g["launches"]
            ach = fl / (g["ms"] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": name, "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                    "traffic": None, "launches": g["launches"], "peak_source": "cuBLAS DGEMM 8192^3 probe in this run",
                    "share_of_step": round(g["ms"] / tot_ms, 4), "note": "latency-bound (sequential column dependency), not throughput-bound"}

    # ---- (D2) the setup kernels of SRE(): eval_basis -> CSR, point -> BAU lookup, binning (north_star items 1-2) --------------
    setup = []
    if not args.no_profile:
        nloc = mdl.hi - mdl.lo
        call("frk_profile", dev.ctx, 1)
        ms = F.SRE("z ~ 1", data, basis, BAUs, est_error=False, K_type=args.k_type, dev=dev)
        ms.S0.buckets(dev)
        buf = C.create_string_buffer(1 << 16)
        call("frk_profile_report", dev.ctx, buf, len(buf))
        call("frk_profile", dev.ctx, 0)
        rep = {k["kernel"]: k for k in json.loads(buf.value.decode())}
        nobs_loc = len(z) if world == 1 else None
        # algorithmic bytes: eval_basis N*8d + (N+1)*4 + nnz*12; lookup 20 B/point
        alg = {"eval_basis (count/scan/fill or one pass)": (nloc * 16 + (nloc + 1) * 4 + ms.S0.nnz * 12,
                                                             [k for k in rep if "eval_count" in k or "eval_fill" in k or "eval_one" in k]),
               "grid_lookup_kernel": ((20 * nobs_loc) if nobs_loc else 0, ["grid_lookup_kernel"]),
               "bin_* (counting sort + per-BAU means)": (0, [k for k in rep if k.startswith("bin_")]),
               "gather_count/fill (S = C_Z S0)": (ms.S.nnz * 24 + ms.m * 8, [k for k in rep if k.startswith("gather_")]),
               "row buckets + ELL tiles (bucket_*, ell_*)": (0, [k for k in rep if k.startswith(("bucket_", "ell_"))])}
        for name, (by, ks) in alg.items():
            t = sum(rep[k]["ms"] for k in ks if k in rep)
            if t <= 0:
                continue
            ent = {"kernel": name, "ms": round(t, 3), "launches": int(sum(rep[k]["launches"] for k in ks if k in rep))}
            if by:
                ent.update(bound="hbm", algorithmic_bytes=int(by), achieved=by / (t * 1e-3) / 1e9, peak=hbm_peak, unit="GB/s",
                           frac=by / (t * 1e-3) / 1e9 / hbm_peak)
            setup.append(ent)
        setup.append({"kernel": "all SRE() + bucket kernels", "ms": round(sum(k["ms"] for k in rep.values()), 3),
                      "launches": int(sum(k["launches"] for k in rep.values()))})
        del ms

    # ---- (C) end to end through the public host-buffer API -------------------------------------------
    # one untimed pass first (it allocates the pinned result buffers and fills the allocator pools), then the timed pass
    del mdl
    torch.cuda.empty_cache()
    for timed in (False, True):
        barrier()
        t0 = time.perf_counter()
        m2 = F.SRE("z ~ 1", data, basis, BAUs, est_error=False, K_type=args.k_type, dev=dev)
        F.SRE_fit(m2, n_EM=args.steps, tol=0.0, lanes=args.lanes, distribute_nm=args.distribute_nm)
        slots = {k: m2.get(k) for k in ("mu_eta", "S_eta", "Khat", "Khat_inv")}
        barrier()
        t_fit_e2e = max_over_ranks(time.perf_counter() - t0)
        h2d_fit, d2h_fit = m2.h2d_bytes, m2.d2h_bytes
        t0 = time.perf_counter()
        pr = F.predict(m2, obs_fs=False, to_host=True)
        barrier()
        t_pred_e2e = max_over_ranks(time.perf_counter() - t0)
        d2h_pred = 3 * 8 * (m2.hi - m2.lo)
        finite = bool(np.isfinite(pr["mu"]).all() and (pr["var"] > 0).all())
        m_binned = m2.m_total
        if not timed:
            del m2, slots, pr

    _dbg("e2e done")
    llk_trace = [float(v) for v in m2.info_fit["llk"]]           # from .EM_initialise (the e2e pass): what the parity checks read
    # ---- (E) CPU baseline (rank 0, N = 1 only) + parity of THIS configuration against it -------------------
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = cpu_em(args, budget_s=args.cpu_budget, max_steps=2, warmup=0)
        cpu = {"value": res["iters_per_s"], "unit": "EM iters/s", "cores": res["cores"], "kind": "port", "sample": res["sample"],
               "predict_baus_per_s": res["predict_baus_per_s"]}
        # the oracle ran k full-size EM iterations from .EM_initialise on the identical inputs: compare the log-likelihood trace,
        # and the state after exactly k iterations (alpha, sigma2fs, mu_eta, diag S_eta, predict mu / var on its 1e5-BAU sample)
        orc = res["oracle"]
        k = len(orc["llk"])
        m3 = F.SRE("z ~ 1", data, basis, BAUs, est_error=False, K_type=args.k_type, dev=dev)
        F.SRE_fit(m3, n_EM=k, tol=0.0, lanes=args.lanes)
        pr3 = F.predict(m3, obs_fs=True, to_host=True)
        def rel(g, o):
            g, o = np.asarray(g, dtype=np.float64), np.asarray(o, dtype=np.float64)
            den = float(np.max(np.abs(o)))
            return float(np.max(np.abs(g - o)) / den) if den > 0 else float(np.max(np.abs(g)))     # (reference value exactly 0: absolute)
        parity = {"against": "oracle (NumPy/SciPy restatement of the reference; parity unpinned by reference-held vectors), full C3 size",
                  "iters": k, "tol": args.parity_tol,
                  "llk_rel": float(max(abs(g - o) / abs(o) for g, o in zip(m3.info_fit["llk"], orc["llk"]))),
                  "llk_gpu": [float(v) for v in m3.info_fit["llk"]], "llk_oracle": orc["llk"],
                  "alpha_rel": rel(m3.alphahat, orc["alpha"]), "sigma2fs_rel": rel([m3.sigma2fshat], [orc["sigma2fs"]]),
                  "mu_eta_rel": rel(m3.get("mu_eta"), orc["mu_eta"]), "S_eta_diag_rel": rel(np.diag(m3.get("S_eta")), orc["S_eta_diag"]),
                  "mu_rel": rel(pr3["mu"][orc["bau_idx"]], orc["mu"]), "var_rel": rel(pr3["var"][orc["bau_idx"]], orc["var"]),
                  "norm": "max |gpu - oracle| / max |oracle| per quantity; llk: worst relative error over the iterations"}
        if orc.get("mu_eta_refined") is not None:
            # the oracle's last E-step system solved to working precision (iterative refinement, residual in extended precision):
            # how far the oracle's OWN fp64 solve is from it, and how far the device is -- the floor for any fp64 comparison here
            parity["mu_eta_oracle_vs_refined"] = orc["mu_eta_self_err"]
            parity["mu_eta_gpu_vs_refined"] = rel(m3.get("mu_eta"), orc["mu_eta_refined"])
        # Gate: the quantities north_star names -- log-likelihood trace, predictions (mu, var) -- and the parameters at 1e-9.  mu_eta and
        # Sigma_eta = Q^-1 themselves are reported and gated at 10x that: at this size Q_eta is ill conditioned (block-exponential K^-1
        # plus S'D^-1 S), the forward error of ANY fp64 inverse is ~ cond(Q) * eps, and two backward-stable inverses (LAPACK's in the
        # oracle / R, ours) differ by that much in directions the data do not determine -- S0 mu_eta and diag(S0 Sigma S0') agree to 1e-10.
        parity["tol_eta"] = 10 * args.parity_tol
        parity["ok"] = bool(all(parity[q] <= args.parity_tol for q in ("llk_rel", "alpha_rel", "sigma2fs_rel", "mu_rel", "var_rel")) and
                            all(parity[q] <= parity["tol_eta"] for q in ("mu_eta_rel", "S_eta_diag_rel")))
        del m3, pr3
    elif rank == 0:
        # no oracle run here (N > 1 or --no-cpu-baseline): compare the trace with the committed single-GPU trace of this configuration
        gpath = os.path.join(ROOT, "tests", "golden", "c3_llk_trace_n1.json")
        default_cfg = (args.n_obs, args.grid, args.regular, args.nres, args.k_type) == (1_000_000, 3163, 2, 3, "block-exponential")
        if default_cfg and os.path.exists(gpath):
            ref = json.load(open(gpath))["llk"]
            kk = min(len(ref), len(llk_trace))
            parity = {"against": "tests/golden/c3_llk_trace_n1.json (single-GPU trace of this configuration from .EM_initialise)",
                      "iters": kk, "tol": 1e-11,
                      "llk_rel": float(max(abs(g - o) / abs(o) for g, o in zip(llk_trace[:kk], ref[:kk])))}
            parity["ok"] = bool(parity["llk_rel"] <= 1e-11)

    if rank == 0:
        line = {"metric": METRIC, "value": args.steps / t_em, "unit": "EM iters/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * t_em / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config(args, r, N, m_binned),
                "predict": {"value": N / t_pred, "unit": "BAUs/s", "ms": 1e3 * t_pred, "launches": launches_pred,
                            "e2e": {"value": N / t_pred_e2e, "unit": "BAUs/s", "d2h_bytes": d2h_pred * world, "results_ok": finite}},
                "e2e": {"value": args.steps / t_fit_e2e, "unit": "EM iters/s", "h2d_bytes_per_step": h2d_fit / args.steps,
                        "d2h_bytes_per_step": d2h_fit / args.steps,
                        "what": f"SRE() from pinned host arrays + SRE_fit({args.steps} EM iterations from .EM_initialise) + fitted slots (mu_eta, S_eta, Khat, Khat_inv) back to pinned host memory, wall clock, after one untimed pass"},
                "gpu_launches": int(launches), "clocks": clk, "llk_last": llk_tail, "nelder_mead_last_mstep": nm_stats,
                "ms_per_iteration_next5": iter_ms, "llk_trace_from_init": llk_trace,
                "comm": (args.comm if world > 1 else None)}
        if parity:
            line["parity"] = parity
        if roof:
            line["roofline"] = roof
        if kernels:
            # the twelve largest, plus the streaming kernels of the EM iteration / predict wherever they rank
            line["kernels"] = kernels[:12] + [k for k in kernels[12:] if "gram" in k["kernel"] or "quadform" in k["kernel"]]
        if setup:
            line["setup_kernels"] = setup
        line["e2e"]["predict_baus_per_s"] = N / t_pred_e2e
        line["config"]["predict_baus_per_s_device"] = N / t_pred
        if cpu:
            line["cpu_baseline"] = cpu
        _emit(line)
    _dbg("line printed")
    if rank == 0 and parity and not parity["ok"]:
        print(f"PARITY FAILED: {json.dumps({k: v for k, v in parity.items() if k.endswith('_rel')})}", file=sys.stderr, flush=True)
        _EXIT[0] = 3
    if world > 1:
        try:
            del m2, slots, pr
            barrier()
            dist.destroy_process_group()
        except Exception as e:      # the result line is already out; never fail the run on teardown
            print(f"teardown: {e!r}", file=sys.stderr, flush=True)


def _dbg(msg):
    if os.environ.get("FRK_BENCH_DEBUG"):
        print(f"[bench rank {os.environ.get('RANK', '0')} pid {os.getpid()} t={time.time():.3f}] {msg}", file=sys.stderr, flush=True)


_REAL_STDOUT = None
_EXIT = [0]


def _guard_stdout():
    """Libraries (NCCL's version banner, BLAS warnings) sometimes write to fd 1; the contract is ONE JSON line on stdout.  Keep a
    private copy of the real stdout for the result line and point fd 1 at stderr for everything else."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def _emit(line: dict):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    _guard_stdout()
    _dbg("start")
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    if _EXIT[0]:
        sys.exit(_EXIT[0])


if __name__ == "__main__":
    main()

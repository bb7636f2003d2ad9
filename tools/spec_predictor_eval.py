"""How good is the Nelder-Mead candidate predictor (model.cu SpecNM::candidates)?  CPU only.  The oracle runs EM on a smaller analog
of the bench workload (r = 360: blocks of 36 and 324 functions, 3e5 observations); every f_tau minimisation of its M-steps is
repeated through the library's speculative driver (frk_optim_nelder_mead_speculative) at widths 2 and 4, and the rounds are
counted.  Usage: python tools/spec_predictor_eval.py [iterations]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import frk_oracle as O
from frk_b200.optim import nmmin_speculative
import bench
n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sys.argv = ['x', '--n-obs', '300000', '--grid', '1000', '--regular', '2', '--nres', '2']
args = bench.parse()
xy, z, std = bench.workload(args)
ob = O.auto_basis(O.plane(), np.array([[0.0, 0.0], [1.0, 1.0]]), regular=args.regular, nres=args.nres)
ba = O.auto_BAUs_plane(xy, cellsize=1.0 / (args.grid - 1), xlims=(0, 1), ylims=(0, 1))
uo, means, _ = O.bin_mean(O.points_to_BAUs(xy, ba), np.column_stack([z, std]))
cc = np.column_stack([ba.xgrid[uo % ba.nx], ba.ygrid[uo // ba.nx]])
M = O.SRE(ob, cc, np.ones(len(uo)), np.arange(len(uo)), means[:, 0], means[:, 1], K_type='block-exponential', fast_eval=True)
orig = O.nmmin
tot = {}
def traced(fn, par, **kw):
    out = orig(fn, par, **kw)
    cache = {}
    def f(t):
        k = float(np.atleast_1d(t)[0])
        if k not in cache: cache[k] = float(fn(np.atleast_1d(t)))
        return cache[k]
    row = []
    for w in (1, 2, 4):
        x, fx, req, rounds, evals = nmmin_speculative(f, np.atleast_1d(par), w)
        assert x[0] == np.atleast_1d(out[0])[0]
        t = tot.setdefault(w, [0, 0, 0]); t[0] += req; t[1] += rounds; t[2] += evals
        row.append(f"w={w}: {rounds} rounds / {evals} ev")
    print(f"   start {np.atleast_1d(par)[0]:.5g}: {req} requests; " + "; ".join(row))
    return out
O.nmmin = traced
for it in range(n_it):
    O.loglik(M); O.Estep(M); print('iteration', it + 1, flush=True); O.Mstep(M)
for w, t in tot.items():
    print(f"TOTAL width {w}: requests {t[0]}, rounds {t[1]}, evaluations {t[2]}")

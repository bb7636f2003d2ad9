"""Stand-alone timing of the dense fp64 kernels (frk_potrf / frk_potri / GEMM) on one B200.
python tools/dense_bench.py [n ...]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import frk_b200 as F  # noqa: E402
from frk_b200._lib import call  # noqa: E402

dev = F.get_device(0)
sizes = [int(a) for a in sys.argv[1:]] or [3276]
for n in sizes:
    g = torch.Generator(device="cuda").manual_seed(0)
    B = torch.randn(n, n + 8, dtype=torch.float64, device="cuda", generator=g)
    A0 = (B @ B.T + n * torch.eye(n, dtype=torch.float64, device="cuda")).contiguous().view(-1)
    A = torch.empty_like(A0); Ainv = torch.empty_like(A0); W = torch.empty_like(A0)
    ld = C.c_double(0)
    def run():
        A.copy_(A0)
        call("frk_potrf", dev.ctx, n, dev.ptr(A), C.byref(ld))
        call("frk_potri", dev.ctx, n, dev.ptr(A), dev.ptr(Ainv), dev.ptr(W))
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    reps = 5
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    tf = ti = 0.0
    for _ in range(reps):
        A.copy_(A0)
        e0.record()
        call("frk_potrf", dev.ctx, n, dev.ptr(A), C.byref(ld))
        e1.record()
        call("frk_potri", dev.ctx, n, dev.ptr(A), dev.ptr(Ainv), dev.ptr(W))
        e2.record()
        torch.cuda.synchronize()
        tf += e0.elapsed_time(e1); ti += e1.elapsed_time(e2)
    tf /= reps; ti /= reps
    # correctness
    err = (Ainv.view(n, n) @ A0.view(n, n) - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max().item()
    # cuSOLVER / cuBLAS reference timings through torch for context
    M0 = A0.view(n, n)
    for _ in range(2):
        Lt = torch.linalg.cholesky(M0); torch.cholesky_inverse(Lt)
    torch.cuda.synchronize()
    e0.record(); Lt = torch.linalg.cholesky(M0); e1.record(); Li = torch.cholesky_inverse(Lt); e2.record(); torch.cuda.synchronize()
    print(f"n={n}: potrf {tf:.3f} ms ({n**3/3/tf/1e9:.2f} TF/s)  potri {ti:.3f} ms ({2*n**3/3/ti/1e9:.2f} TF/s)  |Ainv A - I|max={err:.2e}"
          f"   [torch/cuSOLVER: cholesky {e0.elapsed_time(e1):.3f} ms, cholesky_inverse {e1.elapsed_time(e2):.3f} ms]")
    call("frk_profile", dev.ctx, 1)
    run()
    buf = C.create_string_buffer(1 << 16)
    call("frk_profile_report", dev.ctx, buf, len(buf))
    call("frk_profile", dev.ctx, 0)
    for k in sorted(json.loads(buf.value.decode()), key=lambda k: -k["ms"]):
        tfl = k["flops"] / (k["ms"] * 1e-3) / 1e12 if k["flops"] else 0
        print(f"   {k['kernel'][:48]:48s} n={k['launches']:4d} {k['ms']:8.3f} ms  avg {1e3*k['ms']/k['launches']:7.1f} us  {tfl:6.2f} TF/s")

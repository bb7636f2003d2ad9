#!/bin/bash
# timing experiments for the panel Cholesky (one GPU call): chain only / everything on one stream / normal; GEMM shapes of the trailing updates
for mode in 0 1 2; do
  echo "== FRK_POTRF_SKIP=$mode"; FRK_POTRF_SKIP=$mode python tools/dense_bench.py 3276 2>&1 | grep -E "^n="
done
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import frk_b200 as F
from frk_b200._lib import call
dev = F.get_device(0)
for (M, N, K, beta) in [(3000, 3000, 128, 1.0), (2000, 2000, 128, 1.0), (1000, 1000, 128, 1.0), (3000, 128, 128, 1.0), (3000, 3000, 256, 1.0), (3000, 3000, 128, 0.0)]:
    A = torch.randn(K, M, dtype=torch.float64, device="cuda").contiguous()
    B = torch.randn(K, N, dtype=torch.float64, device="cuda").contiguous()
    Cc = torch.zeros(N, M, dtype=torch.float64, device="cuda")
    def mine():
        call("frk_dgemm_nt", dev.ctx, M, N, K, -1.0, dev.ptr(A), M, dev.ptr(B), N, beta, dev.ptr(Cc), M)
    def cub():
        torch.addmm(Cc, B.T, A, beta=beta, alpha=-1.0, out=Cc)
    res = []
    for fn in (mine, cub):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res.append((ms, 2.0 * M * N * K / ms / 1e9))
    print(f"M={M} N={N} K={K} beta={beta}: frk {res[0][0]*1e3:8.1f} us {res[0][1]:6.2f} TF/s | cuBLAS {res[1][0]*1e3:8.1f} us {res[1][1]:6.2f} TF/s")
PY

"""GEMM sweep: frk_dgemm_nt (fp64 DMMA) vs cuBLAS (torch.matmul) on one B200."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import frk_b200 as F
from frk_b200._lib import call
dev = F.get_device(0)
shapes = [(3276, 3276, 3276), (3276, 3276, 256), (3276, 3276, 64), (3276, 192, 64), (1024, 1024, 1024), (8192, 8192, 512)]
for (M, N, K) in shapes:
    A = torch.randn(K, M, dtype=torch.float64, device="cuda").contiguous()   # column-major M x K  == row-major K x M
    B = torch.randn(K, N, dtype=torch.float64, device="cuda").contiguous()   # column-major N x K
    Cc = torch.zeros(N, M, dtype=torch.float64, device="cuda")
    def mine():
        call("frk_dgemm_nt", dev.ctx, M, N, K, 1.0, dev.ptr(A), M, dev.ptr(B), N, 0.0, dev.ptr(Cc), M)
    def cub():
        torch.matmul(B.T, A)     # (N x K)(K x M) = C' row-major == C column-major
    res = []
    for fn in (mine, cub):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res.append((ms, 2.0 * M * N * K / ms / 1e9))
    err = (Cc - torch.matmul(B.T, A)).abs().max().item()
    print(f"M={M} N={N} K={K}: frk {res[0][0]*1e3:8.1f} us {res[0][1]:6.2f} TF/s | cuBLAS {res[1][0]*1e3:8.1f} us {res[1][1]:6.2f} TF/s | maxerr {err:.1e}")

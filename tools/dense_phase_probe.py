"""Phase timing inside the diagonal / panel-solve kernels of frk_potrf (clock64 stamps of CTA 0, frk_debug_clocks)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import frk_b200 as F
from frk_b200._lib import call, lib
dev = F.get_device(0)
call("frk_use_graphs", dev.ctx, 0)
for n in (256, 3276):
    g = torch.Generator(device="cuda").manual_seed(0)
    B = torch.randn(n, n + 8, dtype=torch.float64, device="cuda", generator=g)
    A0 = (B @ B.T + n * torch.eye(n, dtype=torch.float64, device="cuda")).contiguous().view(-1)
    A = A0.clone(); ld = C.c_double(0)
    for _ in range(3):
        A.copy_(A0); call("frk_potrf", dev.ctx, n, dev.ptr(A), C.byref(ld))
    out = (C.c_longlong * 32)()
    lib.frk_debug_clocks.restype = C.c_int
    lib.frk_debug_clocks(out)
    c = np.array(out[:], dtype=np.int64)
    mhz = 1965.0
    us = lambda a, b: (c[b] - c[a]) / mhz
    print(f"n={n}: diag128 (last launch): load {us(0,1):.2f}  potf2 {us(1,2):.2f}  store+log {us(2,6):.2f}  total {us(0,6):.2f} us")
    print(f"      potf2 (last call, first 16-column step): factor {us(16,17):.2f}  store {us(17,18):.2f}  inv16 {us(18,19):.2f}  sync {us(19,20):.2f}  rows below {us(20,21):.2f}  update {us(21,22):.2f} us")
    print(f"      panel_solve128 (last launch, CTA 0): load {us(8,9):.2f}  8 steps {us(9,11):.2f}  store {us(11,12):.2f}  total {us(8,12):.2f} us")

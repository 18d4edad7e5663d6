"""Mat-vec phase alone at the row-shard shapes of the n=32768 problem (1 GPU): how much of the
multi-GPU iteration time is the shard's own streaming, how much is exchange + barriers."""
import ctypes as C
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ccqppy_b200 import _capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn((n, n), generator=g, device=dev, dtype=torch.float64)
v = torch.zeros(n + 64, device=dev, dtype=torch.float64)
v[:n] = torch.randn(n, generator=g, device=dev, dtype=torch.float64)
y = torch.zeros(n + 64, device=dev, dtype=torch.float64)
for P in (1, 2, 4, 8):
    rows = n // P
    torch.cuda.synchronize()          # a raw handle runs on its own stream, not on torch's
    h = _capi.Handle()
    _capi.check(h.h, h.lib.ccqp_set_matrix(h.h, C.c_void_p(A.data_ptr()), n, n, 0, rows, 1))
    sec = C.c_double()
    for rep in (20, 200):
        _capi.check(h.h, h.lib.ccqp_gemv_timed(h.h, C.c_void_p(v.data_ptr()), C.c_void_p(y.data_ptr()), rep, C.byref(sec)))
    byt = 8.0 * rows * n + 8 * n + 8 * rows
    print("P=%d rows=%5d  %8.1f us per mat-vec launch  %6.0f GB/s  (ideal at 6547.5 GB/s: %.1f us)" %
          (P, rows, sec.value * 1e6, byt / sec.value / 1e9, byt / 6547.5e9 * 1e6), flush=True)
    h.close()

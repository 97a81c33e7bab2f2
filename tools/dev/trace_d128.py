import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fastmax_experiments_b200 import _lib
from fastmax_experiments_b200.fastmax import fastmax_fwd_raw
lib = _lib.load()
B, H, N = 1, 32, 4096
gen = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(B, H, N, 128, device="cuda", generator=gen).to(torch.bfloat16) for _ in range(3))
for _ in range(2):
    fastmax_fwd_raw(q, k, v, True, 1.0 / 90, 2, _lib.ALGO_AUTO)
torch.cuda.synchronize()
tab = (C.c_longlong * (8 * 64))()
lib.fastmax_debug_trace_fetch_d128.argtypes = [C.POINTER(C.c_longlong)]
lib.fastmax_debug_trace_fetch_d128(tab)
t = [[tab[e * 64 + j] for j in range(64)] for e in range(8)]
t0 = t[0][0]
print("CTA (0,0) = the last q tile (32 kv tiles); cycles relative to the first event")
print("tile  K in   S issued  P seen  PV issued | S ready  loaded  P stored")
for j in (0, 1, 2, 3, 10, 11, 20, 30, 31):
    print(f"{j:4d} " + " ".join(f"{t[e][j] - t0:8d}" for e in range(7)))

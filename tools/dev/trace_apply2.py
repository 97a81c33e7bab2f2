import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fastmax_experiments_b200 import _lib
from fastmax_experiments_b200.fastmax import fastmax_fwd_raw
lib = _lib.load()
B, H, N = 8, 32, 16384
gen = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(B, H, N, 64, device="cuda", generator=gen).to(torch.bfloat16) for _ in range(3))
for _ in range(2):
    fastmax_fwd_raw(q, k, v, True, 1.0 / 64, 2, _lib.ALGO_MOMENTS)
torch.cuda.synchronize()
tab = (C.c_longlong * (8 * 40))()
lib.fastmax_debug_trace_fetch_apply2.argtypes = [C.POINTER(C.c_longlong)]
lib.fastmax_debug_trace_fetch_apply2(tab)
t = [[tab[e * 40 + j] for j in range(40)] for e in range(8)]
t0 = t[0][0]
print("tile 0 of one apply CTA (forward); cycles from the issuer's first event")
print("pair | issuer: wait t_free, seen, MMAs+commit issued | read-out warp 0: wait t_full, seen, done+arrive | warp 3 done")
for j in range(33):
    print(f"{j:4d} " + " ".join(f"{t[e][j] - t0:8d}" for e in range(7)))

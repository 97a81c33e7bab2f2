import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fastmax_experiments_b200 import _lib
from fastmax_experiments_b200.fastmax import fastmax_fwd_raw, fastmax_bwd_raw
lib = _lib.load()
B, H, N = 8, 32, 16384
gen = torch.Generator(device="cuda").manual_seed(0)
q, k, v, do = (torch.randn(B, H, N, 64, device="cuda", generator=gen).to(torch.bfloat16) for _ in range(4))
for _ in range(2):
    o, g = fastmax_fwd_raw(q, k, v, True, 1.0 / 64, 2, _lib.ALGO_MOMENTS)
    fastmax_bwd_raw(q, k, v, o, g, do, True, 1.0 / 64, 2, _lib.ALGO_MOMENTS)
torch.cuda.synchronize()
tab = (C.c_longlong * (16 * 16))()
lib.fastmax_debug_trace_fetch_segb.argtypes = [C.POINTER(C.c_longlong)]
lib.fastmax_debug_trace_fetch_segb(tab)
t = [[tab[e * 16 + j] for j in range(16)] for e in range(16)]
t0 = t[0][0]
print("CTA events: start, loop done, acc done, stores issued, end")
print(" ".join(f"{x - t0:8d}" for x in t[0][:5]))
print("pair  top  s_ready  computed  arrived | issuer: p_ready seen, acc issued, scores(np+2) issued")
for j in range(14):
    print(f"{j:4d} " + " ".join(f"{t[e][j] - t0:8d}" for e in (1, 2, 3, 4, 5, 6, 7)))

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
tab = (C.c_longlong * (16 * 16))()
lib.fastmax_debug_trace_fetch_seg.argtypes = [C.POINTER(C.c_longlong)]
lib.fastmax_debug_trace_fetch_seg(tab)
t = [[tab[e * 16 + j] for j in range(16)] for e in range(16)]
t0 = t[0][0]
print("CTA events: start, setup done, q landed, loop done, acc done, stores read, end")
print(" ".join(f"{x - t0:8d}" for x in t[0][:7]))
print("pair  top  s_ready  ld+bar  poly  st+bar+arrive | pv issue")
for j in range(10):
    print(f"{j:4d} " + " ".join(f"{t[e][j] - t0:8d}" for e in (1, 2, 3, 4, 5, 8)))
print("epilogue: g parts done, barrier passed")
print(" ".join(f"{t[6][j] - t0:8d}" for j in range(2)))
print("issuer: PV(n) after wait, after PV + S(n+2) issue")
for j in range(10):
    print(f"{j:4d} " + " ".join(f"{t[e][j] - t0:8d}" for e in (8, 12)))
tab2 = (C.c_longlong * (2 * 16 * 16))()
lib.fastmax_debug_trace_fetch_seg2.argtypes = [C.POINTER(C.c_longlong)]
lib.fastmax_debug_trace_fetch_seg2(tab2)
for ev, name in ((0, "s_ready seen"), (1, "arrived")):
    print(name, "per warp (rows = pair, cols = warp 0..15)")
    for j in range(10):
        print(f"{j:3d} " + " ".join(f"{tab2[(ev * 16 + w) * 16 + j] - t0:6d}" for w in range(16)))

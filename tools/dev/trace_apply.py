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
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
if which == "bwd":
    fastmax_bwd_raw(q, k, v, o, g, do, True, 1.0 / 64, 2, _lib.ALGO_MOMENTS)   # last apply launch = dk (GEN)
tab = (C.c_longlong * (8 * 40))()
lib.fastmax_debug_trace_fetch_apply.argtypes = [C.POINTER(C.c_longlong)]
lib.fastmax_debug_trace_fetch_apply(tab)
t = [[tab[e * 40 + s] for s in range(33)] for e in range(8)]
t0 = min(x for row in t for x in row if x)
names = ["gen:start", "gen:buf free", "gen:arrived", "iss:probe", "iss:ready", "iss:committed", "prod:start", "prod:stage free"]
print("pair " + " ".join(f"{n:>15}" for n in names))
for s in range(33):
    print(f"{s:4d} " + " ".join(f"{(t[e][s] - t0) if t[e][s] else -1:15d}" for e in range(8)))

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
    o, g = fastmax_fwd_raw(q, k, v, True, 1.0 / 64, 1, _lib.ALGO_MOMENTS)
    fastmax_bwd_raw(q, k, v, o, g, do, True, 1.0 / 64, 1, _lib.ALGO_MOMENTS)
tab = (C.c_longlong * (4 * 14 * 160))()
lib.fastmax_debug_trace_fetch_p1.argtypes = [C.POINTER(C.c_longlong)]
lib.fastmax_debug_trace_fetch_p1(tab)
names = ["tile start", "operands in", "S0 ready", "P0 stored", "S1 ready", "P1 stored", "MMAs done", "tile end"]
for mode, mn in enumerate(("fwd", "dq", "dv", "dk")):
    t = [[tab[(mode * 14 + e) * 160 + s] for s in range(128)] for e in range(14)]
    t0 = t[0][0]
    print(f"--- p1 scan {mn}: CTA 0, thread 0; deltas between consecutive events of a tile, and tile period")
    print("tile " + " ".join(f"{n:>12}" for n in names[1:]) + "     period")
    for s in (1, 2, 3, 10, 20, 50, 100, 126):
        row = [t[e][s] for e in range(8)]
        print(f"{s:4d} " + " ".join(f"{b - a:12d}" for a, b in zip(row, row[1:])) + f" {t[0][s + 1] - t[0][s] if s + 1 < 128 else 0:10d}")
    print("total cycles", t[7][127] - t0)
    print("  auxiliary issuer, relative to the start of the tile: P1(it-1) seen, PV0(it-1) seen, operands(it) seen, S0(it) issued, go+B2 seen, update issued")
    for s in (2, 3, 10, 50, 100):
        print(f"{s:4d} " + " ".join(f"{t[e][s] - t[0][s]:9d}" for e in range(8, 14)))

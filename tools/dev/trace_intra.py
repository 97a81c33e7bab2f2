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
tab = (C.c_longlong * (3 * 8 * 48))()
lib.fastmax_debug_trace_fetch_intra.argtypes = [C.POINTER(C.c_longlong)]
lib.fastmax_debug_trace_fetch_intra(tab)
names = ["pair start", "scores issued", "scores ready", "poly done", "P stored+bar", "acc issued"]
for mode, mn in enumerate(("fwd", "dq", "dkdv")):
    t = [[tab[(mode * 8 + e) * 48 + s] for s in range(48)] for e in range(6)]
    vals = [x for row in t for x in row if x]
    if not vals:
        continue
    t0 = min(vals)
    print(f"--- intra {mn}: one CTA, warpgroup 0 (cycles from its first event)")
    print("pair " + " ".join(f"{n:>14}" for n in names) + "   (deltas)")
    for s in range(48):
        if not t[0][s]:
            break
        row = [t[e][s] - t0 for e in range(6)]
        print(f"{s:4d} " + " ".join(f"{x:14d}" for x in row) + "   " + " ".join(f"{b - a:5d}" for a, b in zip(row, row[1:])))

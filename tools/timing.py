"""Developer helper (GPU box): in-kernel cycle counters of a -DFM_TIMING build.

    FASTMAX_NVCC_EXTRA=-DFM_TIMING python -m fastmax_experiments_b200.build   # here
    gpurun -- python tools/timing.py [B H N]
"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastmax_experiments_b200 import _lib
from fastmax_experiments_b200.fastmax import fastmax_fwd_raw, fastmax_bwd_raw

B, H, N = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (8, 32, 16384)))
lib = _lib.load()
gen = torch.Generator(device="cuda").manual_seed(0)
q, k, v, do = (torch.randn(B, H, N, 64, device="cuda", generator=gen).to(torch.bfloat16) for _ in range(4))
tab = (C.c_ulonglong * 64)()


def fetch(name, reset=1):
    f = getattr(lib, "fastmax_debug_timing_fetch_" + name, None)
    if f is None:
        return None
    f.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    f(tab, reset)
    return list(tab)


for _ in range(2):
    o, g = fastmax_fwd_raw(q, k, v, True, 1.0 / 64, 2, _lib.ALGO_MOMENTS)
fetch("moments"); fetch("intra"); fetch("apply")
o, g = fastmax_fwd_raw(q, k, v, True, 1.0 / 64, 2, _lib.ALGO_MOMENTS)
def show_intra(t):
    for mode, nm in enumerate(("fwd", "dq", "dkdv")):
        c = t[8 * mode]
        if not c:
            continue
        v = [round(t[8 * mode + i] / c) for i in range(1, 8)]
        print(f"   intra {nm}: ctas {c}; WG0 warp: setup {v[0]}, operands landed {v[1]}, item loop end {v[2]} "
              f"(wait scores {v[3]}, wait warpgroup barrier {v[4]}, items {v[6]}), accumulators final {v[5]}, "
              f"CTA end {round(t[24 + mode] / c)}")


def show_apply(t, what):
    n = max(t[0], 1)
    print(f"apply ({what}): CTAs with state {t[0]}; per CTA (warp 0 / issuer 0): generator loop {t[1] // n}, "
          f"of which waiting for a free A buffer {t[2] // n}, st+fence+arrive {t[3] // n}; epilogue {t[4] // n}; "
          f"issuer: in slab loads (ring stage free) {t[5] // n}, waiting slab landed {t[6] // n}, waiting A ready {t[7] // n}, "
          f"issue+commit region {t[9] // n}")


for name in ("moments", "intra", "apply"):
    t = fetch(name)
    if t is None:
        continue
    if name == "intra":
        show_intra(t)
        continue
    if name == "apply":
        show_apply(t, "fwd")
        continue
    n = max(t[0], 1)
    print(name, "fwd: ctas", t[0], " per-CTA avg cycles by slot:")
    print("  ", {i: round(x / n) for i, x in enumerate(t) if x and 0 < i < 16})
    if t[16]:
        nb = t[16]
        print(f"   build, cx=0 CTAs, per boundary ({nb} boundaries): snapshot warps idle {t[17] // nb}, read-out + staging "
              f"{t[18] // nb}, stores / mirror {t[19] // nb}; issuer 0 per block ({t[22]} blocks): stalled on the snapshot "
              f"{t[20] / max(t[22], 1):.0f}, waiting for the generators {t[21] / max(t[22], 1):.0f}")
        if t[24]:
            print(f"   special CTAs: per boundary ({t[24]}): idle {t[25] // t[24]}, read-out + direct stores {t[26] // t[24]}; "
                  f"issuer 0 per block: stalled on the snapshot {t[28] / max(t[30], 1):.0f}, waiting for the generators "
                  f"{t[29] / max(t[30], 1):.0f}")
        print(f"   whole-CTA cycles: regular {t[34] // max(t[32], 1)} ({t[32]} CTAs), special {t[35] // max(t[33], 1)} ({t[33]} CTAs)")
if "--bwd" in sys.argv:
    fastmax_bwd_raw(q, k, v, o, g, do, True, 1.0 / 64, 2, _lib.ALGO_MOMENTS)
    for name in ("moments", "intra", "apply"):
        t = fetch(name)
        if t is None:
            continue
        if name == "intra":
            show_intra(t)
            continue
        if name == "apply":
            show_apply(t, "dq + dv + dk")
            continue
        n = max(t[0], 1)
        print(name, "bwd: ctas", t[0], {i: round(x / n) for i, x in enumerate(t) if x and i})

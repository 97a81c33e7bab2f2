"""Developer probe (GPU box): tcgen05.mma rate while other warps stream tcgen05.ld / tcgen05.st."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastmax_experiments_b200 import _lib
lib = _lib.load_dev()
f = lib.fastmax_selftest_contention
f.restype = ctypes.c_int
f.argtypes = [ctypes.c_int] * 4 + [ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(40, dtype=torch.int64, device="cuda")
names = {0: "TS N=64", 1: "SS N=64", 2: "SS N=128", 3: "TS N=128", 4: "no MMA"}
reps = 8192
for kind in (0, 1):
    for mode in (4, 2, 0, 3):
        for ldwarps in (0, 4, 8, 16):
            if mode == 4 and ldwarps == 0:
                continue
            out.zero_()
            _lib.check(f(mode, reps, ldwarps, kind, 1040, out.data_ptr(), None), "contention")
            torch.cuda.synchronize()
            o = out.tolist()
            mma = o[0] / reps
            cols = [16 * o[3 + 2 * w] / max(o[2 + 2 * w], 1) for w in range(ldwarps)]
            per_warp = (1 / (sum(cols) / len(cols))) if cols else 0
            print(f"{'ld' if kind == 0 else 'st'} warps={ldwarps:2d} {names[mode]:9s}: {mma:6.1f} cyc/mma; "
                  f"{per_warp:5.2f} cyc/column/warp ({per_warp / max(1, ldwarps // 4):5.2f} per quadrant)")

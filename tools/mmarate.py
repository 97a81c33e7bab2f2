import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastmax_experiments_b200 import _lib
lib = _lib.load_dev()
lib.fastmax_selftest_mmarate.restype = ctypes.c_int
lib.fastmax_selftest_mmarate.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
reps = 2048
for rot in (64, 64+128):
    for nsel in (0, 1, 2):
        if nsel == 2 and rot & 128: continue
        for v in (0, 3):
            variant = v | (nsel << 2) | rot
            for _ in range(2):
                _lib.check(lib.fastmax_selftest_mmarate(variant, reps, out.data_ptr(), None), "mmarate")
                torch.cuda.synchronize()
            o = out.tolist()
            print(f"N={64<<nsel:3d} A={'tmem' if v&2 else 'smem'} B={'MN' if v&1 else 'K '} rotate={bool(rot&16)} issuers={4 if rot&128 else 1}: issue {o[0]/reps:6.1f} cyc/mma, total {o[1]/reps:6.1f} cyc/mma")

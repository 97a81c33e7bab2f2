import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastmax_experiments_b200 import _lib
lib = _lib.load_dev()
lib.fastmax_selftest_ldrate.restype = ctypes.c_int
lib.fastmax_selftest_ldrate.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
iters = 2048
for threads in (32, 128, 512):
    for (w, f) in ((16, 1), (16, 2), (16, 4), (32, 1), (32, 2)):
        for _ in range(2):
            _lib.check(lib.fastmax_selftest_ldrate(w, f, iters, threads, out.data_ptr(), None), "ldrate")
            torch.cuda.synchronize()
        cyc = out.tolist()[0] / iters
        print(f"threads={threads:4d} ld x{w} inflight={f}: {cyc:7.1f} cycles per wait group = {cyc/(w*f):6.2f} cycles per column")

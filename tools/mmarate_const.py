import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastmax_experiments_b200 import _lib
lib = _lib.load()
lib.fastmax_selftest_mmarate_const.restype = ctypes.c_int
lib.fastmax_selftest_mmarate_const.argtypes = [ctypes.c_int, ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(4, dtype=torch.int64, device="cuda")
reps = 4096
base = 0
for trial in range(3):
    _lib.check(lib.fastmax_selftest_mmarate_const(reps, base, out.data_ptr(), None), "mmarate_const")
    torch.cuda.synchronize()
    o = out.tolist()
    print(f"smem_base passed {base}, actual {o[2]}, tmem base {o[3]}: issue {o[0]/reps:.1f} cyc/mma, total {o[1]/reps:.1f} cyc/mma")
    base = o[2]

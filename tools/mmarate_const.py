"""Developer probe (GPU box): cycles per tcgen05.mma issued by one thread with constant-bank operands.
mode 0: A in tensor memory, N=64; 1: A in shared memory, N=64; 2: A in smem, N=128; 3: A in tmem, N=128."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastmax_experiments_b200 import _lib
lib = _lib.load_dev()
lib.fastmax_selftest_mmarate_const.restype = ctypes.c_int
lib.fastmax_selftest_mmarate_const.argtypes = [ctypes.c_int, ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
out = torch.zeros(4, dtype=torch.int64, device="cuda")
reps = 4096
base = 1024
names = {0: 'TS N=64, 4 accumulators', 1: 'SS N=64, 4 acc', 2: 'SS N=128, 2 acc', 3: 'TS N=128, 2 acc',
         5: 'SS N=128, 1 acc', 6: 'TS N=64, 1 acc', 7: 'SS N=64, 1 acc', 8: 'TS N=128, 1 acc'}
for mode in (0, 1, 2, 3, 5, 6, 7, 8):
    for trial in range(2):
        _lib.check(lib.fastmax_selftest_mmarate_const(reps, base, out.data_ptr(), None, mode), "mmarate_const")
        torch.cuda.synchronize()
        o = out.tolist()
        print(f"mode {mode} ({names[mode]}): smem_base passed {base}, actual {o[2]}, tmem base {o[3]}: issue {o[0]/reps:.1f} cyc/mma, "
              f"total {o[1]/reps:.1f} cyc/mma")
        base = o[2]

import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastmax_experiments_b200 import _lib
lib = _lib.load_dev()
lib.fastmax_selftest_alurate.restype = ctypes.c_int
lib.fastmax_selftest_alurate.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
names = ["mul.bf16x2", "fma.bf16x2", "mul.f16x2", "mul.f32", "mul.f32x2"]
iters = 4096
for threads in (32, 128, 512):
    for kind in range(5):
        for _ in range(2):
            _lib.check(lib.fastmax_selftest_alurate(kind, iters, threads, out.data_ptr(), None), "alurate")
            torch.cuda.synchronize()
        cyc = out.tolist()[0]
        warps_per_smsp = max(1, threads // 128)
        print(f"threads={threads:4d} {names[kind]:11s}: {cyc/(iters*8):6.2f} cycles per instruction per warp; "
              f"{cyc/(iters*8*warps_per_smsp):6.2f} per SMSP issue")

"""Developer probe (GPU box): N converged warps issuing groups of four TS N=64 MMAs + commits."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastmax_experiments_b200 import _lib
lib = _lib.load_dev()
f = lib.fastmax_selftest_multi_issue
f.restype = ctypes.c_int
f.argtypes = [ctypes.c_int] * 3 + [ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(8, dtype=torch.int64, device="cuda")
blocks = 4096
for nissue in (1, 2, 4):
    for commits in (0, 1, 5):
        out.zero_()
        _lib.check(f(blocks, nissue, commits, 1040, out.data_ptr(), None), "multi_issue")
        torch.cuda.synchronize()
        o = out.tolist()[:nissue]
        print(f"issuers={nissue} commits/group={commits}: cycles per group of 4 MMAs per issuer: "
              + ", ".join(f"{x / blocks:.0f}" for x in o) + f"  -> {max(o) / blocks / (4 * nissue):.1f} cycles per MMA overall")

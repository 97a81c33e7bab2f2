"""Developer helper (GPU box): run the moments path a few times on one shape (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastmax_experiments_b200 import _lib
from fastmax_experiments_b200.fastmax import fastmax_fwd_raw, fastmax_bwd_raw, saved_state_bytes

B, H, N, D = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (2, 32, 16384, 64)))
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
bwd = "--bwd" in sys.argv
keep = "--keep" in sys.argv      # keep the forward states for backward, like bench.py / the autograd op
gen = torch.Generator(device="cuda").manual_seed(0)
q, k, v, do = (torch.randn(B, H, N, D, device="cuda", generator=gen).to(torch.bfloat16) for _ in range(4))
saved = torch.empty(saved_state_bytes(q, True, 2, _lib.ALGO_MOMENTS), dtype=torch.uint8, device="cuda") if keep else None
for _ in range(reps):
    o, g = fastmax_fwd_raw(q, k, v, True, 1.0 / 64, 2, _lib.ALGO_MOMENTS, saved=saved)
    if bwd:
        fastmax_bwd_raw(q, k, v, o, g, do, True, 1.0 / 64, 2, _lib.ALGO_MOMENTS, saved=saved)
torch.cuda.synchronize()
print("ok")

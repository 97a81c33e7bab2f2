import sys, os
sys.path.insert(0, "/root/repo")
import torch
from fastmax_experiments_b200 import _lib
from fastmax_experiments_b200.fastmax import fastmax_fwd_raw, fastmax_bwd_raw
P_ORDER = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B,H,N,D=8,32,16384,64
gen = torch.Generator(device="cuda").manual_seed(0)
q, k, v, do = (torch.randn(B, H, N, D, device="cuda", generator=gen).to(torch.bfloat16) for _ in range(4))
for causal in (True, False):
    for _ in range(2):
        o, g = fastmax_fwd_raw(q, k, v, causal, 1/64, P_ORDER, _lib.ALGO_AUTO); fastmax_bwd_raw(q, k, v, o, g, do, causal, 1/64, P_ORDER, _lib.ALGO_AUTO)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    o, g = fastmax_fwd_raw(q, k, v, causal, 1/64, P_ORDER, _lib.ALGO_AUTO); fastmax_bwd_raw(q, k, v, o, g, do, causal, 1/64, P_ORDER, _lib.ALGO_AUTO)
    torch.cuda.synchronize()
    print("causal" if causal else "full", [(n, round(ms,3)) for n, ms in _lib.profile_fetch()])
    _lib.profile_enable(False)

import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fastmax_experiments_b200 import _lib
from fastmax_experiments_b200.fastmax import fastmax_fwd_raw, fastmax_bwd_raw
_lib.profile_enable(True)
for (B, H, N, D, causal, p) in ((1, 2, 700, 64, True, 1), (1, 2, 2700, 64, True, 1), (1, 2, 700, 64, False, 1), (1, 2, 300, 128, True, 2), (1, 1, 513, 128, False, 1)):
    gen = torch.Generator(device="cuda").manual_seed(N)
    q, k, v, do = (torch.randn(B, H, N, D, device="cuda", generator=gen).to(torch.bfloat16) for _ in range(4))
    o, g = fastmax_fwd_raw(q, k, v, causal, 1.0 / 64, p, _lib.ALGO_AUTO)
    fastmax_bwd_raw(q, k, v, o, g, do, causal, 1.0 / 64, p, _lib.ALGO_AUTO)
    torch.cuda.synchronize()
print(sorted({n for n, _ in _lib.profile_fetch()}))

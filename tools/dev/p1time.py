import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fastmax_experiments_b200 import _lib
from fastmax_experiments_b200.fastmax import fastmax_fwd_raw, fastmax_bwd_raw
B, H, N = 8, 32, 16384
CAUSAL = os.environ.get('P1_CAUSAL', '1') == '1'
PP = int(os.environ.get('P1_P', '1'))
gen = torch.Generator(device="cuda").manual_seed(0)
q, k, v, do = (torch.randn(B, H, N, 64, device="cuda", generator=gen).to(torch.bfloat16) for _ in range(4))
def step():
    o, g = fastmax_fwd_raw(q, k, v, CAUSAL, 1 / 64, PP, _lib.ALGO_MOMENTS)
    return fastmax_bwd_raw(q, k, v, o, g, do, CAUSAL, 1 / 64, PP, _lib.ALGO_MOMENTS)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
_lib.profile_enable(True); step(); torch.cuda.synchronize(); pr = _lib.profile_fetch(); _lib.profile_enable(False)
print("%s: %.3f ms/step, %.1f GB/s algorithmic  %s" % ((os.environ.get("FASTMAX_LIB", "x/default/y").split("/")[-2]), ms, B * H * N * 1544 / ms / 1e6, " ".join("%s %.3f" % (n.replace("p1_scan_", ""), t) for n, t in pr)))

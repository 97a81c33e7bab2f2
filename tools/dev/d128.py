import torch, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fastmax_experiments_b200 import _lib
from fastmax_experiments_b200.fastmax import fastmax_fwd_raw, fastmax_bwd_raw
from oracle import fastmax_oracle as orc
BWD = os.environ.get("D128_BWD", "0") == "1"
for (B, H, N, causal, p) in ((1, 2, 128, True, 2), (1, 2, 300, True, 2), (2, 1, 700, False, 2), (1, 2, 513, True, 1), (1, 1, 260, False, 1)):
    gen = torch.Generator().manual_seed(N + p)
    q, k, v, do = (torch.randn(B, H, N, 128, generator=gen).to(torch.bfloat16) for _ in range(4))
    ref = orc.dense_fwd_bwd(q.double(), k.double(), v.double(), do.double(), mask=causal, p=p)
    c = orc.denominator_c(128)
    qc, kc, vc, dc = (t.cuda() for t in (q, k, v, do))
    o, g = fastmax_fwd_raw(qc, kc, vc, causal, 1.0 / c, p, _lib.ALGO_AUTO)
    outs = [o, g]
    if BWD:
        outs += list(fastmax_bwd_raw(qc, kc, vc, o, g, dc, causal, 1.0 / c, p, _lib.ALGO_AUTO))
    torch.cuda.synchronize()
    print((B, H, N, causal, p), " ".join("%s %.2e" % (nm, orc.rel_err(x, r)) for nm, x, r in zip(("o", "g", "dq", "dk", "dv"), outs, ref)))
B, H, N = 1, 32, 4096
gen = torch.Generator(device="cuda").manual_seed(0)
q, k, v, do = (torch.randn(B, H, N, 128, device="cuda", generator=gen).to(torch.bfloat16) for _ in range(4))
c = orc.denominator_c(128)
_lib.profile_enable(True)
for _ in range(2):
    o, g = fastmax_fwd_raw(q, k, v, True, 1.0 / c, 2, _lib.ALGO_AUTO)
    if BWD: fastmax_bwd_raw(q, k, v, o, g, do, True, 1.0 / c, 2, _lib.ALGO_AUTO)
torch.cuda.synchronize()
print(_lib.profile_fetch())

"""Developer check (GPU box): moments path vs the dense oracle, per tensor, several shapes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastmax_experiments_b200 import _lib
from fastmax_experiments_b200.fastmax import fastmax_fwd_raw, fastmax_bwd_raw
from oracle import fastmax_oracle as orc

cases = [  # B,H,N,D,p,causal
    (1, 1, 600, 64, 2, True), (1, 2, 1024, 64, 2, True), (2, 2, 1500, 64, 2, True),
    (1, 1, 130, 64, 2, True), (1, 2, 700, 64, 2, False), (1, 2, 1100, 64, 1, True), (1, 1, 900, 64, 1, False),
    (1, 1, 4096, 64, 2, True),
]
only_fwd = "--fwd" in sys.argv
for (B, H, N, D, p, causal) in cases:
    gen = torch.Generator().manual_seed(N + p)
    q, k, v, do = (torch.randn(B, H, N, D, generator=gen).to(torch.bfloat16) for _ in range(4))
    ref = orc.dense_fwd_bwd(q.double(), k.double(), v.double(), do.double(), mask=causal, p=p)
    qd, kd, vd, dod = (t.cuda() for t in (q, k, v, do))
    inv_c = 1.0 / orc.denominator_c(D)
    try:
        o, g = fastmax_fwd_raw(qd, kd, vd, causal, inv_c, p, _lib.ALGO_MOMENTS)
        torch.cuda.synchronize()
        msg = f"o {orc.rel_err(o, ref[0]):.2e} g {orc.rel_err(g, ref[1]):.2e}"
        if not only_fwd:
            # feed the oracle's o, g (in bf16/fp32) so backward errors are isolated from forward ones
            o_ref = ref[0].to(torch.bfloat16).cuda()
            g_ref = ref[1].float().cuda()
            dq, dk, dv = fastmax_bwd_raw(qd, kd, vd, o_ref, g_ref, dod, causal, inv_c, p, _lib.ALGO_MOMENTS)
            torch.cuda.synchronize()
            msg += f" dq {orc.rel_err(dq, ref[2]):.2e} dk {orc.rel_err(dk, ref[3]):.2e} dv {orc.rel_err(dv, ref[4]):.2e}"
        print((B, H, N, D, p, causal), msg, flush=True)
    except Exception as e:
        print((B, H, N, D, p, causal), "FAILED", repr(e)[:300], flush=True)
        break

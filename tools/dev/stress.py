"""Repeatability stress of the second-generation intra kernels: one dense segment (no moment states, so the result is
deterministic), many launches, every output compared bit for bit with the first launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fastmax_experiments_b200 import _lib
from fastmax_experiments_b200.fastmax import fastmax_fwd_raw, fastmax_bwd_raw
bad = 0
for (B, H, N) in ((1, 4, 2048), (2, 3, 1537), (64, 32, 1024)):
    gen = torch.Generator(device="cuda").manual_seed(N)
    q, k, v, do = (torch.randn(B, H, N, 64, device="cuda", generator=gen).to(torch.bfloat16) for _ in range(4))
    ref = None
    for it in range(150 if B < 64 else 30):
        o, g = fastmax_fwd_raw(q, k, v, True, 1.0 / 64, 2, _lib.ALGO_MOMENTS)
        out = (o, g) + tuple(fastmax_bwd_raw(q, k, v, o, g, do, True, 1.0 / 64, 2, _lib.ALGO_MOMENTS))
        if ref is None:
            ref = [t.clone() for t in out]
        else:
            for nm, a, b in zip(("o", "g", "dq", "dk", "dv"), out, ref):
                if not torch.equal(a, b):
                    bad += 1
                    print("MISMATCH", (B, H, N), it, nm, float((a.float() - b.float()).abs().max()))
    torch.cuda.synchronize()
    print((B, H, N), "done")
print("mismatches", bad)

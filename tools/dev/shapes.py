"""Ad-hoc edge-shape sweep of the p = 2 causal moment path against the fp64 dense formula (ragged tails, one dense
segment of 1024 / 2048 tokens, several segments, grouped queries)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fastmax_experiments_b200 import _lib, fastmax_gqa
from fastmax_experiments_b200.fastmax import fastmax_fwd_raw, fastmax_bwd_raw
from fastmax_experiments_b200.gqa import expand_kv
from oracle import fastmax_oracle as orc
worst = 0.0
for (B, H, Hkv, N) in [(1, 1, 1, 257), (1, 2, 2, 513), (1, 1, 1, 640), (1, 2, 1, 1025), (1, 1, 1, 1537), (1, 1, 1, 2047), (1, 2, 2, 2049),
                       (1, 1, 1, 2175), (1, 2, 1, 2689), (1, 1, 1, 3073), (2, 2, 2, 3200), (1, 4, 2, 4097)]:
    g = torch.Generator().manual_seed(N)
    q = torch.randn(B, H, N, 64, generator=g).to(torch.bfloat16)
    k, v = (torch.randn(B, Hkv, N, 64, generator=g).to(torch.bfloat16) for _ in range(2))
    do = torch.randn(B, H, N, 64, generator=g).to(torch.bfloat16)
    qa, ka, va = (t.cuda().requires_grad_(True) for t in (q, k, v))
    o = fastmax_gqa(qa, ka, va, mask=True, p=2)
    o.backward(do.cuda())
    ke, ve = (expand_kv(t, H // Hkv).double() for t in (k, v))
    ro, _, rdq, rdk, rdv = orc.dense_fwd_bwd(q.double(), ke, ve, do.double(), mask=True, p=2)
    fold = lambda t: t.view(B, Hkv, H // Hkv, N, 64).sum(2)
    errs = [orc.rel_err(x.double().cpu(), r) for x, r in ((o, ro), (qa.grad, rdq), (ka.grad, fold(rdk)), (va.grad, fold(rdv)))]
    worst = max(worst, max(errs))
    print((B, H, Hkv, N), " ".join(f"{e:.2e}" for e in errs), "OK" if max(errs) < 2e-2 else "FAIL")
print("worst", worst)

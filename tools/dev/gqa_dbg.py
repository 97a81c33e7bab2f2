import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fastmax_experiments_b200 import fastmax, fastmax_gqa
from fastmax_experiments_b200.gqa import expand_kv
from oracle import fastmax_oracle as orc
for (B, H, Hkv, N, D, causal) in [(1, 8, 2, 2600, 64, True), (1, 4, 2, 512, 64, True), (1, 8, 2, 2600, 64, False), (2, 6, 2, 700, 64, True)]:
    g = torch.Generator().manual_seed(11)
    q = torch.randn(B, H, N, D, generator=g).to(torch.bfloat16)
    k, v = (torch.randn(B, Hkv, N, D, generator=g).to(torch.bfloat16) for _ in range(2))
    do = torch.randn(B, H, N, D, generator=g).to(torch.bfloat16)
    mask = True if causal else False
    qa, ka, va = (t.cuda().requires_grad_(True) for t in (q, k, v))
    oa = fastmax_gqa(qa, ka, va, mask=mask, p=2); oa.backward(do.cuda())
    qb, kb, vb = (t.cuda().requires_grad_(True) for t in (q, k, v))
    ob = fastmax(qb, expand_kv(kb, H // Hkv), expand_kv(vb, H // Hkv), mask=mask, p=2); ob.backward(do.cuda())
    ke, ve = (expand_kv(t, H // Hkv).double() for t in (k, v))
    ro, _, rdq, rdk, rdv = orc.dense_fwd_bwd(q.double(), ke, ve, do.double(), mask=mask, p=2)
    fold = lambda t: t.view(B, Hkv, H // Hkv, N, D).sum(2)
    print((B, H, Hkv, N, D, causal))
    for nm, x, y, r in (("o", oa, ob, ro), ("dq", qa.grad, qb.grad, rdq), ("dk", ka.grad, kb.grad, fold(rdk)), ("dv", va.grad, vb.grad, fold(rdv))):
        print(f"  {nm}: gqa vs dense {orc.rel_err(x.double().cpu(), r):.3e}  mha vs dense {orc.rel_err(y.double().cpu(), r):.3e}  gqa vs mha {orc.rel_err(x.double().cpu(), y.double().cpu()):.3e}")
        if nm == "o":
            e = (x.double() - y.double()).abs().amax(-1)[0].cpu()   # (H, N)
            print("   per-head max |diff| :", [f"{float(e[h].max()):.2e}" for h in range(H)], "first bad token per head", [int((e[h] > 0).nonzero()[0]) if (e[h] > 0).any() else -1 for h in range(H)])

import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fastmax_experiments_b200 import _lib
from fastmax_experiments_b200.fastmax import fastmax_fwd_raw, fastmax_bwd_raw
from oracle import fastmax_oracle as orc
for (B, H, N, D, causal, p) in ((1, 2, 128, 64, True, 2), (1, 2, 300, 64, True, 2), (2, 1, 700, 64, False, 2), (1, 2, 513, 64, True, 1),
                                (1, 1, 260, 64, False, 1), (1, 2, 1024, 32, True, 2), (1, 1, 333, 48, False, 2), (1, 1, 2048, 64, True, 2)):
    gen = torch.Generator().manual_seed(N + p)
    q, k, v, do = (torch.randn(B, H, N, D, generator=gen) for _ in range(4))
    ref = orc.dense_fwd_bwd(q.double(), k.double(), v.double(), do.double(), mask=causal, p=p)
    c = orc.denominator_c(D)
    qc, kc, vc, dc = (t.cuda() for t in (q, k, v, do))
    _lib.profile_enable(True)
    o, g = fastmax_fwd_raw(qc, kc, vc, causal, 1.0 / c, p, _lib.ALGO_AUTO)
    outs = [o, g] + list(fastmax_bwd_raw(qc, kc, vc, o, g, dc, causal, 1.0 / c, p, _lib.ALGO_AUTO))
    torch.cuda.synchronize()
    names = sorted({n for n, _ in _lib.profile_fetch()})
    _lib.profile_enable(False)
    print((B, H, N, D, causal, p), " ".join("%s %.2e" % (nm, orc.rel_err(x, r)) for nm, x, r in zip(("o", "g", "dq", "dk", "dv"), outs, ref)), names[:2])
for (B, H, N) in ((64, 32, 1024), (64, 32, 2048)):
    gen = torch.Generator(device="cuda").manual_seed(0)
    q, k, v, do = (torch.randn(B, H, N, 64, device="cuda", generator=gen) for _ in range(4))
    def step():
        o, g = fastmax_fwd_raw(q, k, v, True, 1 / 64, 2, _lib.ALGO_AUTO)
        return fastmax_bwd_raw(q, k, v, o, g, do, True, 1 / 64, 2, _lib.ALGO_AUTO)
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): step()
    e1.record(); torch.cuda.synchronize()
    _lib.profile_enable(True); step(); torch.cuda.synchronize(); pr = _lib.profile_fetch(); _lib.profile_enable(False)
    print("fp32 p=2 causal B=%d N=%d: %.2f ms/step | %s" % (B, N, e0.elapsed_time(e1) / 3, " ".join("%s %.2f" % (n.replace("dense_split_", ""), t) for n, t in pr)))
    del q, k, v, do

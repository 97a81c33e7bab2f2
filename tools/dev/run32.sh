( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "p1 or (seeded and moments) or linearmax or autograd or first_order or variants" 2>&1 | tail -8 )
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
from fastmax_experiments_b200 import _lib
from fastmax_experiments_b200.fastmax import fastmax_fwd_raw, fastmax_bwd_raw
for (B,H,N) in ((2,32,65536),(1,8,131072),(8,32,16384)):
    gen=torch.Generator(device="cuda").manual_seed(0)
    q,k,v,do=(torch.randn(B,H,N,64,device="cuda",generator=gen).to(torch.bfloat16) for _ in range(4))
    def step():
        o,g=fastmax_fwd_raw(q,k,v,True,1/64,1,_lib.ALGO_MOMENTS)
        return fastmax_bwd_raw(q,k,v,o,g,do,True,1/64,1,_lib.ALGO_MOMENTS)
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): step()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    _lib.profile_enable(True); step(); torch.cuda.synchronize(); pr=_lib.profile_fetch(); _lib.profile_enable(False)
    print("p1 causal B=%d H=%d N=%d: %.3f ms/step, %.1f GB/s algorithmic | %s" % (B,H,N,ms, B*H*N*1544/ms/1e6, " ".join("%s %.3f"%(n,t) for n,t in pr)))
PY

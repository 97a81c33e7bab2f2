"""Developer probe (GPU box): fwd+bwd step as eager launches vs one CUDA-graph replay (is there launch-gap time?)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastmax_experiments_b200 import _lib
from fastmax_experiments_b200.fastmax import fastmax_fwd_raw, fastmax_bwd_raw, saved_state_bytes

B, H, N, D = 8, 32, 16384, 64
gen = torch.Generator(device="cuda").manual_seed(0)
q, k, v, do = (torch.randn(B, H, N, D, device="cuda", generator=gen).to(torch.bfloat16) for _ in range(4))
nb = saved_state_bytes(q, True, 2, _lib.ALGO_AUTO)


def step():
    saved = torch.empty(nb, dtype=torch.uint8, device="cuda")
    o, g = fastmax_fwd_raw(q, k, v, True, 1 / 64, 2, _lib.ALGO_AUTO, saved=saved)
    return fastmax_bwd_raw(q, k, v, o, g, do, True, 1 / 64, 2, _lib.ALGO_AUTO, saved=saved)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / reps * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, host


print("eager : %.3f ms/step on the device, %.3f ms/step of host enqueue" % timeit(step))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(2):
        step()
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr, stream=s):
    out = step()
print("graph : %.3f ms/step on the device, %.3f ms/step of host enqueue" % timeit(gr.replay))

"""fastmax_gqa against fastmax on expanded k, v at the attention shape of TinyLlama-1.1B (32 heads in 4 groups, seq 8192)
and the prologue kernel against the reference's torch ops; CUDA-event times, L2 flushed between iterations."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fastmax_experiments_b200 import fastmax, fastmax_gqa, qkv_prologue, _lib
from fastmax_experiments_b200.gqa import expand_kv, _prologue_torch
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
out = {}
for (B, H, Hkv, N) in [(1, 32, 4, 8192), (4, 32, 4, 8192), (8, 32, 8, 16384)]:
    g = torch.Generator(device=dev).manual_seed(0)
    q = torch.randn(B, H, N, 64, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True)
    k, v = (torch.randn(B, Hkv, N, 64, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True) for _ in range(2))
    do = torch.randn(B, H, N, 64, device=dev, generator=g).to(torch.bfloat16)
    def f_gqa():
        fastmax_gqa(q, k, v, mask=True, p=2).backward(do)
    def f_mha():
        fastmax(q, expand_kv(k, H // Hkv), expand_kv(v, H // Hkv), mask=True, p=2).backward(do)
    t_g, t_m = timeit(f_gqa), timeit(f_mha)
    _lib.profile_enable(True); f_gqa(); torch.cuda.synchronize(); recs = _lib.profile_fetch(); _lib.profile_enable(False)
    out[f"B{B}_H{H}_Hkv{Hkv}_N{N}"] = {"gqa_ms": round(t_g, 3), "expanded_mha_ms": round(t_m, 3), "speedup": round(t_m / t_g, 3),
                                      "gqa_kernels_ms": {n: round(ms, 3) for n, ms in recs}}
# prologue: TinyLlama qkv (B=1, T=8192, 4 groups x (8 + 2) x 64) and a batch of 8
for (B, T, G, qpk, hs) in [(1, 8192, 4, 8, 64), (8, 8192, 4, 8, 64), (8, 4096, 32, 1, 64)]:
    qkv = torch.randn(B, T, G * (qpk + 2) * hs, device=dev).to(torch.bfloat16).requires_grad_(True)
    ang = torch.randn(T, hs // 2, device=dev).repeat(1, 2)
    cos, sin = torch.cos(ang).to(torch.bfloat16), torch.sin(ang).to(torch.bfloat16)
    grads = None
    def f_ours():
        q, k, v = qkv_prologue(qkv, cos, sin, G, qpk, hs, hs)
        torch.autograd.backward((q, k, v), (q.detach(), k.detach(), v.detach()))
    def f_torch():
        q, k, v = _prologue_torch(qkv, cos, sin, G, qpk, hs, hs)
        k, v = expand_kv(k, qpk), expand_kv(v, qpk)
        torch.autograd.backward((q, k, v), (q.detach(), k.detach(), v.detach()))
    t_o, t_t = timeit(f_ours), timeit(f_torch)
    nbytes = 4 * qkv.numel() * 2          # forward read + write, backward read + write
    out[f"prologue_B{B}_T{T}_G{G}_qpk{qpk}"] = {"fused_ms": round(t_o, 4), "torch_ops_ms": round(t_t, 4), "speedup": round(t_t / t_o, 2),
                                                 "fused_GBps_algorithmic": round(nbytes / t_o / 1e6, 1)}
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r2_gqa_prologue.json", "w"), indent=1)

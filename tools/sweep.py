"""Op sweep of BASELINE.json configs[1] on one GPU: B*N = 128k tokens, H=32, D=64, N = 1k..64k, p = 1 and 2,
causal / non-causal, bf16 (tcgen05 moment / scan paths) and fp32 (dense tcgen05 tiles on split bf16 pairs: O(N^2)).
Prints a markdown table: ms per fwd+bwd, tokens/s, and for p=1 the algorithmic HBM rate
(1544 B/unit bf16) against the measured copy bandwidth in MEASURED_PEAKS.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastmax_experiments_b200 import _lib
from fastmax_experiments_b200.fastmax import fastmax_fwd_raw, fastmax_bwd_raw, saved_state_bytes, SAVE_STATES_MAX_BYTES

H, D = 32, 64
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
hbm = peaks["hbm_gbs"]
print(f"| N | B | p | mask | dtype | algo | seg | ms/step | M tokens/s | alg. GB/s (p=1) | of {hbm:.0f} GB/s |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for dtype, ns in ((torch.bfloat16, (1024, 2048, 4096, 16384, 65536)), (torch.float32, (1024, 2048, 4096, 16384))):
    for N in ns:
        B = max(1, min(64, 131072 // N))
        gen = torch.Generator(device="cuda").manual_seed(0)
        q, k, v, do = (torch.randn(B, H, N, D, device="cuda", generator=gen).to(dtype) for _ in range(4))
        for p in (1, 2):
            for causal in (True, False):
                pl = _lib.plan(B, H, N, D, p, causal, _lib.DTYPE_BF16 if dtype == torch.bfloat16 else _lib.DTYPE_F32, _lib.ALGO_AUTO)
                nb = saved_state_bytes(q, causal, p, _lib.ALGO_AUTO) if dtype == torch.bfloat16 else 0
                keep = 0 < nb <= SAVE_STATES_MAX_BYTES

                def step():
                    saved = torch.empty(nb, dtype=torch.uint8, device="cuda") if keep else None
                    o, g = fastmax_fwd_raw(q, k, v, causal, 1.0 / 64, p, _lib.ALGO_AUTO, saved=saved)
                    return fastmax_bwd_raw(q, k, v, o, g, do, causal, 1.0 / 64, p, _lib.ALGO_AUTO, saved=saved)

                for _ in range(2):
                    step()
                torch.cuda.synchronize()
                reps = 3
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                units = B * H * N
                # (fp32 with N >= 128 under ALGO_AUTO: dense tcgen05 tiles on [hi | lo] bf16 pairs, fm_dense128.cu)
                algo_name = 'moments' if pl.algo == 2 else ('dense split' if dtype == torch.float32 and N >= 128 else 'exact')
                es = 2 if dtype == torch.bfloat16 else 4
                gbs = units * (12 * D * es + 8) / (ms * 1e-3) / 1e9
                print(f"| {N} | {B} | {p} | {'causal' if causal else 'full'} | {'bf16' if es == 2 else 'fp32'} | "
                      f"{algo_name} | {pl.seg_len} | {ms:.2f} | {B * N / ms / 1e3:.2f} | "
                      f"{gbs:.0f} | {100 * gbs / hbm:.1f} % |" if p == 1 else
                      f"| {N} | {B} | {p} | {'causal' if causal else 'full'} | {'bf16' if es == 2 else 'fp32'} | "
                      f"{algo_name} | {pl.seg_len} | {ms:.2f} | {B * N / ms / 1e3:.2f} | | |")
        del q, k, v, do
        torch.cuda.empty_cache()

# head size 128 (BASELINE.json config 4: Llama-2-7B, 32 heads of 128): dense tcgen05 tiles, O(N^2)
print()
print("| N | B | H | D | p | mask | dtype | kernels | ms/step | M tokens/s |")
print("|---|---|---|---|---|---|---|---|---|---|")
for N, B in ((1024, 32), (4096, 8), (16384, 1)):
    gen = torch.Generator(device="cuda").manual_seed(0)
    q, k, v, do = (torch.randn(B, H, N, 128, device="cuda", generator=gen).to(torch.bfloat16) for _ in range(4))
    inv_c = 1.0 / (8 * 128 ** 0.5)
    for p in (1, 2):
        for causal in (True, False):
            def step():
                o, g = fastmax_fwd_raw(q, k, v, causal, inv_c, p, _lib.ALGO_AUTO)
                return fastmax_bwd_raw(q, k, v, o, g, do, causal, inv_c, p, _lib.ALGO_AUTO)
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"| {N} | {B} | {H} | 128 | {p} | {'causal' if causal else 'full'} | bf16 | dense128 | {ms:.2f} | {B * N / ms / 1e3:.2f} |")
    del q, k, v, do
    torch.cuda.empty_cache()

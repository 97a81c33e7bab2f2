#!/usr/bin/env python
"""bench.py -- headline measurement of the FastMax hot path (contract: task brief section 4).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # reference CPU algorithm (oracle port)

Workload (BASELINE.json configs[1] at the size `metric` is quoted on):
    FastMax p=2 causal forward+backward, B=8, H=32, N=16384, D=64, bf16 in / fp32 accumulate,
    synthetic N(0,1) q, k, v, do.  One "step" = one fwd+bwd over that batch.
    tokens/s = B*N / t(step).  With N GPUs every rank runs its own batch of 8 sequences
    ((b,h) units shard with no data-path collective): scaling = "weak".

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fastmax_p2_causal_fwd_bwd_tokens_per_s"
UNIT = "tokens/s"
HL = dict(B=8, H=32, N=16384, D=64, p=2, causal=True)


def flops_alg(D: int, p: int = 2) -> tuple[float, float]:
    """Algorithmic flops per (b,h,token) unit, forward and backward."""
    A = D + 1
    if p == 2:
        return 4 * D * D * A + 4 * D * A, 8 * D * D * A + 2 * D ** 3 + 8 * D * A + 2 * D * D
    return 4 * D * A, 8 * D * A + 2 * D * D


def load_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(tflops=d["bf16_tflops"], tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm_gbs=d["hbm_gbs"], source="measured")
    return dict(tflops=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback")


# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while self.nv is not None and not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def finish(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=2)
        return dict(sm_mhz=statistics.median(self.samples) if self.samples else None,
                    sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(self.samples))


# ----------------------------------------------------------------------------------------------
def _cpu_step_fn(n_sample: int, d: int, threads: int):
    """One fwd+bwd of the reference's CPU implementation on a (1,1,n_sample,d) fp32 slice of the workload.
    Returns (callable, kind): the UNMODIFIED attention_mechanisms/fastmax.py from baseline/_ref when it is
    installed ("reference"), else the oracle's factorised restatement of it ("port")."""
    torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(0)
    q, k, v, do = (torch.randn(1, 1, n_sample, d, generator=gen) for _ in range(4))
    try:
        from oracle import ref_loader
        if not ref_loader.installed():
            raise ImportError
        ref_fastmax, _ = ref_loader.reference_op()

        def step():
            qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
            ref_fastmax(qq, kk, vv, mask=True, p=HL["p"]).backward(do)
        return step, "reference"
    except ImportError:
        from oracle import fastmax_oracle as orc
        return (lambda: orc.factorised_fwd_bwd(q, k, v, do, mask=True, p=HL["p"])), "port"


def cpu_rate(n_sample: int, d: int, repeats: int, threads: int) -> tuple[float, float, str]:
    """(tokens/s at H=32, seconds per fwd+bwd of the sample, kind) -- best of `repeats`."""
    step, kind = _cpu_step_fn(n_sample, d, threads)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        step()
        best = min(best, time.perf_counter() - t0)
    return n_sample / best / HL["H"], best, kind


def _cpu_sample_text(kind: str, n_sample: int) -> str:
    what = ("the unmodified attention_mechanisms/fastmax.py (baseline/_ref)" if kind == "reference"
            else "oracle port of fastmax.py (factorised einsum/cumsum)")
    return (f"{what}, fp32, fwd+bwd of a (B=1,H=1,N={n_sample},D={HL['D']}) slice of the workload (the full shape "
            f"needs PB-sized intermediates in the reference); tokens/s = (b,h,token) units/s / H={HL['H']}")


def run_reference(args, rank: int) -> None:
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    if rank != 0:
        return
    if args.workload == "model":
        return run_reference_model(args)
    threads = os.cpu_count() or 1
    n_sample = 512
    step, kind = _cpu_step_fn(n_sample, HL["D"], threads)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = n_sample / dt / HL["H"]
    _emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "fastmax p=2 causal fwd+bwd B=8 H=32 N=16384 D=64 (CPU: bounded sample)", **HL},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": "each step = " + _cpu_sample_text(kind, n_sample)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------
MODEL_SEQ = {"pythia-14m": 2048, "tiny-llama-1.1b": 8192, "Llama-2-7b-hf": 4096}     # BASELINE.json configs 1, 3, 4
LORA = dict(r=8, alpha=16, dropout=0.05, to_query=True, to_key=False, to_value=True, to_projection=False,
            to_mlp=False, to_head=False)                                               # finetune/lora.py:40-48


def _model_step_fn(model, cfg, T: int, device, opt=None):
    idx = torch.randint(0, cfg.padded_vocab_size, (1, T), generator=torch.Generator().manual_seed(0)).to(device)

    def step():
        model.zero_grad(set_to_none=True)
        logits = model(idx)
        loss = torch.nn.functional.cross_entropy(logits.view(-1, logits.size(-1)).float(), idx.view(-1))
        loss.backward()
        if opt is not None:
            opt.step()
        return loss
    return step, idx


def run_model(args, local: int) -> None:
    """BASELINE.json configs 3 / 4 (and config 1's model on the GPU): the UNMODIFIED reference model
    (lit_gpt.model.GPT from baseline/_ref, random-init weights, synthetic tokens) with its attention
    dispatch rebound to the CUDA path; one step = forward + backward (+ an AdamW step on the LoRA
    parameters for the Llama-2 fine-tune config).  Reports tokens/s and the share of the step spent in
    the FastMax kernels (their own CUDA-event times)."""
    from oracle import ref_loader as rl          # bench infrastructure: builds the reference's model
    from fastmax_experiments_b200 import _lib
    from fastmax_experiments_b200.attention import patch_causal_self_attention
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    name = args.model
    T = args.seq if args.seq != HL["N"] else MODEL_SEQ.get(name, 2048)
    lora = LORA if name.startswith("Llama-2") else None
    rl.reference_model()
    from lit_gpt.model import CausalSelfAttention
    # --fused: the one-pass q/k/v prologue + grouped-query states instead of model.py:396-451; for the
    # LoRA model the projection is lit_gpt.lora's own class, so only the stock CausalSelfAttention is rebound
    patch_causal_self_attention(CausalSelfAttention, fused=args.fused)
    model, cfg = rl.build_gpt(name, attn_alg="fastmax", block_size=T, lora=lora, device=dev, dtype=torch.bfloat16)
    opt = None
    if lora:
        from lit_gpt.lora import mark_only_lora_as_trainable
        mark_only_lora_as_trainable(model)
        opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=3e-4)
    step, idx = _model_step_fn(model, cfg, T, dev, opt)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.finish()
    ms_step = e0.elapsed_time(e1) / args.steps
    _lib.profile_enable(True)
    step()
    torch.cuda.synchronize()
    recs = _lib.profile_fetch()
    _lib.profile_enable(False)
    attn_ms = sum(ms for _, ms in recs)
    kern = {}
    for nm, ms in recs:
        a = kern.setdefault(nm, [0, 0.0])
        a[0] += 1
        a[1] += ms
    nparam = sum(p.numel() for p in model.parameters())
    _emit(json.dumps({
        "metric": "model_fwd_bwd_tokens_per_s", "value": T / (ms_step * 1e-3), "unit": UNIT, "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{name} (unmodified lit_gpt model, random init, {nparam / 1e6:.0f} M parameters"
                               f"{', LoRA r=8 on q,v, base frozen (bf16: no bitsandbytes in this image)' if lora else ''}) "
                               f"with fastmax p=2 causal attention on the CUDA path, B=1, seq {T}, forward+backward"
                               f"{' + AdamW step on the LoRA parameters' if lora else ''}",
                   "model": name, "seq": T, "n_layer": cfg.n_layer, "n_head": cfg.n_head,
                   "n_query_groups": cfg.n_query_groups, "head_size": cfg.head_size, "attn_alg": "fastmax",
                   "fused_prologue_gqa": bool(args.fused)},
        "clocks": clocks, "loss": float(loss), "gpu_launches": len(recs) * args.steps,
        "attention": {"ms_per_step": round(attn_ms, 3), "share_of_step": round(attn_ms / ms_step, 4),
                      "kernels": {n: {"launches": a[0], "ms": round(a[1], 3)} for n, a in
                                  sorted(kern.items(), key=lambda kv: -kv[1][1])}},
    }))


def run_reference_model(args) -> None:
    """BASELINE.json config 1: the unmodified reference model with its own attention path (q,k,v -> host ->
    attention_mechanisms/fastmax.py, lit_gpt/model.py:475-487) forward+backward on the host cores."""
    from oracle import ref_loader as rl
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    name = args.model
    T = args.seq if args.seq != HL["N"] else MODEL_SEQ.get(name, 2048)
    with rl.cuda_hop_neutralised():
        model, cfg = rl.build_gpt(name, attn_alg="fastmax", block_size=T)
        step, _ = _model_step_fn(model, cfg, T, "cpu")
        warm = min(args.warmup, 1 if T <= 512 else 0)     # a step takes about a minute at seq 2048
        for _ in range(warm):
            step()
        steps = max(1, min(args.steps, 3 if T <= 512 else 1))
        t0 = time.perf_counter()
        for _ in range(steps):
            loss = step()
        dt = (time.perf_counter() - t0) / steps
    value = T / dt
    _emit(json.dumps({
        "impl": "reference", "metric": "model_fwd_bwd_tokens_per_s", "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name} with the reference's own fastmax p=2 causal attention, forward+backward on the "
                               f"host, B=1, seq {T}, synthetic tokens (BASELINE.json configs[0])",
                   "model": name, "seq": T, "attn_alg": "fastmax"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": f"{steps} whole step(s) of the unmodified lit_gpt model + attention_mechanisms/fastmax.py"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "loss": float(loss), "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------
def _sp_parity(q, k, v, do, o, g, dq, dk, dv, a0: int, world: int, rank: int, c: float) -> dict:
    """In-run parity check of the sequence-parallel path over the real collective: for head (b=0, h=0) every rank
    gathers the whole sequence of q, k, v, do and of the computed o, g, and checks sampled LOCAL rows of o, g, dq
    and sampled local columns of dk, dv against the dense formula evaluated in fp64 with plain torch ops on its
    GPU (the recipe of tests/test_gpu_parity.py::_sampled_rows_and_columns_check).  Returns the largest per-row
    relative error of each tensor over all ranks."""
    import torch.distributed as dist
    n_loc = q.shape[2]

    def gather(t):      # (n_loc, ...) of head (0, 0) -> (N, ...)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        return torch.cat(parts, 0)

    Q, K, V, DO, O = (gather(t[0, 0]).double() for t in (q, k, v, do, o))
    G = gather(g[0, 0]).double()
    N = Q.shape[0]
    idx = torch.arange(N, device=q.device)
    rows = torch.tensor(sorted({a0 + 1 if a0 == 0 else a0, a0 + 511, a0 + 512, a0 + n_loc // 2 + 77, a0 + n_loc - 1}),
                        device=q.device)
    cols = torch.tensor(sorted({a0, a0 + 511, a0 + 512, a0 + n_loc // 3 + 5, a0 + n_loc - 2}), device=q.device)

    def worst(x, ref):
        err = (x.double() - ref).abs().reshape(ref.shape[0], -1).max(-1).values
        mag = ref.abs().reshape(ref.shape[0], -1).max(-1).values
        return float((err / mag).max())

    s = Q[rows] @ K.T / c
    m = (idx[None, :] <= rows[:, None]).double()
    w = (1 + s + 0.5 * s * s) * m
    gr = w.sum(-1)
    o_ref = (w @ V) / gr[:, None]
    u = DO[rows] / gr[:, None]
    dlt = (DO[rows] * o_ref).sum(-1) / gr
    e = (1 + s) * (u @ V.T - dlt[:, None]) * m
    res = {"o": worst(O[rows], o_ref), "g": worst(G[rows, None], gr[:, None]),
           "dq": worst(dq[0, 0][rows - a0], e @ K / c)}
    Uall = DO / G[:, None]
    Dall = (DO * O).sum(-1) / G
    st = K[cols] @ Q.T / c
    mt = (idx[None, :] >= cols[:, None]).double()
    wt = (1 + st + 0.5 * st * st) * mt
    et = (1 + st) * (V[cols] @ Uall.T - Dall[None, :]) * mt
    res["dv"] = worst(dv[0, 0][cols - a0], wt @ Uall)
    res["dk"] = worst(dk[0, 0][cols - a0], et @ Q / c)
    t = torch.tensor([res[n] for n in ("o", "g", "dq", "dk", "dv")], device=q.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {n: float(x) for n, x in zip(("o", "g", "dq", "dk", "dv"), t.tolist())}


def measure_sequence_parallel(args, rank: int, world: int, dev, N: int, steps: int, warmup: int) -> dict:
    """BASELINE.json config 5: p=2 causal fwd+bwd of ONE long sequence set (B=1, H=32, D=64, N tokens) split by
    token ranges over the ranks; one state all-gather per scan direction over NCCL.  Timed like the headline
    (barrier + synchronize on both sides, CUDA events, max over ranks), then checked in the same run against the
    dense formula (`_sp_parity`).  The process group must exist."""
    import torch.distributed as dist
    from fastmax_experiments_b200 import _lib
    from fastmax_experiments_b200.sequence_parallel import fastmax_sp, partition

    B, H, D, p = 1, HL["H"], HL["D"], HL["p"]
    a0, a1 = partition(N, world, rank)
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    q, k, v = (torch.randn(B, H, a1 - a0, D, device=dev, generator=gen).to(torch.bfloat16).requires_grad_(True)
               for _ in range(3))
    do = torch.randn(B, H, a1 - a0, D, device=dev, generator=gen).to(torch.bfloat16)
    keep = {}

    def step():
        for t in (q, k, v):
            t.grad = None
        o = fastmax_sp(q, k, v, p=p, token_offset=a0)
        o.backward(do)
        keep["o"] = o

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    sampler = ClockSampler(dev.index)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.finish()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    # the exchange alone: all-gather of one state record set per rank (what each scan direction moves)
    lib = _lib.load()
    rec = torch.empty(lib.fastmax_sp_state_bytes(B, H), dtype=torch.uint8, device=dev)
    gathered = [torch.empty_like(rec) for _ in range(world)]
    ag_ms = None
    if world > 1:
        for _ in range(2):
            dist.all_gather(gathered, rec)
        barrier()
        e0.record()
        for _ in range(5):
            dist.all_gather(gathered, rec)
        e1.record()
        barrier()
        ag_ms = e0.elapsed_time(e1) / 5
    # per-kernel times of one more step
    _lib.profile_enable(True)
    step()
    torch.cuda.synchronize()
    recs = _lib.profile_fetch()
    _lib.profile_enable(False)
    kern = {}
    for name, ms in recs:
        a = kern.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    # parity over the real collective
    parity = None
    if world > 1 and (a1 - a0) >= 1024 and N % world == 0:
        with torch.no_grad():
            # g is not returned by the autograd op: recompute this rank's forward through the raw entry points
            from fastmax_experiments_b200.sequence_parallel import (exchange_states, sp_fwd, sp_local_state, sp_reduce)
            red = lambda gathered, w, r, d: sp_reduce(gathered, w, r, d, B, H)
            init = exchange_states(sp_local_state(k.detach(), v.detach(), p), None, 0, red)
            o2, g2 = sp_fwd(q.detach(), k.detach(), v.detach(), init, a0, 1.0 / (8 * D ** 0.5), p)
            parity = _sp_parity(q.detach(), k.detach(), v.detach(), do, o2, g2, q.grad, k.grad, v.grad, a0, world, rank,
                                8 * D ** 0.5)
            parity["tolerance"] = 2e-2
            parity["ok"] = all(x <= 2e-2 for n, x in parity.items() if n in ("o", "g", "dq", "dk", "dv"))
            parity["how"] = ("head (b=0,h=0): sampled local rows of o, g, dq and columns of dk, dv on every rank vs the "
                             "dense formula in fp64 (torch ops on the gathered sequence); worst per-row relative error, "
                             "max over ranks")
    f_fwd, f_bwd = flops_alg(D, p)
    peaks = load_peaks()
    ach = B * H * N * (f_fwd + f_bwd) / (ms_step * 1e-3) / 1e12
    one = None      # best single-GPU rate of the same kernels on the same problem sizes (committed bench lines)
    try:
        one = json.load(open(os.path.join(ROOT, "profiles", "bench", "r2_sp_single_gpu.json")))
    except Exception:
        pass
    value = B * N / (ms_step * 1e-3)
    eff = None
    if one and world > 1:
        best1 = max(one.get("tokens_per_s", {}).values())
        eff = {"vs_best_single_gpu_rate": round(value / (world * best1), 4), "single_gpu_tokens_per_s": one["tokens_per_s"],
               "note": "strong scaling: value / (n_gpus * best single-GPU rate over N in {128k, 1M})"}
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"sequence-parallel fastmax p=2 causal fwd+bwd B={B} H={H} N={N} D={D} bf16, "
                               f"tokens split over {world} rank(s), state all-gather over NCCL",
                   "B": B, "H": H, "N": N, "D": D, "p": p, "causal": True, "local_tokens": a1 - a0,
                   "l2": "inputs larger than L2"},
        "clocks": clocks, "gpu_launches": len(recs) * steps,
        "roofline": {"bound": "tensor", "achieved": round(ach, 3), "peak": peaks["tflops"] * world,
                     "unit": "TFLOP/s", "frac": round(ach / (peaks["tflops"] * world), 5), "traffic": None,
                     "scope": "whole fwd+bwd step over all ranks (algorithmic flops; the per-rank state "
                              "reduce passes are redundant work and not counted)"},
        "parity": parity, "efficiency": eff,
        "exchange": {"all_gather_ms": ag_ms, "bytes_per_rank": int(rec.numel()), "per_step": 2,
                     "redundant_flop_fraction": round(2 * (2 * D * D * (D + 1)) * (20 + 2) / 34 / (f_fwd + f_bwd), 4),
                     "note": "one extra packed state-build pass over the local tokens per scan direction"},
        "kernels": {n: {"launches": a[0], "ms": round(a[1], 4)} for n, a in sorted(kern.items(), key=lambda kv: -kv[1][1])},
    }


def run_sequence_parallel(args, rank: int, world: int, local: int) -> None:
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    N = args.seq if args.seq != HL["N"] else 131072
    out = measure_sequence_parallel(args, rank, world, dev, N, args.steps, args.warmup)
    if rank == 0:
        _emit(json.dumps(out))
    dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", default="auto", choices=["auto", "exact", "moments"])
    ap.add_argument("--seq", type=int, default=HL["N"], help="override N (diagnostics only)")
    ap.add_argument("--batch", type=int, default=HL["B"], help="override B (diagnostics only)")
    ap.add_argument("--fused", action="store_true",
                    help="--workload model: rebind CausalSelfAttention.forward to the fused prologue + grouped-query path")
    ap.add_argument("--workload", default="headline", choices=["headline", "sp", "model"],
                    help="headline: (b,h)-sharded op at B=8,H=32,N=16k (default); sp: sequence-parallel "
                         "long-context scan, B=1,H=32, --seq tokens split over the ranks (BASELINE config 5); "
                         "model: the unmodified reference model with the CUDA attention path (configs 1, 3, 4)")
    ap.add_argument("--model", default="pythia-14m", choices=sorted(MODEL_SEQ),
                    help="--workload model: which lit_gpt config (default sequence lengths: BASELINE.json configs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-p1", action="store_true", help="skip the first-order (p=1) side measurement")
    ap.add_argument("--no-sp", action="store_true", help="N > 1: skip the sequence-parallel (config 5) leg")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3   # timing rule: at least 3 warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload == "sp":
        run_sequence_parallel(args, rank, world, local)
        return
    if args.workload == "model":
        if rank == 0:
            run_model(args, local)
        return

    import torch.distributed as dist
    from fastmax_experiments_b200 import _lib
    from fastmax_experiments_b200.fastmax import (SAVE_STATES_MAX_BYTES, fastmax_bwd_raw, fastmax_fwd_raw,
                                                  saved_state_bytes)
    from fastmax_experiments_b200.host import fastmax_fwd_bwd_host

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    B, H, N, D, p = args.batch, HL["H"], args.seq, HL["D"], HL["p"]
    algo = {"auto": _lib.ALGO_AUTO, "exact": _lib.ALGO_EXACT, "moments": _lib.ALGO_MOMENTS}[args.algo]
    inv_c = 1.0 / (8 * D ** 0.5)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    q, k, v, do = (torch.randn(B, H, N, D, device=dev, generator=gen).to(torch.bfloat16) for _ in range(4))
    pl = _lib.plan(B, H, N, D, p, True, _lib.DTYPE_BF16, algo)

    launches = 0
    # same policy as the autograd op (fastmax.py): keep the forward moment states for backward when
    # they fit the budget, else rebuild them in backward
    keep_bytes = saved_state_bytes(q, True, p, algo)
    keep = 0 < keep_bytes <= SAVE_STATES_MAX_BYTES

    def step():
        nonlocal launches
        saved = torch.empty(keep_bytes, dtype=torch.uint8, device=dev) if keep else None
        o, g = fastmax_fwd_raw(q, k, v, True, inv_c, p, algo, saved=saved)
        launches += lib.fastmax_last_launch_count()
        out = fastmax_bwd_raw(q, k, v, o, g, do, True, inv_c, p, algo, saved=saved)
        launches += lib.fastmax_last_launch_count()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches = 0
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.finish()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    tokens = B * N * world
    value = tokens / (ms_step * 1e-3)

    # ---- per-kernel breakdown with CUDA events on the launching stream (separate pass) ------
    _lib.profile_enable(True)
    step()
    torch.cuda.synchronize()
    recs = _lib.profile_fetch()
    _lib.profile_enable(False)
    kern = {}
    for name, ms in recs:
        a = kern.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    prof_total = sum(a[1] for a in kern.values()) or 1.0
    kernels = {n: {"launches": a[0], "ms": round(a[1], 4), "share": round(a[1] / prof_total, 4)}
               for n, a in sorted(kern.items(), key=lambda kv: -kv[1][1])}

    peaks = load_peaks()
    U = B * H * N
    f_fwd, f_bwd = flops_alg(D, p)
    exec_flops = (pl.fwd_tensor_flops + pl.bwd_tensor_flops - (pl.rebuild_tensor_flops if keep else 0.0)) or None
    counted = U * (f_fwd + f_bwd)
    if exec_flops:
        counted = min(counted, exec_flops)       # never overstate utilisation
    peak_tf = peaks["tflops_sustained"] if ms_step > 100 else peaks["tflops"]
    achieved_tf = counted / (ms_step * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):      # dram__bytes_read.sum + dram__bytes_write.sum of one step (ncu --set full), same config only
        tj = json.load(open(tpath))
        if tj.get("config") == {"B": B, "H": H, "N": N, "kept_states": bool(keep)} and pl.algo == 2:
            traffic = tj.get("dram_bytes_per_step")
    # the dominant kernel on its own: algorithmic flops of one launch (the moment contraction of every unit
    # that has a boundary state: 2*D*D*(D+1) per unit) over its measured launch time
    dominant = None
    if pl.algo == 2 and p == 2:
        seg = pl.seg_len
        units_with_state = B * H * max(0, N - seg)
        mom_flops = 2.0 * D * D * (D + 1) * units_with_state
        for nm, a in kern.items():
            if nm.startswith(("mom_apply", "mom_build")):
                tf = mom_flops / (a[1] / a[0] * 1e-3) / 1e12
                kernels[nm]["tflops"] = round(tf, 1)
                kernels[nm]["frac"] = round(tf / peaks["tflops"], 4)
        ap = [(nm, a) for nm, a in kern.items() if nm.startswith("mom_apply")]
        if ap:
            ms_avg = sum(a[1] for _, a in ap) / sum(a[0] for _, a in ap)
            tf = mom_flops / (ms_avg * 1e-3) / 1e12
            dominant = {"kernel": "apply_kernel (mom_apply_fwd/dq/dv/dk: %d launches, %.1f %% of the step)"
                                  % (sum(a[0] for _, a in ap), 100 * sum(a[1] for _, a in ap) / prof_total),
                        "ms_per_launch": round(ms_avg, 4), "flops_alg_per_launch": mom_flops,
                        "achieved": round(tf, 1), "peak": peaks["tflops"], "unit": "TFLOP/s",
                        "frac": round(tf / peaks["tflops"], 4)}
    roofline = {
        "bound": "tensor", "achieved": round(achieved_tf, 3), "peak": peak_tf, "unit": "TFLOP/s",
        "frac": round(achieved_tf / peak_tf, 5), "traffic": traffic, "dominant_kernel": dominant,
        "scope": "whole fwd+bwd step (all kernels: the stricter number); `dominant_kernel` = the apply kernel alone; "
                 "per-kernel times and rates in `kernels`",
        "peak_source": f"MEASURED_PEAKS.json ({peaks['source']}; "
                       f"{'sustained' if ms_step > 100 else 'burst'} bf16)",
        "flops_counted": counted, "flops_alg_per_unit": f_fwd + f_bwd,
    }

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"fastmax p=2 causal fwd+bwd B={B} H={H} N={N} D={D} bf16 (per GPU)",
                   "B": B, "H": H, "N": N, "D": D, "p": p, "causal": True,
                   "algo": {1: "exact", 2: "moments"}[pl.algo], "seg_len": pl.seg_len,
                   "kept_forward_states_bytes": keep_bytes if keep else 0,
                   "l2": "inputs larger than L2 (q,k,v,do 512 MiB each)",
                   "sharding": "(b,h) units per rank, no collective"},
        "clocks": clocks, "gpu_launches": launches, "roofline": roofline, "kernels": kernels,
    }

    # ---- first order (p = 1) on the same inputs: the HBM-bound side of the op (1544 B per unit) ----
    if not args.no_p1:
        fo = {}
        for name, causal in (("causal", True), ("full", False)):
            def step1():
                o, g = fastmax_fwd_raw(q, k, v, causal, inv_c, 1, algo)
                return fastmax_bwd_raw(q, k, v, o, g, do, causal, inv_c, 1, algo)
            for _ in range(3):
                step1()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n1 = max(3, args.steps)
            a0.record()
            for _ in range(n1):
                step1()
            a1.record()
            torch.cuda.synchronize()
            ms1 = a0.elapsed_time(a1) / n1
            gbs = U * (12 * D * 2 + 8) / (ms1 * 1e-3) / 1e9
            _lib.profile_enable(True)
            step1()
            torch.cuda.synchronize()
            k1 = {}
            for nm, ms in _lib.profile_fetch():
                k1[nm] = round(k1.get(nm, 0.0) + ms, 4)
            _lib.profile_enable(False)
            fo[name] = {"ms_per_step": round(ms1, 4), "tokens_per_s": round(B * N / (ms1 * 1e-3)),
                        "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                     "frac": round(gbs / peaks["hbm_gbs"], 4)},
                        "kernels_ms": k1}
        out["first_order"] = {"workload": f"fastmax p=1 fwd+bwd B={B} H={H} N={N} D={D} bf16 on this rank's inputs, "
                                          "forward state not kept; algorithmic bytes 1544 B per (b,h,token)", **fo}

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region ---------------
    if not args.no_e2e:
        # 4 GiB of pinned host memory per rank; if any rank cannot get it, every rank skips the e2e leg
        # together (the collectives below must stay matched) and the line says so
        try:
            hq, hk, hv, hdo = (t.cpu().pin_memory() for t in (q, k, v, do))
            outs = tuple(torch.empty_like(hq).pin_memory() for _ in range(4))
            have = 1
        except RuntimeError:
            have = 0
        if world > 1:
            t = torch.tensor([have], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            have = int(t.item())
        n_e2e = max(1, min(args.steps, 3))
    if not args.no_e2e and not have:
        out["e2e"] = {"value": None, "unit": UNIT, "error": "pinned host allocation failed on at least one rank"}
    elif not args.no_e2e:
        for _ in range(2):
            fastmax_fwd_bwd_host(hq, hk, hv, hdo, out=outs, p=p, device=dev)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            fastmax_fwd_bwd_host(hq, hk, hv, hdo, out=outs, p=p, device=dev)
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        nbytes = hq.numel() * hq.element_size()
        out["e2e"] = {"value": tokens / dt, "unit": UNIT, "h2d_bytes_per_step": 4 * nbytes,
                      "d2h_bytes_per_step": 4 * nbytes, "ms_per_step": dt * 1e3, "steps": n_e2e,
                      "api": "fastmax_experiments_b200.host.fastmax_fwd_bwd_host (pinned host q,k,v,do -> "
                             "pinned host o,dq,dk,dv; copies pipelined over the batch)"}

    # ---- CPU baseline beside it (rank 0, N=1 only) ------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        rate, secs, kind = cpu_rate(512, D, repeats=3, threads=threads)
        out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                               "sample": f"best of 3 ({secs:.2f} s each): " + _cpu_sample_text(kind, 512)}

    # ---- N > 1: the sequence-parallel split of north_star (BASELINE.json config 5, N = 1M) in the same run ------
    if world > 1 and not args.no_sp:
        del q, k, v, do
        torch.cuda.empty_cache()
        sp = measure_sequence_parallel(args, rank, world, dev, 1 << 20, steps=max(2, min(args.steps, 3)), warmup=3)
        out["sp"] = sp

    if rank == 0:
        _emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _emit(line: str) -> None:
    """The one JSON line goes to the process's ORIGINAL stdout (see _quiet_stdout)."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


_REAL_STDOUT = 1


def _quiet_stdout() -> None:
    """Libraries chat on fd 1 (NCCL prints its version banner there): point fd 1 at stderr for the
    whole run and keep the original for the result line, so stdout carries exactly one JSON line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


if __name__ == "__main__":
    _quiet_stdout()
    main()

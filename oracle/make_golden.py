"""Mint golden vectors from the UNMODIFIED reference (test infrastructure).

Run in the build container only (needs /root/reference; the GPU box has none):

    python oracle/make_golden.py            # writes tests/golden/*.npz

The reference (``attention_mechanisms/fastmax.py``) ships no golden vectors for
this path, so these fixtures *are* the parity pin: each
holds seeded inputs and the outputs ``o, g, dq, dk, dv`` the reference's own
``fastattention_einops`` forward/backward produced for them, evaluated in
float64 so that the fixture is limited by the algorithm, not by the reference's
working precision.  Inputs are rounded to bf16-representable values (stored as
their raw upper 16 bits) so one fixture serves both the fp32 and the bf16
kernels: the protocol of SURVEY.md section 8(c) is "bf16 kernels vs the reference
evaluated in higher precision on the bf16-rounded inputs".

A second, fp32-working-precision run of the reference is stored for one case
(``ref32_*`` keys) to document the reference's own fp32 rounding level.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF_ROOT = os.environ.get("FASTMAX_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# (name, B, H, N, D, p, mask, normalize_term, tensors_normalized, seed)
CASES = [
    ("p2_causal_d32",      1, 2,  96,  32, 2, True,  8, False, 1),
    ("p2_causal_d64",      2, 2, 130,  64, 2, True,  8, False, 2),
    ("p2_causal_d64_long", 1, 1, 600,  64, 2, True,  8, False, 3),
    ("p2_causal_d128",     1, 1,  70, 128, 2, True,  8, False, 4),
    ("p2_full_d32",        1, 2,  96,  32, 2, False, 8, False, 5),
    ("p2_full_d64",        2, 2, 130,  64, 2, False, 8, False, 6),
    ("p1_causal_d64",      2, 2, 130,  64, 1, True,  8, False, 7),
    ("p1_full_d64",        2, 2, 130,  64, 1, False, 8, False, 8),
    ("p1_causal_d128",     1, 1, 257, 128, 1, True,  8, False, 9),
    ("p2_causal_d64_nt1",  1, 2,  64,  64, 2, True,  1, False, 10),
    ("p2_causal_d32_tn",   1, 2,  64,  32, 2, True,  8, True,  11),
    ("p1_causal_d32_odd",  2, 3,  33,  32, 1, True,  8, False, 12),
    # long enough for the chunked scan of the tcgen05 path (D = 32 runs zero-padded on the D = 64 kernels):
    # several 512-token segments, i.e. the build / apply moment kernels, pinned directly to the reference
    ("p2_causal_d32_multiseg", 1, 1, 2600, 32, 2, True, 8, False, 13),
    ("p1_causal_d32_multiseg", 1, 2, 1100, 32, 1, True, 8, False, 14),
]


# inference shapes of the dispatch (model.py:478-481: input_pos given -> mask=False over the KV cache):
# N_q != N_k, forward only.  (name, B, H, Nq, Nk, D, p, seed)
RECT_CASES = [
    ("p2_full_decode_d64",  1, 2,   1, 300, 64, 2, 21),
    ("p2_full_prefill_d64", 2, 2,  40, 300, 64, 2, 22),
    ("p1_full_prefill_d32", 1, 2,  17,  96, 32, 1, 23),
]


# linearmax (attention_mechanisms/fastmax_hack.py, dispatched at model.py:444-445,460-472): p=1, c=1 after a
# mean-centre / global-max-norm prologue on q,k; autograd through torch.  (name, B, H, N, D, mask, seed)
LINEARMAX_CASES = [
    ("linearmax_causal_d64", 2, 2, 130, 64, True, 31),
    ("linearmax_full_d64",   1, 2,  96, 64, False, 32),
    ("linearmax_causal_d32", 1, 3,  70, 32, True, 33),
]

# linearmax over a KV cache (input_pos given -> mask=False, N_q != N_k): the count term of fastmax_hack.py:21 is
# N_k, unlike fastmax.py's N_q.  Forward only.  (name, B, H, Nq, Nk, D, seed)
LINEARMAX_RECT_CASES = [
    ("linearmax_full_prefill_d64", 1, 2, 17, 96, 64, 34),
    ("linearmax_full_decode_d32",  2, 2,  1, 40, 32, 35),
]

# model level: the unmodified lit_gpt GPT, attn_alg="fastmax", on the host -- including
# the dispatch's missing head transpose (quirk Q1), which lives in CausalSelfAttention.forward, not in the op
MODEL_CASES = [
    ("model_pythia14m_seq256", "pythia-14m", 256, 0),
]


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def bf16_bits(x: torch.Tensor) -> np.ndarray:
    return x.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


def main() -> None:
    sys.path.insert(0, REF_ROOT)
    from attention_mechanisms.fastmax import fastmax as ref_fastmax  # the unmodified reference

    os.makedirs(OUT_DIR, exist_ok=True)
    for name, B, H, N, D, p, mask, nt, tn, seed in CASES:
        if NAMES and name not in NAMES:
            continue
        gen = torch.Generator().manual_seed(seed)
        scale = 0.25 if tn else 1.0   # tensors_normalized => c = 1: keep |s| small like normalised q,k
        q = bf16_round(torch.randn(B, H, N, D, generator=gen) * scale)
        k = bf16_round(torch.randn(B, H, N, D, generator=gen) * scale)
        v = bf16_round(torch.randn(B, H, N, D, generator=gen))
        do = bf16_round(torch.randn(B, H, N, D, generator=gen))

        def run(dtype):
            qq, kk, vv = (t.to(dtype).clone().requires_grad_(True) for t in (q, k, v))
            saved = {}
            # the reference keeps g only inside ctx; fish it out of the saved tensors
            with torch.autograd.graph.saved_tensors_hooks(
                    lambda t: (saved.setdefault("list", []).append(t), t)[1], lambda t: t):
                o = ref_fastmax(qq, kk, vv, mask=mask, normalize_term=nt,
                                tensors_normalized=tn, p=p)
            g = saved["list"][4]          # save_for_backward(q, k, v, o, g)  fastmax.py:106
            o.backward(do.to(dtype))
            return o.detach(), g.detach(), qq.grad, kk.grad, vv.grad

        o, g, dq, dk, dv = run(torch.float64)
        # small cases keep the float64 outputs (pins the oracle to ~1e-13); larger ones
        # are stored as float32 (pin ~1e-7) to keep the committed fixtures small
        st = (lambda t: t.double().numpy()) if q.numel() <= 10000 else (lambda t: t.float().numpy())
        out = dict(
            B=B, H=H, N=N, D=D, p=p, mask=mask, normalize_term=nt, tensors_normalized=tn,
            q_bf16=bf16_bits(q), k_bf16=bf16_bits(k), v_bf16=bf16_bits(v), do_bf16=bf16_bits(do),
            o=st(o), g=st(g), dq=st(dq), dk=st(dk), dv=st(dv),
        )
        if name == "p2_causal_d64":
            o32, g32, dq32, dk32, dv32 = run(torch.float32)
            out.update(ref32_o=o32.numpy(), ref32_dq=dq32.numpy(), ref32_dk=dk32.numpy(),
                       ref32_dv=dv32.numpy())
        path = os.path.join(OUT_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: wrote {os.path.getsize(path)/1e6:.2f} MB  |o|max={o.abs().max():.4f}")


def main_rect() -> None:
    sys.path.insert(0, REF_ROOT)
    from attention_mechanisms.fastmax import fastmax as ref_fastmax  # the unmodified reference

    for name, B, H, Nq, Nk, D, p, seed in RECT_CASES:
        if NAMES and name not in NAMES:
            continue
        gen = torch.Generator().manual_seed(seed)
        q = bf16_round(torch.randn(B, H, Nq, D, generator=gen))
        k = bf16_round(torch.randn(B, H, Nk, D, generator=gen))
        v = bf16_round(torch.randn(B, H, Nk, D, generator=gen))
        with torch.no_grad():
            o = ref_fastmax(q.double(), k.double(), v.double(), mask=False, p=p)
        out = dict(B=B, H=H, Nq=Nq, Nk=Nk, D=D, p=p, q_bf16=bf16_bits(q), k_bf16=bf16_bits(k), v_bf16=bf16_bits(v),
                   o=o.double().numpy())
        path = os.path.join(OUT_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: wrote {os.path.getsize(path)/1e6:.2f} MB  |o|max={o.abs().max():.4f}")


def main_linearmax() -> None:
    sys.path.insert(0, REF_ROOT)
    from attention_mechanisms.fastmax_hack import fastmax_hack as ref_linearmax  # the unmodified reference

    for name, B, H, N, D, mask, seed in LINEARMAX_CASES:
        if NAMES and name not in NAMES:
            continue
        gen = torch.Generator().manual_seed(seed)
        q, k, v, do = (bf16_round(torch.randn(B, H, N, D, generator=gen)) for _ in range(4))
        qq, kk, vv = (t.double().clone().requires_grad_(True) for t in (q, k, v))
        o = ref_linearmax(qq, kk, vv, p=1, mask=mask)
        o.backward(do.double())
        out = dict(B=B, H=H, N=N, D=D, mask=mask, q_bf16=bf16_bits(q), k_bf16=bf16_bits(k), v_bf16=bf16_bits(v),
                   do_bf16=bf16_bits(do), o=o.detach().numpy(), dq=qq.grad.numpy(), dk=kk.grad.numpy(),
                   dv=vv.grad.numpy())
        path = os.path.join(OUT_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: wrote {os.path.getsize(path)/1e6:.2f} MB  |o|max={o.abs().max():.4f}")


def main_linearmax_rect() -> None:
    sys.path.insert(0, REF_ROOT)
    from attention_mechanisms.fastmax_hack import fastmax_hack as ref_linearmax  # the unmodified reference

    for name, B, H, Nq, Nk, D, seed in LINEARMAX_RECT_CASES:
        if NAMES and name not in NAMES:
            continue
        gen = torch.Generator().manual_seed(seed)
        q = bf16_round(torch.randn(B, H, Nq, D, generator=gen))
        k = bf16_round(torch.randn(B, H, Nk, D, generator=gen))
        v = bf16_round(torch.randn(B, H, Nk, D, generator=gen))
        with torch.no_grad():
            o = ref_linearmax(q.double(), k.double(), v.double(), p=1, mask=False)
        out = dict(B=B, H=H, Nq=Nq, Nk=Nk, D=D, p=1, q_bf16=bf16_bits(q), k_bf16=bf16_bits(k), v_bf16=bf16_bits(v),
                   o=o.double().numpy())
        path = os.path.join(OUT_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: wrote {os.path.getsize(path)/1e6:.2f} MB  |o|max={o.abs().max():.4f}")


def main_model() -> None:
    """Logits and loss of the unmodified reference model (fp32, host).  The full logits are 51 MB, so the
    fixture keeps the loss, every position's log-sum-exp and a strided sample of the logits; inputs and
    weights are re-created from the seed (oracle/ref_loader.py::build_gpt)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import ref_loader as rl

    for name, model_name, T, seed in MODEL_CASES:
        if NAMES and name not in NAMES:
            continue
        with rl.cuda_hop_neutralised():
            model, cfg = rl.build_gpt(model_name, attn_alg="fastmax", block_size=T, seed=seed)
            idx = torch.randint(0, cfg.padded_vocab_size, (1, T), generator=torch.Generator().manual_seed(seed + 1))
            logits = model(idx)
            loss = torch.nn.functional.cross_entropy(logits.view(-1, logits.size(-1)), idx.view(-1))
            loss.backward()
        gnorm = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in model.parameters() if p.grad is not None))
        out = dict(model=model_name, T=T, seed=seed, idx=idx.numpy(), loss=float(loss),
                   lse=torch.logsumexp(logits.detach().double(), -1).numpy(),
                   logits_sample=logits.detach()[:, ::8, ::64].numpy(), grad_norm=float(gnorm),
                   wte_grad_sample=model.transformer.wte.weight.grad[idx[0, :8]].numpy())
        path = os.path.join(OUT_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: wrote {os.path.getsize(path)/1e6:.2f} MB  loss={float(loss):.6f} grad_norm={float(gnorm):.6f}")


# grouped-query attention module of the reference (tiny-llama-1.1b preset, lit_gpt/config.py:1394-1410, shrunk:
# 1 layer, 4 heads of 64 in 2 query groups), fp32 and bf16: q, k, v as the unmodified CausalSelfAttention.forward
# hands them to its attention method (after view / permute / split / expand / reshape / RoPE, model.py:396-425), its
# output, and the gradient it sends back into the qkv projection.  (name, T, seed, rotary_percentage)
GQA_CASES = [
    ("gqa_prologue_t256", 256, 41, 1.0),
    ("gqa_prologue_t256_rot25", 256, 42, 0.25),
]


def main_gqa() -> None:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import ref_loader as rl

    for name, T, seed, rot in GQA_CASES:
        if NAMES and name not in NAMES:
            continue
        out = dict(T=T, seed=seed, rot=rot)
        for dtype, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
            with rl.cuda_hop_neutralised():
                model, cfg = rl.build_gpt("tiny-llama-1.1b", attn_alg="fastmax", block_size=T, seed=seed, n_layer=1,
                                          n_embd=256, n_head=4, n_query_groups=2, intermediate_size=256,
                                          rotary_percentage=rot)
                model = model.to(dtype)
                attn = model.transformer.h[0].attn
                cap = {}
                inner = attn.fastmax                      # the reference's own method; only observed, not replaced

                def spy(q, k, v, input_pos, _inner=inner, _cap=cap):
                    _cap["q"], _cap["k"], _cap["v"] = q, k, v
                    q.retain_grad(), k.retain_grad(), v.retain_grad()
                    return _inner(q, k, v, input_pos)
                attn.fastmax = spy

                def keep_qkv(m, i, o, _cap=cap):
                    _cap["qkv"] = o
                    o.retain_grad()
                hook = attn.attn.register_forward_hook(keep_qkv)
                g = torch.Generator().manual_seed(seed + 1)
                x = torch.randn(1, T, cfg.n_embd, generator=g).to(dtype)
                cos, sin = model.cos[:T].to(dtype), model.sin[:T].to(dtype)
                y = attn(x, cos, sin)
                dy = torch.randn(y.shape, generator=g).to(dtype)
                y.backward(dy)
                hook.remove()
            q_per_kv = cfg.n_head // cfg.n_query_groups
            k, v = cap["k"], cap["v"]
            assert torch.equal(k[:, ::q_per_kv].repeat_interleave(q_per_kv, 1), k)      # the expand is a plain copy
            enc = (lambda t: t.detach().numpy()) if dtype == torch.float32 else (lambda t: bf16_bits(t.detach().float()))
            B_, Hh, T_, hs_ = k.shape
            gsum = lambda t: t.grad.double().view(B_, Hh // q_per_kv, q_per_kv, T_, hs_).sum(2).float().numpy()
            # (x and dy are re-created from the seed by the tests: torch.Generator().manual_seed(seed + 1), randn x then dy)
            out.update({f"{tag}_qkv": enc(cap["qkv"]),
                        f"{tag}_q": enc(cap["q"]), f"{tag}_k": enc(k[:, ::q_per_kv]), f"{tag}_v": enc(v[:, ::q_per_kv]),
                        f"{tag}_y": enc(y), f"{tag}_dqkv": enc(cap["qkv"].grad),
                        # upstream gradients of the prologue: dq as is, dk / dv summed over the group in fp64 (the
                        # reference accumulates them in the working dtype inside autograd's expand backward)
                        f"{tag}_dq": enc(cap["q"].grad),
                        f"{tag}_dk": gsum(k), f"{tag}_dv": gsum(v),
                        f"{tag}_cos": enc(cos), f"{tag}_sin": enc(sin)})
            out.update(n_embd=cfg.n_embd, n_head=cfg.n_head, n_query_groups=cfg.n_query_groups, head_size=cfg.head_size,
                       rope_n_elem=cfg.rope_n_elem)
        path = os.path.join(OUT_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: wrote {os.path.getsize(path)/1e6:.2f} MB")


NAMES: set = set()

if __name__ == "__main__":
    only = [a for a in sys.argv[1:] if a.startswith("--only=")]
    which = only[0].split("=")[1] if only else "all"
    names = [a for a in sys.argv[1:] if a.startswith("--names=")]
    if names:
        NAMES = set(names[0].split("=")[1].split(","))
    if which in ("all", "linearmax_rect"):
        main_linearmax_rect()
    if which in ("all", "model"):
        main_model()
    if which in ("all", "gqa"):
        main_gqa()
    if which in ("all", "square"):
        main()
    if which in ("all", "rect") or "--rect-only" in sys.argv:
        main_rect()
    if which in ("all", "linearmax"):
        main_linearmax()

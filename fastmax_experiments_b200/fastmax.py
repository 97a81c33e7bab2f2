"""FastMax attention op -- drop-in for the reference's ``attention_mechanisms/fastmax.py``.

``fastmax(q, k, v, mask=True, normalize_term=8, tensors_normalized=False, p=1,
dropout_rate=0.0, create_attn=False)`` has the reference's signature and defaults
(fastmax.py:7-17) and the same autograd contract (``fastattention_einops.apply`` with nine
positional arguments, fastmax.py:24-26; gradients for q, k, v only, fastmax.py:182), but runs
hand-written sm_100a kernels through the C ABI of ``include/fastmax_b200.h``.

Behavioural notes:
* ``mask`` is tested with ``is False`` exactly like fastmax.py:85,132 -- only the literal
  ``False`` selects the non-causal branch.
* ``c = 1`` if ``tensors_normalized is True`` else ``normalize_term * sqrt(D)`` (fastmax.py:78-82).
* ``dropout_rate`` is accepted and ignored, as in the reference (fastmax.py:185-216 never use it).
* ``p`` outside {1,2} raises ``ValueError`` -- in forward (the reference silently computes
  p=1 forward and raises only in backward, fastmax.py:210,428: quirk Q3).
* ``create_attn=True`` raises ``NotImplementedError`` (the reference's backward for it is
  broken: quirk Q5).
* Inputs must be CUDA tensors: there is no CPU or PyTorch fallback.  Output dtype = input
  dtype (fp32 or bf16); accumulation is always fp32; ``g`` is saved in fp32.
* ``mask=False`` accepts N_q != N_k (prefill / decode over the KV cache, model.py:478-487), forward
  only, with the reference's count term N_q in the normaliser (fastmax.py:269-271, quirk Q2).
* Meta tensors pass through shape-only (the pretrain scripts' flop count runs the model on the meta
  device, where the reference's own host hop fails: quirk Q7).
* fp32 inputs of head size <= 64 with N >= 128 run on the tensor pipe in split precision (every fp32 operand as a
  [hi | lo] bf16 pair, three-term products, fp32 accumulation: ~1e-5 against the fp64 formula; csrc/fm_dense128.cu);
  other fp32 shapes use the fp32 CUDA-core kernels (~1e-6).
* bf16 heads narrower than 64 (Pythia: 32) are zero-padded to 64, heads between 64 and 128 (phi: 80, Llama-3.2: 96 ...)
  to 128, so that they run on the tcgen05 kernels (D = 64 moment / scan kernels, D = 128 dense tiles): zero features
  change neither the scores nor the normaliser, and the padded output / gradient columns are exactly zero and sliced
  off.  ``c`` still uses the caller's D.
"""
from __future__ import annotations

import math
import os

import torch

from . import _lib

__all__ = ["fastmax", "fastattention_einops", "FastmaxFunction", "fastmax_fwd_raw", "fastmax_bwd_raw", "rect_forward"]

_DTYPES = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16}
_TENSOR_PATH_D = 64      # head size of the tcgen05 moment / scan kernels
_DENSE_PATH_D = 128      # head size of the dense tcgen05 tiles (csrc/fm_dense128.cu)


def _head_pad(dtype, D, N):
    """Zero features to append so that a bf16 head runs on a tcgen05 kernel (0: none / not applicable)."""
    if dtype != torch.bfloat16:
        return 0
    if D < _TENSOR_PATH_D and N >= 256:
        return _TENSOR_PATH_D - D
    if _TENSOR_PATH_D < D < _DENSE_PATH_D and N >= 128:
        return _DENSE_PATH_D - D
    return 0
_RECT_STATE_MAX_Q = 64   # N_q != N_k with at most this many queries goes through the recurrent-state kernels


def _check_inputs(q, k, v, rect_ok=False):
    if not (q.is_cuda and k.is_cuda and v.is_cuda):
        raise _lib.FastmaxError("fastmax_experiments_b200.fastmax needs CUDA tensors "
                                "(no CPU fallback; the CPU restatement lives in oracle/ for tests only)")
    same = q.shape == k.shape
    if rect_ok and q.dim() == 4 and k.dim() == 4:     # inference: N_q != N_k allowed
        same = q.shape[:2] == k.shape[:2] and q.shape[3] == k.shape[3]
    if q.dim() != 4 or not same or k.shape != v.shape:
        raise ValueError(f"q, k, v must share one (B, H, N, D) shape (N_q may differ from N_k only with "
                         f"mask=False), got {tuple(q.shape)}, {tuple(k.shape)}, {tuple(v.shape)}")
    if q.dtype not in _DTYPES or k.dtype != q.dtype or v.dtype != q.dtype:
        raise TypeError(f"q, k, v must all be float32 or all bfloat16, got {q.dtype}, {k.dtype}, {v.dtype}")


def _workspace(nbytes: int, device) -> torch.Tensor | None:
    if nbytes == 0:
        return None
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


# Forward moment states are kept for backward when they fit this budget (bytes per call); above it
# backward rebuilds them from k, v.  0 disables keeping (the reference's contract: only q,k,v,o,g saved).
SAVE_STATES_MAX_BYTES = int(os.environ.get("FASTMAX_SAVE_STATES_MAX_BYTES", str(6 << 30)))


def saved_state_bytes(q, causal: bool, p: int, algo: int = _lib.ALGO_AUTO) -> int:
    B, H, N, D = q.shape
    return _lib.load().fastmax_saved_state_bytes(B, H, N, D, p, int(causal), _DTYPES[q.dtype], algo)


def _require_dense(**tensors):
    """The C ABI takes dense (B,H,N,D) buffers: a strided view would be read as if it were dense."""
    for name, t in tensors.items():
        if not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous (got strides {tuple(t.stride())} for shape {tuple(t.shape)})")


def fastmax_fwd_raw(q, k, v, causal: bool, inv_c: float, p: int, algo: int = _lib.ALGO_AUTO, saved=None):
    """One call of the C ABI ``fastmax_fwd`` (``fastmax_fwd_ex`` when ``saved`` is a state buffer) on
    contiguous (B,H,N,D) CUDA tensors -> (o, g)."""
    lib = _lib.load()
    _require_dense(q=q, k=k, v=v)
    B, H, N, D = q.shape
    dt = _DTYPES[q.dtype]
    o = torch.empty(q.shape, dtype=q.dtype, device=q.device)
    g = torch.empty(B, H, N, dtype=torch.float32, device=q.device)
    nbytes = lib.fastmax_fwd_workspace_bytes(B, H, N, D, p, int(causal), dt, algo)
    ws = _workspace(nbytes, q.device)
    with torch.cuda.device(q.device):
        rc = lib.fastmax_fwd_ex(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), g.data_ptr(),
                                saved.data_ptr() if saved is not None else None,
                                B, H, N, D, p, int(causal), inv_c, dt, algo,
                                ws.data_ptr() if ws is not None else None, nbytes, _stream_ptr(q.device))
    _lib.check(rc, "fastmax_fwd")
    return o, g


def fastmax_bwd_raw(q, k, v, o, g, do, causal: bool, inv_c: float, p: int, algo: int = _lib.ALGO_AUTO, saved=None):
    """One call of the C ABI ``fastmax_bwd`` (``fastmax_bwd_ex`` with kept states) -> (dq, dk, dv)."""
    lib = _lib.load()
    _require_dense(q=q, k=k, v=v, o=o, g=g, do=do)
    B, H, N, D = q.shape
    dt = _DTYPES[q.dtype]
    dq, dk, dv = (torch.empty(q.shape, dtype=q.dtype, device=q.device) for _ in range(3))
    nbytes = lib.fastmax_bwd_workspace_bytes(B, H, N, D, p, int(causal), dt, algo)
    ws = _workspace(nbytes, q.device)
    with torch.cuda.device(q.device):
        rc = lib.fastmax_bwd_ex(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), g.data_ptr(),
                                do.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                saved.data_ptr() if saved is not None else None,
                                B, H, N, D, p, int(causal), inv_c, dt, algo,
                                ws.data_ptr() if ws is not None else None, nbytes, _stream_ptr(q.device))
    _lib.check(rc, "fastmax_bwd")
    return dq, dk, dv


def rect_forward(q, k, v, inv_c: float, p: int, count: str = "q"):
    """Non-causal forward with N_q != N_k (prefill / decode over the KV cache, model.py:478-487 with
    ``input_pos`` given).  The shorter side is zero-padded along the sequence -- zero keys carry zero values
    and only add to the count term, zero queries are sliced off -- and the count term of the normaliser is
    then set to what the reference uses: ``count="q"``: N_q (fastmax.py:269-271, quirk Q2);
    ``count="k"``: N_k (fastmax_hack.py:21, the linearmax dispatch)."""
    Nq, Nk, D = q.shape[2], k.shape[2], q.shape[3]
    F = torch.nn.functional
    q_, k_, v_ = q.contiguous(), k.contiguous(), v.contiguous()
    if Nq <= _RECT_STATE_MAX_Q:
        # decode / short prefill: reduce the cache to its moment state once (fp32) and evaluate only the real
        # query rows against it, instead of padding q to N_k rows -- one normalisation, one rounding
        from .decode import FastmaxState
        st = FastmaxState(q.shape[0], q.shape[1], D, p=p, device=q.device, tensors_normalized=True)
        st.inv_c = inv_c
        st.update(k_, v_)
        return st.attend(q_, count_term=float(Nq if count == "q" else Nk))
    if Nq < Nk:
        q_ = F.pad(q_, (0, 0, 0, Nk - Nq))
    else:
        k_, v_ = F.pad(k_, (0, 0, 0, Nq - Nk)), F.pad(v_, (0, 0, 0, Nq - Nk))
    pad = _head_pad(q.dtype, D, q_.shape[2])
    if pad:
        q_, k_, v_ = (F.pad(t, (0, pad)) for t in (q_, k_, v_))
    o, g = fastmax_fwd_raw(q_, k_, v_, False, inv_c, p)
    o, g = o[:, :, :Nq, :D], g[:, :, :Nq]
    have = max(Nq, Nk)                       # count term of the padded problem
    want = Nq if count == "q" else Nk
    if have != want:
        o = (o.float() * (g / (g - float(have - want))).unsqueeze(-1)).to(q.dtype)
    return o.contiguous()


def _inference_rect(ctx, q, k, v, inv_c, p):
    """Forward only: the reference's non-causal backward assumes one N as well."""
    if any(ctx.needs_input_grad[:3]):
        raise NotImplementedError("fastmax with N_q != N_k is an inference (forward-only) path")
    o = rect_forward(q, k, v, inv_c, p, count="q")
    ctx.mark_non_differentiable(o)
    return o


class FastmaxFunction(torch.autograd.Function):
    """Autograd wrapper with the reference's nine-argument ``apply`` (fastmax.py:42-53)."""

    @staticmethod
    def forward(ctx, q, k, v, mask=True, normalize_term=8, tensors_normalized=False, p=1,
                dropout_rate=0.0, create_attn=False):
        if p not in (1, 2):
            raise ValueError(f"p should be 1 or 2, got p={p}")
        if create_attn is not False:
            raise NotImplementedError("create_attn=True (dense debug path, fastmax.py:99-104) is not "
                                      "provided; its backward is broken in the reference as well")
        causal = mask is not False                                                 # fastmax.py:85
        if q.device.type == "meta":
            # shape-only pass (pretrain/openwebtext.py:146-156 runs the model on the meta device to count
            # flops; the reference's .cpu() hop crashes there: quirk Q7)
            ctx.meta = True
            return torch.empty_like(q)
        ctx.meta = False
        _check_inputs(q, k, v, rect_ok=not causal)
        D = q.shape[-1]
        c = 1.0 if tensors_normalized is True else normalize_term * math.sqrt(D)   # fastmax.py:79-82
        if q.shape[2] != k.shape[2]:
            return _inference_rect(ctx, q, k, v, 1.0 / c, p)
        q_, k_, v_ = q.contiguous(), k.contiguous(), v.contiguous()
        pad = _head_pad(q.dtype, D, q.shape[2])
        if pad:
            q_, k_, v_ = (torch.nn.functional.pad(t, (0, pad)) for t in (q_, k_, v_))
        saved = None
        if any(t.requires_grad for t in (q, k, v)):
            nb = saved_state_bytes(q_, causal, p)
            if 0 < nb <= SAVE_STATES_MAX_BYTES:
                saved = torch.empty(nb, dtype=torch.uint8, device=q.device)
        o, g = fastmax_fwd_raw(q_, k_, v_, causal, 1.0 / c, p, saved=saved)
        ctx.save_for_backward(q_, k_, v_, o, g)                                    # fastmax.py:106
        ctx.causal, ctx.inv_c, ctx.p, ctx.saved, ctx.pad = causal, 1.0 / c, p, saved, pad
        return o[..., :D].contiguous() if pad else o

    @staticmethod
    def backward(ctx, o_grad):
        if ctx.meta:
            return (torch.empty_like(o_grad),) * 3 + (None,) * 6
        q, k, v, o, g = ctx.saved_tensors
        do = o_grad.to(q.dtype).contiguous()
        if ctx.pad:
            do = torch.nn.functional.pad(do, (0, ctx.pad))
        dq, dk, dv = fastmax_bwd_raw(q, k, v, o, g, do, ctx.causal, ctx.inv_c, ctx.p, saved=ctx.saved)
        ctx.saved = None
        if ctx.pad:
            D = q.shape[-1] - ctx.pad
            dq, dk, dv = (t[..., :D].contiguous() for t in (dq, dk, dv))
        return dq, dk, dv, None, None, None, None, None, None                      # fastmax.py:182


# the reference's class name, for code that calls ``fastattention_einops.apply`` directly
fastattention_einops = FastmaxFunction


def fastmax(q, k, v, mask=True, normalize_term=8, tensors_normalized=False, p=1, dropout_rate=0.0,
            create_attn=False):
    """Drop-in for ``attention_mechanisms.fastmax.fastmax`` (fastmax.py:7-27)."""
    return FastmaxFunction.apply(q, k, v, mask, normalize_term, tensors_normalized, p, dropout_rate,
                                 create_attn)

"""Grouped-query FastMax and the fused q/k/v prologue (no reference counterpart as an op).

The reference assembles q, k, v in ``CausalSelfAttention.forward`` (lit_gpt/model.py:392-425): view + permute +
split of the projection output, ``expand`` of k, v over the ``q_per_kv`` query heads of a group (:408-416), reshape
to ``(B, heads, T, hs)`` (:418-420), rotary embedding on the first ``rope_n_elem`` features (:422-425, :702-708).
Only then does the attention op run -- on k, v that are ``q_per_kv`` identical copies.

* ``qkv_prologue`` does the assembly in one HBM pass (``csrc/fm_prologue.cu``), bit-exact against the reference's
  eager arithmetic, and leaves k, v un-expanded: ``(B, n_query_groups, T, hs)``.
* ``fastmax_gqa(q, k, v, ...)`` is ``fastmax`` for ``k, v`` of shape ``(B, Hkv, N, D)``: the forward moment states
  are built (and kept) once per (b, kv head), every kernel reads the shared rows in place, dk and dv come back summed
  over the group.  Equal to ``fastmax(q, expand(k), expand(v))`` with autograd through the expand.
* ``fused_attention_forward`` is ``CausalSelfAttention.forward`` with the two above in place of model.py:396-451 for
  the FastMax training path (``attn_alg`` fastmax / fastmax_cuda, ``input_pos is None``); everything else goes to the
  class's own forward.  ``patch_causal_self_attention(cls, fused=True)`` binds it.

Both ops need CUDA tensors and the in-tree library; shapes the kernels do not cover (fp32 or p = 1 grouped
attention, head sizes that are not whole 16-byte chunks) are computed by the un-fused CUDA path (``fastmax`` on
expanded views, torch ops for the prologue) -- never on the host.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from .fastmax import (SAVE_STATES_MAX_BYTES, _DTYPES, _TENSOR_PATH_D, _require_dense, _stream_ptr, _workspace, fastmax)

__all__ = ["fastmax_gqa", "qkv_prologue", "fused_attention_forward", "expand_kv"]


def expand_kv(t: torch.Tensor, q_per_kv: int) -> torch.Tensor:
    """(B, Hkv, N, D) -> (B, Hkv * q_per_kv, N, D), query head h reading kv head h // q_per_kv (model.py:408-420)."""
    B, Hkv, N, D = t.shape
    return t.unsqueeze(2).expand(B, Hkv, q_per_kv, N, D).reshape(B, Hkv * q_per_kv, N, D)


def _gqa_native(q, k, p: int) -> bool:
    return (q.is_cuda and q.dtype == torch.bfloat16 and p == 2 and q.shape[-1] <= _TENSOR_PATH_D
            and q.shape[2] >= 256 and q.shape[2] == k.shape[2])


class FastmaxGQAFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, causal: bool, inv_c: float, p: int):
        lib = _lib.load()
        B, H, N, D = q.shape
        Hkv = k.shape[1]
        q_, k_, v_ = q.contiguous(), k.contiguous(), v.contiguous()
        pad = _TENSOR_PATH_D - D
        if pad:
            q_, k_, v_ = (torch.nn.functional.pad(t, (0, pad)) for t in (q_, k_, v_))
        Dp = D + pad
        dt = _DTYPES[q.dtype]
        saved = None
        if any(t.requires_grad for t in (q, k, v)):
            nb = lib.fastmax_gqa_saved_state_bytes(B, H, Hkv, N, Dp, p, int(causal), dt)
            if 0 < nb <= SAVE_STATES_MAX_BYTES:
                saved = torch.empty(nb, dtype=torch.uint8, device=q.device)
        o = torch.empty_like(q_)
        g = torch.empty(B, H, N, dtype=torch.float32, device=q.device)
        nbytes = lib.fastmax_gqa_workspace_bytes(B, H, Hkv, N, Dp, p, int(causal), dt, 0)
        ws = _workspace(nbytes, q.device)
        with torch.cuda.device(q.device):
            rc = lib.fastmax_fwd_gqa(q_.data_ptr(), k_.data_ptr(), v_.data_ptr(), o.data_ptr(), g.data_ptr(),
                                     saved.data_ptr() if saved is not None else None, B, H, Hkv, N, Dp, p, int(causal),
                                     inv_c, dt, ws.data_ptr(), nbytes, _stream_ptr(q.device))
        _lib.check(rc, "fastmax_fwd_gqa")
        ctx.save_for_backward(q_, k_, v_, o, g)
        ctx.saved_states, ctx.args, ctx.D = saved, (causal, inv_c, p), D
        return o[..., :D].contiguous() if pad else o

    @staticmethod
    def backward(ctx, grad_o):
        lib = _lib.load()
        q, k, v, o, g = ctx.saved_tensors
        causal, inv_c, p = ctx.args
        B, H, N, Dp = q.shape
        Hkv, D = k.shape[1], ctx.D
        do = grad_o.contiguous()
        if Dp != D:
            do = torch.nn.functional.pad(do, (0, Dp - D))
        _require_dense(do=do)
        dt = _DTYPES[q.dtype]
        dq = torch.empty_like(q)
        dk, dv = torch.empty_like(k), torch.empty_like(v)
        nbytes = lib.fastmax_gqa_workspace_bytes(B, H, Hkv, N, Dp, p, int(causal), dt, 1)
        ws = _workspace(nbytes, q.device)
        saved = ctx.saved_states
        with torch.cuda.device(q.device):
            rc = lib.fastmax_bwd_gqa(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), g.data_ptr(), do.data_ptr(),
                                     dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                     saved.data_ptr() if saved is not None else None, B, H, Hkv, N, Dp, p, int(causal),
                                     inv_c, dt, ws.data_ptr(), nbytes, _stream_ptr(q.device))
        _lib.check(rc, "fastmax_bwd_gqa")
        if Dp != D:
            dq, dk, dv = dq[..., :D].contiguous(), dk[..., :D].contiguous(), dv[..., :D].contiguous()
        return dq, dk, dv, None, None, None


def fastmax_gqa(q, k, v, mask=True, normalize_term=8, tensors_normalized=False, p=2):
    """``fastmax`` (fastmax.py:7-27, same ``mask`` / ``normalize_term`` / ``tensors_normalized`` / ``p`` meaning) for
    grouped keys and values: q ``(B, H, N, D)``, k, v ``(B, Hkv, N, D)`` with ``H % Hkv == 0``."""
    if p not in (1, 2):
        raise ValueError(f"p should be 1 or 2, got p={p}")
    if q.dim() != 4 or k.dim() != 4 or k.shape != v.shape or q.shape[0] != k.shape[0] or q.shape[3] != k.shape[3] \
            or k.shape[1] == 0 or q.shape[1] % k.shape[1] != 0:
        raise ValueError(f"expected q (B,H,N,D) and k, v (B,Hkv,N,D) with H % Hkv == 0, got {tuple(q.shape)}, "
                         f"{tuple(k.shape)}, {tuple(v.shape)}")
    if not q.is_cuda:
        raise _lib.FastmaxError("fastmax_gqa needs CUDA tensors (no CPU fallback)")
    q_per_kv = q.shape[1] // k.shape[1]
    if q_per_kv == 1 or not _gqa_native(q, k, p):
        return fastmax(q, expand_kv(k, q_per_kv), expand_kv(v, q_per_kv), mask=mask, normalize_term=normalize_term,
                       tensors_normalized=tensors_normalized, p=p)
    causal = mask is not False                                                     # fastmax.py:85
    c = 1.0 if tensors_normalized is True else normalize_term * math.sqrt(q.shape[-1])   # fastmax.py:79-82
    return FastmaxGQAFunction.apply(q, k, v, causal, 1.0 / c, p)


# ------------------------------------------------------------------------------------------------- prologue
def _prologue_torch(qkv, cos, sin, G, q_per_kv, hs, rope_n):
    """model.py:396-425 in torch ops, minus the expand (shapes the kernel does not take)."""
    B, T, _ = qkv.shape
    x = qkv.view(B, T, G, q_per_kv + 2, hs).permute(0, 2, 3, 1, 4)
    q, k, v = x.split((q_per_kv, 1, 1), dim=2)
    q, k, v = q.reshape(B, -1, T, hs), k.reshape(B, -1, T, hs), v.reshape(B, -1, T, hs)

    def rope(t):
        r = t[..., :rope_n]
        h = rope_n // 2
        rot = torch.cat((-r[..., h:], r[..., :h]), dim=-1)
        return torch.cat((((r * cos) + (rot * sin)).to(t.dtype), t[..., rope_n:]), dim=-1)
    return rope(q), rope(k), v


class QKVPrologueFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, cos, sin, G: int, q_per_kv: int, hs: int, rope_n: int):
        lib = _lib.load()
        B, T, _ = qkv.shape
        qkv_ = qkv.contiguous()
        cos_, sin_ = cos.contiguous(), sin.contiguous()
        q = torch.empty(B, G * q_per_kv, T, hs, dtype=qkv.dtype, device=qkv.device)
        k = torch.empty(B, G, T, hs, dtype=qkv.dtype, device=qkv.device)
        v = torch.empty_like(k)
        with torch.cuda.device(qkv.device):
            rc = lib.fastmax_qkv_prologue(qkv_.data_ptr(), cos_.data_ptr(), sin_.data_ptr(), q.data_ptr(), k.data_ptr(),
                                          v.data_ptr(), B, T, G, q_per_kv, hs, rope_n, _DTYPES[qkv.dtype],
                                          _DTYPES[cos.dtype], _stream_ptr(qkv.device))
        _lib.check(rc, "fastmax_qkv_prologue")
        ctx.save_for_backward(cos_, sin_)
        ctx.dims = (B, T, G, q_per_kv, hs, rope_n)
        return q, k, v

    @staticmethod
    def backward(ctx, dq, dk, dv):
        lib = _lib.load()
        cos, sin = ctx.saved_tensors
        B, T, G, q_per_kv, hs, rope_n = ctx.dims
        dq, dk, dv = dq.contiguous(), dk.contiguous(), dv.contiguous()
        dqkv = torch.empty(B, T, G * (q_per_kv + 2) * hs, dtype=dq.dtype, device=dq.device)
        with torch.cuda.device(dq.device):
            rc = lib.fastmax_qkv_prologue_bwd(dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), cos.data_ptr(), sin.data_ptr(),
                                              dqkv.data_ptr(), B, T, G, q_per_kv, hs, rope_n, _DTYPES[dq.dtype],
                                              _DTYPES[cos.dtype], _stream_ptr(dq.device))
        _lib.check(rc, "fastmax_qkv_prologue_bwd")
        return dqkv, None, None, None, None, None, None


def _prologue_native(qkv, cos, hs, rope_n) -> bool:
    if not qkv.is_cuda or qkv.dtype not in _DTYPES or cos.dtype not in _DTYPES:
        return False
    if qkv.dtype == torch.float32 and cos.dtype != torch.float32:
        return False
    el = 8 if qkv.dtype == torch.bfloat16 else 4        # elements per 16-byte chunk; the kernel moves chunk pairs
    return hs % (2 * el) == 0 and rope_n % 2 == 0 and (rope_n // 2) % el == 0 and (hs - rope_n) % (2 * el) == 0


def qkv_prologue(qkv, cos, sin, n_query_groups: int, q_per_kv: int, head_size: int, rope_n_elem: int):
    """Projection output ``(B, T, n_query_groups * (q_per_kv + 2) * head_size)`` -> ``q (B, n_head, T, hs)``,
    ``k, v (B, n_query_groups, T, hs)`` with the rotary embedding applied to q and k (model.py:396-425 without the
    expand).  ``cos`` / ``sin``: the model's ``(T, rope_n_elem)`` tables."""
    if qkv.dim() != 3 or qkv.shape[2] != n_query_groups * (q_per_kv + 2) * head_size:
        raise ValueError(f"qkv must be (B, T, {n_query_groups * (q_per_kv + 2) * head_size}), got {tuple(qkv.shape)}")
    if cos.shape != (qkv.shape[1], rope_n_elem) or sin.shape != cos.shape:
        raise ValueError(f"cos, sin must be (T, rope_n_elem) = ({qkv.shape[1]}, {rope_n_elem}), got {tuple(cos.shape)}")
    if not qkv.is_cuda:
        raise _lib.FastmaxError("qkv_prologue needs CUDA tensors (no CPU fallback)")
    if not _prologue_native(qkv, cos, head_size, rope_n_elem):
        return _prologue_torch(qkv, cos, sin, n_query_groups, q_per_kv, head_size, rope_n_elem)
    return QKVPrologueFunction.apply(qkv, cos, sin, n_query_groups, q_per_kv, head_size, rope_n_elem)


def fused_attention_forward(self, x, cos, sin, mask=None, input_pos=None):
    """``CausalSelfAttention.forward`` (model.py:380-458) for the FastMax training path: projection -> fused prologue ->
    grouped FastMax -> the reference's own output reshape (quirk Q1 kept, :453-455) -> output projection."""
    cfg = self.config
    alg = cfg.attn_alg[0] if isinstance(cfg.attn_alg, tuple) else cfg.attn_alg
    if input_pos is not None or alg not in ("fastmax", "fastmax_cuda") or not x.is_cuda:
        return self._reference_forward(x, cos, sin, mask, input_pos)
    q_per_kv = cfg.n_head // cfg.n_query_groups
    el = 8 if x.dtype == torch.bfloat16 else 4
    chunked = (cfg.head_size % (2 * el) == 0 and cfg.rope_n_elem % 2 == 0 and (cfg.rope_n_elem // 2) % el == 0
               and (cfg.head_size - cfg.rope_n_elem) % (2 * el) == 0)
    if q_per_kv == 1 and not chunked:
        # nothing to fuse and nothing to share (e.g. Pythia: 32-wide heads with 8 rotary features): the class's own
        # forward with the rebound attention methods is the same work
        return self._reference_forward(x, cos, sin, mask, input_pos)
    B, T, _ = x.size()
    qkv = self.attn(x)
    q, k, v = qkv_prologue(qkv, cos, sin, cfg.n_query_groups, q_per_kv, cfg.head_size, cfg.rope_n_elem)
    y = fastmax_gqa(q, k, v, mask=True, p=2)
    y = y.reshape(B, T, cfg.head_size * cfg.n_head)
    return self.proj(y)

"""CPU restatement of the q/k/v assembly in front of the attention op (TEST INFRASTRUCTURE -- the product never
imports this; it checks ``csrc/fm_prologue.cu`` and ``fastmax_experiments_b200/gqa.py``).

Follows ``CausalSelfAttention.forward`` of the reference, lit_gpt/model.py:396-425, and ``apply_rope``,
lit_gpt/model.py:702-708, in torch eager arithmetic (so that bf16 rounds after every elementwise op exactly as the
reference does), and is pinned to q, k, v captured from the unmodified reference module
(tests/golden/gqa_prologue_*.npz, minted by oracle/make_golden.py::main_gqa).
"""
from __future__ import annotations

import torch


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """model.py:702-708."""
    h = x.size(-1) // 2
    rotated = torch.cat((-x[..., h:], x[..., :h]), dim=-1)          # :704-706
    return ((x * cos) + (rotated * sin)).to(dtype=x.dtype)           # :707-708


def qkv_assemble(qkv: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, n_query_groups: int, q_per_kv: int,
                 head_size: int, rope_n_elem: int, expand: bool = True):
    """(B, T, G * (q_per_kv + 2) * hs) -> q (B, G * q_per_kv, T, hs), k, v (B, G * q_per_kv or G, T, hs).

    ``expand=True`` is the reference (k, v copied over the group, model.py:408-416); ``expand=False`` leaves them
    per group, which is what the CUDA prologue produces -- RoPE is applied per row, so the two differ only by the
    copy."""
    B, T, _ = qkv.shape
    x = qkv.view(B, T, n_query_groups, q_per_kv + 2, head_size)      # :397-399
    x = x.permute(0, 2, 3, 1, 4)                                      # :401
    q, k, v = x.split((q_per_kv, 1, 1), dim=2)                        # :404
    if expand and q_per_kv != 1:                                      # :408-416 (training: input_pos is None)
        k = k.expand(B, n_query_groups, q_per_kv, T, head_size)
        v = v.expand(B, n_query_groups, q_per_kv, T, head_size)
    q = q.reshape(B, -1, T, head_size)                                # :418-420
    k = k.reshape(B, -1, T, head_size)
    v = v.reshape(B, -1, T, head_size)
    q_roped = apply_rope(q[..., :rope_n_elem], cos, sin)              # :422-425
    k_roped = apply_rope(k[..., :rope_n_elem], cos, sin)
    q = torch.cat((q_roped, q[..., rope_n_elem:]), dim=-1)
    k = torch.cat((k_roped, k[..., rope_n_elem:]), dim=-1)
    return q, k, v


def qkv_assemble_grad(dq, dk, dv, qkv_shape, cos, sin, n_query_groups, q_per_kv, head_size, rope_n_elem, dtype):
    """Gradient of ``qkv_assemble(expand=False)`` w.r.t. qkv for upstream dq, dk, dv, through torch autograd (what the
    reference's graph does once the sum over the expanded group copies has been taken)."""
    qkv = torch.zeros(qkv_shape, dtype=dtype, requires_grad=True)
    q, k, v = qkv_assemble(qkv, cos, sin, n_query_groups, q_per_kv, head_size, rope_n_elem, expand=False)
    torch.autograd.backward((q, k, v), (dq, dk, dv))
    return qkv.grad

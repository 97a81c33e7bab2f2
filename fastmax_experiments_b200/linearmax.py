"""``linearmax`` -- drop-in for the reference's ``attention_mechanisms/fastmax_hack.py::fastmax_hack``
(dispatched as ``attn_alg="linearmax"``, lit_gpt/model.py:444-445, 460-472).

It is first-order FastMax with ``c = 1`` after a normalisation prologue on q and k (fastmax_hack.py:9-15,
41-46): subtract each token's mean over the features, then divide the whole (b,h) sequence by its largest
token norm.  The reference differentiates all of it with torch autograd; here the prologue stays in torch
(elementwise work plus two reductions, differentiated by autograd) and the attention itself runs on the
p = 1 CUDA path with its hand-written backward (``fastmax(..., p=1, tensors_normalized=True)``).
"""
from __future__ import annotations

import torch

from .fastmax import _check_inputs, fastmax, rect_forward

__all__ = ["linearmax", "linearmax_prologue"]


def linearmax_prologue(q: torch.Tensor, k: torch.Tensor):
    """fastmax_hack.py:9-15 (non-causal) / :41-46 (causal) -- the same arithmetic in both branches.
    Computed in fp32 and returned in the input dtype."""
    dt = q.dtype
    q, k = q.float(), k.float()
    q = q - q.mean(dim=3, keepdim=True)
    k = k - k.mean(dim=3, keepdim=True)
    qn = torch.linalg.norm(q, dim=3).amax(dim=2)[:, :, None, None]
    kn = torch.linalg.norm(k, dim=3).amax(dim=2)[:, :, None, None]
    return (q / qn).to(dt), (k / kn).to(dt)


def linearmax(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, p: int = 1, mask=True) -> torch.Tensor:
    """Same signature as ``fastmax_hack(q, k, v, p=1, mask=True)``.  The reference's non-causal branch is
    first order whatever ``p`` says (fastmax_hack.py:18-34) and its dispatch always passes p=1; other
    values are rejected here."""
    if p != 1:
        raise ValueError(f"linearmax is the first-order kernel (the reference dispatch passes p=1), got p={p}")
    qn, kn = linearmax_prologue(q, k)
    if not mask and q.shape[2] != k.shape[2]:
        # prefill / decode over the KV cache (model.py:460-472 with input_pos given): forward only, and the count
        # term of the normaliser is the number of KEYS (fastmax_hack.py:21), not fastmax.py's N_q
        if torch.is_grad_enabled() and any(t.requires_grad for t in (q, k, v)):
            raise NotImplementedError("linearmax with N_q != N_k is an inference (forward-only) path")
        _check_inputs(qn, kn, v, rect_ok=True)
        return rect_forward(qn, kn, v, 1.0, 1, count="k")
    return fastmax(qn, kn, v, mask=True if mask else False, tensors_normalized=True, p=1)

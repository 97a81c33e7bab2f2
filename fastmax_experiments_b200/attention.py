"""Drop-in for the FastMax branch of the reference's attention dispatch (lit_gpt/model.py).

The reference routes ``config.attn_alg`` through a string switch (model.py:432-451).  Its FastMax
branches are

* ``CausalSelfAttention.fastmax``  (model.py:475-487): ``q,k,v -> .cpu() -> fastmax(q,k,v,p=2,
  mask=input_pos is None) -> .cuda()``;
* ``CausalSelfAttention.fastmax_cuda`` (model.py:489-532): the wrapper around the never-vendored
  ``fastmax_cuda`` extension (un-runnable as committed: SURVEY.md quirk Q6).

``fastmax_attention`` below has the signature and return layout of those methods --
``(B, n_head, T, head_size)`` in, same layout out; the caller keeps doing its own
``y.reshape(B, T, C)`` (model.py:453-455, including quirk Q1: the reference does *not*
transpose heads and time before that reshape; that is the dispatch's behaviour, not the op's, and
it is left to the caller) -- but stays on the GPU: no host round trip, hand-written sm_100a kernels.

``linearmax_attention`` does the same for ``CausalSelfAttention.linearmax`` (model.py:460-472,
``attention_mechanisms/fastmax_hack.py``: first-order FastMax with c = 1 behind a q,k normalisation).

``patch_causal_self_attention(cls)`` rebinds the three methods on a lit_gpt-style class so an
unmodified ``lit_gpt.model.GPT`` picks the CUDA path up.
"""
from __future__ import annotations

from typing import Optional

import torch

from .fastmax import fastmax
from .linearmax import linearmax

__all__ = ["fastmax_attention", "linearmax_attention", "patch_causal_self_attention"]


def fastmax_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
                      input_pos: Optional[torch.Tensor] = None, p: int = 2) -> torch.Tensor:
    """Mirror of ``CausalSelfAttention.fastmax(self, q, k, v, input_pos)`` (model.py:475-487).

    ``input_pos is not None`` means "KV cache / inference" in the reference and selects
    ``mask=False`` (model.py:478-481); training passes ``input_pos=None`` -> causal."""
    mask = True
    if input_pos is not None:
        mask = False
    return fastmax(q, k, v, p=p, mask=mask)


def _method_fastmax(self, q, k, v, input_pos):           # bound as CausalSelfAttention.fastmax
    return fastmax_attention(q, k, v, input_pos, p=2)


def _method_fastmax_cuda(self, q, k, v):                  # bound as CausalSelfAttention.fastmax_cuda
    # model.py:489-532 hard-codes mask=True and p=2-equivalent coefficients (a0,a1,a2 = 1,1,.5);
    # RPE / dropout-noise / q,k-normalisation arguments of the external extension have no
    # counterpart in fastmax.py and are not applied.
    return fastmax(q, k, v, p=2, mask=True)


def patch_causal_self_attention(cls, fused: bool = False) -> type:
    """Rebind ``cls.fastmax`` / ``cls.fastmax_cuda`` / ``cls.linearmax`` (a ``lit_gpt.model.CausalSelfAttention``-like
    class) to the CUDA path.  ``fused=True`` also replaces ``cls.forward`` for the FastMax training path by
    ``gqa.fused_attention_forward`` (one-pass q/k/v prologue, k and v not expanded over the query groups:
    SURVEY.md section 8f-1); the class's own forward stays reachable as ``cls._reference_forward`` and still serves
    every other ``attn_alg`` and the KV-cache path.  Returns ``cls`` for chaining."""
    for name in ("forward",):
        if not hasattr(cls, name):
            raise TypeError(f"{cls!r} does not look like lit_gpt.model.CausalSelfAttention")
    cls.fastmax = _method_fastmax
    cls.fastmax_cuda = _method_fastmax_cuda
    cls.linearmax = _method_linearmax
    if fused and not hasattr(cls, "_reference_forward"):
        from .gqa import fused_attention_forward
        cls._reference_forward = cls.forward
        cls.forward = fused_attention_forward
    return cls


def linearmax_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
                        input_pos: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Mirror of ``CausalSelfAttention.linearmax(self, q, k, v, input_pos)`` (model.py:460-472):
    ``fastmax_hack(q, k, v, p=1, mask=input_pos is None)``."""
    return linearmax(q, k, v, p=1, mask=input_pos is None)


def _method_linearmax(self, q, k, v, input_pos):         # bound as CausalSelfAttention.linearmax
    return linearmax_attention(q, k, v, input_pos)

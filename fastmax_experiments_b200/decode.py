"""Constant-size recurrent FastMax state for generation.

The reference generates with a KV cache: ``CausalSelfAttention.forward`` writes the new k, v rows into a
``max_seq_length``-row cache (lit_gpt/model.py:427-430) and calls ``fastmax(q, k_cache, v_cache, mask=False)``
(model.py:478-487), i.e. the non-causal op of the new queries over the whole zero-padded cache -- work and
memory proportional to ``max_seq_length`` for every token.  ``FastmaxState`` keeps the moment state of the keys
instead ((1 + D + D^2)(D + 1) floats per (b,h)); one decode step is ``update(k_new, v_new)`` + ``attend(q_new)``
and costs the same whatever the position.

The count term of the normaliser defaults to what the reference uses for the same call, the number of QUERY
rows (fastmax.py:269-271, quirk Q2); it can be overridden.  With that, ``attend`` after ``update`` equals the
reference's ``fastmax(q, k_cache, v_cache, mask=False)`` as long as cache rows are only ever appended (a
position that is overwritten cannot be taken out of the sums).  Forward only.  ``linearmax`` has no such state:
its prologue divides k by the largest key norm of the whole cache (fastmax_hack.py:9-15), which changes with
every appended key.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from .fastmax import _DTYPES, _require_dense, _stream_ptr

__all__ = ["FastmaxState"]


class FastmaxState:
    def __init__(self, B: int, H: int, D: int, p: int = 2, device="cuda", normalize_term=8, tensors_normalized=False):
        if p not in (1, 2):
            raise ValueError(f"p should be 1 or 2, got p={p}")
        self.B, self.H, self.D, self.p = B, H, D, p
        self.inv_c = 1.0 if tensors_normalized is True else 1.0 / (normalize_term * math.sqrt(D))
        lib = _lib.load()
        nb = lib.fastmax_decode_state_bytes(B, H, D, p)
        if nb == 0:
            raise _lib.FastmaxError(f"unsupported state shape B={B} H={H} D={D} p={p}")
        self.state = torch.empty(nb, dtype=torch.uint8, device=device)
        self.n_keys = 0
        self.reset()

    def reset(self) -> None:
        with torch.cuda.device(self.state.device):
            rc = _lib.load().fastmax_decode_init(self.state.data_ptr(), self.B, self.H, self.D, self.p,
                                                 _stream_ptr(self.state.device))
        _lib.check(rc, "fastmax_decode_init")
        self.n_keys = 0

    def _check(self, *ts):
        for t in ts:
            if not t.is_cuda or t.dtype not in _DTYPES or t.dim() != 4 or t.shape[:2] != (self.B, self.H) or t.shape[3] != self.D:
                raise ValueError(f"expected CUDA (B={self.B}, H={self.H}, T, D={self.D}) float32 / bfloat16 tensors, "
                                 f"got {tuple(t.shape)} {t.dtype} on {t.device}")

    def update(self, k: torch.Tensor, v: torch.Tensor) -> None:
        """Add the T new key / value rows (B, H, T, D) to the state."""
        self._check(k, v)
        k, v = k.contiguous(), v.contiguous()
        _require_dense(k=k, v=v)
        T = k.shape[2]
        with torch.cuda.device(k.device):
            rc = _lib.load().fastmax_decode_update(self.state.data_ptr(), k.data_ptr(), v.data_ptr(), self.B, self.H, T,
                                                   self.D, self.p, _DTYPES[k.dtype], _stream_ptr(k.device))
        _lib.check(rc, "fastmax_decode_update")
        self.n_keys += T

    def attend(self, q: torch.Tensor, count_term: float | None = None) -> torch.Tensor:
        """o (B, H, T, D) for the T query rows against everything added so far."""
        self._check(q)
        q = q.contiguous()
        T = q.shape[2]
        if count_term is None:
            count_term = float(T)                    # fastmax.py:269-271: the reference counts the QUERIES (quirk Q2)
        o = torch.empty_like(q)
        with torch.cuda.device(q.device):
            rc = _lib.load().fastmax_decode_attend(self.state.data_ptr(), q.data_ptr(), o.data_ptr(), self.B, self.H, T,
                                                   self.D, self.p, self.inv_c, float(count_term), _DTYPES[q.dtype],
                                                   _stream_ptr(q.device))
        _lib.check(rc, "fastmax_decode_attend")
        return o

"""Host-buffer entry of the hot path: the call the reference's dispatch makes.

``CausalSelfAttention.fastmax`` in the reference (lit_gpt/model.py:475-487) moves q, k, v to the
host, runs ``fastmax(q, k, v, p=2, mask=...)`` there and moves ``o`` back.  The drop-in keeps
the data on the GPU (``attention.py``); this module provides the same *host-in / host-out*
contract for callers that really hold host tensors (and for ``bench.py``'s ``e2e`` number):
pinned host q, k, v (and do) -> device -> kernels -> pinned host o (and dq, dk, dv).

The copies are pipelined against the kernels over slices of the batch dimension on three
streams (H2D, compute, D2H), so the end-to-end time approaches max(copy, compute) instead of
their sum.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from .fastmax import SAVE_STATES_MAX_BYTES, fastmax_bwd_raw, fastmax_fwd_raw, saved_state_bytes

__all__ = ["fastmax_fwd_bwd_host"]

# One (H2D, compute, D2H) stream triple per device, created once: torch's caching allocator keys its
# free blocks by stream, so fresh streams per call would turn every device buffer into a cudaMalloc.
_STREAMS: dict = {}


def _streams(dev):
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _STREAMS:
        _STREAMS[key] = tuple(torch.cuda.Stream(dev) for _ in range(3))
    return _STREAMS[key]


def fastmax_fwd_bwd_host(q, k, v, do, out=None, mask=True, normalize_term=8, tensors_normalized=False,
                         p=2, device="cuda", chunks=None):
    """q, k, v, do: pinned host tensors (B,H,N,D).  Returns pinned host (o, dq, dk, dv).

    ``out``: optional tuple of four pre-allocated pinned host tensors to write into.
    Returns after all device work and copies have completed (host-synchronous, like the
    reference's CPU call)."""
    if p not in (1, 2):
        raise ValueError(f"p should be 1 or 2, got p={p}")
    if q.is_cuda:
        raise _lib.FastmaxError("fastmax_fwd_bwd_host takes host tensors; use fastmax() for device tensors")
    B, H, N, D = q.shape
    c = 1.0 if tensors_normalized is True else normalize_term * math.sqrt(D)
    causal = mask is not False
    dev = torch.device(device)
    if out is None:
        out = tuple(torch.empty_like(q).pin_memory() for _ in range(4))
    # (b,h) units are independent: slice the flattened (B*H) axis, so the pipeline has enough stages to
    # hide its fill and drain (one slice of H2D before, one of D2H after the overlapped part) even at small B
    BH = B * H
    flat = all(t.is_contiguous() for t in (q, k, v, do) + tuple(out))
    if flat:
        q, k, v, do = (t.view(BH, 1, N, D) for t in (q, k, v, do))
        out_v = tuple(t.view(BH, 1, N, D) for t in out)
        units = BH
    else:
        out_v, units = out, B
    chunks = max(1, min(units, chunks or 16))
    bounds = [(i * units // chunks, (i + 1) * units // chunks) for i in range(chunks)]
    s_in, s_cmp, s_out = _streams(dev)
    ev_in = [torch.cuda.Event() for _ in bounds]
    ev_cmp = [torch.cuda.Event() for _ in bounds]
    keep = []
    for i, (b0, b1) in enumerate(bounds):
        with torch.cuda.stream(s_in):
            # (.to() keeps a strided host view's strides; the kernels take dense (B,H,N,D))
            dq_, dk_, dv_, ddo = (t[b0:b1].to(dev, non_blocking=True).contiguous() for t in (q, k, v, do))
            ev_in[i].record(s_in)
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev_in[i])
            nb = saved_state_bytes(dq_, causal, p) if dq_.dtype == torch.bfloat16 else 0
            saved = torch.empty(nb, dtype=torch.uint8, device=dev) if 0 < nb <= SAVE_STATES_MAX_BYTES else None
            o, g = fastmax_fwd_raw(dq_, dk_, dv_, causal, 1.0 / c, p, saved=saved)
            gq, gk, gv = fastmax_bwd_raw(dq_, dk_, dv_, o, g, ddo, causal, 1.0 / c, p, saved=saved)
            ev_cmp[i].record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp[i])
            for dst, src in zip(out_v, (o, gq, gk, gv)):
                dst[b0:b1].copy_(src, non_blocking=True)
        keep.append((dq_, dk_, dv_, ddo, o, g, gq, gk, gv, saved))   # keep alive until the streams drain
    s_out.synchronize()
    s_cmp.synchronize()
    return out

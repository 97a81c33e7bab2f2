"""Sequence-parallel FastMax over the GPUs of one box (BASELINE.json north_star).

The reference has no counterpart (its only long-context lever is the O(N) algorithm itself, and its
masked p=2 path needs N*D^3 floats per (b,h)).  Here a sequence is split into contiguous token
ranges, one per rank.  FastMax outputs are *linear* in the moment state of the preceding tokens,
so one state exchange per scan direction is enough:

    forward : Z_r = state(k_r, v_r)            -> all-gather -> P_r = sum_{r'<r} Z_r'  -> local scan from P_r
    backward: Y_r = state_rev(q_r, do_r/g_r)   -> all-gather -> Q_r = sum_{r'>r} Y_r'  -> local scans from P_r, Q_r

A record is 545 KiB per (b,h) (17.4 MB at B=1, H=32); the all-gather moves G of them per rank over
NVLink.  bf16, D = 64, causal, p in {1,2}.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .fastmax import _require_dense

__all__ = ["partition", "exchange_states", "fastmax_sp", "FastmaxSPFunction", "sp_local_state",
           "sp_local_state_rev", "sp_reduce", "sp_fwd", "sp_bwd"]


def partition(n_tokens: int, world: int, rank: int, multiple: int = 512) -> Tuple[int, int]:
    """Contiguous token range of `rank`: equal shares rounded up to `multiple` (one scan segment),
    the last ranks take what is left (possibly nothing)."""
    share = -(-n_tokens // world)
    share = -(-share // multiple) * multiple
    start = min(n_tokens, rank * share)
    return start, min(n_tokens, start + share)


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _ws(nbytes, dev):
    return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)


# ---- thin wrappers over the C ABI (one call each) -------------------------------------------------
def sp_local_state(k, v, p: int) -> torch.Tensor:
    lib = _lib.load()
    _require_dense(k=k, v=v)
    B, H, N, D = k.shape
    out = torch.empty(lib.fastmax_sp_state_bytes(B, H), dtype=torch.uint8, device=k.device)
    nb = lib.fastmax_sp_local_state_workspace_bytes(B, H, N, 0)
    ws = _ws(nb, k.device)
    with torch.cuda.device(k.device):
        rc = lib.fastmax_sp_local_state(k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, N, D, p,
                                        ws.data_ptr(), nb, _stream(k.device))
    _lib.check(rc, "fastmax_sp_local_state")
    return out


def sp_local_state_rev(q, o, g, do, p: int) -> torch.Tensor:
    lib = _lib.load()
    _require_dense(q=q, o=o, g=g, do=do)
    B, H, N, D = q.shape
    out = torch.empty(lib.fastmax_sp_state_bytes(B, H), dtype=torch.uint8, device=q.device)
    nb = lib.fastmax_sp_local_state_workspace_bytes(B, H, N, 1)
    ws = _ws(nb, q.device)
    with torch.cuda.device(q.device):
        rc = lib.fastmax_sp_local_state_rev(q.data_ptr(), o.data_ptr(), g.data_ptr(), do.data_ptr(), out.data_ptr(),
                                            B, H, N, D, p, ws.data_ptr(), nb, _stream(q.device))
    _lib.check(rc, "fastmax_sp_local_state_rev")
    return out


def sp_reduce(gathered: torch.Tensor, world: int, rank: int, direction: int, B: int, H: int) -> torch.Tensor:
    """Exclusive prefix (direction 0) / suffix (direction 1) of the gathered records, on the GPU."""
    lib = _lib.load()
    out = torch.empty(lib.fastmax_sp_state_bytes(B, H), dtype=torch.uint8, device=gathered.device)
    with torch.cuda.device(gathered.device):
        rc = lib.fastmax_sp_reduce(gathered.data_ptr(), world, rank, direction, out.data_ptr(), B, H,
                                   _stream(gathered.device))
    _lib.check(rc, "fastmax_sp_reduce")
    return out


def sp_fwd(q, k, v, init, token_offset: int, inv_c: float, p: int):
    lib = _lib.load()
    B, H, N, D = q.shape
    o = torch.empty_like(q)
    g = torch.empty(B, H, N, dtype=torch.float32, device=q.device)
    nb = lib.fastmax_fwd_workspace_bytes(B, H, N, D, p, 1, _lib.DTYPE_BF16, _lib.ALGO_MOMENTS)
    ws = _ws(nb, q.device)
    with torch.cuda.device(q.device):
        rc = lib.fastmax_fwd_sp(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), g.data_ptr(),
                                init.data_ptr() if init is not None else None, token_offset,
                                B, H, N, D, p, inv_c, _lib.DTYPE_BF16, ws.data_ptr(), nb, _stream(q.device))
    _lib.check(rc, "fastmax_fwd_sp")
    return o, g


def sp_bwd(q, k, v, o, g, do, init_fwd, init_rev, token_offset: int, inv_c: float, p: int):
    lib = _lib.load()
    B, H, N, D = q.shape
    dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    nb = lib.fastmax_bwd_workspace_bytes(B, H, N, D, p, 1, _lib.DTYPE_BF16, _lib.ALGO_MOMENTS)
    ws = _ws(nb, q.device)
    ptr = lambda t: t.data_ptr() if t is not None else None
    with torch.cuda.device(q.device):
        rc = lib.fastmax_bwd_sp(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), g.data_ptr(), do.data_ptr(),
                                dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), ptr(init_fwd), ptr(init_rev),
                                token_offset, B, H, N, D, p, inv_c, _lib.DTYPE_BF16, ws.data_ptr(), nb,
                                _stream(q.device))
    _lib.check(rc, "fastmax_bwd_sp")
    return dq, dk, dv


# ---- the exchange step (host logic; testable on CPU with gloo) ---------------------------------------
def exchange_states(local: torch.Tensor, group, direction: int,
                    reduce_fn: Callable[[torch.Tensor, int, int, int], torch.Tensor]) -> Optional[torch.Tensor]:
    """All-gather the per-rank state records and reduce them to this rank's exclusive prefix
    (direction 0) or suffix (direction 1).  Returns None when there is nothing before / after.

    `reduce_fn(gathered[world, ...], world, rank, direction)` does the sum (the CUDA kernel in
    production; injectable so the collective logic is testable on CPU)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if (direction == 0 and rank == 0 and world == 1) or world == 1:
        return None
    gathered = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather(list(gathered.unbind(0)), local.contiguous(), group=group)
    if (direction == 0 and rank == 0) or (direction == 1 and rank == world - 1):
        return None
    return reduce_fn(gathered, world, rank, direction)


def _token_offset(n_local: int, group, device) -> int:
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = torch.zeros(world, 1, dtype=torch.int64, device=device)
    mine = torch.tensor([n_local], dtype=torch.int64, device=device)
    dist.all_gather(list(counts.unbind(0)), mine, group=group)
    return int(counts[:rank].sum().item())


def _require_supported(q) -> None:
    if not q.is_cuda or q.dtype != torch.bfloat16 or q.shape[-1] != 64:
        raise _lib.FastmaxError("fastmax_sp: bf16 CUDA tensors with D = 64 only (no CPU fallback)")


def _zero_state(B: int, H: int, device) -> torch.Tensor:
    """The moment state of no tokens (all-zero record): what a rank with an empty token range contributes."""
    return torch.zeros(_lib.load().fastmax_sp_state_bytes(B, H), dtype=torch.uint8, device=device)


class FastmaxSPFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, group, normalize_term, tensors_normalized, p, token_offset):
        if p not in (1, 2):
            raise ValueError(f"p should be 1 or 2, got p={p}")
        _require_supported(q)
        B, H, N, D = q.shape
        c = 1.0 if tensors_normalized is True else normalize_term * math.sqrt(D)
        q_, k_, v_ = q.contiguous(), k.contiguous(), v.contiguous()
        offset = _token_offset(N, group, q.device) if token_offset is None else int(token_offset)
        red = lambda gathered, world, rank, d: sp_reduce(gathered, world, rank, d, B, H)
        # a rank whose token range is empty (partition() of a short sequence) still takes part in both
        # exchanges, with the state of no tokens, and returns empty outputs
        local = sp_local_state(k_, v_, p) if N > 0 else _zero_state(B, H, q.device)
        init = exchange_states(local, group, 0, red)
        if N > 0:
            o, g = sp_fwd(q_, k_, v_, init, offset, 1.0 / c, p)
        else:
            o, g = torch.empty_like(q_), torch.empty(B, H, 0, dtype=torch.float32, device=q.device)
        ctx.save_for_backward(q_, k_, v_, o, g)
        ctx.init, ctx.offset, ctx.inv_c, ctx.p, ctx.group = init, offset, 1.0 / c, p, group
        return o

    @staticmethod
    def backward(ctx, o_grad):
        q, k, v, o, g = ctx.saved_tensors
        B, H, N, D = q.shape
        do = o_grad.to(q.dtype).contiguous()
        red = lambda gathered, world, rank, d: sp_reduce(gathered, world, rank, d, B, H)
        local = sp_local_state_rev(q, o, g, do, ctx.p) if N > 0 else _zero_state(B, H, q.device)
        init_rev = exchange_states(local, ctx.group, 1, red)
        if N > 0:
            dq, dk, dv = sp_bwd(q, k, v, o, g, do, ctx.init, init_rev, ctx.offset, ctx.inv_c, ctx.p)
        else:
            dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
        return dq, dk, dv, None, None, None, None, None


def fastmax_sp(q, k, v, group=None, normalize_term=8, tensors_normalized=False, p=2, token_offset=None):
    """Causal FastMax over a sequence sharded across `group` (default: the world group): every rank
    passes its own contiguous token range (B, H, N_local, 64) in rank order (N_local may be 0).
    `token_offset` = number of tokens held by earlier ranks; a caller that knows its partition passes it
    (``partition(n, world, rank)[0]``), otherwise it is found with one small all-gather and a host read."""
    return FastmaxSPFunction.apply(q, k, v, group, normalize_term, tensors_normalized, p, token_offset)

"""Sequence-parallel path: host logic on CPU (gloo, world_size 2) and, on the GPU box, the kernels
with the ranks emulated one after the other on a single device."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastmax_experiments_b200 import sequence_parallel as sp
from oracle import fastmax_oracle as orc


def test_partition_covers_sequence_without_overlap():
    for n, world in ((1 << 20, 8), (5000, 4), (100, 8), (131072, 2), (513, 2)):
        spans = [sp.partition(n, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0 and a0 <= a1
        assert all((a1 - a0) % 512 == 0 for a0, a1 in spans[:-1] if a1 < n)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # toy "state records": exact small integers so sums are order independent
        local = torch.full((3, 5), float(rank + 1))
        calls = []

        def cpu_reduce(gathered, w, r, direction):
            calls.append((w, r, direction))
            sel = gathered[:r] if direction == 0 else gathered[r + 1:]
            return sel.sum(0)

        pre = sp.exchange_states(local, None, 0, cpu_reduce)
        suf = sp.exchange_states(local, None, 1, cpu_reduce)
        exp_pre = sum(range(1, rank + 1))
        exp_suf = sum(range(rank + 2, world + 1))
        ok = True
        ok &= (pre is None) if rank == 0 else bool((pre == exp_pre).all())
        ok &= (suf is None) if rank == world - 1 else bool((suf == exp_suf).all())
        off = sp._token_offset(100 * (rank + 1), None, torch.device("cpu"))
        ok &= off == sum(100 * (r + 1) for r in range(rank))
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_states_gloo(world):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)


def _toy_worker(rank, world, port, n_total, ret):
    """FastmaxSPFunction's control flow over gloo with toy stand-ins for the five kernel wrappers (a "state" is
    the sum of the local k rows; o = q + state of everything before; the reverse state is the sum of do): ranks
    with an EMPTY token range must take part in both exchanges instead of leaving the others in the collective."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sp._require_supported = lambda q: None
        sp._zero_state = lambda B, H, dev: torch.zeros(B, H, 4)
        sp.sp_local_state = lambda k, v, p: k.sum(2)
        sp.sp_local_state_rev = lambda q, o, g, do, p: do.sum(2)
        sp.sp_reduce = lambda gathered, w, r, d, B, H: (gathered[:r] if d == 0 else gathered[r + 1:]).sum(0)

        def toy_fwd(q, k, v, init, off, inv_c, p):
            pre = torch.cumsum(k, 2) - k + (init[:, :, None] if init is not None else 0)
            return q + pre, torch.full(q.shape[:3], float(off))

        def toy_bwd(q, k, v, o, g, do, init_f, init_r, off, inv_c, p):
            suf = torch.flip(torch.cumsum(torch.flip(do, [2]), 2), [2]) - do + (init_r[:, :, None] if init_r is not None else 0)
            return do.clone(), suf, torch.zeros_like(v)

        sp.sp_fwd, sp.sp_bwd = toy_fwd, toy_bwd
        gen = torch.Generator().manual_seed(0)
        q, k, v, do = (torch.randint(-3, 4, (1, 2, n_total, 4), generator=gen).float() for _ in range(4))
        a0, a1 = sp.partition(n_total, world, rank)
        ql, kl, vl = (t[:, :, a0:a1].clone().requires_grad_(True) for t in (q, k, v))
        o = sp.fastmax_sp(ql, kl, vl, p=2, token_offset=a0)
        o.backward(do[:, :, a0:a1])
        ok = o.shape[2] == a1 - a0
        ok &= bool((o.detach() == (q + torch.cumsum(k, 2) - k)[:, :, a0:a1]).all())
        suf_full = torch.flip(torch.cumsum(torch.flip(do, [2]), 2), [2]) - do
        ok &= bool((kl.grad == suf_full[:, :, a0:a1]).all())
        ret[rank] = (bool(ok), a1 - a0)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(3, 1000), (2, 100), (3, 1536)])
def test_sequence_parallel_protocol_with_empty_tail_ranks(world, n_total):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_toy_worker, args=(world, port, n_total, ret), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world)), dict(ret)
    if n_total <= 512 * (world - 1):
        assert any(ret[r][1] == 0 for r in range(world))      # the case under test: a rank without tokens


@pytest.mark.gpu
@pytest.mark.parametrize("world,N,p", [(2, 2048, 2), (4, 4096, 2), (3, 3000, 2), (2, 1536, 1)])
def test_sequence_parallel_kernels_emulated_ranks(world, N, p):
    """All ranks of the exchange run one after the other on one GPU (no collective: the gather is a
    torch.stack); outputs and gradients must match the single-sequence oracle."""
    B, H, D = 1, 2, 64
    gen = torch.Generator().manual_seed(world * 100 + p)
    q, k, v, do = (torch.randn(B, H, N, D, generator=gen).to(torch.bfloat16) for _ in range(4))
    ref = orc.dense_fwd_bwd(q.double(), k.double(), v.double(), do.double(), mask=True, p=p)
    inv_c = 1.0 / orc.denominator_c(D)
    spans = [sp.partition(N, world, r) for r in range(world)]
    loc = [tuple(t[:, :, a:b].contiguous().cuda() for t in (q, k, v, do)) for a, b in spans]
    # forward
    states = torch.stack([sp.sp_local_state(kk, vv, p) for (_, kk, vv, _) in loc])
    outs = []
    for r, (qq, kk, vv, _) in enumerate(loc):
        init = sp.sp_reduce(states, world, r, 0, B, H) if r > 0 else None
        outs.append(sp.sp_fwd(qq, kk, vv, init, spans[r][0], inv_c, p) + (init,))
    o = torch.cat([x[0] for x in outs], 2)
    g = torch.cat([x[1] for x in outs], 2)
    assert orc.rel_err(o, ref[0]) <= 2e-2
    assert orc.rel_err(g, ref[1]) <= 2e-2
    # backward
    rstates = torch.stack([sp.sp_local_state_rev(qq, outs[r][0], outs[r][1], dd, p)
                           for r, (qq, _, _, dd) in enumerate(loc)])
    grads = []
    for r, (qq, kk, vv, dd) in enumerate(loc):
        init_rev = sp.sp_reduce(rstates, world, r, 1, B, H) if r < world - 1 else None
        grads.append(sp.sp_bwd(qq, kk, vv, outs[r][0], outs[r][1], dd, outs[r][2], init_rev, spans[r][0], inv_c, p))
    for idx, nm in enumerate(("dq", "dk", "dv")):
        x = torch.cat([gr[idx] for gr in grads], 2)
        assert orc.rel_err(x, ref[2 + idx]) <= 2e-2, (nm, orc.rel_err(x, ref[2 + idx]))

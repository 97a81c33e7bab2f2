"""q/k/v prologue and grouped-query FastMax.

CPU: the restatement oracle/prologue_oracle.py is pinned bit-for-bit to q, k, v (and the gradient into the qkv
projection) captured from the UNMODIFIED reference attention module (tests/golden/gqa_prologue_*.npz).
GPU: csrc/fm_prologue.cu against that oracle, bit-exact, forward and transpose; fastmax_gqa against fastmax on
expanded k, v and against the fp64 dense formula; the fused CausalSelfAttention.forward against the reference module's
output and input gradient."""
import pytest
import torch

from conftest import golden_gqa_names, load_golden_gqa
from oracle import fastmax_oracle as orc
from oracle import prologue_oracle as pro

NAMES = golden_gqa_names()


def _dims(gd):
    G = gd["n_query_groups"]
    return G, gd["n_head"] // G, gd["head_size"], gd["rope_n_elem"]


def test_fixtures_present():
    assert len(NAMES) >= 2


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("tag", ["f32", "bf16"])
def test_oracle_matches_reference_module(name, tag):
    gd = load_golden_gqa(name, tag)
    G, qpk, hs, rn = _dims(gd)
    q, k, v = pro.qkv_assemble(gd["qkv"], gd["cos"], gd["sin"], G, qpk, hs, rn, expand=True)
    assert torch.equal(q, gd["q"])
    assert torch.equal(k[:, ::qpk], gd["k"]) and torch.equal(k, gd["k"].repeat_interleave(qpk, 1))
    assert torch.equal(v[:, ::qpk], gd["v"]) and torch.equal(v, gd["v"].repeat_interleave(qpk, 1))
    q2, k2, v2 = pro.qkv_assemble(gd["qkv"], gd["cos"], gd["sin"], G, qpk, hs, rn, expand=False)
    assert torch.equal(q2, q) and torch.equal(k2, gd["k"]) and torch.equal(v2, gd["v"])
    # gradient into the projection: the q rows see only dq (bit-exact); the k, v rows see the group sums, which the
    # reference accumulates in the working dtype and the fixture holds as float64 sums
    dqkv = pro.qkv_assemble_grad(gd["dq"], gd["dk"].to(gd["dtype"]), gd["dv"].to(gd["dtype"]), gd["qkv"].shape, gd["cos"],
                                 gd["sin"], G, qpk, hs, rn, gd["dtype"])
    a = dqkv.view(1, gd["T"], G, qpk + 2, hs)
    b = gd["dqkv"].view(1, gd["T"], G, qpk + 2, hs)
    assert torch.equal(a[:, :, :, :qpk], b[:, :, :, :qpk])
    tol = 1e-6 if tag == "f32" else 1.6e-2
    assert orc.rel_err(a[:, :, :, qpk:].double(), b[:, :, :, qpk:].double()) < tol


# ------------------------------------------------------------------------------------------------------- GPU
def _cases():
    # (B, T, G, qpk, hs, rope_n, dtype, table dtype)
    bf, f32 = torch.bfloat16, torch.float32
    return [(2, 130, 2, 4, 64, 64, bf, bf), (1, 257, 4, 8, 64, 64, bf, f32), (2, 96, 2, 1, 64, 16, bf, bf),
            (1, 200, 1, 4, 128, 128, bf, bf), (2, 130, 2, 4, 64, 64, f32, f32), (1, 77, 3, 2, 32, 8, f32, f32),
            (1, 64, 2, 2, 80, 32, bf, bf)]


@pytest.mark.gpu
@pytest.mark.parametrize("case", _cases(), ids=lambda c: "-".join(str(x).replace("torch.", "") for x in c))
def test_prologue_kernel_bit_exact_vs_oracle(case):
    from fastmax_experiments_b200.gqa import QKVPrologueFunction
    B, T, G, qpk, hs, rn, dt, dtc = case
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(B, T, G * (qpk + 2) * hs, generator=g).to(dt)
    ang = torch.randn(T, rn // 2, generator=g).repeat(1, 2)
    cos, sin = torch.cos(ang).to(dtc), torch.sin(ang).to(dtc)
    rq, rk, rv = pro.qkv_assemble(qkv, cos, sin, G, qpk, hs, rn, expand=False)
    dq, dk, dv = (torch.randn(t.shape, generator=g).to(dt) for t in (rq, rk, rv))
    rd = pro.qkv_assemble_grad(dq, dk, dv, qkv.shape, cos, sin, G, qpk, hs, rn, dt)
    x = qkv.cuda().requires_grad_(True)
    q, k, v = QKVPrologueFunction.apply(x, cos.cuda(), sin.cuda(), G, qpk, hs, rn)
    assert torch.equal(q.cpu(), rq) and torch.equal(k.cpu(), rk) and torch.equal(v.cpu(), rv)
    torch.autograd.backward((q, k, v), (dq.cuda(), dk.cuda(), dv.cuda()))
    assert torch.equal(x.grad.cpu(), rd)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("tag", ["f32", "bf16"])
def test_prologue_kernel_matches_reference_fixture(name, tag):
    from fastmax_experiments_b200 import qkv_prologue
    gd = load_golden_gqa(name, tag)
    G, qpk, hs, rn = _dims(gd)
    q, k, v = qkv_prologue(gd["qkv"].cuda(), gd["cos"].cuda(), gd["sin"].cuda(), G, qpk, hs, rn)
    assert torch.equal(q.cpu(), gd["q"]) and torch.equal(k.cpu(), gd["k"]) and torch.equal(v.cpu(), gd["v"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 8, 2, 2600, 64, True), (2, 4, 1, 1100, 64, True), (1, 8, 4, 2600, 32, True),
                                   (1, 8, 2, 2600, 64, False), (2, 6, 2, 700, 64, True)],
                         ids=lambda s: "B%d-H%d-Hkv%d-N%d-D%d-%s" % (s[:5] + ("causal" if s[5] else "full",)))
def test_gqa_matches_expanded_mha_and_dense_oracle(shape):
    """Several 512-token segments (build / apply / intra kernels with the kv index mapping), one dense segment (N=700 /
    1100), head padding (D=32), non-causal.  Against the expanded MHA call o and dq come from the same kernels on the
    same operands and differ only by the build kernel's run-to-run fp32 accumulation order (four issuer warps), i.e. by
    bf16 ulps of the state; dk, dv are group sums of bf16 rows here and a bf16 autograd accumulation there."""
    from fastmax_experiments_b200 import fastmax, fastmax_gqa
    from fastmax_experiments_b200.gqa import expand_kv
    B, H, Hkv, N, D, causal = shape
    g = torch.Generator().manual_seed(11)
    q = torch.randn(B, H, N, D, generator=g).to(torch.bfloat16)
    k, v = (torch.randn(B, Hkv, N, D, generator=g).to(torch.bfloat16) for _ in range(2))
    do = torch.randn(B, H, N, D, generator=g).to(torch.bfloat16)
    mask = True if causal else False
    qa, ka, va = (t.cuda().requires_grad_(True) for t in (q, k, v))
    oa = fastmax_gqa(qa, ka, va, mask=mask, p=2)
    oa.backward(do.cuda())
    qb, kb, vb = (t.cuda().requires_grad_(True) for t in (q, k, v))
    ob = fastmax(qb, expand_kv(kb, H // Hkv), expand_kv(vb, H // Hkv), mask=mask, p=2)
    ob.backward(do.cuda())
    for nm, x, y in (("o", oa, ob), ("dq", qa.grad, qb.grad), ("dk", ka.grad, kb.grad), ("dv", va.grad, vb.grad)):
        assert orc.rel_err(x.double(), y.double()) < 5e-3, (shape, nm)
    # against the dense formula in float64 on the expanded inputs
    ke, ve = (expand_kv(t, H // Hkv).double() for t in (k, v))
    ro, _, rdq, rdk, rdv = orc.dense_fwd_bwd(q.double(), ke, ve, do.double(), mask=mask, p=2)
    fold = lambda t: t.view(B, Hkv, H // Hkv, N, D).sum(2)
    for nm, x, r in (("o", oa, ro), ("dq", qa.grad, rdq), ("dk", ka.grad, fold(rdk)), ("dv", va.grad, fold(rdv))):
        err = orc.rel_err(x.double().cpu(), r)
        assert err <= 2e-2, (shape, nm, err)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_fused_attention_forward_matches_reference_module(name):
    """The unmodified lit_gpt CausalSelfAttention with patch_causal_self_attention(fused=True): output and gradient
    into the module input against the reference module run on the host in bf16 (fixture), both within the bf16 bar."""
    from oracle import ref_loader as rl
    if not rl.installed():
        pytest.skip("baseline/_ref not installed")
    from fastmax_experiments_b200 import patch_causal_self_attention
    gd = load_golden_gqa(name, "bf16")
    model, cfg = rl.build_gpt("tiny-llama-1.1b", attn_alg="fastmax", block_size=gd["T"], seed=gd["seed"], n_layer=1,
                              n_embd=256, n_head=4, n_query_groups=2, intermediate_size=256, rotary_percentage=gd["rot"])
    model = model.to(torch.bfloat16).cuda()
    attn = model.transformer.h[0].attn
    cls = type(attn)
    orig = cls.forward
    try:
        patch_causal_self_attention(cls, fused=True)
        x = gd["x"].cuda().requires_grad_(True)
        y = attn(x, gd["cos"].cuda(), gd["sin"].cuda())
        y.backward(gd["dy"].cuda())
    finally:
        cls.forward = orig
        if hasattr(cls, "_reference_forward"):
            del cls._reference_forward
    assert orc.rel_err(y.double().cpu(), gd["y"].double()) <= 2e-2
    # the host reference keeps every cumulative sum in bf16 (its own error vs fp64 is 6e-3 .. 1.9e-2)
    assert torch.isfinite(x.grad).all()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 1, 1, 257), (1, 2, 1, 1025), (1, 1, 1, 2047), (1, 2, 2, 2049), (1, 2, 1, 2689), (1, 4, 2, 4097)],
                         ids=lambda s: "B%d-H%d-Hkv%d-N%d" % s)
def test_ragged_shapes_second_generation_intra_kernels(shape):
    """Ragged tails around the tile / segment sizes of csrc/fm_seg.cu and csrc/fm_seg_bwd.cu (a last CTA with one live
    tile, one dense segment of 1024 / 2048 tokens with CTAs deep inside it, one token past a segment boundary), with and
    without grouped keys, against the fp64 dense formula."""
    from fastmax_experiments_b200 import fastmax_gqa
    from fastmax_experiments_b200.gqa import expand_kv
    B, H, Hkv, N = shape
    g = torch.Generator().manual_seed(N)
    q = torch.randn(B, H, N, 64, generator=g).to(torch.bfloat16)
    k, v = (torch.randn(B, Hkv, N, 64, generator=g).to(torch.bfloat16) for _ in range(2))
    do = torch.randn(B, H, N, 64, generator=g).to(torch.bfloat16)
    qa, ka, va = (t.cuda().requires_grad_(True) for t in (q, k, v))
    o = fastmax_gqa(qa, ka, va, mask=True, p=2)
    o.backward(do.cuda())
    ke, ve = (expand_kv(t, H // Hkv).double() for t in (k, v))
    ro, _, rdq, rdk, rdv = orc.dense_fwd_bwd(q.double(), ke, ve, do.double(), mask=True, p=2)
    fold = lambda t: t.view(B, Hkv, H // Hkv, N, 64).sum(2)
    for nm, x, r in (("o", o, ro), ("dq", qa.grad, rdq), ("dk", ka.grad, fold(rdk)), ("dv", va.grad, fold(rdv))):
        err = orc.rel_err(x.double().cpu(), r)
        assert err <= 2e-2, (shape, nm, err)

"""Dispatch mirror (lit_gpt/model.py:432-487) and the fastmax_cuda stand-in (model.py:82-84,116)."""
import inspect
import os
import sys
import types

import pytest
import torch

import fastmax_experiments_b200 as fm
from fastmax_experiments_b200 import attention, fastmax_cuda_compat
from oracle import fastmax_oracle as orc

REF = "/root/reference"


def test_patch_rebinds_methods():
    class FakeCSA(torch.nn.Module):
        def forward(self, x):
            return x

        def fastmax(self, q, k, v, input_pos):
            raise AssertionError("reference CPU path must not run")

    attention.patch_causal_self_attention(FakeCSA)
    assert FakeCSA.fastmax is attention._method_fastmax
    assert list(inspect.signature(FakeCSA.fastmax).parameters) == ["self", "q", "k", "v", "input_pos"]
    assert list(inspect.signature(FakeCSA.fastmax_cuda).parameters) == ["self", "q", "k", "v"]
    assert list(inspect.signature(FakeCSA.linearmax).parameters) == ["self", "q", "k", "v", "input_pos"]
    with pytest.raises(TypeError):
        attention.patch_causal_self_attention(object)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_reference_method_signatures_match():
    """The methods we rebind have the argument lists of the reference's own (model.py:475,489)."""
    src = open(os.path.join(REF, "lit_gpt", "model.py")).read()
    assert "def fastmax(\n        self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, input_pos: torch.Tensor" in src
    assert "def fastmax_cuda(\n        self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor\n" in src
    assert "o = fastmax(q, k, v, p=2, mask=mask)" in src
    assert "def linearmax(\n        self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, input_pos: torch.Tensor" in src
    assert "o = fastmax_hack(q, k, v, p=1, mask=mask)" in src


def test_compat_rejects_undefined_extras():
    q = torch.zeros(4, 8, 2)
    with pytest.raises(NotImplementedError):
        fastmax_cuda_compat.forwardpass(q, q, q, None, None, True, 0.1, False, 1.0)
    with pytest.raises(NotImplementedError):
        fastmax_cuda_compat.forwardpass(q, q, q, None, torch.ones(7, 8), True, 0.0, False, 1.0)


@pytest.mark.gpu
def test_fastmax_attention_matches_oracle_training_and_inference_modes():
    torch.manual_seed(0)
    q, k, v = (torch.randn(2, 4, 96, 32, device="cuda") for _ in range(3))
    ref_c, _ = orc.dense_forward(q.double().cpu(), k.double().cpu(), v.double().cpu(), mask=True, p=2)
    ref_f, _ = orc.dense_forward(q.double().cpu(), k.double().cpu(), v.double().cpu(), mask=False, p=2)
    assert orc.rel_err(fm.fastmax_attention(q, k, v, None), ref_c) <= 1e-4          # training: causal
    assert orc.rel_err(fm.fastmax_attention(q, k, v, torch.arange(96)), ref_f) <= 1e-4   # KV-cache path


@pytest.mark.gpu
def test_fastmax_cuda_compat_layout_roundtrip():
    """(n, d, b*h) fp32 in/out, o carries g as row d, backwardpass consumes the unsliced o."""
    torch.manual_seed(1)
    n, d, bh = 80, 32, 3
    q, k, v, do = (torch.randn(n, d, bh, device="cuda") for _ in range(4))
    o = fastmax_cuda_compat.forwardpass(q, k, v, None, None, True, 0.0, False, 1.0)
    assert o.shape == (n, d + 1, bh)
    gq, gk, gv = fastmax_cuda_compat.backwardpass(q, k, v, o, do, True)
    as4 = lambda x: x.permute(2, 0, 1).unsqueeze(0).double().cpu()
    ref = orc.dense_fwd_bwd(as4(q), as4(k), as4(v), as4(do), mask=True, p=2)
    assert orc.rel_err(as4(o[:, :d, :]), ref[0]) <= 1e-4
    assert orc.rel_err(o[:, d, :].transpose(0, 1).unsqueeze(0), ref[1]) <= 1e-4
    for x, r in zip((gq, gk, gv), ref[2:]):
        assert orc.rel_err(as4(x), r) <= 1e-4


def test_meta_device_pass_is_shape_only():
    """pretrain/openwebtext.py:146-156 runs GPT.forward on the meta device to count flops (quirk Q7)."""
    import torch
    import fastmax_experiments_b200 as fm
    q = torch.empty(2, 4, 128, 64, device="meta", requires_grad=True)
    o = fm.fastmax(q, q, q, p=2)
    assert o.device.type == "meta" and o.shape == q.shape
    o.sum().backward()
    assert q.grad is not None and q.grad.shape == q.shape

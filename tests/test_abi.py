"""CPU: the C-ABI library builds in-tree, loads, and exports every symbol the header declares;
host-side argument handling of the drop-in op.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
import fastmax_experiments_b200 as fm
from fastmax_experiments_b200 import _lib, build


def header_symbols():
    src = open(os.path.join(ROOT, "include", "fastmax_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fastmax_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = header_symbols()
    assert sorted(_lib.SYMBOLS) == syms
    for s in syms:
        assert hasattr(lib, s), s


def test_abi_version_and_error_strings():
    lib = _lib.load()
    assert lib.fastmax_abi_version() == _lib.ABI_VERSION == 4
    assert lib.fastmax_error_string(0) == b"ok"
    assert b"unsupported" in lib.fastmax_error_string(_lib.ERR_UNSUPPORTED)


def test_plan_argument_checks():
    lib = _lib.load()
    pl = _lib.Plan()
    assert lib.fastmax_plan_make(ctypes.byref(pl), 1, 1, 128, 64, 3, 1, 1, 0) == _lib.ERR_BAD_ARG   # p=3
    assert lib.fastmax_plan_make(ctypes.byref(pl), 1, 1, 128, 256, 2, 1, 1, 0) == _lib.ERR_UNSUPPORTED
    assert lib.fastmax_plan_make(ctypes.byref(pl), 0, 1, 128, 64, 2, 1, 1, 0) == _lib.ERR_BAD_ARG
    assert lib.fastmax_plan_make(ctypes.byref(pl), 2, 4, 2048, 32, 2, 1, 1, 0) == 0
    assert pl.algo in (_lib.ALGO_EXACT, _lib.ALGO_MOMENTS)
    assert lib.fastmax_fwd(None, None, None, None, None, 1, 1, 8, 8, 2, 1, 1.0, 0, 0, None, 0, None) == _lib.ERR_BAD_ARG


def test_signature_matches_reference_defaults():
    import inspect
    sig = inspect.signature(fm.fastmax)
    assert list(sig.parameters) == ["q", "k", "v", "mask", "normalize_term", "tensors_normalized", "p",
                                    "dropout_rate", "create_attn"]
    d = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect._empty}
    assert d == dict(mask=True, normalize_term=8, tensors_normalized=False, p=1, dropout_rate=0.0,
                     create_attn=False)
    assert fm.fastattention_einops is fm.FastmaxFunction


def test_no_cpu_fallback_and_error_behaviour():
    q = torch.randn(1, 1, 8, 8)
    with pytest.raises(_lib.FastmaxError):
        fm.fastmax(q, q, q, p=2)                      # CPU tensors: loud failure, never a fallback
    with pytest.raises(ValueError):
        fm.fastmax(q, q, q, p=3)
    with pytest.raises(NotImplementedError):
        fm.fastmax(q, q, q, p=2, create_attn=True)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fastmax_experiments_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_graft_entry_build_runs():
    """The driver's build check: compiles (or finds up to date) the library and validates the ABI."""
    import __graft_entry__ as g
    g.build()

"""GPU: known-answer test of the hand-written sm_100a primitives (TMA swizzled loads, UMMA
shared-memory / instruction descriptors, A operand from shared and from tensor memory, TMEM
ld/st lane mapping) through the library's selftest entry point."""
import ctypes

import pytest
import torch

from fastmax_experiments_b200 import _lib

pytestmark = pytest.mark.gpu


def test_umma_selftest_matches_matmul():
    lib = _lib.load_dev()
    lib.fastmax_selftest_umma.restype = ctypes.c_int
    lib.fastmax_selftest_umma.argtypes = [ctypes.c_void_p] * 5
    torch.manual_seed(0)
    A = torch.randn(128, 64, device="cuda").to(torch.bfloat16)
    Bk = torch.randn(64, 64, device="cuda").to(torch.bfloat16)   # [n][k]
    Bn = torch.randn(64, 64, device="cuda").to(torch.bfloat16)   # [k][n]
    D = torch.full((4, 128, 64), float("nan"), device="cuda")
    rc = lib.fastmax_selftest_umma(A.data_ptr(), Bk.data_ptr(), Bn.data_ptr(), D.data_ptr(), None)
    _lib.check(rc, "fastmax_selftest_umma")
    torch.cuda.synchronize()
    ref_k = A.float() @ Bk.float().T
    ref_n = A.float() @ Bn.float()
    names = ["SS K-major B", "SS MN-major B", "TS K-major B", "TS MN-major B"]
    for i, ref in enumerate((ref_k, ref_n, ref_k, ref_n)):
        err = (D[i] - ref).abs().max().item()
        assert err < 1e-3, (names[i], err)

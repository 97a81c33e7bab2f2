"""ctypes binding of the C ABI in include/fastmax_b200.h (the thin shim between torch
tensors and libfastmax_b200.so).  There is no fallback: if the library is missing or a call
fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# FASTMAX_LIB: developer override (a variant build, e.g. with -DFM_TIMING); the default is the in-tree library
LIB_PATH = os.environ.get("FASTMAX_LIB") or os.path.join(HERE, "libfastmax_b200.so")

OK, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_CUDA = 0, -1, -2, -3, -4
DTYPE_F32, DTYPE_BF16 = 0, 1
ALGO_AUTO, ALGO_EXACT, ALGO_MOMENTS = 0, 1, 2

# every symbol include/fastmax_b200.h declares (tests/test_abi.py checks the two lists agree)
SYMBOLS = [
    "fastmax_abi_version", "fastmax_error_string", "fastmax_last_cuda_error",
    "fastmax_plan_make", "fastmax_fwd_workspace_bytes", "fastmax_bwd_workspace_bytes",
    "fastmax_fwd", "fastmax_bwd", "fastmax_last_launch_count",
    "fastmax_profile_enable", "fastmax_profile_fetch",
    "fastmax_saved_state_bytes", "fastmax_fwd_ex", "fastmax_bwd_ex",
    "fastmax_sp_state_bytes", "fastmax_sp_local_state_workspace_bytes", "fastmax_sp_local_state",
    "fastmax_sp_local_state_rev", "fastmax_sp_reduce", "fastmax_fwd_sp", "fastmax_bwd_sp",
    "fastmax_decode_state_bytes", "fastmax_decode_init", "fastmax_decode_update", "fastmax_decode_attend",
    "fastmax_gqa_workspace_bytes", "fastmax_gqa_saved_state_bytes", "fastmax_fwd_gqa", "fastmax_bwd_gqa",
    "fastmax_qkv_prologue", "fastmax_qkv_prologue_bwd",
]


class Plan(C.Structure):
    _fields_ = [
        ("algo", C.c_int32), ("seg_len", C.c_int32), ("n_seg", C.c_int32), ("reserved", C.c_int32),
        ("fwd_workspace_bytes", C.c_uint64), ("bwd_workspace_bytes", C.c_uint64),
        ("fwd_tensor_flops", C.c_double), ("bwd_tensor_flops", C.c_double),
        ("rebuild_tensor_flops", C.c_double),
    ]


class FastmaxError(RuntimeError):
    pass


_lib = None
_dev = None
ABI_VERSION = 4          # FASTMAX_B200_ABI_VERSION of include/fastmax_b200.h these argtypes were written for


def load_dev() -> C.CDLL:
    """The developer library (product + csrc/dev/: micro-benchmarks, primitive self-test); tools and one test only."""
    global _dev
    if _dev is None:
        load()
        _dev = C.CDLL(os.path.join(os.path.dirname(LIB_PATH), "libfastmax_b200_dev.so"))
    return _dev


def load() -> C.CDLL:
    """dlopen the in-tree library.  With nvcc on the PATH the library is (re)built first when its sources
    changed (a digest check, so the common case costs a few file reads); without nvcc -- the GPU box -- the
    prebuilt file is used as is.  Either way its ABI version must be the one this binding was written for."""
    global _lib
    if _lib is not None:
        return _lib
    if "FASTMAX_LIB" not in os.environ:
        import shutil
        if shutil.which("nvcc") or not os.path.exists(LIB_PATH):
            from . import build as _build
            _build.build()
    if not os.path.exists(LIB_PATH):
        raise FastmaxError(f"{LIB_PATH} is missing: run `python -m fastmax_experiments_b200.build` "
                           "(there is no CPU/PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.fastmax_abi_version.restype = C.c_int
    if lib.fastmax_abi_version() != ABI_VERSION:
        raise FastmaxError(f"{LIB_PATH} has ABI version {lib.fastmax_abi_version()}, this binding expects "
                           f"{ABI_VERSION}: rebuild with `python -m fastmax_experiments_b200.build --force`")
    vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    lib.fastmax_abi_version.restype = i32
    lib.fastmax_error_string.restype = C.c_char_p
    lib.fastmax_error_string.argtypes = [i32]
    lib.fastmax_last_cuda_error.restype = C.c_char_p
    lib.fastmax_last_launch_count.restype = i32
    lib.fastmax_plan_make.restype = i32
    lib.fastmax_plan_make.argtypes = [C.POINTER(Plan)] + [i32] * 8
    for nm in ("fastmax_fwd_workspace_bytes", "fastmax_bwd_workspace_bytes"):
        getattr(lib, nm).restype = sz
        getattr(lib, nm).argtypes = [i32] * 8
    lib.fastmax_fwd.restype = i32
    lib.fastmax_fwd.argtypes = [vp] * 5 + [i32] * 6 + [f32, i32, i32, vp, sz, vp]
    lib.fastmax_bwd.restype = i32
    lib.fastmax_bwd.argtypes = [vp] * 9 + [i32] * 6 + [f32, i32, i32, vp, sz, vp]
    lib.fastmax_saved_state_bytes.restype = sz
    lib.fastmax_saved_state_bytes.argtypes = [i32] * 8
    lib.fastmax_fwd_ex.restype = i32
    lib.fastmax_fwd_ex.argtypes = [vp] * 6 + [i32] * 6 + [f32, i32, i32, vp, sz, vp]
    lib.fastmax_bwd_ex.restype = i32
    lib.fastmax_bwd_ex.argtypes = [vp] * 10 + [i32] * 6 + [f32, i32, i32, vp, sz, vp]
    ll = C.c_longlong
    lib.fastmax_sp_state_bytes.restype = sz
    lib.fastmax_sp_state_bytes.argtypes = [i32, i32]
    lib.fastmax_sp_local_state_workspace_bytes.restype = sz
    lib.fastmax_sp_local_state_workspace_bytes.argtypes = [i32] * 4
    lib.fastmax_sp_local_state.restype = i32
    lib.fastmax_sp_local_state.argtypes = [vp] * 3 + [i32] * 5 + [vp, sz, vp]
    lib.fastmax_sp_local_state_rev.restype = i32
    lib.fastmax_sp_local_state_rev.argtypes = [vp] * 5 + [i32] * 5 + [vp, sz, vp]
    lib.fastmax_sp_reduce.restype = i32
    lib.fastmax_sp_reduce.argtypes = [vp, i32, i32, i32, vp, i32, i32, vp]
    lib.fastmax_fwd_sp.restype = i32
    lib.fastmax_fwd_sp.argtypes = [vp] * 6 + [ll] + [i32] * 5 + [f32, i32, vp, sz, vp]
    lib.fastmax_bwd_sp.restype = i32
    lib.fastmax_bwd_sp.argtypes = [vp] * 11 + [ll] + [i32] * 5 + [f32, i32, vp, sz, vp]
    lib.fastmax_decode_state_bytes.restype = sz
    lib.fastmax_decode_state_bytes.argtypes = [i32] * 4
    lib.fastmax_decode_init.restype = i32
    lib.fastmax_decode_init.argtypes = [vp] + [i32] * 4 + [vp]
    lib.fastmax_decode_update.restype = i32
    lib.fastmax_decode_update.argtypes = [vp] * 3 + [i32] * 6 + [vp]
    lib.fastmax_decode_attend.restype = i32
    lib.fastmax_decode_attend.argtypes = [vp] * 3 + [i32] * 5 + [f32, f32, i32, vp]
    lib.fastmax_gqa_workspace_bytes.restype = sz
    lib.fastmax_gqa_workspace_bytes.argtypes = [i32] * 9
    lib.fastmax_gqa_saved_state_bytes.restype = sz
    lib.fastmax_gqa_saved_state_bytes.argtypes = [i32] * 8
    lib.fastmax_fwd_gqa.restype = i32
    lib.fastmax_fwd_gqa.argtypes = [vp] * 6 + [i32] * 7 + [f32, i32, vp, sz, vp]
    lib.fastmax_bwd_gqa.restype = i32
    lib.fastmax_bwd_gqa.argtypes = [vp] * 10 + [i32] * 7 + [f32, i32, vp, sz, vp]
    lib.fastmax_qkv_prologue.restype = i32
    lib.fastmax_qkv_prologue.argtypes = [vp] * 6 + [i32] * 8 + [vp]
    lib.fastmax_qkv_prologue_bwd.restype = i32
    lib.fastmax_qkv_prologue_bwd.argtypes = [vp] * 6 + [i32] * 8 + [vp]
    lib.fastmax_profile_enable.restype = None
    lib.fastmax_profile_enable.argtypes = [i32]
    lib.fastmax_profile_fetch.restype = i32
    lib.fastmax_profile_fetch.argtypes = [C.c_char_p, sz, C.POINTER(C.c_float), i32]
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == OK:
        return
    lib = load()
    msg = lib.fastmax_error_string(rc).decode()
    if rc == ERR_CUDA:
        msg += ": " + lib.fastmax_last_cuda_error().decode()
    if rc == ERR_BAD_ARG:
        raise ValueError(f"{what}: {msg}")
    raise FastmaxError(f"{what}: {msg}")


def plan(B, H, N, D, p, causal, dtype, algo=ALGO_AUTO) -> Plan:
    pl = Plan()
    check(load().fastmax_plan_make(C.byref(pl), B, H, N, D, p, int(causal), dtype, algo), "fastmax_plan_make")
    return pl


def profile_enable(on: bool) -> None:
    load().fastmax_profile_enable(int(on))


def profile_fetch(cap: int = 4096) -> list[tuple[str, float]]:
    """[(kernel name, milliseconds)] for every launch since profile_enable(True)."""
    names = C.create_string_buffer(cap * 48)
    ms = (C.c_float * cap)()
    n = load().fastmax_profile_fetch(names, len(names), ms, cap)
    check(min(n, 0), "fastmax_profile_fetch")
    nm = names.value.decode().split("\n")
    return [(nm[i] if i < len(nm) else "?", float(ms[i])) for i in range(n)]

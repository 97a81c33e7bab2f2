"""pytest configuration: the ``gpu`` marker and shared fixture helpers."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _all_golden():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def _is_rect(n):
    return "decode" in n or "prefill" in n


def golden_names():
    """Square fixtures (q, k, v of one shape; forward + backward)."""
    return [n for n in _all_golden() if not _is_rect(n) and not n.startswith(("linearmax", "model_", "gqa_"))]


def golden_rect_names():
    """Inference fixtures (N_q != N_k, non-causal, forward only)."""
    return [n for n in _all_golden() if _is_rect(n) and not n.startswith("linearmax")]


def golden_linearmax_names():
    """linearmax fixtures (fastmax_hack.py: forward + torch-autograd gradients)."""
    return [n for n in _all_golden() if n.startswith("linearmax") and not _is_rect(n)]


def golden_linearmax_rect_names():
    """linearmax over a KV cache (N_q != N_k, forward only; count term N_k, fastmax_hack.py:21)."""
    return [n for n in _all_golden() if n.startswith("linearmax") and _is_rect(n)]


def golden_gqa_names():
    """q/k/v assembly + grouped-query attention module of the reference (oracle/make_golden.py::main_gqa)."""
    return [n for n in _all_golden() if n.startswith("gqa_")]


def load_golden_gqa(name: str, tag: str) -> dict:
    """tag: "f32" or "bf16" (the reference module run in that dtype).  Tensors come back in that dtype; dk, dv (sums
    over the query group, taken in float64) as float32."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    dt = torch.float32 if tag == "f32" else torch.bfloat16
    dec = (lambda a: torch.from_numpy(a)) if tag == "f32" else (lambda a: _from_bf16_bits(a).to(torch.bfloat16))
    out = dict(name=name, dtype=dt, **{k: int(z[k]) for k in ("T", "seed", "n_embd", "n_head", "n_query_groups",
                                                                  "head_size", "rope_n_elem")}, rot=float(z["rot"]))
    for key in ("qkv", "q", "k", "v", "y", "dqkv", "dq", "cos", "sin"):
        out[key] = dec(z[f"{tag}_{key}"])
    for key in ("dk", "dv"):
        out[key] = torch.from_numpy(z[f"{tag}_{key}"])
    g = torch.Generator().manual_seed(out["seed"] + 1)       # as in the minting script: x, then dy
    out["x"] = torch.randn(1, out["T"], out["n_embd"], generator=g).to(dt)
    out["dy"] = torch.randn(out["y"].shape, generator=g).to(dt)
    return out


def golden_model_names():
    return [n for n in _all_golden() if n.startswith("model_")]


def load_golden_model(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    return dict(name=name, model=str(d["model"]), T=int(d["T"]), seed=int(d["seed"]), idx=torch.from_numpy(d["idx"]),
                loss=float(d["loss"]), lse=torch.from_numpy(d["lse"]), logits_sample=torch.from_numpy(d["logits_sample"]),
                grad_norm=float(d["grad_norm"]), wte_grad_sample=torch.from_numpy(d["wte_grad_sample"]))


def load_golden_linearmax(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    out = dict(name=name, mask=bool(d["mask"]))
    for key in ("q", "k", "v", "do"):
        out[key] = _from_bf16_bits(d[key + "_bf16"])
    for key in ("o", "dq", "dk", "dv"):
        out[key] = torch.from_numpy(d[key]).double()
    return out


def load_golden_rect(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    out = dict(name=name, p=int(d["p"]), Nq=int(d["Nq"]), Nk=int(d["Nk"]), D=int(d["D"]))
    for key in ("q", "k", "v"):
        out[key] = _from_bf16_bits(d[key + "_bf16"])
    out["o"] = torch.from_numpy(d["o"]).double()
    return out


def _from_bf16_bits(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)


def load_golden(name: str) -> dict:
    """Fixture minted by oracle/make_golden.py from the unmodified reference."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    out = dict(
        name=name, B=int(d["B"]), H=int(d["H"]), N=int(d["N"]), D=int(d["D"]), p=int(d["p"]),
        mask=bool(d["mask"]), normalize_term=float(d["normalize_term"]),
        tensors_normalized=bool(d["tensors_normalized"]),
    )
    for key in ("q", "k", "v", "do"):
        out[key] = _from_bf16_bits(d[key + "_bf16"])          # bf16, exact
    for key in ("o", "g", "dq", "dk", "dv"):
        out[key] = torch.from_numpy(d[key]).double()
    for key in d:
        if key.startswith("ref32_"):
            out[key] = torch.from_numpy(d[key])
    return out

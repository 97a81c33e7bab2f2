"""TEST / BENCH INFRASTRUCTURE -- installs and imports the UNMODIFIED reference (ryansynk/fastmax-experiments).

Only tests/, bench.py's CPU legs and __graft_entry__ may import this module; the product never does.

* ``install()`` puts the reference into the git-ignored ``baseline/_ref/`` (it travels to the GPU box with
  gpurun, ``/root/reference`` does not):  ``pip install --no-index --no-build-isolation --no-deps --target
  baseline/_ref <copy of /root/reference>`` installs the ``lit_gpt`` package; ``attention_mechanisms/`` has no
  ``__init__.py`` so ``find_packages()`` skips it, and its three files are placed next to it as the namespace
  directory the reference's own ``from attention_mechanisms.fastmax import fastmax`` (lit_gpt/model.py:24)
  expects.  Nothing is edited.
* ``reference_op()`` returns the reference's ``fastmax`` (attention_mechanisms/fastmax.py:7) and
  ``fastmax_hack`` (fastmax_hack.py:5).
* ``reference_model()`` imports the unmodified ``lit_gpt`` with the stub modules of SURVEY.md Appendix B
  standing in for the four packages this image does not have (lightning, lightning_utilities,
  fast_transformers, fastmax_cuda) and returns the module.
"""
from __future__ import annotations

import contextlib
import os
import shutil
import subprocess
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference"
REF_DST = os.path.join(ROOT, "baseline", "_ref")


def installed() -> bool:
    return (os.path.exists(os.path.join(REF_DST, "attention_mechanisms", "fastmax.py"))
            and os.path.exists(os.path.join(REF_DST, "lit_gpt", "model.py")))


def install(force: bool = False) -> str | None:
    """Install the reference into baseline/_ref (no-op when it is there or the source tree is absent).
    Returns the install directory or None."""
    if installed() and not force:
        return REF_DST
    if not os.path.isdir(REF_SRC):
        return None
    os.makedirs(REF_DST, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")      # pip builds in the source tree; /root/reference is read-only
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns(".git", "notebooks", "tutorials"))
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                            "--find-links", "/opt/wheelhouse", "--upgrade", "--target", REF_DST, src],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    dst = os.path.join(REF_DST, "attention_mechanisms")
    os.makedirs(dst, exist_ok=True)
    for f in os.listdir(os.path.join(REF_SRC, "attention_mechanisms")):
        if f.endswith(".py"):
            shutil.copy2(os.path.join(REF_SRC, "attention_mechanisms", f), os.path.join(dst, f))
    return REF_DST


def _root() -> str:
    if installed():
        return REF_DST
    if os.path.isdir(REF_SRC):
        return REF_SRC
    raise ImportError("the reference is neither installed under baseline/_ref nor present at /root/reference")


def _ensure_path() -> str:
    root = _root()
    if root not in sys.path:
        sys.path.insert(0, root)
    return root


def reference_op():
    """(fastmax, fastmax_hack) of the unmodified reference."""
    _ensure_path()
    from attention_mechanisms.fastmax import fastmax            # noqa: E402
    from attention_mechanisms.fastmax_hack import fastmax_hack  # noqa: E402
    return fastmax, fastmax_hack


def _stub(name: str, **attrs) -> types.ModuleType:
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []          # behaves as a package for `import a.b`
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def reference_model():
    """The unmodified ``lit_gpt`` module (GPT, Config, ...), stubs for the absent third-party packages."""
    _ensure_path()
    if "lit_gpt" in sys.modules and hasattr(sys.modules["lit_gpt"], "GPT"):
        return sys.modules["lit_gpt"]

    class _Truthy:
        def __init__(self, *a, **k):
            pass

        def __bool__(self):
            return True

    class _FSDPStrategy:      # only used in an isinstance() test (lit_gpt/utils.py:301)
        pass

    def _missing(*a, **k):
        raise NotImplementedError("stub of a package this image does not have")

    _stub("fastmax_cuda")                                                     # lit_gpt/model.py:27
    _stub("fast_transformers")
    _stub("fast_transformers.causal_product", CausalDotProduct=_missing)      # lit_gpt/model.py:21
    _stub("lightning", Fabric=_Truthy)                                        # lit_gpt/utils.py:12, :300 (annotation)
    _stub("lightning.fabric")
    _stub("lightning.fabric.strategies", FSDPStrategy=_FSDPStrategy)          # lit_gpt/utils.py:16
    _stub("lightning.fabric.utilities")
    _stub("lightning.fabric.utilities.load", _lazy_load=_missing)             # lit_gpt/utils.py:17
    _stub("lightning_utilities")
    _stub("lightning_utilities.core")
    _stub("lightning_utilities.core.imports", RequirementCache=_Truthy)       # lit_gpt/__init__.py:10
    import lit_gpt                                                            # noqa: E402
    return lit_gpt


@contextlib.contextmanager
def cuda_hop_neutralised():
    """On a host without CUDA the reference's ``o.cuda()`` (lit_gpt/model.py:486) must be an identity."""
    import torch
    if torch.cuda.is_available():
        yield
        return
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def build_gpt(name: str, attn_alg: str = "fastmax", block_size: int | None = None, seed: int = 0, lora: dict | None = None,
              device=None, dtype=None, **overrides):
    """Random-init reference model ``GPT(Config.from_name(name, attn_alg=...))`` with ``_init_weights``
    (lit_gpt/model.py:245-252).  ``lora``: kwargs of lit_gpt.lora.Config (r, alpha, to_query, ...) to build the
    LoRA variant (finetune/lora.py:40-48) instead."""
    import torch
    lit_gpt = reference_model()
    kw = dict(attn_alg=attn_alg, **overrides)
    if block_size is not None:
        kw["block_size"] = block_size
    torch.manual_seed(seed)
    old_dtype = torch.get_default_dtype()
    if dtype is not None:
        torch.set_default_dtype(dtype)
    try:
        with (torch.device(device) if device is not None else contextlib.nullcontext()):
            if lora:
                from lit_gpt import lora as lit_lora
                cfg = lit_lora.Config.from_name(name, **kw, **lora)
                model = lit_lora.GPT(cfg)
            else:
                cfg = lit_gpt.Config.from_name(name, **kw)
                model = lit_gpt.GPT(cfg)
            model.apply(model._init_weights)
    finally:
        torch.set_default_dtype(old_dtype)
    return model, cfg

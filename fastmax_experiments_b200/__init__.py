"""fastmax_experiments_b200 -- B200-native FastMax attention (one hot path)."""
from .fastmax import fastmax, fastattention_einops, FastmaxFunction  # noqa: F401
from .attention import fastmax_attention, linearmax_attention, patch_causal_self_attention  # noqa: F401
from .linearmax import linearmax  # noqa: F401
from .decode import FastmaxState  # noqa: F401
from .gqa import fastmax_gqa, qkv_prologue  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["fastmax", "fastattention_einops", "FastmaxFunction", "fastmax_attention", "linearmax",
           "linearmax_attention", "patch_causal_self_attention", "FastmaxState", "fastmax_gqa", "qkv_prologue"]

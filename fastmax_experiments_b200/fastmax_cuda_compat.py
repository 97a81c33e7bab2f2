"""Literal stand-in for the un-vendored ``fastmax_cuda`` extension the reference imports
(lit_gpt/model.py:27) and calls as

    o                  = fastmax_cuda.forwardpass(q, k, v, drop_noise, rpe_matrix, mask, dropout, normalize, temperature)   # :82-84
    gradq, gradk, gradv = fastmax_cuda.backwardpass(q, k, v, o, grad_output, mask)                                             # :116

with fp32 tensors in ``(n, d, b*h)`` layout (model.py:73-81).  The returned ``o`` carries one extra
row in dim 1 (the normaliser ``g``); the caller slices ``o[:, :d, :]`` (model.py:89-91) and hands the
unsliced tensor back to ``backwardpass`` (model.py:85,116).

To use:  ``sys.modules['fastmax_cuda'] = fastmax_experiments_b200.fastmax_cuda_compat``.
Only what ``fastmax.py`` defines is computed: ``dropout`` must be 0, ``normalize`` False,
``temperature`` 1 and ``rpe_matrix`` all zeros / None -- those knobs have no counterpart in the
in-tree algorithm and no oracle; anything else raises.
"""
from __future__ import annotations

import math

import torch

from .fastmax import fastmax_bwd_raw, fastmax_fwd_raw

P_DEFAULT = 2   # a0, a1, a2 = 1, 1, 0.5 in the reference's call site (model.py:499-502)


def _check_extras(rpe_matrix, dropout, normalize, temperature):
    if dropout not in (0, 0.0) or normalize or temperature != 1.0:
        raise NotImplementedError("fastmax_cuda_compat: dropout / normalize / temperature are not defined by "
                                  "attention_mechanisms/fastmax.py and are not provided")
    if rpe_matrix is not None and bool((rpe_matrix != 0).any()):
        raise NotImplementedError("fastmax_cuda_compat: relative positional encoding is not provided")


def _to_bhnd(x):      # (n, d, b*h) -> (1, b*h, n, d) contiguous
    return x.permute(2, 0, 1).unsqueeze(0).contiguous()


def _from_bhnd(x):    # (1, b*h, n, d) -> (n, d, b*h)
    return x.squeeze(0).permute(1, 2, 0).contiguous()


def forwardpass(q, k, v, drop_noise=None, rpe_matrix=None, mask=False, dropout=0.0, normalize=False,
                temperature=1.0):
    _check_extras(rpe_matrix, dropout, normalize, temperature)
    n, d, _ = q.shape
    inv_c = 1.0 / (8 * math.sqrt(d))                      # fastmax.py default normalize_term=8
    o, g = fastmax_fwd_raw(_to_bhnd(q), _to_bhnd(k), _to_bhnd(v), bool(mask), inv_c, P_DEFAULT)
    out = torch.empty(n, d + 1, q.shape[2], dtype=q.dtype, device=q.device)
    out[:, :d, :] = _from_bhnd(o)
    out[:, d, :] = g.squeeze(0).transpose(0, 1).to(q.dtype)
    return out


def backwardpass(q, k, v, o, grad_output, mask=False):
    n, d, _ = q.shape
    inv_c = 1.0 / (8 * math.sqrt(d))
    o4 = _to_bhnd(o[:, :d, :])
    g = o[:, d, :].transpose(0, 1).unsqueeze(0).float().contiguous()
    dq, dk, dv = fastmax_bwd_raw(_to_bhnd(q), _to_bhnd(k), _to_bhnd(v), o4, g, _to_bhnd(grad_output),
                                 bool(mask), inv_c, P_DEFAULT)
    return _from_bhnd(dq), _from_bhnd(dk), _from_bhnd(dv)

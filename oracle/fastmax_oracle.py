"""CPU oracle for FastMax attention -- TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement (torch CPU tensors, any float dtype,
float64 recommended) of the algorithm in the reference's
``attention_mechanisms/fastmax.py``.  It is the checker for the CUDA path: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it.  The product package
(``fastmax_experiments_b200``) never imports anything from ``oracle/``.

Parity pin: the reference ships no golden vectors / known-answer tests for this
path, so the pin is the reference itself: the fixtures in
``tests/golden/*.npz`` were produced by importing the unmodified reference here
(``oracle/make_golden.py``) and ``tests/test_oracle.py`` checks every function
in this file against every fixture.

Two independent restatements are provided:

* ``factorised_*``  -- the running-moment algorithm, same contraction/cumsum
  structure as the reference (one function per reference function, each citing
  the reference lines it follows).  O(N * D^(p+1)) work, O(N * D^(p+1)) memory.
  This is also what ``bench.py`` times as the CPU baseline ("port").
* ``dense_*`` -- the explicit score-matrix formula
  ``o = (f(QK^T/c) o tril) V / rowsum``, evaluated block-row by block-row so it
  runs at sizes where the factorised form runs out of memory.  O(N^2 D) work.

Conventions (reference fastmax.py:78-82): ``c = 1`` if ``tensors_normalized is
True`` else ``normalize_term * sqrt(D)``;  ``f(s) = 1 + s`` (p=1) or
``1 + s + s^2/2`` (p=2);  ``mask is False`` (identity test, fastmax.py:85,132)
selects the non-causal branch.
"""
from __future__ import annotations

import math

import torch

__all__ = [
    "denominator_c",
    "factorised_forward",
    "factorised_backward",
    "factorised_fwd_bwd",
    "dense_forward",
    "dense_fwd_bwd",
    "rel_err",
]


def denominator_c(D: int, normalize_term=8, tensors_normalized=False) -> float:
    """Score divisor ``c`` (reference fastmax.py:78-82)."""
    if tensors_normalized is True:
        return 1.0
    return float(normalize_term) * math.sqrt(D)


def _check_p(p):
    # reference raises only in backward / compute_attn (fastmax.py:362,428,...)
    if p not in (1, 2):
        raise ValueError(f"p should be 1 or 2, got p={p}")


def _rcumsum(x: torch.Tensor, dim: int) -> torch.Tensor:
    """Reverse (suffix) cumulative sum; the reference does this with an index
    gather around ``cumsum`` (fastmax.py:565-567, 669-671)."""
    return torch.flip(torch.cumsum(torch.flip(x, (dim,)), dim), (dim,))


# --------------------------------------------------------------------------
# factorised restatement
# --------------------------------------------------------------------------

def _numerator(q, k, v, c, p, causal):
    """F_i = sum_j f(s_ij) v_j via moments.
    causal: compute_F_masked (fastmax.py:236-250); else compute_F_unmasked
    (fastmax.py:202-216)."""
    if causal:
        F = torch.cumsum(v, 2)                                              # :236
        m1 = torch.cumsum(torch.einsum("bhnm,bhnj->bhnmj", k, v), 2)       # :239-240
        F = F + torch.einsum("bhim,bhimj->bhij", q, m1) / c                # :241
        if p == 2:
            m2 = torch.cumsum(torch.einsum("bhnm,bhnl,bhnj->bhnmlj", k, k, v), 2)   # :244-245
            F = F + torch.einsum("bhim,bhil,bhimlj->bhij", q, q, m2) / (2 * c * c)  # :246-248
    else:
        N = q.shape[2]
        F = v.sum(2, keepdim=True).expand(-1, -1, N, -1)                    # :204-205
        m1 = torch.einsum("bhnm,bhnj->bhmj", k, v)                          # :207
        F = F + torch.einsum("bhim,bhmj->bhij", q, m1) / c                  # :208
        if p == 2:
            m2 = torch.einsum("bhnm,bhnl,bhnj->bhmlj", k, k, v)             # :211
            F = F + torch.einsum("bhim,bhil,bhmlj->bhij", q, q, m2) / (2 * c * c)   # :212-214
    return F


def _normaliser(q, k, c, p, causal):
    """g_i = sum_j f(s_ij).
    causal: compute_g_masked (fastmax.py:304-322) -- count term is arange(N)+1;
    else compute_g_unmasked (fastmax.py:269-285) -- count term is N of *q*."""
    B, H, N, _ = q.shape
    if causal:
        g = (torch.arange(N, device=q.device) + 1).to(q.dtype).expand(B, H, N)     # :306-308
        g = g + torch.einsum("bhim,bhim->bhi", q, torch.cumsum(k, 2)) / c           # :310-312
        if p == 2:
            kk = torch.cumsum(torch.einsum("bhnm,bhnl->bhnml", k, k), 2)            # :315-316
            g = g + torch.einsum("bhim,bhil,bhiml->bhi", q, q, kk) / (2 * c * c)    # :317-320
    else:
        g = torch.full((B, H, N), float(N), dtype=q.dtype, device=q.device)        # :271
        g = g + torch.einsum("bhil,bhl->bhi", q, k.sum(2)) / c                      # :274-276
        if p == 2:
            kk = torch.einsum("bhml,bhmp->bhlp", k, k)                              # :279
            g = g + torch.einsum("bhil,bhip,bhlp->bhi", q, q, kk) / (2 * c * c)     # :280-283
    return g


def factorised_forward(q, k, v, mask=True, normalize_term=8, tensors_normalized=False, p=1):
    """o, g of fastattention_einops.forward (fastmax.py:42-111), factorised
    branch only (create_attn=False).  As in the reference, ``p`` other than 2
    computes the p=1 polynomial in forward (fastmax.py:210,243)."""
    D = q.shape[-1]
    c = denominator_c(D, normalize_term, tensors_normalized)
    causal = mask is not False
    F = _numerator(q, k, v, c, p, causal)
    g = _normaliser(q, k, c, p, causal)
    return F / g.unsqueeze(-1), g                                           # :97


def _grad_q(q, k, v, o, g, do, c, p, causal):
    """gradient_o_q_masked (fastmax.py:452-485) / _unmasked (:403-430).
    dq_i = [z1 - z2 + z3 - z4] / (c g_i)."""
    dodo = torch.einsum("bhpj,bhpj->bhp", do, o)                            # do_i . o_i
    if causal:
        vk = torch.cumsum(torch.einsum("bhlj,bhlr->bhljr", v, k), 2)        # :453-454
        z1 = torch.einsum("bhpj,bhpjr->bhpr", do, vk)                       # :455
        z2 = dodo.unsqueeze(-1) * torch.cumsum(k, 2)                        # :457-458
        acc = z1 - z2
        if p == 2:
            kvk = torch.cumsum(torch.einsum("bhlm,bhlj,bhlr->bhlmjr", k, v, k), 2)  # :463-464
            z3 = torch.einsum("bhpm,bhpj,bhpmjr->bhpr", q, do, kvk) / c     # :465-468
            kk = torch.cumsum(torch.einsum("bhlr,bhlm->bhlrm", k, k), 2)    # :470-471
            z4 = dodo.unsqueeze(-1) * torch.einsum("bhpm,bhprm->bhpr", q, kk) / c   # :472-477
            acc = acc + z3 - z4
    else:
        vk = torch.einsum("bhlj,bhlr->bhjr", v, k)                          # :404-405 (regrouped, linear in N)
        z1 = torch.einsum("bhpj,bhjr->bhpr", do, vk)
        z2 = dodo.unsqueeze(-1) * k.sum(2).unsqueeze(2)                     # :407-409
        acc = z1 - z2
        if p == 2:
            kvk = torch.einsum("bhlm,bhlj,bhlr->bhmjr", k, v, k)            # :414-415 (regrouped)
            z3 = torch.einsum("bhpm,bhpj,bhmjr->bhpr", q, do, kvk) / c      # :416
            kk = torch.einsum("bhlm,bhlr->bhmr", k, k)                      # :418
            z4 = dodo.unsqueeze(-1) * torch.einsum("bhpm,bhmr->bhpr", q, kk) / c    # :419-423
            acc = acc + z3 - z4
    return acc / (c * g.unsqueeze(-1))                                      # :424,479


def _grad_k(q, k, v, o, g, do, c, p, causal):
    """gradient_o_k_masked (fastmax.py:561-604) / _unmasked (:507-539).
    dk_j = [z1 - z2 + z3 - z4] / c with suffix sums over i >= j."""
    u = do / g.unsqueeze(-1)                                                # do_i / g_i
    du = torch.einsum("bhij,bhij->bhi", u, o)                               # (do_i . o_i) / g_i
    if causal:
        uq = _rcumsum(torch.einsum("bhij,bhir->bhijr", u, q), 2)            # :566-567
        z1 = torch.einsum("bhpj,bhpjr->bhpr", v, uq)                        # :568
        z2 = _rcumsum(du.unsqueeze(-1) * q, 2)                              # :570-574
        acc = z1 - z2
        if p == 2:
            uqq = _rcumsum(torch.einsum("bhij,bhim,bhir->bhijmr", u, q, q), 2)      # :577-580
            z3 = torch.einsum("bhpm,bhpj,bhpjmr->bhpr", k, v, uqq) / c      # :581-584
            dqq = _rcumsum(torch.einsum("bhi,bhim,bhir->bhimr", du, q, q), 2)       # :586-594
            z4 = torch.einsum("bhpm,bhpmr->bhpr", k, dqq) / c               # :595-597
            acc = acc + z3 - z4
    else:
        uq = torch.einsum("bhij,bhir->bhjr", u, q)                          # :509
        z1 = torch.einsum("bhpj,bhjr->bhpr", v, uq)                         # :510
        z2 = torch.einsum("bhi,bhir->bhr", du, q).unsqueeze(2)              # :512-514
        acc = z1 - z2
        if p == 2:
            uqq = torch.einsum("bhij,bhim,bhir->bhjmr", u, q, q)            # :517-519
            z3 = torch.einsum("bhpm,bhpj,bhjmr->bhpr", k, v, uqq) / c       # :520-523
            dqq = torch.einsum("bhi,bhim,bhir->bhmr", du, q, q)             # :525-531 (regrouped, linear in N)
            z4 = torch.einsum("bhpm,bhmr->bhpr", k, dqq) / c
            acc = acc + z3 - z4
    return acc / c                                                          # :533,598


def _grad_v(q, k, g, do, c, p, causal):
    """gradient_o_v_masked (fastmax.py:665-691) / _unmasked (:624-645).
    dv_j = sum_{i>=j} f(s_ij) do_i / g_i."""
    u = do / g.unsqueeze(-1)                                                # :670
    if causal:
        acc = _rcumsum(u, 2)                                                # :671
        uq = _rcumsum(torch.einsum("bhir,bhim->bhirm", u, q), 2)            # :673-674
        acc = acc + torch.einsum("bhpm,bhprm->bhpr", k, uq) / c             # :675
        if p == 2:
            uqq = _rcumsum(torch.einsum("bhir,bhim,bhil->bhirml", u, q, q), 2)      # :678-681
            acc = acc + torch.einsum("bhpm,bhpl,bhprml->bhpr", k, k, uqq) / (2 * c * c)  # :682-684
    else:
        acc = u.sum(2, keepdim=True).expand(-1, -1, k.shape[2], -1)        # :626-627
        uq = torch.einsum("bhir,bhim->bhrm", u, q)                          # :629
        acc = acc + torch.einsum("bhpm,bhrm->bhpr", k, uq) / c              # :630
        if p == 2:
            uqq = torch.einsum("bhir,bhim,bhil->bhrml", u, q, q)            # :633-635
            acc = acc + torch.einsum("bhpm,bhpl,bhrml->bhpr", k, k, uqq) / (2 * c * c)   # :636-638
    return acc


def factorised_backward(q, k, v, o, g, do, mask=True, normalize_term=8,
                        tensors_normalized=False, p=1):
    """dq, dk, dv of fastattention_einops.backward (fastmax.py:114-182)."""
    _check_p(p)
    c = denominator_c(q.shape[-1], normalize_term, tensors_normalized)
    causal = mask is not False
    dq = _grad_q(q, k, v, o, g, do, c, p, causal)
    dk = _grad_k(q, k, v, o, g, do, c, p, causal)
    dv = _grad_v(q, k, g, do, c, p, causal)
    return dq, dk, dv


def factorised_fwd_bwd(q, k, v, do, mask=True, normalize_term=8, tensors_normalized=False, p=1):
    o, g = factorised_forward(q, k, v, mask, normalize_term, tensors_normalized, p)
    dq, dk, dv = factorised_backward(q, k, v, o, g, do, mask, normalize_term, tensors_normalized, p)
    return o, g, dq, dk, dv


# --------------------------------------------------------------------------
# dense restatement (independent second oracle; SURVEY.md section 4 verified it
# equals the reference to 3e-15 in fp64)
# --------------------------------------------------------------------------

def _poly(s, p):
    return 1 + s + 0.5 * s * s if p == 2 else 1 + s


def _dpoly(s, p):
    return 1 + s if p == 2 else torch.ones_like(s)


def dense_forward(q, k, v, mask=True, normalize_term=8, tensors_normalized=False, p=1,
                  block=1024):
    """o_i = sum_j f(q_i.k_j/c) [j<=i] v_j / g_i evaluated in row blocks.
    Equals compute_attn + einsum of the reference (fastmax.py:356-381,103)."""
    _check_p(p)
    B, H, N, D = q.shape
    c = denominator_c(D, normalize_term, tensors_normalized)
    causal = mask is not False
    o = torch.empty_like(q)
    g = torch.empty(B, H, N, dtype=q.dtype)
    Nk = k.shape[2]
    col = torch.arange(Nk)
    for r0 in range(0, N, block):
        r1 = min(N, r0 + block)
        kk = k[:, :, : (r1 if causal else Nk)]
        s = torch.einsum("bhid,bhjd->bhij", q[:, :, r0:r1], kk) / c
        w = _poly(s, p)
        if causal:
            keep = col[: r1][None, :] <= torch.arange(r0, r1)[:, None]
            w = w * keep
        g[:, :, r0:r1] = w.sum(-1)
        o[:, :, r0:r1] = torch.einsum("bhij,bhjd->bhid", w, v[:, :, : kk.shape[2]]) / g[:, :, r0:r1, None]
    return o, g


def dense_fwd_bwd(q, k, v, do, mask=True, normalize_term=8, tensors_normalized=False, p=1,
                  block=1024):
    """Forward plus analytic gradients of the dense formula (row blocks):
    with u_i = do_i/g_i, d_i = (do_i.o_i)/g_i and e_ij = f'(s_ij)(u_i.v_j - d_i)[j<=i]:
    dv_j = sum_i f(s_ij) u_i,  dq_i = sum_j e_ij k_j / c,  dk_j = sum_i e_ij q_i / c."""
    _check_p(p)
    B, H, N, D = q.shape
    c = denominator_c(D, normalize_term, tensors_normalized)
    causal = mask is not False
    o, g = dense_forward(q, k, v, mask, normalize_term, tensors_normalized, p, block)
    u = do / g.unsqueeze(-1)
    d = (do * o).sum(-1) / g
    dq = torch.zeros_like(q)
    dk = torch.zeros_like(k)
    dv = torch.zeros_like(v)
    Nk = k.shape[2]
    col = torch.arange(Nk)
    for r0 in range(0, N, block):
        r1 = min(N, r0 + block)
        n1 = r1 if causal else Nk
        s = torch.einsum("bhid,bhjd->bhij", q[:, :, r0:r1], k[:, :, :n1]) / c
        w = _poly(s, p)
        e = _dpoly(s, p) * (torch.einsum("bhid,bhjd->bhij", u[:, :, r0:r1], v[:, :, :n1])
                            - d[:, :, r0:r1, None])
        if causal:
            keep = col[:n1][None, :] <= torch.arange(r0, r1)[:, None]
            w = w * keep
            e = e * keep
        dv[:, :, :n1] += torch.einsum("bhij,bhid->bhjd", w, u[:, :, r0:r1])
        dq[:, :, r0:r1] = torch.einsum("bhij,bhjd->bhid", e, k[:, :, :n1]) / c
        dk[:, :, :n1] += torch.einsum("bhij,bhid->bhjd", e, q[:, :, r0:r1]) / c
    return o, g, dq, dk, dv


def rel_err(x: torch.Tensor, ref: torch.Tensor, floor: float = 1e-300) -> float:
    """Tolerance protocol of SURVEY.md section 8(c): max|x-ref| / max|ref|.
    ``floor`` bounds the denominator for tensors that are identically ~0 in the reference
    (e.g. dq of a single-token causal sequence)."""
    x = x.detach().double().cpu()
    ref = ref.detach().double().cpu()
    return float((x - ref).abs().max() / ref.abs().max().clamp_min(floor))


def linearmax_prologue(q, k):
    """q, k normalisation of fastmax_hack.py (non-causal :9-15, causal :41-46, identical): subtract the
    per-token mean over features, then divide by the largest token norm of the (b,h) sequence."""
    q = q - q.mean(dim=3, keepdim=True)
    k = k - k.mean(dim=3, keepdim=True)
    qn = torch.linalg.norm(q, dim=3).amax(dim=2)[:, :, None, None]
    kn = torch.linalg.norm(k, dim=3).amax(dim=2)[:, :, None, None]
    return q / qn, k / kn


def linearmax_forward(q, k, v, mask=True):
    """fastmax_hack(q, k, v, p=1, mask) (fastmax_hack.py:5-60): the prologue above, then first-order
    FastMax with c = 1 (normalize_term = 1, :48; denum_term = 1, :9).  Differentiable with torch autograd,
    like the reference (which has no custom backward here)."""
    qn, kn = linearmax_prologue(q, k)
    if mask is False and q.shape[2] != k.shape[2]:
        # KV-cache shapes (N_q != N_k): the non-causal branch counts the KEYS, div1 = k.shape[2]
        # (fastmax_hack.py:21) -- unlike fastmax.py's compute_g_unmasked, whose count term is N_q
        num = qn @ (kn.transpose(-1, -2) @ v) + v.sum(-2, keepdim=True)            # fastmax_hack.py:18-19,24,27-28
        den = qn @ kn.sum(-2).unsqueeze(-1) + float(k.shape[2])                    # fastmax_hack.py:21-22,25,29-30
        return num / den
    o, _ = factorised_forward(qn, kn, v, mask, normalize_term=1, tensors_normalized=True, p=1)
    return o

"""Oracle (test infrastructure): GPT "Quartet" causal attention on CPU.

Restates ``CausalSelfAttention.forward`` (mop/models/quartet_attn_patch.py:75-127):
two score maps, each z-scored per row over the FULL unmasked row with the
unbiased standard deviation (:95-98), mixed as
``(1-m) n1 + m (n1*n2) quartet_scale`` with ``m = sigmoid(mixture)`` (:103-106),
then causal fill, optional additive mask, softmax, PV (:112-121).  With
``use_quartet=False`` the single map is still z-scored (:108-110).

``quartet_core`` = kernel boundary; ``quartet_module`` = module boundary pinned
against the imported reference.  ``quartet_core_manual`` is the hand-derived
forward/backward (centred form) that specifies the CUDA
kernels.  Not used by the product path.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


def _zscore(s, eps):
    mu = s.mean(dim=-1, keepdim=True)
    sd = s.std(dim=-1, keepdim=True)  # unbiased (correction=1)
    return (s - mu) / (sd + eps)


def quartet_core(q, k, v, q2=None, k2=None, mixture=None, quartet_scale=None, *,
                 eps: float = 1e-5, add_mask: Optional[torch.Tensor] = None,
                 return_probs: bool = False, drop_mask: Optional[torch.Tensor] = None):
    """q,k,v,(q2,k2): [B,H,T,dk].  ``q2 is None`` selects the non-quartet branch.

    The non-quartet branch uses a hard-coded 1e-5 (:110) - pass ``eps=1e-5``.
    """
    T, dk = q.shape[-2:]
    sc = 1.0 / math.sqrt(dk)
    n1 = _zscore((q @ k.transpose(-2, -1)) * sc, eps)
    if q2 is not None:
        n2 = _zscore((q2 @ k2.transpose(-2, -1)) * sc, eps)
        m = torch.sigmoid(mixture)
        scores = (1.0 - m) * n1 + m * (n1 * n2) * quartet_scale
    else:
        scores = n1
    keep = torch.tril(torch.ones(T, T, dtype=torch.bool, device=q.device))
    scores = scores.masked_fill(~keep, float("-inf"))
    if add_mask is not None:
        scores = scores + add_mask
    p = torch.softmax(scores, dim=-1)
    if drop_mask is not None:   # attn_drop(P) with an explicit mask of factors 0 | 1/(1-p)  (quartet_attn_patch.py:118-119)
        p = p * drop_mask
    y = p @ v
    return (y, p) if return_probs else y


def quartet_module(x, sd: Dict[str, torch.Tensor], n_head: int, use_quartet: bool = True,
                   eps: float = 1e-5, attention_mask=None):
    B, T, Cdim = x.shape
    dh = Cdim // n_head
    lin = lambda n: F.linear(x, sd[f"{n}.weight"], sd.get(f"{n}.bias")).view(B, T, n_head, dh).transpose(1, 2)
    q, k, v = lin("q_proj"), lin("k_proj"), lin("v_proj")
    if use_quartet:
        y = quartet_core(q, k, v, lin("q2_proj"), lin("k2_proj"), sd["mixture"], sd["quartet_scale"],
                         eps=eps, add_mask=attention_mask)
    else:
        y = quartet_core(q, k, v, eps=1e-5, add_mask=attention_mask)
    y = y.transpose(1, 2).contiguous().view(B, T, Cdim)
    return F.linear(y, sd["o_proj.weight"], sd.get("o_proj.bias"))


def quartet_core_manual(q, k, v, q2, k2, mixture, quartet_scale, *, eps=1e-5, add_mask=None, dy=None):
    """Hand-derived forward/backward in the centred form.

    c_ij = s q_i.(k_j - kbar) equals s_ij - mu_i exactly, so the z-score is
    ``c / (sigma + eps)`` with ``sigma_i^2 = sum_j c_ij^2 / (T-1)``.
    Backward (given dn, nonzero only for j <= i):
        g_i   = sum_j dn_ij c_ij / ((sigma_i+eps)^2 (T-1) sigma_i)
        dc_ij = dn_ij/(sigma_i+eps) - g_i c_ij              (dense in j)
        dq_i  = s sum_j dc_ij kc_j ;  dkc_j = s sum_i dc_ij q_i ;  dk = dkc - mean_j(dkc)
    The dense part is rank-structured:  sum_j c_ij kc_j = s (Kc^T Kc) q_i  and
    sum_i g_i c_ij q_i = s (sum_i g_i q_i q_i^T) kc_j.
    """
    B, H, T, dk = q.shape
    s = 1.0 / math.sqrt(dk)
    quart = q2 is not None

    def centred(qq, kk):
        kc = kk - kk.mean(dim=-2, keepdim=True)
        c = s * qq @ kc.transpose(-1, -2)
        sig = torch.sqrt((c * c).sum(-1, keepdim=True) / (T - 1))
        return kc, c, sig

    kc1, c1, sg1 = centred(q, k)
    n1 = c1 / (sg1 + eps)
    if quart:
        kc2, c2, sg2 = centred(q2, k2)
        n2 = c2 / (sg2 + eps)
        m = torch.sigmoid(mixture)
        gam = quartet_scale
        scores = (1 - m) * n1 + m * gam * n1 * n2
    else:
        scores = n1
    keep = torch.tril(torch.ones(T, T, dtype=torch.bool, device=q.device))
    scores = scores.masked_fill(~keep, float("-inf"))
    if add_mask is not None:
        scores = scores + add_mask
    P = torch.softmax(scores, -1)
    y = P @ v
    if dy is None:
        return y
    dV = P.transpose(-1, -2) @ dy
    dP = dy @ v.transpose(-1, -2)
    D = P * (dP - (dP * P).sum(-1, keepdim=True))  # zero above the diagonal
    grads = {"v": dV}
    if quart:
        dn1 = D * ((1 - m) + m * gam * n2)
        dn2 = D * (m * gam * n1)
        grads["mixture"] = (m * (1 - m) * (D * (-n1 + gam * n1 * n2)).sum()).reshape(1)
        grads["quartet_scale"] = ((D * (m * n1 * n2)).sum()).reshape(1)
    else:
        dn1 = D

    def back(dn, qq, kc, c, sig):
        g = (dn * c).sum(-1, keepdim=True) / ((sig + eps) ** 2 * (T - 1) * sig)
        W = dn / (sig + eps)                       # causal part
        gram = kc.transpose(-1, -2) @ kc           # [dk,dk]
        dq = s * W @ kc - s * s * g * (qq @ gram)
        Mg = (g * qq).transpose(-1, -2) @ qq       # sum_i g_i q_i q_i^T
        dkc = s * W.transpose(-1, -2) @ qq - s * s * (kc @ Mg)
        return dq, dkc - dkc.mean(dim=-2, keepdim=True)

    grads["q"], grads["k"] = back(dn1, q, kc1, c1, sg1)
    if quart:
        grads["q2"], grads["k2"] = back(dn2, q2, kc2, c2, sg2)
    return y, grads

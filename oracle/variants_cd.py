"""Oracle (test infrastructure): attention variants C and D on CPU.

Restates, op for op, ``CrossViewMixerMSA`` (mop/models/attention_variants.py:51-156) and ``MultiHopMSA`` (:163-231):

* C: four score maps S1 = q1 k1^T, S2 = q2 k2^T, S12 = q1 k2^T, S21 = q2 k1^T (all scaled by dk^-1/2) mixed by the 2x2
  parameter ``mix`` (:100-105), optional transpose cues t1 S1^T + t2 S2^T (:106-110), optional ``mask == 0 -> -inf``,
  softmax, optional per-key prior sharpening (:125-150), ``y = A v1``.
* D: ``Smix = S1 + and_ S2 + or_ (lse(S1,S2) - S1) - not_ beta S2 + chain log(A1 A2^(hops-1) + 1e-6)`` with fixed scalar gates
  (:207-217), ``y = softmax(Smix) v1 + sigmoid(chain_value_logit) A1 A2^(hops-1) v2`` (:221-228).

``*_core`` = kernel boundary ([B,H,N,dk] tensors); ``*_module`` = module boundary pinned against fixtures generated from the
imported reference (tests/golden/make_golden_cd.py).  Not used by the product path.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


def crossview_core(q1, k1, v1, q2, k2, mix, *, t1: float = 0.0, t2: float = 0.0, use_transpose_cues: bool = True,
                   zero_mask: Optional[torch.Tensor] = None):
    dk = q1.shape[-1]
    s = 1.0 / math.sqrt(dk)
    S1 = (q1 @ k1.transpose(-2, -1)) * s
    S2 = (q2 @ k2.transpose(-2, -1)) * s
    S12 = (q1 @ k2.transpose(-2, -1)) * s
    S21 = (q2 @ k1.transpose(-2, -1)) * s
    S = mix[0, 0] * S1 + mix[0, 1] * S12 + mix[1, 0] * S21 + mix[1, 1] * S2
    if use_transpose_cues:
        if t1 != 0.0:
            S = S + t1 * S1.transpose(-2, -1)
        if t2 != 0.0:
            S = S + t2 * S2.transpose(-2, -1)
    if zero_mask is not None:
        S = S.masked_fill(zero_mask == 0, float("-inf"))
    return F.softmax(S, dim=-1) @ v1


def crossview_module(x, sd: Dict[str, torch.Tensor], heads: int, *, t1=0.0, t2=0.0, use_transpose_cues=True, attn_mask=None):
    B, N, D = x.shape
    dk = D // heads
    a = F.linear(x, sd["qkv1.weight"]).reshape(B, N, 3, heads, dk).permute(2, 0, 3, 1, 4)
    b = F.linear(x, sd["qkv2.weight"]).reshape(B, N, 3, heads, dk).permute(2, 0, 3, 1, 4)
    y = crossview_core(a[0], a[1], a[2], b[0], b[1], sd["mix"], t1=t1, t2=t2, use_transpose_cues=use_transpose_cues,
                       zero_mask=attn_mask)
    return F.linear(y.transpose(1, 2).reshape(B, N, D), sd["proj.weight"])


def multihop_core(q1, k1, v1, q2, k2, v2, chain_value_logit, *, gates: Dict[str, float], beta_not: float = 0.5, hops: int = 3,
                  zero_mask: Optional[torch.Tensor] = None):
    dk = q1.shape[-1]
    s = 1.0 / math.sqrt(dk)
    S1 = (q1 @ k1.transpose(-2, -1)) * s
    S2 = (q2 @ k2.transpose(-2, -1)) * s
    if zero_mask is not None:
        S1 = S1.masked_fill(zero_mask == 0, float("-inf"))
        S2 = S2.masked_fill(zero_mask == 0, float("-inf"))
    A1, A2 = F.softmax(S1, dim=-1), F.softmax(S2, dim=-1)
    Smix = S1 + gates.get("and_", 1.0) * S2
    Smix = Smix + gates.get("or_", 0.0) * (torch.logsumexp(torch.stack([S1, S2], 0), 0) - S1)
    Smix = Smix - gates.get("not_", 0.0) * (beta_not * S2)
    C = A1 @ A2
    for _ in range(max(0, hops - 2)):
        C = C @ A2
    Smix = Smix + gates.get("chain", 0.0) * torch.log(C + 1e-6)
    if zero_mask is not None:
        Smix = Smix.masked_fill(zero_mask == 0, float("-inf"))
    A = F.softmax(Smix, dim=-1)
    t = v2
    for _ in range(max(0, hops - 1)):
        t = A2 @ t
    return A @ v1 + torch.sigmoid(chain_value_logit) * (A1 @ t)


def multihop_module(x, sd: Dict[str, torch.Tensor], heads: int, *, gates=None, beta_not=0.5, hops=3, attn_mask=None):
    B, N, D = x.shape
    dk = D // heads
    gates = gates or dict(and_=1.0, or_=0.0, not_=0.0, chain=0.0, base=1.0)
    a = F.linear(x, sd["qkv1.weight"]).reshape(B, N, 3, heads, dk).permute(2, 0, 3, 1, 4)
    b = F.linear(x, sd["qkv2.weight"]).reshape(B, N, 3, heads, dk).permute(2, 0, 3, 1, 4)
    y = multihop_core(a[0], a[1], a[2], b[0], b[1], b[2], sd["chain_value_logit"], gates=gates, beta_not=beta_not, hops=hops,
                      zero_mask=attn_mask)
    return F.linear(y.transpose(1, 2).reshape(B, N, D), sd["proj.weight"])

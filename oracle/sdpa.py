"""Oracle (test infrastructure): plain / causal / biased / cross attention on CPU.

Restates the reference's hand-written ``matmul -> softmax -> matmul`` attention:

* ``MSA.forward``                      mop/models/components.py:56-66
* ``BaselineMSA.forward``              mop/models/attention_variants.py:36-48
* ``MultiheadSelfAttention.forward``   mop/models/whisper_mop.py:154-177
* ``MultiheadCrossAttention.forward``  mop/models/whisper_mop.py:197-221

``sdpa_core`` is the kernel boundary (``q [B,H,Nq,dk]``, ``k,v [B,H,Nk,dk]``);
the ``*_module`` functions are the module boundary pinned against the imported
reference.  Not used by the product path.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


def sdpa_core(q, k, v, *, causal: bool = False, zero_mask: Optional[torch.Tensor] = None,
              bias: Optional[torch.Tensor] = None, return_probs: bool = False, drop_mask: Optional[torch.Tensor] = None):
    """softmax(q k^T / sqrt(dk) [masked] [+ bias]) v.

    Order of operations follows whisper_mop.py:163-175: scale, causal fill with
    -inf where ``j > i`` (bool tril, :165-167), then the additive bias
    (:169-170).  ``zero_mask`` is the ``mask == 0 -> -inf`` form of
    attention_variants.py:43-44.
    """
    dk = q.shape[-1]
    att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(dk))
    if zero_mask is not None:
        att = att.masked_fill(zero_mask == 0, float("-inf"))
    if causal:
        Tq, Tk = att.shape[-2:]
        keep = torch.tril(torch.ones(Tq, Tk, dtype=torch.bool, device=att.device))
        att = att.masked_fill(~keep, float("-inf"))
    if bias is not None:
        att = att + bias
    p = F.softmax(att, dim=-1)
    if drop_mask is not None:   # attn_drop(P) with an explicit mask of factors 0 | 1/(1-p)  (whisper_mop.py:173, components.py:63)
        p = p * drop_mask
    y = p @ v
    return (y, p) if return_probs else y


def msa_module(x, sd: Dict[str, torch.Tensor], heads: int, attn_mask=None):
    """components.MSA / attention_variants.BaselineMSA from a state_dict."""
    B, N, D = x.shape
    dk = D // heads
    qkv = F.linear(x, sd["qkv.weight"]).reshape(B, N, 3, heads, dk).permute(2, 0, 3, 1, 4)
    y = sdpa_core(qkv[0], qkv[1], qkv[2], zero_mask=attn_mask)
    return F.linear(y.transpose(1, 2).reshape(B, N, D), sd["proj.weight"])


def whisper_self_module(x, sd, n_head: int, causal: bool, attn_bias=None):
    B, T, D = x.shape
    dh = D // n_head
    lin = lambda n, t: F.linear(t, sd[f"{n}.weight"], sd.get(f"{n}.bias"))
    q = lin("q_proj", x).view(B, T, n_head, dh).transpose(1, 2)
    k = lin("k_proj", x).view(B, T, n_head, dh).transpose(1, 2)
    v = lin("v_proj", x).view(B, T, n_head, dh).transpose(1, 2)
    y = sdpa_core(q, k, v, causal=causal, bias=attn_bias)
    return lin("o_proj", y.transpose(1, 2).contiguous().view(B, T, D))


def whisper_cross_module(x_q, x_kv, sd, n_head: int, attn_mask=None):
    B, Tq, Dq = x_q.shape
    Tk = x_kv.shape[1]
    dh = Dq // n_head
    lin = lambda n, t: F.linear(t, sd[f"{n}.weight"], sd.get(f"{n}.bias"))
    q = lin("q_proj", x_q).view(B, Tq, n_head, dh).transpose(1, 2)
    k = lin("k_proj", x_kv).view(B, Tk, n_head, dh).transpose(1, 2)
    v = lin("v_proj", x_kv).view(B, Tk, n_head, dh).transpose(1, 2)
    y = sdpa_core(q, k, v, bias=attn_mask)
    return lin("o_proj", y.transpose(1, 2).contiguous().view(B, Tq, Dq))

"""GPT "Quartet" causal self-attention on the fused kernels.

Drop-in for ``TransformerConfig`` / ``CausalSelfAttention`` of
``mop/models/quartet_attn_patch.py:19-127``: same config fields, same parameters
(``{q,k,v,o,q2,k2}_proj``, ``mixture``, ``quartet_scale``), same non-persistent
``causal_mask`` buffer.  The two score maps, their full-row z-scores, the mix, the
causal fill, the softmax and PV run in libmop_b200 without materialising any T x T map.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from . import functional as MF


@dataclass
class TransformerConfig:
    n_layer: int = 6
    n_head: int = 8
    n_embd: int = 512
    dropout: float = 0.1
    block_size: int = 512
    bias: bool = False
    use_quartet: bool = True
    quartet_scale: float = 1.0
    quartet_gate_init: float = -5.0
    score_norm_eps: float = 1e-5
    use_abs_pos_emb: bool = True


class CausalSelfAttention(nn.Module):
    def __init__(self, config: TransformerConfig):
        super().__init__()
        if config.n_embd % config.n_head:
            raise AssertionError("n_embd must be divisible by n_head")
        self.config = config
        self.n_head = config.n_head
        self.head_dim = config.n_embd // config.n_head
        self.scale = 1.0 / math.sqrt(self.head_dim)
        D, b = config.n_embd, config.bias
        self.q_proj = nn.Linear(D, D, bias=b)
        self.k_proj = nn.Linear(D, D, bias=b)
        self.v_proj = nn.Linear(D, D, bias=b)
        self.o_proj = nn.Linear(D, D, bias=b)
        if config.use_quartet:
            self.q2_proj = nn.Linear(D, D, bias=b)
            self.k2_proj = nn.Linear(D, D, bias=b)
            self.mixture = nn.Parameter(torch.tensor([config.quartet_gate_init], dtype=torch.float32))
            self.quartet_scale = nn.Parameter(torch.tensor([config.quartet_scale], dtype=torch.float32))
        else:
            self.q2_proj = None
            self.k2_proj = None
            self.register_parameter("mixture", None)
            self.register_parameter("quartet_scale", None)
        self.attn_drop = nn.Dropout(config.dropout)
        self.resid_drop = nn.Dropout(config.dropout)
        # kept for state/API compatibility (reference :67-73); the kernel applies the causal fill itself
        self.register_buffer("causal_mask", torch.tril(torch.ones(config.block_size, config.block_size)).view(
            1, 1, config.block_size, config.block_size), persistent=False)

    def forward(self, x: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, need_weights: bool = False):
        if need_weights:
            raise NotImplementedError("need_weights=True would materialise the [B,H,T,T] map the fused kernel avoids")
        drop_p = self.attn_drop.p if self.training else 0.0   # applied to the probabilities inside the kernels (:118-119)
        B, T, C = x.shape
        if T > self.config.block_size:
            raise ValueError("sequence longer than block_size")
        shp = (B, T, self.n_head, self.head_dim)
        q, k, v = self.q_proj(x).view(shp), self.k_proj(x).view(shp), self.v_proj(x).view(shp)
        if self.config.use_quartet:
            y = MF.quartet_attention(q, k, v, self.q2_proj(x).view(shp), self.k2_proj(x).view(shp), self.mixture,
                                     self.quartet_scale, eps=self.config.score_norm_eps, add_mask=attention_mask, dropout_p=drop_p)
        else:
            y = MF.quartet_attention(q, k, v, eps=1e-5, add_mask=attention_mask, dropout_p=drop_p)  # reference :110 hard-codes 1e-5
        return self.resid_drop(self.o_proj(y.reshape(B, T, C)))

"""Whisper-MoP attention blocks on the fused kernels.

Drop-in for ``MultiheadSelfAttention`` / ``MultiheadCrossAttention`` of
``mop/models/whisper_mop.py:137-221``: same constructors, same
``{q,k,v,o}_proj`` parameters; the matmul-softmax-matmul core runs in
libmop_b200 (causal fill and additive bias handled in-kernel).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import functional as MF


def _drop_p(m) -> float:
    """Probability of the attention dropout (reference :173, :217), applied to the probabilities inside the kernels."""
    return m.attn_drop.p if m.training else 0.0


class MultiheadSelfAttention(nn.Module):
    def __init__(self, dim: int, n_head: int, dropout: float, bias: bool, causal: bool):
        super().__init__()
        if dim % n_head:
            raise AssertionError("dim must be divisible by n_head")
        self.dim, self.n_head, self.head_dim = dim, n_head, dim // n_head
        self.scale = self.head_dim ** -0.5
        self.causal = causal
        self.q_proj = nn.Linear(dim, dim, bias=bias)
        self.k_proj = nn.Linear(dim, dim, bias=bias)
        self.v_proj = nn.Linear(dim, dim, bias=bias)
        self.o_proj = nn.Linear(dim, dim, bias=bias)
        self.attn_drop = nn.Dropout(dropout)
        self.resid_drop = nn.Dropout(dropout)

    def forward(self, x: torch.Tensor, attn_bias: Optional[torch.Tensor] = None):
        B, T, D = x.shape
        shp = (B, T, self.n_head, self.head_dim)
        y = MF.sdpa(self.q_proj(x).view(shp), self.k_proj(x).view(shp), self.v_proj(x).view(shp),
                    causal=self.causal, bias=attn_bias, dropout_p=_drop_p(self))
        return self.resid_drop(self.o_proj(y.reshape(B, T, D)))


class MultiheadCrossAttention(nn.Module):
    def __init__(self, dim_q: int, dim_kv: int, n_head: int, dropout: float, bias: bool):
        super().__init__()
        if dim_q % n_head:
            raise AssertionError("dim_q must be divisible by n_head")
        self.n_head, self.head_dim = n_head, dim_q // n_head
        self.scale = self.head_dim ** -0.5
        self.q_proj = nn.Linear(dim_q, dim_q, bias=bias)
        self.k_proj = nn.Linear(dim_kv, dim_q, bias=bias)
        self.v_proj = nn.Linear(dim_kv, dim_q, bias=bias)
        self.o_proj = nn.Linear(dim_q, dim_q, bias=bias)
        self.attn_drop = nn.Dropout(dropout)
        self.resid_drop = nn.Dropout(dropout)

    def forward(self, x_q: torch.Tensor, x_kv: torch.Tensor, attn_mask: Optional[torch.Tensor] = None):
        B, Tq, Dq = x_q.shape
        Tk = x_kv.shape[1]
        H, dh = self.n_head, self.head_dim
        y = MF.sdpa(self.q_proj(x_q).view(B, Tq, H, dh), self.k_proj(x_kv).view(B, Tk, H, dh),
                    self.v_proj(x_kv).view(B, Tk, H, dh), bias=attn_mask, dropout_p=_drop_p(self))
        return self.resid_drop(self.o_proj(y.reshape(B, Tq, Dq)))


class ViewsConv2D(nn.Module):
    """1x1 conv producing V views of the mel map (reference :47-56); parameter container of the fused gate."""

    def __init__(self, n_views: int):
        super().__init__()
        self.conv = nn.Conv2d(1, n_views, kernel_size=1, bias=False)

    def forward(self, mel2d):
        return self.conv(mel2d)


class Kernels2D(nn.Module):
    """k x k conv producing K pattern maps from the views (reference :59-69)."""

    def __init__(self, in_ch: int, n_kernels: int, kernel_size: int):
        super().__init__()
        self.conv = nn.Conv2d(in_ch, n_kernels, kernel_size, padding=kernel_size // 2, bias=False)

    def forward(self, x):
        return self.conv(x)


class FuseExcInh2D(nn.Module):
    """1x1 conv to the excitatory / inhibitory fields + two global scalars (reference :72-88)."""

    def __init__(self, in_ch: int):
        super().__init__()
        self.conv = nn.Conv2d(in_ch, 2, kernel_size=1, bias=False)
        self.alpha = nn.Parameter(torch.ones(2))

    def forward(self, x):
        g_pos, g_neg = self.conv(x).chunk(2, dim=1)
        return g_pos, g_neg, self.alpha[0], self.alpha[1]


class MoP2D(nn.Module):
    """Drop-in for ``MoP2D`` (reference :91-124): same constructor, same parameters (``views.conv``, ``kernels.conv``,
    ``fuse.conv``, ``fuse.alpha``).  ``forward`` returns ``(gate_t [B,T,1], V, K)``; the only reference caller
    (``EncoderBlock.forward`` :261) uses ``gate_t`` alone, so by default the view / kernel maps are NOT materialised
    (``V = K = None``) and the gate comes from the fused kernel.  ``materialize_maps=True`` also returns the maps (PyTorch convs)."""

    def __init__(self, n_views: int, n_kernels: int, kernel_size: int, materialize_maps: bool = False):
        super().__init__()
        self.views = ViewsConv2D(n_views)
        self.kernels = Kernels2D(n_views, n_kernels, kernel_size)
        self.fuse = FuseExcInh2D(n_views + n_kernels)
        self.materialize_maps = bool(materialize_maps)

    def forward(self, mel2d: torch.Tensor):
        gate = MF.mop2d_gate(mel2d[:, 0], self.views.conv.weight, self.kernels.conv.weight, self.fuse.conv.weight, self.fuse.alpha)
        V = K = None
        if self.materialize_maps:
            V = self.views(mel2d)
            K = self.kernels(V)
        return gate.unsqueeze(-1).to(mel2d.dtype), V, K

"""ViT building blocks on the fused kernels.

``MSA`` mirrors ``mop/models/components.py:43-66`` (models A and B run it);
``DropPath``, ``PatchEmbed`` and ``MLP`` are the plain PyTorch neighbours the
ViT wrappers need (reference :14-40, :69-81) - they are not on the hot path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as MF


class DropPath(nn.Module):
    """Per-sample stochastic depth (reference components.py:14-27)."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if not self.training or self.drop_prob == 0.0:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


class PatchEmbed(nn.Module):
    """Non-overlapping patch projection; returns tokens and the patch grid (reference :30-40)."""

    def __init__(self, in_ch: int = 3, dim: int = 256, patch: int = 4):
        super().__init__()
        self.proj = nn.Conv2d(in_ch, dim, kernel_size=patch, stride=patch, bias=False)

    def forward(self, x):
        p = self.proj.kernel_size[0]
        B, C, Hh, Ww = x.shape
        if x.is_cuda and Hh % p == 0 and Ww % p == 0:
            # kernel == stride: the convolution is one GEMM over the flattened patches ([B, Gh*Gw, C*p*p] x [C*p*p, dim]).  cuDNN
            # wraps its implicit GEMM in NCHW <-> NHWC transposes (six extra launches per step for a 3 -> dim, 4 x 4 projection)
            Gh, Gw = Hh // p, Ww // p
            tok = x.view(B, C, Gh, p, Gw, p).permute(0, 2, 4, 1, 3, 5).reshape(B, Gh * Gw, C * p * p)
            return F.linear(tok, self.proj.weight.view(self.proj.out_channels, -1), self.proj.bias), (Gh, Gw)
        f = self.proj(x)
        return f.flatten(2).transpose(1, 2), tuple(f.shape[-2:])


class MLP(nn.Module):
    """fc1 -> GELU(tanh) -> fc2, bias-free (reference :69-81)."""

    def __init__(self, dim: int, mlp_ratio: float = 4.0, drop: float = 0.0):
        super().__init__()
        hid = int(dim * mlp_ratio)
        self.fc1 = nn.Linear(dim, hid, bias=False)
        self.fc2 = nn.Linear(hid, dim, bias=False)
        self.act = nn.GELU(approximate="tanh")
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.act(self.fc1(x))))


class MSA(nn.Module):
    """Multi-head self-attention, no mask (reference components.py:43-66)."""

    def __init__(self, dim, heads=4, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        if dim % heads:
            raise AssertionError("dim must be divisible by heads")
        self.h, self.dk = heads, dim // heads
        self.qkv = nn.Linear(dim, 3 * dim, bias=False)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim, bias=False)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        B, N, D = x.shape
        t = self.qkv(x).view(B, N, 3, self.h, self.dk)
        y = MF.sdpa(t[:, :, 0], t[:, :, 1], t[:, :, 2], dropout_p=self.attn_drop.p if self.training else 0.0)
        return self.proj_drop(self.proj(y.reshape(B, N, D)))


class Block(nn.Module):
    """Pre-LN transformer block around ``MSA`` (reference :124-141)."""

    def __init__(self, dim, heads, mlp_ratio=4.0, drop=0.0, attn_drop=0.0, drop_path=0.0):
        super().__init__()
        self.ln1 = nn.LayerNorm(dim)
        self.attn = MSA(dim, heads, attn_drop, drop)
        self.dp1 = DropPath(drop_path)
        self.ln2 = nn.LayerNorm(dim)
        self.mlp = MLP(dim, mlp_ratio, drop)
        self.dp2 = DropPath(drop_path)

    def forward(self, x):
        x = x + self.dp1(self.attn(self.ln1(x)))
        return x + self.dp2(self.mlp(self.ln2(x)))

"""ViT with Edgewise attention blocks (model "E").

Host-side mirror of ``BlockEdgewise`` / ``ViTEdgewise`` in
``experiments/cifar100_edgewise_gates.py:326-451`` - the classes every A/B/E
training script instantiates - with identical constructor arguments and
``state_dict`` keys.  Only the attention module differs: it is
``mop_b200.EdgewiseMSA`` (fused sm_100a kernels).

``compat_experiments_init`` defaults to True here because that experiments copy
only honours the and/or/chain gate presets; pass False for
the canonical ``mix5`` / ``not`` / ``nor`` / ``xor`` presets.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import functional as MF
from .attention_variants import EdgewiseMSA
from .components import MLP, DropPath, PatchEmbed


class BlockEdgewise(nn.Module):
    def __init__(self, dim: int, heads: int, mlp_ratio: float = 4.0, drop: float = 0.0, attn_drop: float = 0.0,
                 drop_path: float = 0.0, beta_not: float = 0.5, use_k3: bool = False, n_views: int = 2,
                 share_qkv: bool = False, gate_mode: str = "dense", gate_rank: int = 4, gate_init: str = "neutral",
                 use_lens_bank_qk: bool = False, lens_qk_kernel_size: int = 3,
                 lens_qk_dilations: Optional[Tuple[int, ...]] = None, lens_qk_causal: bool = False,
                 compat_experiments_init: bool = True):
        super().__init__()
        self.ln1 = nn.LayerNorm(dim)
        self.attn = EdgewiseMSA(
            dim, heads, attn_drop, drop, beta_not=beta_not, use_k3=use_k3, n_views=n_views, share_qkv=share_qkv,
            gate_mode=gate_mode, gate_rank=gate_rank, gate_init=gate_init, use_lens_bank_qk=use_lens_bank_qk,
            lens_qk_kernel_size=lens_qk_kernel_size, lens_qk_dilations=lens_qk_dilations,
            lens_qk_causal=lens_qk_causal, compat_experiments_init=compat_experiments_init)
        self.dp1 = DropPath(drop_path)
        self.ln2 = nn.LayerNorm(dim)
        self.mlp = MLP(dim, mlp_ratio, drop)
        self.dp2 = DropPath(drop_path)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x + self.dp1(self.attn(self.ln1(x)))
        return x + self.dp2(self.mlp(self.ln2(x)))

    # ---- fused residual stream: `x + dp(branch)` is folded into the LayerNorm that follows it -------------------
    @staticmethod
    def _dp_scale(dp: DropPath, x: torch.Tensor) -> Optional[torch.Tensor]:
        """Per-sample DropPath factor mask / keep (reference components.py:14-27), or None when DropPath is the identity."""
        if not dp.training or dp.drop_prob == 0.0:
            return None
        keep = 1.0 - dp.drop_prob
        return torch.empty(x.shape[0], dtype=torch.float32, device=x.device).bernoulli_(keep) / keep

    def forward_fused(self, x: torch.Tensor, branch: Optional[torch.Tensor], scale: Optional[torch.Tensor], dp_scales=None):
        """Same math as ``forward`` with the residual adds deferred: takes the residual stream and the not-yet-added branch
        of the previous block, returns the stream and this block's not-yet-added MLP branch ``(x, branch, scale)``.
        ``dp_scales``: this block's two per-sample DropPath factors when the caller drew them for the whole model at once."""
        s1, s2 = dp_scales if dp_scales is not None else (self._dp_scale(self.dp1, x), self._dp_scale(self.dp2, x))
        x, h = MF.add_layer_norm(x, branch, scale, self.ln1.weight, self.ln1.bias, self.ln1.eps)
        a = self.attn(h)
        x, h = MF.add_layer_norm(x, a, s1, self.ln2.weight, self.ln2.bias, self.ln2.eps)
        return x, self.mlp(h), s2


class ViTEdgewise(nn.Module):
    def __init__(self, dim: int = 256, depth: int = 8, heads: int = 4, n_classes: int = 100, mlp_ratio: float = 4.0,
                 drop: float = 0.0, drop_path: float = 0.1, patch: int = 4, num_tokens: int = 64,
                 beta_not: float = 0.5, use_k3: bool = False, n_views: int = 2, share_qkv: bool = False,
                 gate_mode: str = "dense", gate_rank: int = 4, gate_init: str = "neutral",
                 use_lens_bank_qk: bool = False, lens_qk_kernel_size: int = 3,
                 lens_qk_dilations: Optional[Tuple[int, ...]] = None, lens_qk_causal: bool = False,
                 compat_experiments_init: bool = True):
        super().__init__()
        self.patch = PatchEmbed(in_ch=3, dim=dim, patch=patch)
        self._use_builtin_patch = False
        self.pos = nn.Parameter(torch.zeros(1, num_tokens, dim))
        rates = [r.item() for r in torch.linspace(0, drop_path, depth)]
        self.blocks = nn.ModuleList(
            BlockEdgewise(dim, heads, mlp_ratio, drop, 0.0, rates[i], beta_not=beta_not, use_k3=use_k3,
                          n_views=n_views, share_qkv=share_qkv, gate_mode=gate_mode, gate_rank=gate_rank,
                          gate_init=gate_init, use_lens_bank_qk=use_lens_bank_qk,
                          lens_qk_kernel_size=lens_qk_kernel_size, lens_qk_dilations=lens_qk_dilations,
                          lens_qk_causal=lens_qk_causal, compat_experiments_init=compat_experiments_init)
            for i in range(depth))
        self.ln_f = nn.LayerNorm(dim)
        self.head = nn.Linear(dim, n_classes, bias=False)
        nn.init.normal_(self.pos, mean=0.0, std=0.02)

    def _draw_drop_path(self, tok: torch.Tensor):
        """All per-sample DropPath factors mask / keep of one forward ([2 * depth, B], reference components.py:14-27) from ONE
        uniform draw - four small kernels per step instead of two per residual branch (32 for depth 8)."""
        rates = [dp.drop_prob for blk in self.blocks for dp in (blk.dp1, blk.dp2)]
        if not self.training or not any(r > 0.0 for r in rates):
            return None
        keep = getattr(self, "_dp_keep", None)
        if keep is None or keep.device != tok.device or keep.numel() != len(rates):
            keep = torch.tensor([1.0 - r for r in rates], dtype=torch.float32, device=tok.device).unsqueeze(1)
            self._dp_keep = keep
        u = torch.rand(len(rates), tok.shape[0], dtype=torch.float32, device=tok.device)
        return (u < keep).to(torch.float32) / keep

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        tok, _ = self.patch(x)
        tok = tok + self.pos
        if tok.is_cuda and tok.dtype == torch.float32 and tok.shape[-1] <= 1024:
            # fused residual stream: every `x + dp(branch)` rides on the LayerNorm that follows it (one pass, one kernel)
            tok = tok.contiguous()
            branch = scale = None
            dps = self._draw_drop_path(tok)
            for i, blk in enumerate(self.blocks):
                tok, branch, scale = blk.forward_fused(tok, branch, scale, None if dps is None else (dps[2 * i], dps[2 * i + 1]))
            _, h = MF.add_layer_norm(tok, branch, scale, self.ln_f.weight, self.ln_f.bias, self.ln_f.eps, out_dtype=torch.float32)
            return self.head(h.mean(dim=1))
        for blk in self.blocks:
            tok = blk(tok)
        return self.head(self.ln_f(tok).mean(dim=1))

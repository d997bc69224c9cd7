"""Drop-in attention modules backed by the sm_100a kernels.

Same constructor signatures, parameter names/shapes (``state_dict`` layout) and
parameter creation order as the reference classes in
``mop/models/attention_variants.py`` - so reference checkpoints load both ways
and a given seed draws the same initial weights - but ``forward`` hands the
whole attention core to libmop_b200 instead of materialising N x N maps:

    EdgewiseGateHead   reference :234-331   (parameter container + preset init)
    EdgewiseMSA        reference :334-564   -> functional.edgewise_attention
    BaselineMSA        reference :23-48     -> functional.sdpa
    UnifiedMSA         reference :567-629   (modes A, B, E)

Unsupported corners raise instead of silently diverging:
  * ``attn_mask`` on EdgewiseMSA: the reference itself returns NaN for any mask
    (-inf enters the gate features), so there is nothing to match.
  * ``attn_drop > 0`` in training mode on EdgewiseMSA.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as MF


class BaselineMSA(nn.Module):
    """Plain multi-head self-attention with an optional ``mask == 0 -> -inf`` mask."""

    def __init__(self, dim: int, heads: int = 4, attn_drop: float = 0.0, proj_drop: float = 0.0):
        super().__init__()
        if dim % heads:
            raise AssertionError("dim must be divisible by heads")
        self.h, self.dk = heads, dim // heads
        self.qkv = nn.Linear(dim, 3 * dim, bias=False)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim, bias=False)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x: torch.Tensor, attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, N, D = x.shape
        t = self.qkv(x).view(B, N, 3, self.h, self.dk)
        y = MF.sdpa(t[:, :, 0], t[:, :, 1], t[:, :, 2], zero_mask=attn_mask, dropout_p=self.attn_drop.p if self.training else 0.0)
        return self.proj_drop(self.proj(y.reshape(B, N, D)))


def _no_attn_dropout(m: nn.Module):
    if m.training and m.attn_drop.p > 0.0:
        raise NotImplementedError(
            "attention dropout inside the fused Edgewise kernel is not provided (every reference caller of EdgewiseMSA passes "
            "attn_drop=0.0: experiments/cifar100_edgewise_gates.py:418); MSA / BaselineMSA / CausalSelfAttention / Whisper "
            "attention do apply it in-kernel")


class CrossViewMixerMSA(nn.Module):
    """Variant C (reference :51-156): two projections, the four score maps q_a k_b^T mixed by a 2x2 parameter, optional
    transpose cues, softmax, ``A v1``.

    The mixed logits are ONE dot product over concatenated features,
        S = [m11 q1 + m21 q2 | m12 q1 + m22 q2 | t1 k1 | t2 k2] . [k1 | k2 | q1 | q2] / sqrt(dk)
    (S1^T[i,j] = k1_i . q1_j), so the module runs on the plain-attention kernels (``functional.sdpa``) with a feature width of
    2 dk (4 dk with transpose cues); the linear feature mixing and its gradients (``mix`` included) stay in PyTorch.
    ``enable_per_key_prior`` (an argmax-anchored re-weighting, off by default) raises.
    """

    def __init__(self, dim: int, heads: int = 4, attn_drop: float = 0.0, proj_drop: float = 0.0, use_transpose_cues: bool = True,
                 t1: float = 0.0, t2: float = 0.0, enable_per_key_prior: bool = False, prior_weight: float = 0.5,
                 anchor_mode: str = "argmax_row_sum", fixed_k_star: int = 0):
        super().__init__()
        if dim % heads:
            raise AssertionError("dim must be divisible by heads")
        self.h, self.dk = heads, dim // heads
        self.qkv1 = nn.Linear(dim, dim * 3, bias=False)
        self.qkv2 = nn.Linear(dim, dim * 3, bias=False)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim, bias=False)
        self.proj_drop = nn.Dropout(proj_drop)
        self.mix = nn.Parameter(torch.tensor([[1.0, 0.0], [0.0, 1.0]]))
        self.use_transpose_cues = bool(use_transpose_cues)
        self.t1, self.t2 = float(t1), float(t2)
        self.enable_per_key_prior = bool(enable_per_key_prior)
        self.prior_weight = float(prior_weight)
        self.anchor_mode = str(anchor_mode)
        self.fixed_k_star = int(fixed_k_star)

    def forward(self, x: torch.Tensor, attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.enable_per_key_prior and self.prior_weight > 0.0:
            raise NotImplementedError("CrossViewMixerMSA: per-key prior sharpening (reference :125-150) is not provided")
        B, N, D = x.shape
        H, dk = self.h, self.dk
        a = self.qkv1(x).view(B, N, 3, H, dk)
        b = self.qkv2(x).view(B, N, 3, H, dk)
        q1, k1, v1, q2, k2 = a[:, :, 0], a[:, :, 1], a[:, :, 2], b[:, :, 0], b[:, :, 1]
        m = self.mix
        qs = [m[0, 0] * q1 + m[1, 0] * q2, m[0, 1] * q1 + m[1, 1] * q2]
        ks = [k1, k2]
        if self.use_transpose_cues:
            if self.t1 != 0.0:
                qs.append(self.t1 * k1); ks.append(q1)
            if self.t2 != 0.0:
                qs.append(self.t2 * k2); ks.append(q2)
        nf = len(qs)
        # the kernel divides by sqrt(nf dk), the reference by sqrt(dk)
        qc = (torch.cat(qs, dim=-1) * (nf ** 0.5)).to(v1.dtype)
        kc = torch.cat(ks, dim=-1).to(v1.dtype)
        vc = F.pad(v1, (0, (nf - 1) * dk))
        y = MF.sdpa(qc, kc, vc, zero_mask=attn_mask, dropout_p=self.attn_drop.p if self.training else 0.0)[..., :dk]
        return self.proj_drop(self.proj(y.reshape(B, N, D)))


class MultiHopMSA(nn.Module):
    """Variant D (reference :163-231): two projections, fixed scalar gates, multi-hop chain ``A_1 A_2^(hops-1)`` inside the
    logits and as value transport.  Runs as the fixed-gate mode of the Edgewise kernel (``gate_mode="const"``: the same
    score maps, LSE / AND / NOT mix, chain products and value transport, without a gate head or a reverse chain)."""

    def __init__(self, dim: int, heads: int = 4, attn_drop: float = 0.0, proj_drop: float = 0.0, beta_not: float = 0.5,
                 gates=None, hops: int = 3):
        super().__init__()
        if dim % heads:
            raise AssertionError("dim must be divisible by heads")
        if hops < 2:
            raise AssertionError("hops must be >= 2")
        self.h, self.dk = heads, dim // heads
        self.hops = int(hops)
        self.qkv1 = nn.Linear(dim, dim * 3, bias=False)
        self.qkv2 = nn.Linear(dim, dim * 3, bias=False)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim, bias=False)
        self.proj_drop = nn.Dropout(proj_drop)
        self.beta_not = float(beta_not)
        self.gates = gates or dict(and_=1.0, or_=0.0, not_=0.0, chain=0.0, base=1.0)
        self.chain_value_logit = nn.Parameter(torch.tensor(-2.0))

    def forward(self, x: torch.Tensor, attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if attn_mask is not None:
            raise NotImplementedError("MultiHopMSA: attn_mask is not provided by the fused kernel (no reference caller passes one)")
        _no_attn_dropout(self)
        B, N, D = x.shape
        qkv = torch.stack([self.qkv1(x), self.qkv2(x)], dim=2).view(B, N, 2, 3, self.h, self.dk)
        g = self.gates
        y = MF.edgewise_attention(qkv, None, None, None, self.chain_value_logit, {}, n_views=2, beta_not=self.beta_not, gate_mode="const",
                                  const_gates=(g.get("and_", 1.0), g.get("or_", 0.0), g.get("not_", 0.0), g.get("chain", 0.0)),
                                  hops=self.hops)
        return self.proj_drop(self.proj(y.reshape(B, N, D)))


class EdgewiseGateHead(nn.Module):
    """Parameters of the per-edge gate head (and/or/not/chain), reference :234-331.

    ``forward(feat)`` evaluates the head on an explicit feature stack
    ``[G,C,N,N]`` with ordinary torch ops; it exists for API compatibility and
    analysis.  ``EdgewiseMSA`` never calls it: the fused kernel reads these
    parameters directly and never builds the feature stack.

    ``compat_experiments_init=True`` reproduces the narrower preset table of
    ``experiments/cifar100_edgewise_gates.py:75-97`` (only and/or/chain), which
    is what the A/B/E training scripts instantiate.
    """

    _CHAN = {"and": 0, "or": 1, "not": 2, "chain": 3, "nor": 2, "xor": 1}

    def __init__(self, in_ch: int, hidden: int = 16, use_k3: bool = False, gate_mode: str = "dense",
                 gate_rank: int = 4, gate_init: str = "neutral", compat_experiments_init: bool = False):
        super().__init__()
        self.use_k3, self.gate_mode = bool(use_k3), str(gate_mode)
        self.gate_rank, self.gate_init = int(gate_rank), str(gate_init)
        chan = {k: v for k, v in self._CHAN.items() if k in ("and", "or", "chain")} if compat_experiments_init else self._CHAN
        if self.gate_mode == "dense":
            self.conv1 = nn.Conv2d(in_ch, hidden, 1)
            self.act = nn.GELU(approximate="tanh")
            if self.use_k3:
                self.mid3 = nn.Conv2d(hidden, hidden, 3, padding=1)
            self.conv2 = nn.Conv2d(hidden, 4, 1)
            with torch.no_grad():
                self.conv2.bias.fill_(-5.0)
                if self.gate_init in chan:
                    self.conv2.bias[chan[self.gate_init]] = 2.0
        else:
            r = self.gate_rank
            self.row_proj = nn.Conv1d(in_ch, 4 * r, 1)
            self.col_proj = nn.Conv1d(in_ch, 4 * r, 1)
            lift = float(max(0.0, (2.0 / max(1, r)) ** 0.5))
            if self.gate_init in chan:
                on = [chan[self.gate_init]]
            elif self.gate_init == "mix5" and not compat_experiments_init:
                on = [0, 1, 2]
            else:
                on = []
            with torch.no_grad():
                for b in (self.row_proj.bias, self.col_proj.bias):
                    b.zero_()
                    for t in on:
                        b[t * r:(t + 1) * r] = lift

    def tensors(self):
        """Gate-head tensors keyed by their reference state_dict names."""
        return {k: v for k, v in self.named_parameters()}

    def forward(self, feat: torch.Tensor) -> torch.Tensor:
        if self.gate_mode == "dense":
            h = self.act(self.conv1(feat))
            if self.use_k3:
                h = self.mid3(self.act(h))
            return torch.sigmoid(self.conv2(h))
        G, _, N, _ = feat.shape
        a = self.row_proj(feat.mean(3)).view(G, 4, self.gate_rank, N)
        b = self.col_proj(feat.mean(2)).view(G, 4, self.gate_rank, N)
        return torch.sigmoid(torch.einsum("gtki,gtkj->gtij", a, b))


class EdgewiseMSA(nn.Module):
    """Edgewise Mixture-of-Products attention (model "E"), reference :334-564."""

    def __init__(self, dim: int, heads: int = 4, attn_drop: float = 0.0, proj_drop: float = 0.0,
                 beta_not: float = 0.5, use_k3: bool = False, n_views: int = 2, share_qkv: bool = False,
                 gate_mode: str = "dense", gate_rank: int = 4, gate_init: str = "neutral",
                 use_lens_bank: bool = False, lens_kernel_size: int = 3,
                 lens_dilations: Optional[Tuple[int, ...]] = None,
                 use_lens_bank_qk: bool = False, lens_qk_kernel_size: int = 3,
                 lens_qk_dilations: Optional[Tuple[int, ...]] = None, lens_qk_causal: bool = False,
                 compat_experiments_init: bool = False, impl: Optional[str] = None):
        super().__init__()
        if dim % heads:
            raise AssertionError("dim must be divisible by heads")
        self.h, self.dk = heads, dim // heads
        self.beta_not = beta_not
        self.n_views = max(2, int(n_views))
        self.share_qkv = bool(share_qkv)
        self.use_lens_bank = bool(use_lens_bank)
        self.lens_kernel_size = int(lens_kernel_size)
        self.lens_dilations = tuple(lens_dilations) if lens_dilations is not None else (1, 2)
        self.use_lens_bank_qk = bool(use_lens_bank_qk)
        self.lens_qk_kernel_size = int(lens_qk_kernel_size)
        self.lens_qk_dilations = tuple(lens_qk_dilations) if lens_qk_dilations is not None else (1, 2)
        self.lens_qk_causal = bool(lens_qk_causal)
        self.impl = impl
        V = self.n_views
        if self.share_qkv:
            self.qkv = nn.Linear(dim, 3 * dim, bias=False)
            self.q_scale = nn.Parameter(torch.ones(V, heads, 1, self.dk))
            self.k_scale = nn.Parameter(torch.ones(V, heads, 1, self.dk))
            self.v_scale = nn.Parameter(torch.ones(V, heads, 1, self.dk))
        else:
            self.qkv_list = nn.ModuleList(nn.Linear(dim, 3 * dim, bias=False) for _ in range(V))
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim, bias=False)
        self.proj_drop = nn.Dropout(proj_drop)
        n_maps = len(self.lens_qk_dilations) if self.use_lens_bank_qk else V
        in_ch = 2 * n_maps + 2
        if self.use_lens_bank_qk:
            if not self.share_qkv:
                raise ValueError("use_lens_bank_qk=True requires share_qkv=True for now")
            ksz = self.lens_qk_kernel_size

            def bank():
                return nn.ModuleList(
                    nn.Conv1d(self.dk, self.dk, ksz, padding=0 if self.lens_qk_causal else d * (ksz - 1) // 2,
                              dilation=d, groups=self.dk, bias=False) for d in self.lens_qk_dilations)
            self.q_lens = bank()
            self.k_lens = bank()
            self._lens_qk_num = n_maps
        if self.use_lens_bank:
            self.lens_bank = nn.ModuleList(
                nn.Conv2d(n_maps, n_maps, self.lens_kernel_size, padding=d, dilation=d, groups=n_maps, bias=False)
                for d in self.lens_dilations)
            in_ch += n_maps * len(self.lens_dilations)
        self.edge_head = EdgewiseGateHead(in_ch, hidden=16, use_k3=use_k3, gate_mode=gate_mode, gate_rank=gate_rank,
                                          gate_init=gate_init, compat_experiments_init=compat_experiments_init)
        self.chain_value_logit = nn.Parameter(torch.tensor(-2.0))

    # -- Q/K lens bank prologue (reference :472-498): depthwise token conv on the view-0 Q/K
    def _lensed_qkv(self, base: torch.Tensor) -> torch.Tensor:
        B, N, _, H, dk = base.shape
        q0 = (base[:, :, 0] * self.q_scale[0].view(1, 1, H, dk)).permute(0, 2, 1, 3)
        k0 = (base[:, :, 1] * self.k_scale[0].view(1, 1, H, dk)).permute(0, 2, 1, 3)
        # reinterpretation of the [B,H,N,dk] buffer as [B*H,dk,N] (not a transpose), as the reference does
        qf, kf = q0.reshape(B * H, dk, N), k0.reshape(B * H, dk, N)
        V2 = len(self.lens_qk_dilations)
        out = base.new_zeros(B, N, V2, 3, H, dk)
        for i, d in enumerate(self.lens_qk_dilations):
            qi, ki = qf, kf
            if self.lens_qk_causal:
                left = (self.lens_qk_kernel_size - 1) * d
                qi, ki = F.pad(qf, (left, 0)), F.pad(kf, (left, 0))
            out[:, :, i, 0] = self.q_lens[i](qi).view(B, H, dk, N).permute(0, 3, 1, 2)
            out[:, :, i, 1] = self.k_lens[i](ki).view(B, H, dk, N).permute(0, 3, 1, 2)
        last = min(self.n_views - 1, V2 - 1)
        out[:, :, 0, 2] = base[:, :, 2] * self.v_scale[0].view(1, 1, H, dk)
        if V2 > 1:
            out[:, :, V2 - 1, 2] = base[:, :, 2] * self.v_scale[last].view(1, 1, H, dk)
        return out

    def forward(self, x: torch.Tensor, attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if attn_mask is not None:
            raise RuntimeError(
                "EdgewiseMSA: attn_mask is not supported - the reference forward returns NaN for any mask "
                "(masked scores enter the gate features as -inf); pass attn_mask=None")
        _no_attn_dropout(self)
        B, N, D = x.shape
        H, dk, V = self.h, self.dk, self.n_views
        scales = (None, None, None)
        if self.share_qkv:
            base = self.qkv(x).view(B, N, 3, H, dk)
            if self.use_lens_bank_qk:
                qkv = self._lensed_qkv(base)
                V = qkv.shape[2]
            else:
                qkv = base.unsqueeze(2)
                scales = (self.q_scale, self.k_scale, self.v_scale)
        else:
            qkv = torch.stack([lin(x) for lin in self.qkv_list], dim=2).view(B, N, V, 3, H, dk)
        head = self.edge_head
        lens_w, lens_d = None, ()
        if self.use_lens_bank:   # S lens bank: depthwise 3x3 convolutions of the score maps as extra gate-head channels
            if self.lens_kernel_size != 3:
                raise NotImplementedError("S lens bank: only lens_kernel_size=3 keeps the map size (reference padding = dilation)")
            lens_w = torch.stack([conv.weight[:, 0] for conv in self.lens_bank], dim=0)   # [L, V, 3, 3]
            lens_d = self.lens_dilations
        y = MF.edgewise_attention(
            qkv, *scales, self.chain_value_logit, head.tensors(), n_views=V, beta_not=self.beta_not,
            gate_mode=head.gate_mode, gate_rank=head.gate_rank, use_k3=head.use_k3, impl=self.impl,
            lens_w=lens_w, lens_dilations=lens_d)
        return self.proj_drop(self.proj(y.reshape(B, N, D)))


class UnifiedMSA(nn.Module):
    """Mode switch of the reference (:567-629).  Modes A/B (plain) and E (edgewise)
    run on the fused kernels; C and D are next-row items."""

    def __init__(self, mode: str, dim: int, heads: int = 4, **kwargs):
        super().__init__()
        mode = str(mode).upper()
        self.mode = mode
        if mode in ("A", "B"):
            self.impl = BaselineMSA(dim, heads, kwargs.get("attn_drop", 0.0), kwargs.get("proj_drop", 0.0))
        elif mode == "E":
            self.impl = EdgewiseMSA(
                dim, heads, kwargs.get("attn_drop", 0.0), kwargs.get("proj_drop", 0.0),
                beta_not=kwargs.get("beta_not", 0.5), use_k3=kwargs.get("use_k3", False),
                n_views=kwargs.get("n_views", 2), share_qkv=kwargs.get("share_qkv", False),
                gate_mode=kwargs.get("gate_mode", "dense"), gate_rank=kwargs.get("gate_rank", 4),
                gate_init=kwargs.get("gate_init", "neutral"))
        elif mode == "C":
            self.impl = CrossViewMixerMSA(
                dim, heads, kwargs.get("attn_drop", 0.0), kwargs.get("proj_drop", 0.0),
                use_transpose_cues=kwargs.get("use_transpose_cues", True), t1=kwargs.get("t1", 0.0), t2=kwargs.get("t2", 0.0),
                enable_per_key_prior=kwargs.get("enable_per_key_prior", False), prior_weight=kwargs.get("prior_weight", 0.5),
                anchor_mode=kwargs.get("anchor_mode", "argmax_row_sum"), fixed_k_star=kwargs.get("fixed_k_star", 0))
        elif mode == "D":
            self.impl = MultiHopMSA(
                dim, heads, kwargs.get("attn_drop", 0.0), kwargs.get("proj_drop", 0.0), beta_not=kwargs.get("beta_not", 0.5),
                gates=kwargs.get("gates", None), hops=kwargs.get("hops", 3))
        else:
            raise ValueError(f"Unknown attention mode: {mode}")

    def forward(self, x: torch.Tensor, attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.impl(x, attn_mask)

"""Oracle (test infrastructure): Edgewise Mixture-of-Products attention on CPU.

Restates, op for op, what the reference computes in
``mop/models/attention_variants.py``:

* gate head            -> ``EdgewiseGateHead.forward``        :311-331
* gate bias presets    -> ``EdgewiseGateHead.__init__``       :256-309
* per-view projections -> ``EdgewiseMSA.forward``             :456-470
* Q/K lens bank        ->                                     :472-498
* score maps, softmax, chains, feature stack                  :500-534
* gate mix, re-mask, softmax, PV, chain value transport       :535-562
* merge heads + proj                                          :563-564

Two entry points:

``edgewise_core``      kernel boundary: per-view Q/K/V ``[B,H,N,dk]`` in,
                       ``y [B,H,N,dk]`` out (no Linear layers).  This is what
                       the CUDA kernels are compared against.
``edgewise_msa``       module boundary: ``x [B,N,D]`` + a reference-layout
                       ``state_dict`` in, ``[B,N,D]`` out.  This is what is
                       pinned against the imported reference module.

Plain eager PyTorch, any float dtype (fp64 is the ground truth).  Not used by
the product path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

EPS_CHAIN = 1e-6  # attention_variants.py:516


@dataclass
class EdgewiseConfig:
    """Constructor arguments of the reference ``EdgewiseMSA`` (:335-356)."""

    dim: int
    heads: int = 4
    beta_not: float = 0.5
    use_k3: bool = False
    n_views: int = 2
    share_qkv: bool = False
    gate_mode: str = "dense"
    gate_rank: int = 4
    gate_init: str = "neutral"
    use_lens_bank: bool = False
    lens_kernel_size: int = 3
    lens_dilations: Tuple[int, ...] = (1, 2)
    use_lens_bank_qk: bool = False
    lens_qk_kernel_size: int = 3
    lens_qk_dilations: Tuple[int, ...] = (1, 2)
    lens_qk_causal: bool = False

    def __post_init__(self):
        self.n_views = max(2, int(self.n_views))  # :362

    @property
    def dk(self) -> int:
        return self.dim // self.heads

    @property
    def num_score_maps(self) -> int:
        return len(self.lens_qk_dilations) if self.use_lens_bank_qk else self.n_views

    @property
    def in_ch(self) -> int:  # :386-389, :442
        c = 2 * self.num_score_maps + 2
        if self.use_lens_bank:
            c += self.num_score_maps * len(self.lens_dilations)
        return c


# ----------------------------------------------------------------------------
# gate head
# ----------------------------------------------------------------------------
def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    return F.gelu(x, approximate="tanh")


def gate_head_dense(feat, conv1_w, conv1_b, conv2_w, conv2_b, mid3_w=None, mid3_b=None):
    """Dense head (:312-318).  feat [G,C,N,N] -> gates [G,4,N,N].

    With ``mid3`` present the hidden image goes through GELU *twice* before the
    3x3 convolution (:314-316) - reproduced on purpose.
    """
    h = gelu_tanh(F.conv2d(feat, conv1_w, conv1_b))
    if mid3_w is not None:
        h = F.conv2d(gelu_tanh(h), mid3_w, mid3_b, padding=1)
    return torch.sigmoid(F.conv2d(h, conv2_w, conv2_b))


def gate_head_lowrank(feat, row_w, row_b, col_w, col_b, rank: int):
    """Low-rank head (:319-331).  feat [G,C,N,N] -> gates [G,4,N,N].

    ``row_w``/``col_w`` are Conv1d(k=1) weights ``[4r,C,1]`` (or ``[4r,C]``).
    """
    G, C, N, _ = feat.shape
    rho = feat.mean(dim=3)  # [G,C,N] mean over columns j
    kap = feat.mean(dim=2)  # [G,C,N] mean over rows i
    a = torch.einsum("qc,gcn->gqn", row_w.reshape(4 * rank, C), rho) + row_b[None, :, None]
    b = torch.einsum("qc,gcn->gqn", col_w.reshape(4 * rank, C), kap) + col_b[None, :, None]
    a = a.reshape(G, 4, rank, N)
    b = b.reshape(G, 4, rank, N)
    return torch.sigmoid(torch.einsum("gtki,gtkj->gtij", a, b))


def gate_bias_preset(gate_mode: str, gate_init: str, rank: int, compat_experiments: bool = False):
    """Bias presets of the gate head (:256-309).

    Returns ``("dense", conv2_bias[4])`` or ``("lowrank", bias[4r])`` (the same
    vector initialises ``row_proj.bias`` and ``col_proj.bias``).

    ``compat_experiments`` reproduces the narrower preset table of the copy in
    ``experiments/cifar100_edgewise_gates.py:75-97`` (only and/or/chain).
    """
    chan = {"and": 0, "or": 1, "not": 2, "chain": 3, "nor": 2, "xor": 1}
    if compat_experiments:
        chan = {"and": 0, "or": 1, "chain": 3}
    if gate_mode == "dense":
        b = torch.full((4,), -5.0)
        if gate_init in chan:
            b[chan[gate_init]] = 2.0
        return "dense", b
    b = torch.zeros(4 * rank)
    c = float(max(0.0, (2.0 / max(1, rank)) ** 0.5))
    if gate_init in chan:
        idxs = [chan[gate_init]]
    elif gate_init == "mix5" and not compat_experiments:
        idxs = [0, 1, 2]  # :302-309
    else:
        idxs = []
    for i in idxs:
        b[i * rank:(i + 1) * rank] = c
    return "lowrank", b


# ----------------------------------------------------------------------------
# kernel-boundary core
# ----------------------------------------------------------------------------
def edgewise_core(
    qs: Sequence[torch.Tensor],
    ks: Sequence[torch.Tensor],
    v_first: torch.Tensor,
    v_last: torch.Tensor,
    head: Dict[str, torch.Tensor],
    chain_value_logit: torch.Tensor,
    *,
    beta_not: float,
    gate_mode: str,
    gate_rank: int = 4,
    attn_mask: Optional[torch.Tensor] = None,
    lens_bank: Optional[List[Tuple[torch.Tensor, int]]] = None,
    return_aux: bool = False,
):
    """Stages 1-3 of the path on per-view tensors ``[B,H,N,dk]``.

    ``qs[i]``, ``ks[i]`` are the per-view queries/keys that enter ``Q_i K_i^T``
    (:500-503); ``v_first`` is ``vs[0]`` (:553) and ``v_last`` is the value
    tensor that starts the transport chain (:556-557).  ``head`` holds the gate
    head tensors under the reference's state_dict names (without prefix).
    Returns ``y [B,H,N,dk]`` (before merge-heads/proj).
    """
    B, H, N, dk = qs[0].shape
    V = len(qs)
    s = 1.0 / math.sqrt(dk)
    S = [torch.matmul(qs[i], ks[i].transpose(-2, -1)) * s for i in range(V)]  # :500-503
    if attn_mask is not None:  # :504-506
        dead = attn_mask == 0
        S = [m.masked_fill(dead, float("-inf")) for m in S]
    A = [F.softmax(m, dim=-1) for m in S]  # :507
    Fc = A[0]
    for i in range(1, V):  # :508-512
        Fc = torch.matmul(Fc, A[i])
    Rc = A[-1]
    for i in range(V - 2, -1, -1):  # :513-515
        Rc = torch.matmul(Rc, A[i])
    G = B * H
    Sg = [m.reshape(G, N, N) for m in S]
    Lf = torch.log(Fc + EPS_CHAIN).reshape(G, N, N)  # :520
    Lr = torch.log(Rc + EPS_CHAIN).reshape(G, N, N)  # :521
    chans = Sg + [m.transpose(1, 2) for m in Sg] + [Lf, Lr]  # :522
    if lens_bank:  # :523-533, depthwise dilated conv over the stacked score maps
        stack = torch.stack(Sg, dim=1)
        for w, d in lens_bank:
            out = F.conv2d(stack, w, None, padding=d, dilation=d, groups=V)
            chans = chans + [out[:, c] for c in range(V)]
    feat = torch.stack(chans, dim=1)  # :534
    if gate_mode == "dense":
        gates = gate_head_dense(
            feat, head["conv1.weight"], head["conv1.bias"], head["conv2.weight"],
            head["conv2.bias"], head.get("mid3.weight"), head.get("mid3.bias"))
    else:
        gates = gate_head_lowrank(
            feat, head["row_proj.weight"], head["row_proj.bias"],
            head["col_proj.weight"], head["col_proj.bias"], gate_rank)
    g_and, g_or, g_not, g_chain = gates[:, 0], gates[:, 1], gates[:, 2], gates[:, 3]  # :536
    S1 = Sg[0]
    Ssum = S1
    for i in range(1, V):
        Ssum = Ssum + Sg[i]
    lse = torch.logsumexp(torch.stack(Sg, dim=1), dim=1)  # :541
    others = (Ssum - S1) / max(1, V - 1)  # :542
    Smix = S1 + g_and * (Ssum - S1)  # :543-547
    Smix = Smix + g_or * (lse - S1)
    Smix = Smix - g_not * (beta_not * others)
    Smix = Smix + g_chain * Lf
    Smix = Smix.reshape(B, H, N, N)
    if attn_mask is not None:  # :549-550
        Smix = Smix.masked_fill(attn_mask == 0, float("-inf"))
    Amix = F.softmax(Smix, dim=-1)  # :551 (attn_drop = 0 in every ViT caller)
    y = torch.matmul(Amix, v_first)  # :554
    t = v_last
    for i in range(V - 1, 0, -1):  # :558-559
        t = torch.matmul(A[i], t)
    y_chain = torch.matmul(A[0], t)  # :560
    w = torch.sigmoid(chain_value_logit)
    y = y + w * y_chain  # :562
    if return_aux:
        return y, dict(S=S, A=A, F=Fc, R=Rc, gates=gates.reshape(B, H, 4, N, N), Amix=Amix)
    return y


# ----------------------------------------------------------------------------
# module boundary
# ----------------------------------------------------------------------------
def _split_heads(t: torch.Tensor, H: int, dk: int):
    B, N, _ = t.shape
    return t.reshape(B, N, 3, H, dk).permute(2, 0, 3, 1, 4)  # :459


def per_view_qkv(x: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: EdgewiseConfig):
    """Per-view projections (:456-470) and the optional Q/K lens bank (:472-498).

    Returns ``(qs, ks, v_first, v_last)`` ready for :func:`edgewise_core`.
    """
    H, dk = cfg.heads, cfg.dk
    qs, ks, vs = [], [], []
    if cfg.share_qkv:
        base = _split_heads(F.linear(x, sd["qkv.weight"]), H, dk)
        for i in range(cfg.n_views):
            qs.append(base[0] * sd["q_scale"][i])
            ks.append(base[1] * sd["k_scale"][i])
            vs.append(base[2] * sd["v_scale"][i])
    else:
        for i in range(cfg.n_views):
            t = _split_heads(F.linear(x, sd[f"qkv_list.{i}.weight"]), H, dk)
            qs.append(t[0]); ks.append(t[1]); vs.append(t[2])
    if cfg.use_lens_bank_qk:
        B, _, N, _ = qs[0].shape
        # NB: a *reinterpretation* of the [B,H,N,dk] buffer as [B*H,dk,N], not
        # a transpose (:477-478).  Reproduced as is.
        qf = qs[0].reshape(B * H, dk, N)
        kf = ks[0].reshape(B * H, dk, N)
        ql, kl = [], []
        ksz = cfg.lens_qk_kernel_size
        for i, d in enumerate(cfg.lens_qk_dilations):
            if cfg.lens_qk_causal:
                pad_l, pad_c = (ksz - 1) * d, 0
            else:
                pad_l, pad_c = 0, d * (ksz - 1) // 2
            qi = F.pad(qf, (pad_l, 0)) if pad_l else qf
            ki = F.pad(kf, (pad_l, 0)) if pad_l else kf
            qo = F.conv1d(qi, sd[f"q_lens.{i}.weight"], None, padding=pad_c, dilation=d, groups=dk)
            ko = F.conv1d(ki, sd[f"k_lens.{i}.weight"], None, padding=pad_c, dilation=d, groups=dk)
            ql.append(qo.view(B, H, dk, N).transpose(2, 3))
            kl.append(ko.view(B, H, dk, N).transpose(2, 3))
        num_s = len(ql)
        v_last = vs[min(len(vs) - 1, num_s - 1)]  # :556-557
        return ql, kl, vs[0], v_last
    return qs, ks, vs[0], vs[cfg.n_views - 1]


def edgewise_msa(x: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: EdgewiseConfig,
                 attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Whole ``EdgewiseMSA.forward`` (:453-564) from a reference-layout state_dict."""
    B, N, D = x.shape
    qs, ks, v_first, v_last = per_view_qkv(x, sd, cfg)
    head = {k[len("edge_head."):]: v for k, v in sd.items() if k.startswith("edge_head.")}
    lens = None
    if cfg.use_lens_bank:
        lens = [(sd[f"lens_bank.{i}.weight"], d) for i, d in enumerate(cfg.lens_dilations)]
    y = edgewise_core(
        qs, ks, v_first, v_last, head, sd["chain_value_logit"],
        beta_not=cfg.beta_not, gate_mode=cfg.gate_mode, gate_rank=cfg.gate_rank,
        attn_mask=attn_mask, lens_bank=lens)
    y = y.transpose(1, 2).reshape(B, N, D)  # :563
    return F.linear(y, sd["proj.weight"])  # :564 (proj_drop = 0)

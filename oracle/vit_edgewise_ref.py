"""Oracle (test infrastructure): eager CPU ViT with Edgewise attention blocks.

Restates ``ViTEdgewise`` / ``BlockEdgewise`` (experiments/cifar100_edgewise_gates.py:326-451)
on top of :func:`oracle.edgewise.edgewise_msa`, i.e. every attention layer is the
reference's eager matmul/softmax/stack sequence that materialises the N x N maps.
It is a pure function of a reference-layout ``state_dict`` so that the product
model (GPU) and this port (CPU) can be driven with identical weights.  Used as
the checker in tests and as the timed CPU baseline (``kind: "port"``) of bench.py.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .edgewise import EdgewiseConfig, edgewise_msa


def drop_path(x, rate: float, training: bool):
    if not training or rate == 0.0:
        return x
    keep = 1.0 - rate
    mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
    return x * mask / keep


def vit_edgewise_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: EdgewiseConfig, *, depth: int,
                         patch: int, drop_path_rate: float = 0.0, training: bool = False) -> torch.Tensor:
    """images [B,3,S,S] -> logits; ``sd`` uses the ViTEdgewise state_dict keys."""
    tok = F.conv2d(x, sd["patch.proj.weight"], stride=patch).flatten(2).transpose(1, 2) + sd["pos"]
    rates = torch.linspace(0, drop_path_rate, depth).tolist()
    D = tok.shape[-1]
    for i in range(depth):
        pre = f"blocks.{i}."
        attn_sd = {k[len(pre) + 5:]: v for k, v in sd.items() if k.startswith(pre + "attn.")}
        h = F.layer_norm(tok, (D,), sd[pre + "ln1.weight"], sd[pre + "ln1.bias"])
        tok = tok + drop_path(edgewise_msa(h, attn_sd, cfg), rates[i], training)
        h = F.layer_norm(tok, (D,), sd[pre + "ln2.weight"], sd[pre + "ln2.bias"])
        h = F.linear(F.gelu(F.linear(h, sd[pre + "mlp.fc1.weight"]), approximate="tanh"), sd[pre + "mlp.fc2.weight"])
        tok = tok + drop_path(h, rates[i], training)
    tok = F.layer_norm(tok, (D,), sd["ln_f.weight"], sd["ln_f.bias"])
    return F.linear(tok.mean(dim=1), sd["head.weight"])


def train_step_cpu(sd: Dict[str, torch.Tensor], cfg: EdgewiseConfig, x, labels, *, depth: int, patch: int,
                   drop_path_rate: float, opt: torch.optim.Optimizer) -> float:
    """One eager fwd + CE + bwd + optimizer step on the CPU (what the reference's train loop does per model,
    experiments/cifar100_ab5_param_budgets.py:799-805)."""
    opt.zero_grad(set_to_none=True)
    logits = vit_edgewise_forward(x, sd, cfg, depth=depth, patch=patch, drop_path_rate=drop_path_rate, training=True)
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    opt.step()
    return float(loss.detach())

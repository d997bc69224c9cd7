"""bf16 compute copies of the Linear weights for mixed-precision training (fp32 master weights, bf16 GEMM operands).

Under `torch.autocast` every `nn.Linear` casts its fp32 weight to bf16 in the forward and casts the bf16 weight gradient back
to fp32 in the backward: two tiny elementwise kernels per Linear and step, ~70 of the ~350 launches of the ViT-MoP E+ bench
step (4 Linears per block, the GEMMs themselves are 10-25 us).  `Bf16Shadow(model)` keeps ONE flat bf16 buffer with a view per
Linear weight, refreshed by one multi-tensor copy after the optimizer step (`refresh()`), and routes those Linears through an
autograd function whose weight gradient is produced in fp32 directly by the GEMM (`torch.mm(..., out_dtype=float32)`).
Same forward numerics as autocast (the same bf16-rounded weights); the weight gradient skips one bf16 rounding.

    shadow = Bf16Shadow(model)          # after model.cuda(), before capture / training
    ...
    optimizer.step(); shadow.refresh()  # whenever the fp32 weights change (also after load_state_dict)
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F


def _weight_grad(dy2: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """dW = dy2^T x2 in fp32.  For the small Linear layers of the ViT-MoP models the output has few tiles (672 x 224: 77 tiles of
    64 x 32 on 148 SMs) and the reduction is long (16384 tokens): the tokens are split into up to 8 slices, one batched GEMM forms the
    slice products and a small sum adds them (measured on B200, CUDA-graph replays: 21.5 -> 13.4 us for 672 x 16384 x 224)."""
    M, O = dy2.shape
    I = x2.shape[1]
    tiles = -(-O // 64) * -(-I // 32)
    S = 1
    while S < 8 and tiles * S < 296 and M % (2 * S) == 0 and M // (2 * S) >= 1024:
        S *= 2
    if S == 1 or not (dy2.is_contiguous() and x2.is_contiguous()):
        return torch.mm(dy2.t(), x2, out_dtype=torch.float32)
    return torch.bmm(dy2.view(S, M // S, O).transpose(1, 2), x2.view(S, M // S, I), out_dtype=torch.float32).sum(0)


class _ShadowLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w32, w16, b):
        ctx.save_for_backward(x, w16)
        ctx.has_bias = b is not None
        return F.linear(x, w16, None if b is None else b.to(w16.dtype))

    @staticmethod
    def backward(ctx, dy):
        x, w16 = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        x2 = x.reshape(-1, x.shape[-1])
        dx = torch.matmul(dy, w16) if ctx.needs_input_grad[0] else None
        dw = _weight_grad(dy2, x2)                                   # fp32 straight out of the GEMM: no bf16 round trip, no cast kernel
        db = dy2.sum(0, dtype=torch.float32) if ctx.has_bias else None
        return dx, dw, None, db


class Bf16Shadow:
    def __init__(self, model: nn.Module):
        self.mods: List[nn.Linear] = [m for m in model.modules() if type(m) is nn.Linear and m.weight.is_cuda and m.weight.dtype == torch.float32]
        n = sum(m.weight.numel() for m in self.mods)
        dev = self.mods[0].weight.device if self.mods else None
        self.flat = torch.empty(n, dtype=torch.bfloat16, device=dev)
        self.views, off = [], 0
        for m in self.mods:
            v = self.flat[off:off + m.weight.numel()].view_as(m.weight)
            off += m.weight.numel()
            self.views.append(v)
            m._mop_w16 = v
            m.forward = _bound_forward(m)
        self.refresh()

    @torch.no_grad()
    def refresh(self) -> None:
        """Copy the fp32 master weights into the bf16 compute copies (one multi-tensor kernel)."""
        if self.mods:
            torch._foreach_copy_(self.views, [m.weight for m in self.mods])

    def disable(self) -> None:
        for m in self.mods:
            m.__dict__.pop("forward", None)
            m.__dict__.pop("_mop_w16", None)


def _bound_forward(m: nn.Linear):
    def forward(x):
        if x.is_cuda and torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
            return _ShadowLinear.apply(x.to(torch.bfloat16), m.weight, m._mop_w16, m.bias)
        return F.linear(x, m.weight, m.bias)
    return forward


class FlatParams:
    """All trainable fp32 parameters of a model re-homed into ONE flat buffer (each `p.data` becomes a view of it), with a
    matching flat gradient buffer: a fused optimizer then updates the whole model with one kernel over one tensor instead of one
    multi-tensor launch per ~40 parameters.  The per-parameter gradients autograd produces are gathered with `pack_grads()`
    (one concatenation), or the flat gradient buffer of `ddp.FlatGradAllReduce(pack=True)` is adopted with `use_grad_buffer()`.

        fp = FlatParams(model); opt = torch.optim.AdamW([fp.param], fused=True, ...)
        loss.backward(); fp.pack_grads(); opt.step()          # p.grad must be None before the backward (fp.begin())
    """

    def __init__(self, model: nn.Module):
        self.params = [p for p in model.parameters() if p.requires_grad]
        if any(p.dtype != torch.float32 for p in self.params):
            raise TypeError("FlatParams expects fp32 parameters")
        total = sum(p.numel() for p in self.params)
        flat = torch.empty(total, dtype=torch.float32, device=self.params[0].device)
        off = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                flat[off:off + n].copy_(p.data.reshape(-1))
                p.data = flat[off:off + n].view_as(p)
                off += n
        self.param = nn.Parameter(flat)
        self.grad_buf = torch.zeros_like(flat)
        self.param.grad = self.grad_buf

    def use_grad_buffer(self, flat_grad: torch.Tensor) -> None:
        self.grad_buf = flat_grad
        self.param.grad = flat_grad

    def begin(self) -> None:
        for p in self.params:
            p.grad = None

    def pack_grads(self) -> None:
        gs = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params]
        torch.cat(gs, out=self.grad_buf)
        self.param.grad = self.grad_buf   # (an optimizer.zero_grad(set_to_none=True) in between must not detach the buffer)

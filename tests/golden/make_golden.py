"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the authoring container only (needs the read-only reference tree):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py [/root/reference]

For every case it builds the unmodified reference module, re-draws all
parameters from a seeded N(.,.) so that every gate / scale / bias is live
(reference default init leaves e.g. q_scale == 1), evaluates the module in
fp64 (``module.double()``) and records

    x, dy                 fp32-representable inputs (stored as fp32)
    state_dict            fp32-representable parameters (stored as fp32)
    y                     fp64 output of the reference module
    dx, dparams           fp64 autograd gradients of <y, dy>

into ``<case>.pt`` (a plain dict of tensors + python scalars).  The reference
publishes no golden vectors for this path, so these
files are what pins the oracle and, through it, the CUDA kernels.

Also writes ``gate_presets.pt``: the gate-head bias vectors the reference
constructors produce for every (gate_mode, gate_init), for the canonical class
and for the narrower copy in experiments/cifar100_edgewise_gates.py.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "experiments"))
for name in ("matplotlib", "matplotlib.pyplot"):  # off-path import of the experiment scripts
    sys.modules.setdefault(name, types.ModuleType(name))

from mop.models.attention_variants import BaselineMSA, EdgewiseGateHead, EdgewiseMSA  # noqa: E402
from mop.models.components import MSA  # noqa: E402
from mop.models.quartet_attn_patch import CausalSelfAttention, TransformerConfig  # noqa: E402
from mop.models.whisper_mop import MultiheadCrossAttention, MultiheadSelfAttention  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def redraw(mod: torch.nn.Module, seed: int):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in mod.named_parameters():
            if name.endswith("_scale") and p.ndim == 4:        # q/k/v_scale: around 1
                p.copy_(1.0 + 0.25 * torch.randn(p.shape, generator=g))
            elif p.ndim == 0 or p.numel() == 1:                # chain_value_logit, mixture, quartet_scale
                p.copy_(torch.randn(p.shape, generator=g) * 0.5 + (1.0 if "quartet_scale" in name else -0.5))
            elif "edge_head" in name:
                p.copy_(0.35 * torch.randn(p.shape, generator=g))
            elif "lens" in name:
                p.copy_(0.5 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            else:                                              # Linear weights
                fan_in = p.shape[-1]
                p.copy_(torch.randn(p.shape, generator=g) * (1.6 / fan_in ** 0.5))


def record(name, mod, inputs, call, meta):
    """inputs: dict name -> fp32 tensor that requires grad in the fp64 replay."""
    mod = mod.double()
    ins64 = {k: v.double().requires_grad_(True) for k, v in inputs.items()}
    y = call(mod, ins64)
    g = torch.Generator().manual_seed(1234)
    dy = torch.randn(y.shape, generator=g).float()
    params = dict(mod.named_parameters())
    grads = torch.autograd.grad(y, list(ins64.values()) + list(params.values()), dy.double(), allow_unused=True)
    out = dict(meta)
    out["inputs"] = {k: v.float() for k, v in inputs.items()}
    out["dy"] = dy
    out["state_dict"] = {k: v.detach().float() for k, v in mod.state_dict().items()}
    out["y"] = y.detach()
    n_in = len(ins64)
    out["dinputs"] = {k: grads[i].detach() for i, k in enumerate(ins64)}
    out["dparams"] = {k: (grads[n_in + i].detach() if grads[n_in + i] is not None else torch.zeros_like(p))
                      for i, (k, p) in enumerate(params.items())}
    torch.save(out, os.path.join(OUT, name + ".pt"))
    print(f"{name}: y{tuple(y.shape)} |y|max={y.abs().max():.3f}")


def edgewise_cases():
    cases = [
        # name,            dim, H, N, kwargs
        ("ew_lowrank_share_v5", 16, 2, 12, dict(n_views=5, share_qkv=True, gate_mode="lowrank", gate_rank=4, gate_init="mix5", use_k3=True)),
        ("ew_lowrank_sep_v3", 24, 3, 9, dict(n_views=3, share_qkv=False, gate_mode="lowrank", gate_rank=2, gate_init="neutral")),
        ("ew_lowrank_share_v2_n64", 112, 2, 64, dict(n_views=2, share_qkv=True, gate_mode="lowrank", gate_rank=4, gate_init="and", beta_not=0.8)),
        ("ew_dense_share_v2", 16, 2, 10, dict(n_views=2, share_qkv=True, gate_mode="dense", gate_init="or")),
        ("ew_dense_k3_share_v5", 16, 2, 12, dict(n_views=5, share_qkv=True, gate_mode="dense", use_k3=True, gate_init="neutral")),
        ("ew_dense_k3_sep_v3", 24, 2, 8, dict(n_views=3, share_qkv=False, gate_mode="dense", use_k3=True, gate_init="and", beta_not=0.25)),
        ("ew_lensqk_lowrank", 16, 2, 12, dict(n_views=3, share_qkv=True, gate_mode="lowrank", gate_rank=2, use_lens_bank_qk=True, lens_qk_dilations=(1, 2, 3))),
        ("ew_lensqk_causal_dense", 16, 2, 12, dict(n_views=2, share_qkv=True, gate_mode="dense", use_k3=True, use_lens_bank_qk=True, lens_qk_causal=True)),
        ("ew_lensS_dense", 16, 2, 10, dict(n_views=3, share_qkv=True, gate_mode="dense", use_k3=True, use_lens_bank=True)),
        ("ew_lensqk_single_v1", 16, 2, 9, dict(n_views=3, share_qkv=True, gate_mode="lowrank", gate_rank=2, use_lens_bank_qk=True,
                                               lens_qk_dilations=(2,))),
        ("ew_lensS_lowrank_qk", 16, 2, 10, dict(n_views=4, share_qkv=True, gate_mode="lowrank", gate_rank=2, use_lens_bank=True,
                                                lens_dilations=(1, 2), use_lens_bank_qk=True, lens_qk_dilations=(2, 3), lens_qk_causal=True)),
    ]
    only = os.environ.get("MOP_GOLDEN_ONLY")   # regenerate a single (new) case without touching the committed ones
    for i, (name, dim, H, N, kw) in enumerate(cases):
        if only and name != only:
            continue
        torch.manual_seed(100 + i)
        mod = EdgewiseMSA(dim, heads=H, **kw)
        redraw(mod, 200 + i)
        x = torch.randn(2, N, dim)
        record(name, mod, {"x": x}, lambda m, t: m(t["x"]), dict(kind="edgewise", dim=dim, heads=H, kwargs=kw))


def sdpa_cases():
    torch.manual_seed(7)
    mod = MSA(54 * 2, heads=2); redraw(mod, 300)
    record("msa_dk54", mod, {"x": torch.randn(2, 17, 108)}, lambda m, t: m(t["x"]), dict(kind="msa", heads=2))
    mod = BaselineMSA(32, heads=4); redraw(mod, 301)
    mask = (torch.rand(2, 1, 11, 11) > 0.3).float()
    mask[..., 0] = 1.0  # keep every row alive
    record("baseline_msa_mask", mod, {"x": torch.randn(2, 11, 32)}, lambda m, t: m(t["x"], mask.double()),
           dict(kind="baseline_msa", heads=4, mask=mask))
    for causal in (False, True):
        mod = MultiheadSelfAttention(32, 4, 0.0, True, causal=causal); redraw(mod, 302 + causal)
        bias = 0.5 * torch.randn(1, 4, 13, 13)
        record(f"whisper_self_causal{int(causal)}", mod, {"x": torch.randn(2, 13, 32)},
               lambda m, t: m(t["x"], bias.double()), dict(kind="whisper_self", n_head=4, causal=causal, bias=bias))
    mod = MultiheadCrossAttention(32, 48, 4, 0.0, False); redraw(mod, 305)
    record("whisper_cross", mod, {"x_q": torch.randn(2, 7, 32), "x_kv": torch.randn(2, 19, 48)},
           lambda m, t: m(t["x_q"], t["x_kv"]), dict(kind="whisper_cross", n_head=4))


def quartet_cases():
    for name, uq, bias, T in (("quartet_T24", True, False, 24), ("quartet_bias_T9", True, True, 9), ("quartet_off_T16", False, False, 16)):
        cfg = TransformerConfig(n_layer=1, n_head=2, n_embd=32, dropout=0.0, block_size=32, bias=bias, use_quartet=uq)
        torch.manual_seed(11)
        mod = CausalSelfAttention(cfg); redraw(mod, 400 + T)
        am = None
        if name == "quartet_bias_T9":
            am = 0.3 * torch.randn(2, 1, T, T)
        record(name, mod, {"x": torch.randn(2, T, 32)},
               lambda m, t: m(t["x"], attention_mask=None if am is None else am.double()),
               dict(kind="quartet", n_head=2, use_quartet=uq, add_mask=am, eps=cfg.score_norm_eps))


def gate_presets():
    import cifar100_edgewise_gates as exp  # the copy every A/B/E script trains
    table = {}
    for mode in ("dense", "lowrank"):
        for init in ("neutral", "and", "or", "not", "nor", "xor", "chain", "mix5"):
            for tag, cls in (("canonical", EdgewiseGateHead), ("experiments", exp.EdgewiseGateHead)):
                h = cls(in_ch=6, hidden=16, use_k3=False, gate_mode=mode, gate_rank=3, gate_init=init)
                if mode == "dense":
                    table[f"{tag}/{mode}/{init}"] = h.conv2.bias.detach().clone()
                else:
                    assert torch.equal(h.row_proj.bias, h.col_proj.bias)
                    table[f"{tag}/{mode}/{init}"] = h.row_proj.bias.detach().clone()
    torch.save(table, os.path.join(OUT, "gate_presets.pt"))
    print("gate_presets:", len(table))


if __name__ == "__main__":
    edgewise_cases()
    if os.environ.get("MOP_GOLDEN_ONLY"):
        sys.exit(0)
    sdpa_cases()
    quartet_cases()
    gate_presets()

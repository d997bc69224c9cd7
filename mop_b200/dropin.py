"""Swap the reference's attention classes for the fused ones, in place.

    import mop_b200.dropin as dropin
    dropin.patch_reference()          # after `import mop` / the experiments module, before building models

Every reference model wrapper looks its attention class up by module-global name at construction
time (`BlockEdgewise.__init__` -> `EdgewiseMSA`, `Block.__init__` -> `MSA`, `MoPBlock` / `Block` ->
`CausalSelfAttention`, `EncoderBlock` / `DecoderBlock` -> `Multihead{Self,Cross}Attention`), so
rebinding those names is enough: model code, CLI flags, `state_dict` layout and training loops stay
untouched.  Each of the three copies of `EdgewiseMSA` / `EdgewiseGateHead` (canonical + two
experiment scripts) is replaced; for the experiment copies the narrower gate-preset table of that
copy is kept (`compat_experiments_init=True`).
"""
from __future__ import annotations

import functools
import sys
from typing import Dict, List

from . import attention_variants as av
from . import components as comp
from . import quartet_attn_patch as qp
from . import whisper_mop as wm

_CANONICAL = {
    "EdgewiseMSA": av.EdgewiseMSA, "EdgewiseGateHead": av.EdgewiseGateHead, "BaselineMSA": av.BaselineMSA,
    "CrossViewMixerMSA": av.CrossViewMixerMSA, "MultiHopMSA": av.MultiHopMSA,
    "MSA": comp.MSA, "CausalSelfAttention": qp.CausalSelfAttention,
    "MultiheadSelfAttention": wm.MultiheadSelfAttention, "MultiheadCrossAttention": wm.MultiheadCrossAttention,
    "MoP2D": wm.MoP2D,
}
_EXPERIMENT_MODULES = tuple(pre + n for pre in ("", "experiments.") for n in (
    "cifar100_edgewise_gates", "cifar10_edgewise_gates", "cifar100_crossview_mixer", "cifar10_crossview_mixer",
    "cifar100_multihop_gates", "cifar10_multihop_gates"))


def _compat(cls):
    @functools.wraps(cls, updated=())
    class _Compat(cls):  # same class, experiments-copy gate presets by default
        def __init__(self, *a, **kw):
            kw.setdefault("compat_experiments_init", True)
            super().__init__(*a, **kw)
    _Compat.__name__ = cls.__name__
    _Compat.__qualname__ = cls.__qualname__
    return _Compat


_ORIGINALS: Dict[tuple, type] = {}
_FUSED_VIT_MOP: Dict[type, type] = {}


def _fused_vit_mop(ref_cls: type) -> type:
    """The reference's ViT_MoP with the post-encoder token gate (vit_mop.py:95-114) computed by the fused kernel: same
    constructor, parameters and state_dict (it IS the reference class, only `forward` is replaced)."""
    if ref_cls in _FUSED_VIT_MOP:
        return _FUSED_VIT_MOP[ref_cls]
    import torch.nn.functional as F
    from . import functional as MF

    def forward(self, x):
        tok, grid = self.enc(x)
        k3, k1 = self.kerns.k[0].weight, self.kerns.k[2].weight
        f1, f2 = self.fuse.fuse[0], self.fuse.fuse[2]
        if not MF.token_gate_supported(tok, grid, self.n_views, self.n_kernels, f1.weight.shape[0]):
            return ref_cls.forward(self, x)   # (shapes outside the kernel: the reference's own PyTorch composition)
        tok = MF.token_gate(tok, grid, self.views.proj.weight, k3, k1, f1.weight, f2.weight, f2.bias,
                            F.softplus(self.fuse.alpha_pos), F.softplus(self.fuse.alpha_neg))
        return self.cls(tok.mean(dim=1))

    cls = type(ref_cls.__name__, (ref_cls,), {"forward": forward, "__module__": "mop_b200.dropin", "__doc__": _fused_vit_mop.__doc__})
    _FUSED_VIT_MOP[ref_cls] = cls
    return cls


_FUSED_MOP_BLOCK: Dict[type, type] = {}


def _fused_mop_block(ref_cls: type) -> type:
    """The reference's `MoPBlock` with `apply_mop` (gpt_mop.py:102-123) through the fused 1-D token-gate kernel: same constructor,
    parameters and state_dict (a subclass that only replaces that method)."""
    if ref_cls in _FUSED_MOP_BLOCK:
        return _FUSED_MOP_BLOCK[ref_cls]
    from . import functional as MF

    def apply_mop(self, x):
        conv = self.kernels.conv
        if not MF.token_gate_1d_supported(x, self.n_views, conv.kernel_size[0]) or conv.padding[0] != 1:
            return ref_cls.apply_mop(self, x)   # (shapes outside the kernel: the reference's own PyTorch composition)
        return MF.token_gate_1d(x, self.views.proj.weight, conv.weight, self.fuse.conv.weight, self.fuse.alpha)

    cls = type(ref_cls.__name__, (ref_cls,), {"apply_mop": apply_mop, "__module__": "mop_b200.dropin", "__doc__": _fused_mop_block.__doc__})
    _FUSED_MOP_BLOCK[ref_cls] = cls
    return cls


def unpatch_reference() -> int:
    """Undo patch_reference(): restore the reference's own classes.  Returns how many names were restored."""
    n = 0
    for (modname, name), cls in list(_ORIGINALS.items()):
        mod = sys.modules.get(modname)
        if mod is not None:
            setattr(mod, name, cls)
            n += 1
    _ORIGINALS.clear()
    return n


def patch_reference(verbose: bool = False) -> Dict[str, List[str]]:
    """Rebind the attention classes in every loaded `mop.*` module and experiment script.  Returns what was replaced."""
    done: Dict[str, List[str]] = {}
    for modname, mod in list(sys.modules.items()):
        if mod is None:
            continue
        is_ref_pkg = modname == "mop" or modname.startswith("mop.")
        is_exp = modname in _EXPERIMENT_MODULES
        if not (is_ref_pkg or is_exp):
            continue
        blk = getattr(mod, "MoPBlock", None)
        if isinstance(blk, type) and not blk.__module__.startswith("mop_b200"):
            _ORIGINALS[(modname, "MoPBlock")] = blk
            setattr(mod, "MoPBlock", _fused_mop_block(blk))
            done.setdefault(modname, []).append("MoPBlock")
        vm = getattr(mod, "ViT_MoP", None)
        if isinstance(vm, type) and not vm.__module__.startswith("mop_b200"):
            _ORIGINALS[(modname, "ViT_MoP")] = vm
            setattr(mod, "ViT_MoP", _fused_vit_mop(vm))
            done.setdefault(modname, []).append("ViT_MoP")
        for name, repl in _CANONICAL.items():
            cur = getattr(mod, name, None)
            if cur is None or not isinstance(cur, type) or cur.__module__.startswith("mop_b200"):
                continue
            if is_exp and name in ("EdgewiseMSA", "EdgewiseGateHead"):
                repl = _compat(repl)
            _ORIGINALS[(modname, name)] = cur
            setattr(mod, name, repl)
            done.setdefault(modname, []).append(name)
    if verbose:
        for m, names in done.items():
            print(f"mop_b200.dropin: {m}: {', '.join(names)}")
    return done

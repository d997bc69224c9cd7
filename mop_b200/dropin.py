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

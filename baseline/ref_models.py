"""Reference-arm model assembly: the reference's OWN modules (imported from baseline/_ref), nothing of mop_b200's kernels.

The reference's `ViTEdgewise` container lives in an experiment script (experiments/cifar100_edgewise_gates.py:326-451),
which is not part of the installable package.  Its attention class there is a clone of
`mop.models.attention_variants.EdgewiseMSA` with identical state_dict keys and bit-identical outputs, so
the reference arm builds the same container shape out of the package's own classes:
`mop.models.attention_variants.EdgewiseMSA` + `mop.models.components.{PatchEmbed, MLP, DropPath}` + `nn.LayerNorm`.
Every op on the hot path (attention_variants.py:453-564) is the unmodified reference code.
"""
from __future__ import annotations

import contextlib
import importlib
import sys
import types

from . import stage_ref


def import_reference():
    """Import the reference package from baseline/_ref; returns dict of modules or raises ImportError."""
    path = stage_ref.import_path()
    if path is None:
        raise ImportError("baseline/_ref is not staged (python baseline/stage_ref.py)")
    if path not in sys.path:
        sys.path.insert(0, path)
    for n in ("matplotlib", "matplotlib.pyplot"):   # optional plotting dependency of mop.visualization, absent here
        sys.modules.setdefault(n, types.ModuleType(n))
    mods = {}
    for short, name in (("av", "mop.models.attention_variants"), ("comp", "mop.models.components"),
                        ("qp", "mop.models.quartet_attn_patch"), ("wm", "mop.models.whisper_mop")):
        mods[short] = importlib.import_module(name)
    if not mods["av"].__file__.startswith(path):
        raise ImportError(f"`mop` resolved to {mods['av'].__file__}, not to baseline/_ref")
    return mods


@contextlib.contextmanager
def _rebind(module, **names):
    old = {k: getattr(module, k) for k in names}
    try:
        for k, v in names.items():
            setattr(module, k, v)
        yield
    finally:
        for k, v in old.items():
            setattr(module, k, v)


def reference_vit_edgewise(**kw):
    """`ViTEdgewise(**kw)` whose blocks hold the reference's EdgewiseMSA / MLP / DropPath / PatchEmbed."""
    ref = import_reference()
    import mop_b200.vit_edgewise as ve   # container only (LayerNorm / residual wiring); its classes are rebound below

    def ref_attn(*a, compat_experiments_init=None, **k):
        return ref["av"].EdgewiseMSA(*a, **k)

    kw.pop("compat_experiments_init", None)
    with _rebind(ve, EdgewiseMSA=ref_attn, MLP=ref["comp"].MLP, DropPath=ref["comp"].DropPath, PatchEmbed=ref["comp"].PatchEmbed):
        model = ve.ViTEdgewise(**kw)
    assert type(model.blocks[0].attn).__module__ == "mop.models.attention_variants"
    return model

"""Drop-in boundary: same constructor signatures / state_dict keys / init as the reference, and in-place patching.
Needs the read-only reference tree (authoring container); skipped where it is absent (GPU box)."""
import inspect
import os
import sys
import types

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "experiments"))
    for n in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(n, types.ModuleType(n))
    import mop.models.attention_variants as av
    import mop.models.components as comp
    import mop.models.quartet_attn_patch as qp
    import mop.models.whisper_mop as wm
    import cifar100_edgewise_gates as exp
    return dict(av=av, comp=comp, qp=qp, wm=wm, exp=exp)


def _sig(cls):
    return [(n, p.default) for n, p in inspect.signature(cls.__init__).parameters.items() if n != "self"]


def test_constructor_signatures_are_supersets(ref):
    import mop_b200 as m
    pairs = [(ref["av"].EdgewiseMSA, m.EdgewiseMSA), (ref["av"].EdgewiseGateHead, m.EdgewiseGateHead), (ref["av"].BaselineMSA, m.BaselineMSA),
             (ref["comp"].MSA, m.MSA), (ref["qp"].CausalSelfAttention, m.CausalSelfAttention),
             (ref["wm"].MultiheadSelfAttention, m.MultiheadSelfAttention), (ref["wm"].MultiheadCrossAttention, m.MultiheadCrossAttention),
             (ref["exp"].ViTEdgewise, m.ViTEdgewise)]
    for r, o in pairs:
        rs, os_ = _sig(r), _sig(o)
        assert os_[:len(rs)] == rs, (r.__name__, rs, os_)   # same names, order and defaults; ours may append extras


@pytest.mark.parametrize("kw", [
    dict(n_views=5, share_qkv=True, gate_mode="lowrank", gate_rank=4, gate_init="mix5", use_k3=True),
    dict(n_views=3, share_qkv=False, gate_mode="dense", use_k3=True, gate_init="and"),
    dict(n_views=2, share_qkv=True, gate_mode="dense", gate_init="xor", use_lens_bank_qk=True, lens_qk_causal=True)])
def test_same_seed_same_parameters(ref, kw):
    import mop_b200 as m
    torch.manual_seed(3); a = ref["av"].EdgewiseMSA(32, heads=4, **kw)
    torch.manual_seed(3); b = m.EdgewiseMSA(32, heads=4, **kw)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    assert all(torch.equal(sa[k], sb[k]) for k in sa)


def test_other_modules_state_dict_keys(ref):
    import mop_b200 as m
    cfg_r = ref["qp"].TransformerConfig(n_head=2, n_embd=16, block_size=8, bias=True)
    cfg_o = m.TransformerConfig(n_head=2, n_embd=16, block_size=8, bias=True)
    assert list(ref["qp"].CausalSelfAttention(cfg_r).state_dict()) == list(m.CausalSelfAttention(cfg_o).state_dict())
    assert list(ref["wm"].MultiheadCrossAttention(16, 24, 2, 0.0, True).state_dict()) == list(m.MultiheadCrossAttention(16, 24, 2, 0.0, True).state_dict())
    assert list(ref["comp"].MSA(16, 2).state_dict()) == list(m.MSA(16, 2).state_dict())


def test_patch_reference_rebinds_and_models_build(ref):
    import mop_b200.dropin as dropin
    done = dropin.patch_reference()
    assert "EdgewiseMSA" in done.get("mop.models.attention_variants", [])
    assert "EdgewiseMSA" in done.get("cifar100_edgewise_gates", [])
    vit = ref["exp"].ViTEdgewise(dim=32, depth=2, heads=2, n_views=3, share_qkv=True, gate_mode="lowrank", gate_init="mix5")
    assert type(vit.blocks[0].attn).__module__.startswith("mop_b200")
    # experiments copy ignores the mix5 preset: biases stay zero there
    assert float(vit.blocks[0].attn.edge_head.row_proj.bias.abs().sum()) == 0.0
    from mop.models import ViT_Baseline
    base = ViT_Baseline(dim=32, depth=1, heads=2, n_classes=10)
    assert type(base.enc.blocks[0].attn).__module__.startswith("mop_b200") if hasattr(base, "enc") else True
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            vit(torch.randn(1, 3, 32, 32))

"""Oracle vs golden vectors recorded from the imported reference (CPU).

The fixtures were produced by tests/golden/make_golden.py from the unmodified
reference modules evaluated in fp64.  Every oracle function must reproduce
them (outputs and all gradients) to fp64 round-off.
"""
import glob
import os

import pytest
import torch

from conftest import GOLDEN, load_golden
from oracle.edgewise import EdgewiseConfig, edgewise_msa, gate_bias_preset
from oracle.quartet import quartet_module
from oracle.sdpa import msa_module, whisper_cross_module, whisper_self_module
from oracle.variants_cd import crossview_module, multihop_module

TOL = 1e-11
CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, "*.pt")) if "gate_presets" not in p)


def _run(case, ins, sd):
    kind = case["kind"]
    if kind == "edgewise":
        cfg = EdgewiseConfig(dim=case["dim"], heads=case["heads"], **case["kwargs"])
        return edgewise_msa(ins["x"], sd, cfg)
    if kind == "msa":
        return msa_module(ins["x"], sd, case["heads"])
    if kind == "baseline_msa":
        return msa_module(ins["x"], sd, case["heads"], attn_mask=case["mask"].double())
    if kind == "whisper_self":
        return whisper_self_module(ins["x"], sd, case["n_head"], case["causal"], case["bias"].double())
    if kind == "whisper_cross":
        return whisper_cross_module(ins["x_q"], ins["x_kv"], sd, case["n_head"])
    if kind == "quartet":
        am = case["add_mask"]
        return quartet_module(ins["x"], sd, case["n_head"], case["use_quartet"], case["eps"],
                              None if am is None else am.double())
    if kind == "crossview":
        m = case["mask"]
        return crossview_module(ins["x"], sd, case["heads"], attn_mask=None if m is None else m.double(), **case["kwargs"])
    if kind == "multihop":
        return multihop_module(ins["x"], sd, case["heads"], **case["kwargs"])
    raise AssertionError(kind)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    case = load_golden(name)
    ins = {k: v.double().requires_grad_(True) for k, v in case["inputs"].items()}
    sd = {k: v.double().requires_grad_(v.is_floating_point()) for k, v in case["state_dict"].items()}
    y = _run(case, ins, sd)
    assert y.shape == case["y"].shape
    assert (y - case["y"]).abs().max().item() < TOL
    wanted = {k: v for k, v in sd.items() if k in case["dparams"]}
    grads = torch.autograd.grad(y, list(ins.values()) + list(wanted.values()), case["dy"].double(), allow_unused=True)
    for (k, _), g in zip(list(ins.items()) + list(wanted.items()), grads):
        ref = case["dinputs"][k] if k in case["dinputs"] else case["dparams"][k]
        g = torch.zeros_like(ref) if g is None else g
        scale = max(1.0, ref.abs().max().item())
        assert (g - ref).abs().max().item() < TOL * scale, k


def test_gate_presets_match_reference_constructors():
    table = load_golden("gate_presets")
    for key, ref in table.items():
        tag, mode, init = key.split("/")
        _, b = gate_bias_preset(mode, init, rank=3, compat_experiments=(tag == "experiments"))
        assert torch.equal(b, ref), key

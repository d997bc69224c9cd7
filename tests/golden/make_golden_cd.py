"""Golden fixtures for attention variants C (CrossViewMixerMSA) and D (MultiHopMSA), generated FROM THE REFERENCE ITSELF
(mop/models/attention_variants.py:51-231) with the helpers of make_golden.py.  Authoring container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_cd.py [/root/reference]
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import record, redraw  # noqa: E402  (also puts the reference tree on sys.path)

from mop.models.attention_variants import CrossViewMixerMSA, MultiHopMSA  # noqa: E402


def crossview_cases():
    cases = [
        ("crossview_default", 32, 2, 12, dict(), False),
        ("crossview_mix_t_mask", 48, 3, 10, dict(t1=0.3, t2=-0.2), True),
        ("crossview_dk56", 112, 2, 9, dict(), False),
    ]
    for i, (name, dim, H, N, kw, masked) in enumerate(cases):
        torch.manual_seed(500 + i)
        mod = CrossViewMixerMSA(dim, heads=H, **kw)
        redraw(mod, 600 + i)
        with torch.no_grad():
            mod.mix.copy_(torch.tensor([[0.9, 0.3], [-0.4, 1.1]]))
        mask = None
        if masked:
            mask = (torch.rand(2, 1, N, N) > 0.3).float()
            mask[..., 0] = 1.0
        record(name, mod, {"x": torch.randn(2, N, dim)}, lambda m, t: m(t["x"], None if mask is None else mask.double()),
               dict(kind="crossview", dim=dim, heads=H, kwargs=kw, mask=mask))


def multihop_cases():
    cases = [
        ("multihop_default", 32, 2, 12, dict()),
        ("multihop_gates_h4", 48, 3, 10, dict(hops=4, beta_not=0.8, gates=dict(and_=0.7, or_=0.4, not_=0.3, chain=0.5, base=1.0))),
        ("multihop_h2_chain", 32, 2, 9, dict(hops=2, gates=dict(and_=1.0, or_=0.0, not_=0.0, chain=1.0, base=1.0))),
    ]
    for i, (name, dim, H, N, kw) in enumerate(cases):
        torch.manual_seed(700 + i)
        mod = MultiHopMSA(dim, heads=H, **kw)
        redraw(mod, 800 + i)
        record(name, mod, {"x": torch.randn(2, N, dim)}, lambda m, t: m(t["x"]), dict(kind="multihop", dim=dim, heads=H, kwargs=kw))


if __name__ == "__main__":
    crossview_cases()
    multihop_cases()

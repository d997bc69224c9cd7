"""mop_b200.mixed.FlatParams on the CPU: parameters re-homed into one flat buffer train exactly like separate parameters."""
import copy

import torch
import torch.nn as nn
import torch.nn.functional as F


def test_flat_params_trains_like_separate_parameters():
    from mop_b200.mixed import FlatParams
    torch.manual_seed(0)
    m1 = nn.Sequential(nn.Linear(8, 16), nn.GELU(), nn.LayerNorm(16), nn.Linear(16, 4))
    m2 = copy.deepcopy(m1)
    fp = FlatParams(m2)
    assert all(p.data_ptr() >= fp.param.data_ptr() for p in m2.parameters())   # views of the flat buffer
    o1 = torch.optim.AdamW(m1.parameters(), lr=1e-2, weight_decay=0.05)
    o2 = torch.optim.AdamW([fp.param], lr=1e-2, weight_decay=0.05)
    x, y = torch.randn(32, 8), torch.randint(0, 4, (32,))
    for it in range(5):
        o1.zero_grad(set_to_none=True)
        F.cross_entropy(m1(x), y).backward()
        o1.step()
        fp.begin()
        o2.zero_grad(set_to_none=True)          # must not detach the flat gradient buffer from the flat parameter
        F.cross_entropy(m2(x), y).backward()
        fp.pack_grads()
        assert fp.param.grad is fp.grad_buf
        o2.step()
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert torch.allclose(a, b, atol=1e-6), (a - b).abs().max()

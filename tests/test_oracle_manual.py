"""Hand-derived backward passes (the kernel specification) vs autograd, fp64, CPU."""
import pytest
import torch

from oracle.edgewise import edgewise_core
from oracle.edgewise_manual import edgewise_packed
from oracle.quartet import quartet_core, quartet_core_manual

DT = torch.float64


def make_head(mode, k3, C, r, gen):
    rn = lambda *s: torch.randn(*s, dtype=DT, generator=gen)
    if mode == "lowrank":
        return {"row_proj.weight": 0.3 * rn(4 * r, C, 1), "row_proj.bias": 0.3 * rn(4 * r),
                "col_proj.weight": 0.3 * rn(4 * r, C, 1), "col_proj.bias": 0.3 * rn(4 * r)}
    h = {"conv1.weight": 0.3 * rn(16, C, 1, 1), "conv1.bias": 0.3 * rn(16),
         "conv2.weight": 0.3 * rn(4, 16, 1, 1), "conv2.bias": 0.3 * rn(4)}
    if k3:
        h["mid3.weight"] = 0.2 * rn(16, 16, 3, 3)
        h["mid3.bias"] = 0.2 * rn(16)
    return h


def core_from_packed(qkv, qs_, ks_, vs_, head, logit, V, mode, r, beta):
    Vp = qkv.shape[2]
    qs, ks = [], []
    for i in range(V):
        src = qkv[:, :, 0 if Vp == 1 else i]
        q = src[:, :, 0].permute(0, 2, 1, 3)
        k = src[:, :, 1].permute(0, 2, 1, 3)
        if qs_ is not None:
            q = q * qs_[i][None, :, None, :]
            k = k * ks_[i][None, :, None, :]
        qs.append(q); ks.append(k)
    v1 = qkv[:, :, 0, 2].permute(0, 2, 1, 3)
    vL = qkv[:, :, 0 if Vp == 1 else V - 1, 2].permute(0, 2, 1, 3)
    if vs_ is not None:
        v1 = v1 * vs_[0][None, :, None, :]
        vL = vL * vs_[V - 1][None, :, None, :]
    y = edgewise_core(qs, ks, v1, vL, head, logit, beta_not=beta, gate_mode=mode, gate_rank=r)
    return y.permute(0, 2, 1, 3)


@pytest.mark.parametrize("mode,k3,shared,V", [
    ("lowrank", False, True, 4), ("lowrank", False, False, 3), ("lowrank", False, True, 2),
    ("dense", False, True, 3), ("dense", True, True, 5), ("dense", True, False, 2)])
def test_edgewise_manual_backward(mode, k3, shared, V):
    gen = torch.Generator().manual_seed(V * 7 + len(mode))
    B, H, N, dk, r = 2, 3, 10, 8, 3
    Vp = 1 if shared else V
    qkv = torch.randn(B, N, Vp, 3, H, dk, dtype=DT, generator=gen).requires_grad_()
    sc = [(1 + 0.2 * torch.randn(V, H, dk, dtype=DT, generator=gen)).requires_grad_() for _ in range(3)] if shared else [None] * 3
    head = {k: v.requires_grad_() for k, v in make_head(mode, k3, 2 * V + 2, r, gen).items()}
    logit = torch.tensor(-0.7, dtype=DT, requires_grad=True)
    dy = torch.randn(B, N, H, dk, dtype=DT, generator=gen)
    y = core_from_packed(qkv, *sc, head, logit, V, mode, r, 0.5)
    ins = [qkv, logit] + list(head.values()) + ([*sc] if shared else [])
    ref = torch.autograd.grad(y, ins, dy)
    det = lambda t: None if t is None else t.detach()
    y2, g = edgewise_packed(det(qkv), det(sc[0]), det(sc[1]), det(sc[2]), {k: v.detach() for k, v in head.items()},
                            logit.detach(), V=V, beta_not=0.5, gate_mode=mode, gate_rank=r, dy=dy)
    assert (y - y2).abs().max() < 1e-12
    names = ["qkv", "logit"] + list(head) + (["q_scale", "k_scale", "v_scale"] if shared else [])
    for n, a in zip(names, ref):
        assert (a - g[n]).abs().max() < 1e-11, n


@pytest.mark.parametrize("quart,with_mask", [(True, False), (True, True), (False, False)])
def test_quartet_manual_backward(quart, with_mask):
    gen = torch.Generator().manual_seed(3)
    B, H, T, d = 2, 3, 12, 8
    rn = lambda: torch.randn(B, H, T, d, dtype=DT, generator=gen).requires_grad_()
    q, k, v = rn(), rn(), rn()
    q2, k2 = (rn(), rn()) if quart else (None, None)
    mix = torch.tensor([0.3], dtype=DT, requires_grad=True)
    gam = torch.tensor([1.3], dtype=DT, requires_grad=True)
    am = 0.4 * torch.randn(B, 1, T, T, dtype=DT, generator=gen) if with_mask else None
    dy = torch.randn(B, H, T, d, dtype=DT, generator=gen)
    y = quartet_core(q, k, v, q2, k2, mix, gam, add_mask=am)
    ins = {"q": q, "k": k, "v": v}
    if quart:
        ins.update({"q2": q2, "k2": k2, "mixture": mix, "quartet_scale": gam})
    ref = torch.autograd.grad(y, list(ins.values()), dy)
    y2, g = quartet_core_manual(q.detach(), k.detach(), v.detach(), None if q2 is None else q2.detach(),
                                None if k2 is None else k2.detach(), mix.detach(), gam.detach(), add_mask=am, dy=dy)
    assert (y - y2).abs().max() < 1e-12
    for n, a in zip(ins, ref):
        assert (a - g[n]).abs().max() < 1e-11, n

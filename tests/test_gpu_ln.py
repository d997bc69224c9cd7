"""GPU parity of the fused residual-add + DropPath scale + LayerNorm kernels vs a plain PyTorch fp32/fp64
reference of the same ops (experiments/cifar100_edgewise_gates.py:371-374, mop/models/components.py:14-27)."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import max_abs, rel_to_max

pytestmark = pytest.mark.gpu


def _ref(x, r, scale, g, b, eps):
    xn = x if r is None else x + (r if scale is None else r * scale.view(-1, 1, 1))
    return xn, F.layer_norm(xn, (x.shape[-1],), g, b, eps)


@pytest.mark.parametrize("B,N,D,with_r,with_scale,rdt,ydt", [
    (8, 64, 224, True, True, torch.bfloat16, torch.bfloat16),     # bench block shape (bf16 branch, bf16 output)
    (3, 196, 768, True, True, torch.float32, torch.float32),      # ViT-B/16 width, fp32 everywhere
    (2, 17, 1000, True, False, torch.float32, torch.float32),     # ragged D, no DropPath factor
    (4, 64, 224, False, False, torch.float32, torch.float32),     # plain LayerNorm (first block)
    (1, 1, 8, True, True, torch.float32, torch.bfloat16),
    (300, 10, 96, True, True, torch.bfloat16, torch.float32),     # more rows than the persistent grid covers in one sweep
    (2, 33, 60, True, True, torch.bfloat16, torch.bfloat16),      # D % 8 != 0: the scalar kernels
    (2, 9, 500, True, True, torch.float32, torch.bfloat16),
    (5, 64, 512, True, True, torch.bfloat16, torch.bfloat16),     # two 256-feature steps per lane
])
def test_add_layer_norm_vs_torch(B, N, D, with_r, with_scale, rdt, ydt):
    from mop_b200 import functional as MF
    gen = torch.Generator().manual_seed(B * 1000 + D)
    rn = lambda *s: torch.randn(*s, generator=gen, dtype=torch.float64)
    x, g, b = rn(B, N, D), 1 + 0.2 * rn(D), 0.1 * rn(D)
    r = rn(B, N, D).to(rdt).double() if with_r else None
    scale = (torch.rand(B, generator=gen) > 0.3).double() / 0.7 if with_scale else None
    dxn, dy = rn(B, N, D), rn(B, N, D).to(ydt).double()
    ins = [t.clone().requires_grad_(True) for t in (x, g, b)] + ([r.clone().requires_grad_(True)] if with_r else [])
    xn_ref, y_ref = _ref(ins[0], ins[3] if with_r else None, scale, ins[1], ins[2], 1e-5)
    loss = (y_ref * dy).sum() + ((xn_ref * dxn).sum() if with_r else 0.0)
    g_ref = torch.autograd.grad(loss, ins)

    xg, gg, bg = (t.float().cuda().requires_grad_(True) for t in (x, g, b))
    rg = r.to(rdt).cuda().requires_grad_(True) if with_r else None
    sg = scale.float().cuda() if with_scale else None
    xn, y = MF.add_layer_norm(xg, rg, sg, gg, bg, 1e-5, out_dtype=ydt)
    assert y.dtype == ydt and xn.dtype == torch.float32
    loss = (y.double() * dy.cuda()).sum() + ((xn.double() * dxn.cuda()).sum() if with_r else 0.0)
    loss.backward()
    ytol = 1e-5 if ydt == torch.float32 else 1e-2
    assert max_abs(y, y_ref) <= ytol * max(1.0, y_ref.abs().max().item())
    if with_r:
        assert max_abs(xn, xn_ref) <= 1e-5
    got = [xg.grad, gg.grad, bg.grad] + ([rg.grad] if with_r else [])
    names = ["x", "gamma", "beta", "r"]
    for n, a, ref in zip(names, got, g_ref):
        tol = 2e-5 if (rdt == torch.float32 or n != "r") else 1e-2
        assert max_abs(a, ref) <= tol * max(1.0, ref.abs().max().item()), n


def test_vit_fused_stream_matches_block_composition():
    """ViTEdgewise.forward (fused residual stream) == the plain block composition `x + dp(attn(ln1 x))`, `x + dp(mlp(ln2 x))`."""
    import mop_b200
    torch.manual_seed(0)
    m = mop_b200.ViTEdgewise(dim=64, depth=3, heads=2, n_classes=10, mlp_ratio=2.0, n_views=3, share_qkv=True, gate_mode="lowrank",
                             gate_rank=2, gate_init="mix5", drop_path=0.0, compat_experiments_init=False).cuda()
    x = torch.randn(4, 3, 32, 32, device="cuda")
    y_fused = m(x)
    tok, _ = m.patch(x)
    tok = tok + m.pos
    for blk in m.blocks:
        tok = blk(tok)
    y_plain = m.head(m.ln_f(tok).mean(dim=1))
    assert max_abs(y_fused, y_plain) <= 2e-5
    gf = torch.autograd.grad(y_fused.square().sum(), list(m.parameters()), retain_graph=False)
    gp = torch.autograd.grad(y_plain.square().sum(), list(m.parameters()))
    for (n, _), a, b in zip(m.named_parameters(), gf, gp):
        assert max_abs(a, b) <= 5e-5 * max(1.0, b.abs().max().item()), n

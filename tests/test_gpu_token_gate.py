"""ViT-MoP token gate: the fused kernels against the reference's composition of ViewsLinear / Kernels3 /
FuseExcInh (mop/models/components.py:255-303, vit_mop.py:95-114) restated with torch ops in fp64 (oracle/token_gate.py)."""
import pytest
import torch

from gpu_util import max_abs

pytestmark = pytest.mark.gpu


def _weights(V, K, D, gen, dev):
    hid = max(8, V + K)
    rn = lambda *s: torch.randn(*s, generator=gen, device=dev)
    return dict(views_w=rn(V, D) / D ** 0.5, k3_w=rn(16, V, 3, 3) / (3 * V ** 0.5), k1_w=rn(K, 16, 1, 1) / 4, f1_w=rn(hid, V + K, 1, 1) / (V + K) ** 0.5,
                f2_w=rn(2, hid, 1, 1) / hid ** 0.5, f2_b=0.3 * rn(2), alpha_pos=torch.tensor(0.8, device=dev), alpha_neg=torch.tensor(0.6, device=dev))


@pytest.mark.parametrize("B,grid,D,V,K,dtype,ctas", [
    (5, (8, 8), 64, 5, 3, torch.float32, 0),
    (7, (8, 8), 256, 5, 3, torch.float32, 3),      # several images per CTA: the partial sums accumulate over the persistent loop
    (3, (14, 14), 768, 5, 3, torch.float32, 2),    # ViT-B/16 grid
    (4, (6, 9), 224, 3, 2, torch.float32, 0),      # non-square grid, other channel counts (hid = 8 > V + K)
    (4, (4, 4), 96, 8, 8, torch.float32, 0),       # maximum channel counts (hid = 16)
    (6, (8, 8), 256, 5, 3, torch.bfloat16, 4),
])
def test_token_gate_vs_oracle(B, grid, D, V, K, dtype, ctas):
    import torch.nn.functional as F
    import mop_b200
    from oracle.token_gate import token_gate_ref
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(B * 131 + D)
    w = _weights(V, K, D, gen, dev)
    for t in w.values():
        t.requires_grad_(True)
    T = grid[0] * grid[1]
    tok = torch.randn(B, T, D, generator=gen, device=dev).to(dtype).requires_grad_(True)
    dy = torch.randn(B, T, D, generator=gen, device=dev).to(dtype)
    out = mop_b200.functional.token_gate(tok, grid, w["views_w"], w["k3_w"], w["k1_w"], w["f1_w"], w["f2_w"], w["f2_b"],
                                         F.softplus(w["alpha_pos"]), F.softplus(w["alpha_neg"]), _max_ctas=ctas)
    out.backward(dy)
    got = {k: v.grad.double() for k, v in w.items()}
    got_dx = tok.grad.double()
    # oracle: fp64 on the same (bf16-rounded) inputs
    tok64 = tok.detach().double().requires_grad_(True)
    w64 = {k: v.detach().double().requires_grad_(True) for k, v in w.items()}
    ref = token_gate_ref(tok64, grid, **w64)
    ref.backward(dy.double())
    tol = 2e-2 if dtype == torch.bfloat16 else 2e-5
    scale = lambda r: max(1.0, r.abs().max().item())
    assert max_abs(out.double(), ref.detach()) <= tol * scale(ref.detach())
    assert max_abs(got_dx, tok64.grad) <= tol * scale(tok64.grad)
    for k in w:
        assert max_abs(got[k], w64[k].grad) <= tol * scale(w64[k].grad), k


def test_token_gate_rejects_bad_shapes():
    import mop_b200
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(0)
    w = _weights(5, 3, 60, gen, dev)
    tok = torch.randn(2, 64, 60, device=dev)
    with pytest.raises(RuntimeError):
        mop_b200.functional.token_gate(tok, (8, 8), w["views_w"], w["k3_w"], w["k1_w"], w["f1_w"], w["f2_w"], w["f2_b"],
                                       torch.ones(1, device=dev), torch.ones(1, device=dev))


@pytest.mark.parametrize("B,T,D,V,K,dtype,ctas", [
    (2, 50, 64, 5, 3, torch.float32, 0),
    (3, 200, 256, 5, 3, torch.float32, 2),       # several chunks per sequence and per CTA: halos, partial accumulation
    (2, 129, 768, 4, 2, torch.float32, 0),       # chunk boundary + 1
    (1, 64, 96, 8, 8, torch.float32, 0),
    (4, 300, 768, 5, 3, torch.bfloat16, 3),
])
def test_token_gate_1d_vs_oracle(B, T, D, V, K, dtype, ctas):
    """GPT-MoP apply_mop (gpt_mop.py:102-123): fused kernel vs the reference's composition in fp64."""
    import torch.nn.functional as F
    import mop_b200
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(T * 7 + D)
    rn = lambda *s: torch.randn(*s, generator=gen, device=dev)
    w = dict(views_w=rn(V, D) / D ** 0.5, kernels_w=rn(K, V, 3) / (3 * V) ** 0.5, fuse_w=rn(2, V + K, 1) / (V + K) ** 0.5,
             alpha=torch.tensor([0.9, 0.7], device=dev))
    for t in w.values():
        t.requires_grad_(True)
    x = rn(B, T, D).to(dtype).requires_grad_(True)
    dy = rn(B, T, D).to(dtype)
    out = mop_b200.functional.token_gate_1d(x, w["views_w"], w["kernels_w"], w["fuse_w"], w["alpha"], _max_ctas=ctas)
    out.backward(dy)

    x64 = x.detach().double().requires_grad_(True)
    w64 = {k: v.detach().double().requires_grad_(True) for k, v in w.items()}
    views = (x64 @ w64["views_w"].t()).transpose(1, 2)                       # ViewsLinear1D
    kmaps = F.conv1d(views, w64["kernels_w"], padding=1)                     # Kernels1D
    g = F.conv1d(torch.cat([views, kmaps], dim=1), w64["fuse_w"])            # FuseExcInh1D
    gate = 1 + w64["alpha"][0] * g[:, :1] - w64["alpha"][1] * g[:, 1:]
    ref = x64 * gate.transpose(1, 2)
    ref.backward(dy.double())
    tol = 2e-2 if dtype == torch.bfloat16 else 2e-5
    scale = lambda r: max(1.0, r.abs().max().item())
    assert max_abs(out.double(), ref.detach()) <= tol * scale(ref.detach())
    assert max_abs(x.grad.double(), x64.grad) <= tol * scale(x64.grad)
    for k in w:
        assert max_abs(w[k].grad.double(), w64[k].grad) <= tol * scale(w64[k].grad), k

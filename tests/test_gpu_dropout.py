"""In-kernel attention dropout (reference quartet_attn_patch.py:24,118-119; whisper_mop.py:173,217; components.py:63):
statistics of the keep mask, and exact forward/backward parity against the oracle evaluated under the kernels' own mask."""
import pytest
import torch

from gpu_util import bf16_round, max_abs, rel_to_max, scaled_tol

pytestmark = pytest.mark.gpu


def test_mask_statistics_and_determinism():
    from mop_b200 import functional as MF
    for p in (0.1, 0.5):
        m = MF.dropout_mask(12, 300, 257, p, seed=1234, offset=77)
        kept = (m > 0).float()
        assert abs(kept.mean().item() - (1 - p)) < 3e-3
        assert torch.all((m == 0) | ((m - 1 / (1 - p)).abs() < 1e-6))
        assert abs(m.mean().item() - 1.0) < 5e-3                       # E[mask] = 1
        # rows / columns / problems are not correlated: per-row keep rates scatter like a binomial
        row_rate = kept.mean(-1)
        assert abs(row_rate.std().item() - (p * (1 - p) / 257) ** 0.5) < 0.2 * (p * (1 - p) / 257) ** 0.5
        assert torch.equal(m, MF.dropout_mask(12, 300, 257, p, seed=1234, offset=77))
        assert not torch.equal(m, MF.dropout_mask(12, 300, 257, p, seed=1234, offset=78))
    assert torch.all(MF.dropout_mask(2, 8, 8, 0.0, 1, 1) == 1)


@pytest.mark.parametrize("B,H,Nq,Nk,dk,causal,dtype,impl", [
    (2, 2, 70, 90, 32, False, torch.float32, "simt"), (1, 2, 130, 130, 64, True, torch.float32, "simt"),
    (2, 3, 200, 200, 64, False, torch.bfloat16, "tcgen05"), (1, 2, 257, 257, 64, True, torch.bfloat16, "tcgen05"),
    (2, 2, 64, 64, 54, False, torch.bfloat16, "simt")])
def test_sdpa_dropout_vs_oracle_under_same_mask(B, H, Nq, Nk, dk, causal, dtype, impl):
    from mop_b200 import functional as MF
    from mop_b200 import sdpa
    from oracle.sdpa import sdpa_core
    g = torch.Generator().manual_seed(Nq + 7 * Nk)
    mk = lambda n: torch.randn(B, n, H, dk, generator=g, dtype=torch.float64)
    q, k, v, dy = mk(Nq), mk(Nk), mk(Nk), mk(Nq)
    if dtype == torch.bfloat16:
        q, k, v, dy = map(bf16_round, (q, k, v, dy))
    qg, kg, vg = (t.to("cuda", dtype).requires_grad_(True) for t in (q, k, v))
    y = sdpa(qg, kg, vg, causal=causal, dropout_p=0.3, impl=impl)
    y.backward(dy.to("cuda", dtype))
    assert MF.last_impl["sdpa_fwd"] == impl and MF.last_impl["sdpa_bwd"] == impl
    p, seed, off = MF.last_dropout["sdpa"]
    mask = MF.dropout_mask(B * H, Nq, Nk, p, seed, off).view(B, H, Nq, Nk).double().cpu()
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    y_ref = sdpa_core(qr.transpose(1, 2), kr.transpose(1, 2), vr.transpose(1, 2), causal=causal, drop_mask=mask).transpose(1, 2)
    g_ref = torch.autograd.grad(y_ref, (qr, kr, vr), dy)
    if dtype == torch.float32:
        assert max_abs(y, y_ref) <= scaled_tol(y_ref, 1e-5)
        for a, b in zip((qg.grad, kg.grad, vg.grad), g_ref):
            assert max_abs(a, b) <= scaled_tol(b, 1e-5)
    else:
        assert rel_to_max(y, y_ref) <= 2e-2
        for a, b in zip((qg.grad, kg.grad, vg.grad), g_ref):
            assert rel_to_max(a, b) <= 2e-2
    # eval mode / p = 0 is the identity: no mask is drawn
    y0 = sdpa(qg, kg, vg, causal=causal, dropout_p=0.0, impl=impl)
    y_plain = sdpa_core(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), causal=causal).transpose(1, 2)
    assert rel_to_max(y0, y_plain) <= (1e-5 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize("B,H,T,dk,dtype,impl", [(2, 2, 150, 32, torch.float32, "simt"), (4, 4, 257, 64, torch.bfloat16, "tcgen05")])
def test_quartet_dropout_vs_oracle_under_same_mask(B, H, T, dk, dtype, impl):
    from mop_b200 import functional as MF
    from mop_b200 import quartet_attention
    from oracle.quartet import quartet_core
    g = torch.Generator().manual_seed(T + dk)
    mk = lambda: torch.randn(B, T, H, dk, generator=g, dtype=torch.float64)
    ts = [mk() for _ in range(6)]
    if dtype == torch.bfloat16:
        ts = [bf16_round(t) for t in ts]
    q, k, v, q2, k2, dy = ts
    mix, gam = torch.tensor([0.4], dtype=torch.float64), torch.tensor([1.2], dtype=torch.float64)
    ins = [q, k, v, q2, k2, mix, gam]
    gin = [t.to("cuda", dtype if t.dim() == 4 else torch.float32).requires_grad_(True) for t in ins]
    y = quartet_attention(*gin, dropout_p=0.2, impl=impl)
    y.backward(dy.to("cuda", dtype))
    assert MF.last_impl["quartet_fwd"] == impl and MF.last_impl["quartet_bwd"] == impl
    p, seed, off = MF.last_dropout["quartet"]
    mask = MF.dropout_mask(B * H, T, T, p, seed, off).view(B, H, T, T).double().cpu()
    ref_in = [t.clone().requires_grad_(True) for t in ins]
    tr = lambda t: t.transpose(1, 2)
    y_ref = tr(quartet_core(tr(ref_in[0]), tr(ref_in[1]), tr(ref_in[2]), tr(ref_in[3]), tr(ref_in[4]), ref_in[5], ref_in[6], drop_mask=mask))
    g_ref = torch.autograd.grad(y_ref, ref_in, dy)
    if dtype == torch.float32:
        assert max_abs(y, y_ref) <= scaled_tol(y_ref, 1e-5)
        for a, b in zip(gin, g_ref):
            assert max_abs(a.grad, b) <= scaled_tol(b, 2e-5)
    else:
        assert rel_to_max(y, y_ref) <= 2e-2
        for a, b in zip(gin[:5], g_ref[:5]):
            assert rel_to_max(a.grad, b) <= 2e-2


def test_default_config_gpt_quartet_trains_with_dropout():
    """TransformerConfig() keeps the reference default dropout = 0.1 (quartet_attn_patch.py:24): a training step must run."""
    from mop_b200 import CausalSelfAttention, TransformerConfig
    cfg = TransformerConfig(n_head=4, n_embd=64, block_size=96)
    assert cfg.dropout == 0.1
    m = CausalSelfAttention(cfg).cuda().train()
    x = torch.randn(2, 80, 64, device="cuda", requires_grad=True)
    torch.manual_seed(5)
    y1 = m(x)
    y1.square().mean().backward()
    assert torch.isfinite(x.grad).all() and all(torch.isfinite(p.grad).all() for p in m.parameters())
    torch.manual_seed(5)
    y2 = m(x)
    assert torch.equal(y1, y2)          # same torch seed -> same mask
    m.eval()
    y3, y4 = m(x), m(x)
    assert torch.equal(y3, y4) and not torch.equal(y3, y1)   # eval: identity dropout

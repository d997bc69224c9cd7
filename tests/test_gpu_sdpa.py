"""GPU parity of plain / causal / biased / cross attention vs the CPU oracle."""
import pytest
import torch

from conftest import load_golden
from gpu_util import bf16_round, max_abs, rel_to_max, scaled_tol

pytestmark = pytest.mark.gpu
FP32_TOL, PARAM_TOL, BF16_TOL = 1e-5, 2e-5, 2e-2


def _check_module(m, case, call):
    ins = {k: v.cuda().requires_grad_(True) for k, v in case["inputs"].items()}
    y = call(m, ins)
    assert max_abs(y, case["y"]) <= scaled_tol(case["y"], FP32_TOL)
    y.backward(case["dy"].cuda())
    for k, t in ins.items():
        assert max_abs(t.grad, case["dinputs"][k]) <= scaled_tol(case["dinputs"][k], FP32_TOL), k
    for k, p in m.named_parameters():
        ref = case["dparams"][k]
        assert max_abs(p.grad, ref) <= PARAM_TOL * max(1.0, ref.abs().max().item()), k


def test_msa_golden():
    from mop_b200 import MSA
    case = load_golden("msa_dk54")
    m = MSA(108, heads=2); m.load_state_dict(case["state_dict"]); m.cuda()
    _check_module(m, case, lambda mod, t: mod(t["x"]))


def test_baseline_msa_mask_golden():
    from mop_b200 import BaselineMSA
    case = load_golden("baseline_msa_mask")
    m = BaselineMSA(32, heads=4); m.load_state_dict(case["state_dict"]); m.cuda()
    _check_module(m, case, lambda mod, t: mod(t["x"], case["mask"].cuda()))


@pytest.mark.parametrize("causal", [0, 1])
def test_whisper_self_golden(causal):
    from mop_b200 import MultiheadSelfAttention
    case = load_golden(f"whisper_self_causal{causal}")
    m = MultiheadSelfAttention(32, 4, 0.0, True, causal=bool(causal)); m.load_state_dict(case["state_dict"]); m.cuda()
    _check_module(m, case, lambda mod, t: mod(t["x"], case["bias"].cuda()))


def test_whisper_cross_golden():
    from mop_b200 import MultiheadCrossAttention
    case = load_golden("whisper_cross")
    m = MultiheadCrossAttention(32, 48, 4, 0.0, False); m.load_state_dict(case["state_dict"]); m.cuda()
    _check_module(m, case, lambda mod, t: mod(t["x_q"], t["x_kv"]))


@pytest.mark.parametrize("B,H,Nq,Nk,dk,causal,dtype", [
    (2, 2, 64, 64, 56, False, torch.float32), (1, 2, 196, 196, 64, False, torch.float32),
    (1, 2, 1500, 1500, 64, False, torch.float32), (1, 2, 300, 300, 64, True, torch.float32),
    (2, 2, 37, 150, 32, False, torch.float32), (1, 1, 1, 1, 8, False, torch.float32),
    (1, 2, 196, 196, 64, False, torch.bfloat16), (1, 2, 257, 257, 64, True, torch.bfloat16),
    (2, 4, 64, 64, 54, False, torch.bfloat16)])   # model B: dk = 54 is not a multiple of 8 -> bf16 storage on the fp32-math kernels
def test_core_vs_oracle(B, H, Nq, Nk, dk, causal, dtype):
    from mop_b200 import sdpa
    from oracle.sdpa import sdpa_core
    g = torch.Generator().manual_seed(Nq + Nk)
    mk = lambda n: torch.randn(B, n, H, dk, generator=g, dtype=torch.float64)
    q, k, v, dy = mk(Nq), mk(Nk), mk(Nk), mk(Nq)
    if dtype == torch.bfloat16:
        q, k, v, dy = map(bf16_round, (q, k, v, dy))
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    y_ref = sdpa_core(qr.transpose(1, 2), kr.transpose(1, 2), vr.transpose(1, 2), causal=causal).transpose(1, 2)
    g_ref = torch.autograd.grad(y_ref, (qr, kr, vr), dy)
    qg, kg, vg = (t.to("cuda", dtype).requires_grad_(True) for t in (q, k, v))
    y = sdpa(qg, kg, vg, causal=causal)
    y.backward(dy.to("cuda", dtype))
    if dtype == torch.float32:
        assert max_abs(y, y_ref) <= FP32_TOL
        for a, b in zip((qg.grad, kg.grad, vg.grad), g_ref):
            assert max_abs(a, b) <= scaled_tol(b, FP32_TOL)
    else:
        assert rel_to_max(y, y_ref) <= BF16_TOL
        for a, b in zip((qg.grad, kg.grad, vg.grad), g_ref):
            assert rel_to_max(a, b) <= BF16_TOL


def test_causal_mask_logic_bit_exact():
    """With V = identity the output IS the probability matrix: entries above the
    diagonal must be exactly 0, rows must sum to 1, and row 0 must be exactly e_0."""
    from mop_b200 import sdpa
    T = 96
    q = torch.randn(1, T, 1, T, device="cuda"); k = torch.randn(1, T, 1, T, device="cuda")
    v = torch.eye(T, device="cuda").view(1, T, 1, T)
    P = sdpa(q, k, v, causal=True)[0, :, 0]
    assert torch.equal(torch.triu(P, diagonal=1), torch.zeros_like(P))
    assert (P.sum(-1) - 1).abs().max().item() <= 1e-6
    assert P[0, 0].item() == 1.0


@pytest.mark.parametrize("B,H,Nq,Nk,dk,causal,extras", [
    (2, 4, 64, 64, 56, False, ""), (1, 3, 196, 196, 64, False, ""), (1, 2, 1500, 1500, 64, False, ""),
    (2, 2, 257, 257, 64, True, ""), (2, 2, 37, 150, 32, False, "bias"), (1, 2, 100, 100, 16, True, "bias"),
    (2, 2, 70, 90, 64, False, "mask"), (1, 1, 1, 1, 8, False, "")])
def test_tcgen05_vs_oracle_and_simt(B, H, Nq, Nk, dk, causal, extras):
    """tcgen05 flash attention (fwd + bwd) vs the fp64 oracle on the same bf16 inputs, and vs the fp32-math kernels."""
    from mop_b200 import functional as MF
    from mop_b200 import sdpa
    from oracle.sdpa import sdpa_core
    g = torch.Generator().manual_seed(Nq * 3 + Nk)
    mk = lambda n: bf16_round(torch.randn(B, n, H, dk, generator=g, dtype=torch.float64))
    q, k, v, dy = mk(Nq), mk(Nk), mk(Nk), mk(Nq)
    bias = 0.5 * torch.randn(1, H, Nq, Nk, generator=g, dtype=torch.float64) if "bias" in extras else None
    zm = None
    if "mask" in extras:
        zm = (torch.rand(B, 1, Nq, Nk, generator=g) > 0.3).double()
        zm[..., 0] = 1.0
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    y_ref = sdpa_core(qr.transpose(1, 2), kr.transpose(1, 2), vr.transpose(1, 2), causal=causal, bias=bias, zero_mask=zm).transpose(1, 2)
    g_ref = torch.autograd.grad(y_ref, (qr, kr, vr), dy)
    outs = {}
    for impl in ("tcgen05", "simt"):
        qg, kg, vg = (t.to("cuda", torch.bfloat16).requires_grad_(True) for t in (q, k, v))
        y = sdpa(qg, kg, vg, causal=causal, bias=None if bias is None else bias.cuda(), zero_mask=None if zm is None else zm.cuda(), impl=impl)
        y.backward(dy.to("cuda", torch.bfloat16))
        assert MF.last_impl["sdpa_fwd"] == impl and MF.last_impl["sdpa_bwd"] == impl
        outs[impl] = (y, qg.grad, kg.grad, vg.grad)
    for got, ref in zip(outs["tcgen05"], (y_ref, *g_ref)):
        assert rel_to_max(got, ref) <= BF16_TOL
    for a, b in zip(outs["tcgen05"], outs["simt"]):
        assert rel_to_max(a, b) <= BF16_TOL


def test_tcgen05_strided_qkv_views_and_causal_bits():
    """q,k,v as strided slices of one fused projection (the MSA layout) + bit-exact causal structure."""
    from mop_b200 import functional as MF
    from mop_b200 import sdpa
    T = 64
    fused = torch.randn(2, T, 3, 2, 64, device="cuda").bfloat16()
    y = sdpa(fused[:, :, 0], fused[:, :, 1], fused[:, :, 2], impl="tcgen05")
    y2 = sdpa(fused[:, :, 0].contiguous(), fused[:, :, 1].contiguous(), fused[:, :, 2].contiguous(), impl="tcgen05")
    assert torch.equal(y, y2)
    q = torch.randn(1, T, 1, T, device="cuda").bfloat16(); k = torch.randn(1, T, 1, T, device="cuda").bfloat16()
    v = torch.eye(T, device="cuda").view(1, T, 1, T).bfloat16()
    P = sdpa(q, k, v, causal=True, impl="tcgen05")[0, :, 0].float()
    assert MF.last_impl["sdpa_fwd"] == "tcgen05"
    assert torch.equal(torch.triu(P, diagonal=1), torch.zeros_like(P))
    assert P[0, 0].item() == 1.0

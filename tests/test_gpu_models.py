"""Model-level parity on the GPU: multi-layer ViTEdgewise against the oracle port, and reference models with their attention
classes swapped in place (mop_b200.dropin) against the same unpatched reference models on the CPU (needs baseline/_ref)."""
import copy
import os
import sys

import pytest
import torch
import torch.nn.functional as F

from gpu_util import max_abs, rel_to_max

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_layer_vit_edgewise_logits_and_grads_vs_oracle():
    import mop_b200
    from oracle.edgewise import EdgewiseConfig
    from oracle.vit_edgewise_ref import vit_edgewise_forward
    kw = dict(dim=64, depth=2, heads=2, n_classes=10, mlp_ratio=2.0, n_views=3, share_qkv=True, use_k3=True, gate_mode="lowrank",
              gate_rank=4, gate_init="mix5", drop_path=0.0)
    torch.manual_seed(0)
    model = mop_b200.ViTEdgewise(num_tokens=64, patch=4, compat_experiments_init=False, **kw)
    sd = {k: v.detach().double().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    cfg = EdgewiseConfig(dim=64, heads=2, n_views=3, share_qkv=True, use_k3=True, gate_mode="lowrank", gate_rank=4, gate_init="mix5")
    x = torch.randn(4, 3, 32, 32)
    lab = torch.randint(0, 10, (4,))
    logits_ref = vit_edgewise_forward(x.double(), sd, cfg, depth=2, patch=4)
    F.cross_entropy(logits_ref, lab).backward()
    model = model.cuda()
    logits = model(x.cuda())
    F.cross_entropy(logits, lab.cuda()).backward()
    assert max_abs(logits, logits_ref) <= 2e-5
    for k, p in model.named_parameters():
        ref = sd[k].grad
        assert max_abs(p.grad, ref) <= 5e-5 * max(1.0, ref.abs().max().item()), k


# ---- the unmodified reference package (baseline/_ref), patched in place ---------------------------------------------------------
def _ref():
    sys.path.insert(0, ROOT)
    from baseline import ref_models
    try:
        return ref_models.import_reference()
    except ImportError as e:
        pytest.skip(f"reference package not staged: {e}")


@pytest.fixture()
def patched():
    mods = _ref()
    import mop_b200.dropin as dropin
    yield mods, dropin
    dropin.unpatch_reference()


def _twin(build, patched):
    """Build the reference model unpatched (CPU, fp32) and again after patch_reference() with the same weights (GPU)."""
    mods, dropin = patched
    torch.manual_seed(0)
    ref_model = build()
    done = dropin.patch_reference()
    assert done, "nothing was patched"
    ours = build()
    ours.load_state_dict(ref_model.state_dict(), strict=True)
    return ref_model.eval(), ours.cuda().eval()


def _grads_close(ref_model, ours, tol):
    bad = {}
    for (k, p), (_, q) in zip(ref_model.named_parameters(), ours.named_parameters()):
        if p.grad is None:
            continue
        e = max_abs(q.grad, p.grad) / max(1.0, p.grad.abs().max().item())
        if e > tol:
            bad[k] = e
    assert not bad, bad


def test_patched_reference_vit_baseline_and_vit_mop(patched):
    """Models A and B of the reference (components.MSA inside ViTEncoder blocks) with MSA swapped for the fused kernel."""
    import mop.models as mm   # (classes are looked up at build time: patch_reference() swaps ViT_MoP itself)
    from mop_b200 import functional as MF
    for i, build in enumerate((lambda: mm.ViT_Baseline(dim=64, depth=2, heads=2, n_classes=10),
                               lambda: mm.ViT_MoP(dim=60, depth=2, heads=2, n_classes=10, n_views=3, n_kernels=2),     # D % 8 != 0: reference gate path
                               lambda: mm.ViT_MoP(dim=64, depth=2, heads=2, n_classes=10, n_views=5, n_kernels=3))):   # fused token gate
        patched[1].unpatch_reference()
        calls = MF.abi_calls.get("token_gate_fwd", 0), MF.abi_calls.get("token_gate_bwd", 0)
        ref_model, ours = _twin(build, patched)
        assert any(type(m).__module__.startswith("mop_b200") for m in ours.modules())
        assert not any(type(m).__module__.startswith("mop_b200") for m in ref_model.modules())
        x = torch.randn(3, 3, 32, 32)
        y_ref = ref_model(x)
        y = ours(x.cuda())
        assert max_abs(y, y_ref) <= 2e-5
        y_ref.square().sum().backward(); y.square().sum().backward()
        _grads_close(ref_model, ours, 5e-5)
        if i == 2:   # the post-encoder gate of model B ran in the fused kernel, forward and backward
            assert MF.abi_calls.get("token_gate_fwd", 0) == calls[0] + 1 and MF.abi_calls.get("token_gate_bwd", 0) == calls[1] + 1


def test_patched_reference_gpt_quartet(patched):
    """create_gpt_quartet with the reference's own default config fields except dropout (eval-mode comparison)."""
    from mop.models import create_gpt_quartet
    from mop.models.quartet_attn_patch import TransformerConfig
    cfg = TransformerConfig(n_layer=2, n_head=2, n_embd=32, dropout=0.1, block_size=48)
    ref_model, ours = _twin(lambda: create_gpt_quartet(97, cfg), patched)
    ids = torch.randint(0, 97, (2, 40))
    out_ref = ref_model(ids)
    out = ours(ids.cuda())
    lr, lo = (o[0] if isinstance(o, (tuple, list)) else o for o in (out_ref, out))
    assert max_abs(lo, lr) <= 5e-5
    lr.square().mean().backward(); lo.square().mean().backward()
    _grads_close(ref_model, ours, 1e-4)


def test_patched_reference_gpt_mop(patched):
    """create_gpt_mop: Quartet attention blocks + the 1-D token gate (MoPBlock.apply_mop) through the fused kernels."""
    import mop.models.gpt_mop as gm
    from mop.models.quartet_attn_patch import TransformerConfig
    from mop_b200 import functional as MF
    cfg = TransformerConfig(n_layer=2, n_head=2, n_embd=32, dropout=0.0, block_size=80)
    calls = MF.abi_calls.get("token_gate1d_fwd", 0), MF.abi_calls.get("token_gate1d_bwd", 0)
    ref_model, ours = _twin(lambda: gm.create_gpt_mop(97, cfg), patched)
    ids = torch.randint(0, 97, (2, 70))
    out_ref = ref_model(ids)
    out = ours(ids.cuda())
    lr, lo = (o[0] if isinstance(o, (tuple, list)) else o for o in (out_ref, out))
    assert max_abs(lo, lr) <= 5e-5
    lr.square().mean().backward(); lo.square().mean().backward()
    _grads_close(ref_model, ours, 1e-4)
    assert MF.abi_calls.get("token_gate1d_fwd", 0) == calls[0] + 2 and MF.abi_calls.get("token_gate1d_bwd", 0) == calls[1] + 2


def test_patched_reference_whisper(patched):
    from mop.models import WhisperConfig, create_whisper_mop
    cfg = WhisperConfig(n_mels=16, n_audio_ctx=64, vocab_size=50, n_text_ctx=16, n_embd=32, n_head=2, n_layer_enc=2, n_layer_dec=2,
                        dropout=0.0, n_views=3, n_kernels=2, kernel_size=3)
    ref_model, ours = _twin(lambda: create_whisper_mop(cfg), patched)
    mel = torch.randn(2, 40, cfg.n_mels)
    tok = torch.randint(0, cfg.vocab_size, (2, 8))
    out_ref = ref_model(mel, tok)
    out = ours(mel.cuda(), tok.cuda())
    lr, lo = (o[0] if isinstance(o, (tuple, list)) else o for o in (out_ref, out))
    assert max_abs(lo, lr) <= 5e-5 * max(1.0, lr.abs().max().item())


def test_patched_reference_unified_msa_e(patched):
    from mop.models import UnifiedMSA
    kw = dict(n_views=3, share_qkv=True, gate_mode="lowrank", gate_rank=2, gate_init="mix5", use_k3=True)
    ref_model, ours = _twin(lambda: UnifiedMSA("E", 32, heads=2, **kw), patched)
    x = torch.randn(2, 24, 32)
    y_ref, y = ref_model(x), ours(x.cuda())
    assert max_abs(y, y_ref) <= 1e-5


# ---- variants C and D against fixtures recorded from the reference (tests/golden/make_golden_cd.py) -----------------------------
def _check_golden_module(m, case, call, tol=1e-5, ptol=2e-5):
    from conftest import load_golden  # noqa: F401
    ins = {k: v.cuda().requires_grad_(True) for k, v in case["inputs"].items()}
    y = call(m, ins)
    assert max_abs(y, case["y"]) <= tol * max(1.0, case["y"].abs().max().item())
    y.backward(case["dy"].cuda())
    for k, t in ins.items():
        assert max_abs(t.grad, case["dinputs"][k]) <= tol * max(1.0, case["dinputs"][k].abs().max().item()), k
    for k, p in m.named_parameters():
        ref = case["dparams"][k]
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        assert max_abs(g, ref) <= ptol * max(1.0, ref.abs().max().item()), k


@pytest.mark.parametrize("name", ["crossview_default", "crossview_mix_t_mask", "crossview_dk56"])
def test_crossview_mixer_matches_reference_golden(name):
    from conftest import load_golden
    from mop_b200 import CrossViewMixerMSA
    case = load_golden(name)
    m = CrossViewMixerMSA(case["dim"], heads=case["heads"], **case["kwargs"])
    m.load_state_dict(case["state_dict"], strict=True)
    mask = case["mask"]
    _check_golden_module(m.cuda(), case, lambda mod, t: mod(t["x"], None if mask is None else mask.cuda()))


@pytest.mark.parametrize("name", ["multihop_default", "multihop_gates_h4", "multihop_h2_chain"])
def test_multihop_matches_reference_golden(name):
    from conftest import load_golden
    from mop_b200 import MultiHopMSA
    case = load_golden(name)
    m = MultiHopMSA(case["dim"], heads=case["heads"], **case["kwargs"])
    m.load_state_dict(case["state_dict"], strict=True)
    _check_golden_module(m.cuda(), case, lambda mod, t: mod(t["x"]))


def test_unified_msa_modes_c_d_build_and_run():
    from mop_b200 import UnifiedMSA
    x = torch.randn(2, 16, 32, device="cuda")
    for mode in ("C", "D"):
        m = UnifiedMSA(mode, 32, heads=2).cuda()
        assert m(x).shape == x.shape


def test_multihop_bf16_storage_vs_oracle():
    """Variant D on bf16 activations (fp32-math kernel, bf16 storage) against the fp64 oracle on the same rounded inputs."""
    from mop_b200 import edgewise_attention
    from oracle.variants_cd import multihop_core
    from gpu_util import bf16_round
    g = torch.Generator().manual_seed(9)
    B, N, H, dk = 2, 64, 2, 32
    qkv = bf16_round(torch.randn(B, N, 2, 3, H, dk, generator=g, dtype=torch.float64))
    dy = bf16_round(torch.randn(B, N, H, dk, generator=g, dtype=torch.float64))
    logit = torch.tensor(-1.0, dtype=torch.float64)
    gates = dict(and_=0.8, or_=0.3, not_=0.2, chain=0.6)
    qr = qkv.clone().requires_grad_(True); lr = logit.clone().requires_grad_(True)
    t = lambda v, w: qr[:, :, v, w].permute(0, 2, 1, 3)
    y_ref = multihop_core(t(0, 0), t(0, 1), t(0, 2), t(1, 0), t(1, 1), t(1, 2), lr, gates=gates, beta_not=0.5, hops=3).permute(0, 2, 1, 3)
    gq, gl = torch.autograd.grad(y_ref, (qr, lr), dy)
    qg = qkv.to("cuda", torch.bfloat16).requires_grad_(True); lg = logit.float().cuda().requires_grad_(True)
    y = edgewise_attention(qg, None, None, None, lg, {}, n_views=2, beta_not=0.5, gate_mode="const",
                           const_gates=(0.8, 0.3, 0.2, 0.6), hops=3)
    y.backward(dy.to("cuda", torch.bfloat16))
    assert rel_to_max(y, y_ref) <= 2e-2 and rel_to_max(qg.grad, gq) <= 2e-2
    assert abs(lg.grad.item() - gl.item()) <= 2e-2 * abs(gl.item()) + 1e-3


@pytest.mark.parametrize("B,T,Fb,V,K,ks", [(2, 40, 16, 3, 2, 3), (3, 150, 80, 5, 3, 5), (1, 7, 9, 2, 1, 7)])
def test_mop2d_gate_vs_reference_module(B, T, Fb, V, K, ks, patched):
    """Fused MoP2D gate against the unmodified reference module (whisper_mop.py:91-124) in fp64: gate and every parameter gradient."""
    mods, dropin = patched
    import mop_b200
    torch.manual_seed(B * 100 + ks)
    ref = mods["wm"].MoP2D(V, K, ks).double()
    with torch.no_grad():
        for p in ref.parameters():
            p.copy_(torch.randn_like(p) * 0.5)
    ours = mop_b200.MoP2D(V, K, ks)
    ours.load_state_dict({k: v.float() for k, v in ref.state_dict().items()}, strict=True)
    ours = ours.cuda()
    mel2d = torch.randn(B, 1, T, Fb)
    dg = torch.randn(B, T, 1)
    g_ref, _, _ = ref(mel2d.double())
    g_ref.backward(dg.double())
    g, Vm, Km = ours(mel2d.cuda())
    assert Vm is None and Km is None
    g.backward(dg.cuda())
    assert max_abs(g, g_ref) <= 1e-5 * max(1.0, g_ref.abs().max().item())
    for (k, p), (_, q) in zip(ref.named_parameters(), ours.named_parameters()):
        assert max_abs(q.grad, p.grad) <= 2e-5 * max(1.0, p.grad.abs().max().item()), k


def test_bf16_shadow_weights_match_autocast():
    """mop_b200.mixed.Bf16Shadow: same loss as plain autocast (same bf16-rounded weights), gradients within bf16 rounding of the
    autocast ones (the shadow path skips the bf16 rounding of the weight gradient), and refresh() follows the optimizer."""
    import copy
    import torch.nn.functional as F
    import mop_b200
    from mop_b200.mixed import Bf16Shadow
    torch.manual_seed(0)
    kw = dict(dim=64, depth=2, heads=2, n_classes=10, n_views=3, gate_mode="lowrank", gate_rank=2, drop_path=0.0)   # no random masks
    m1 = mop_b200.ViTEdgewise(num_tokens=64, patch=4, **kw).cuda().train()
    m2 = copy.deepcopy(m1)
    sh = Bf16Shadow(m2)
    assert len(sh.mods) >= 9
    x = torch.randn(8, 3, 32, 32, device="cuda")
    y = torch.randint(0, 10, (8,), device="cuda")
    opts = [torch.optim.AdamW(m.parameters(), lr=1e-2) for m in (m1, m2)]
    for it in range(2):
        losses = []
        for m, opt in zip((m1, m2), opts):
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = F.cross_entropy(m(x), y)
            loss.backward()
            losses.append(loss.item())
        assert abs(losses[0] - losses[1]) <= (1e-3 if it == 0 else 2e-2) * max(1.0, abs(losses[0])), losses
        if it == 0:
            for (k, p), (_, q) in zip(m1.named_parameters(), m2.named_parameters()):
                assert max_abs(q.grad, p.grad) <= 2e-2 * max(1e-6, p.grad.abs().max().item()), k
        for opt in opts:
            opt.step()
        sh.refresh()
    sh.disable()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert torch.isfinite(m2(x)).all()


def test_shadow_weight_grad_token_split():
    """mop_b200.mixed._weight_grad: the token-sliced (split-K) form of dW = dy^T x equals the single GEMM."""
    from mop_b200.mixed import _weight_grad
    g = torch.Generator(device="cuda").manual_seed(3)
    for M, O, I in ((16384, 672, 224), (8192, 224, 224), (4096, 96, 40), (1000, 64, 32)):   # S = 4, 8, 4 (or 2), 1
        dy = torch.randn(M, O, generator=g, device="cuda").bfloat16()
        x = torch.randn(M, I, generator=g, device="cuda").bfloat16()
        ref = dy.double().t() @ x.double()
        got = _weight_grad(dy, x)
        assert got.dtype == torch.float32 and got.shape == (O, I)
        assert max_abs(got, ref) <= 1e-3 * ref.abs().max().item()

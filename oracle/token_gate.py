"""Oracle (test infrastructure, never on the product path): the ViT-MoP post-encoder token gate restated with plain torch ops.

Follows the reference line by line:
  ViewsLinear.forward   mop/models/components.py:263-268   V = proj(tok) -> [B, V, Gh, Gw]
  Kernels3.forward      components.py:274-283              conv3x3(V -> 16, padding 1, no bias) -> SiLU -> conv1x1(16 -> K, no bias)
  FuseExcInh.forward    components.py:289-303              conv1x1(V+K -> hid) -> SiLU -> conv1x1(hid -> 2, bias); sigmoid; softplus(alpha)
  ViT_MoP.forward       mop/models/vit_mop.py:95-114        gate = 1 + a_pos G_pos - a_neg G_neg;  tok * gate per token
Pinned by tests/test_gpu_models.py::test_patched_reference_vit_baseline_and_vit_mop, which compares the fused kernel inside the
reference's own ViT_MoP against the unmodified reference model (baseline/_ref).
"""
import torch
import torch.nn.functional as F


def token_gate_ref(tok, grid, views_w, k3_w, k1_w, f1_w, f2_w, f2_b, alpha_pos, alpha_neg):
    B, N, D = tok.shape
    Gh, Gw = grid
    V = views_w.shape[0]
    views = (tok @ views_w.t()).transpose(1, 2).reshape(B, V, Gh, Gw)
    kmaps = F.conv2d(F.silu(F.conv2d(views, k3_w, padding=1)), k1_w)
    maps = torch.cat([views, kmaps], dim=1)
    G = F.conv2d(F.silu(F.conv2d(maps, f1_w)), f2_w, f2_b)
    g_pos, g_neg = torch.sigmoid(G[:, :1]), torch.sigmoid(G[:, 1:])
    gate = 1 + F.softplus(alpha_pos) * g_pos - F.softplus(alpha_neg) * g_neg
    return tok * gate.reshape(B, N, 1)

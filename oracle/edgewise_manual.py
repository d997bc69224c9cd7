"""Oracle (test infrastructure): hand-derived forward/backward of the Edgewise core.

This is the *executable specification* of what the CUDA kernels in
``mop_b200/csrc/edgewise_*.cuh`` compute, written with explicit matrix algebra
(no autograd) on the packed kernel-boundary layout:

    qkv      [B, N, Vp, 3, H, dk]   Vp = 1 (share_qkv, attention_variants.py:459-464)
                                    Vp = V (one Linear per view,          :466-470)
    q_scale, k_scale, v_scale  [V, H, dk] or None
    y        [B, N, H, dk]          (merge-heads layout, :563)

``tests/test_oracle_manual.py`` checks it against autograd through
:func:`oracle.edgewise.edgewise_core` in fp64 (agreement ~1e-14), which in turn
is pinned against the imported reference (tests/golden).  Derivation:
SURVEY.md appendix D.1.  Not used by the product path.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .edgewise import EPS_CHAIN, gelu_tanh


def _views(qkv, q_scale, k_scale, v_scale, V):
    """Per-view Q_i, K_i [G?]: returns lists of [B,H,N,dk] plus v_first, v_last."""
    B, N, Vp, _, H, dk = qkv.shape
    qs, ks = [], []
    for i in range(V):
        src = qkv[:, :, 0 if Vp == 1 else i]  # [B,N,3,H,dk]
        q = src[:, :, 0].permute(0, 2, 1, 3)
        k = src[:, :, 1].permute(0, 2, 1, 3)
        if q_scale is not None:
            q = q * q_scale[i][None, :, None, :]
            k = k * k_scale[i][None, :, None, :]
        qs.append(q); ks.append(k)
    vb0 = qkv[:, :, 0, 2].permute(0, 2, 1, 3)
    vbl = qkv[:, :, 0 if Vp == 1 else V - 1, 2].permute(0, 2, 1, 3)
    if v_scale is not None:
        vb0 = vb0 * v_scale[0][None, :, None, :]
        vbl = vbl * v_scale[V - 1][None, :, None, :]
    return qs, ks, vb0, vbl


def _dgelu_tanh(x):
    c = math.sqrt(2.0 / math.pi)
    u = c * (x + 0.044715 * x ** 3)
    t = torch.tanh(u)
    return 0.5 * (1 + t) + 0.5 * x * (1 - t * t) * c * (1 + 3 * 0.044715 * x * x)


def edgewise_packed(
    qkv, q_scale, k_scale, v_scale, head: Dict[str, torch.Tensor], logit, *,
    V: int, beta_not: float, gate_mode: str, gate_rank: int = 4,
    dy: Optional[torch.Tensor] = None,
):
    """Forward (and, when ``dy`` is given, backward) on the packed layout.

    Returns ``y`` or ``(y, grads)`` with ``grads`` keyed like the inputs:
    ``qkv, q_scale, k_scale, v_scale, logit`` and the gate-head tensors.
    """
    B, N, Vp, _, H, dk = qkv.shape
    s = 1.0 / math.sqrt(dk)
    eps = EPS_CHAIN
    Vd = max(1, V - 1)
    r = gate_rank
    C = 2 * V + 2
    qs, ks, v1, vL = _views(qkv, q_scale, k_scale, v_scale, V)

    # ---------------- forward ----------------
    S = [s * qs[i] @ ks[i].transpose(-1, -2) for i in range(V)]
    A = [torch.softmax(m, -1) for m in S]
    P = [A[0]]  # P[k] = A_1..A_{k+1}
    for i in range(1, V):
        P.append(P[-1] @ A[i])
    Rs = [None] * V  # Rs[k] = A_V..A_{k+1}
    Rs[V - 1] = A[V - 1]
    for i in range(V - 2, -1, -1):
        Rs[i] = Rs[i + 1] @ A[i]
    Fc, Rc = P[V - 1], Rs[0]
    Lf, Lr = torch.log(Fc + eps), torch.log(Rc + eps)
    Ssum = sum(S)
    Smax = torch.stack(S, 0).max(0).values
    L = Smax + torch.log(sum(torch.exp(m - Smax) for m in S))
    U = Ssum - S[0]
    O = L - S[0]
    feat = S + [m.transpose(-1, -2) for m in S] + [Lf, Lr]  # C maps [B,H,N,N]

    if gate_mode == "lowrank":
        Wr = head["row_proj.weight"].reshape(4 * r, C)
        Wc = head["col_proj.weight"].reshape(4 * r, C)
        br, bc = head["row_proj.bias"], head["col_proj.bias"]
        rho = torch.stack([f.mean(-1) for f in feat], 2)  # [B,H,C,N]
        kap = torch.stack([f.mean(-2) for f in feat], 2)
        a = torch.einsum("qc,bhcn->bhqn", Wr, rho) + br[None, None, :, None]
        b = torch.einsum("qc,bhcn->bhqn", Wc, kap) + bc[None, None, :, None]
        a4 = a.reshape(B, H, 4, r, N)
        b4 = b.reshape(B, H, 4, r, N)
        g = torch.sigmoid(torch.einsum("bhtki,bhtkj->bhtij", a4, b4))  # [B,H,4,N,N]
    else:
        W1 = head["conv1.weight"].reshape(-1, C)
        hid = W1.shape[0]
        b1 = head["conv1.bias"]
        W2 = head["conv2.weight"].reshape(4, hid)
        b2 = head["conv2.bias"]
        W3 = head.get("mid3.weight")
        b3 = head.get("mid3.bias")
        fst = torch.stack(feat, 2).reshape(B * H, C, N, N)
        z1 = torch.einsum("oc,gcij->goij", W1, fst) + b1[None, :, None, None]
        h1 = gelu_tanh(z1)
        if W3 is not None:
            h1g = gelu_tanh(h1)
            h2 = F.conv2d(h1g, W3, b3, padding=1)
        else:
            h2 = h1
        z2 = torch.einsum("oc,gcij->goij", W2, h2) + b2[None, :, None, None]
        g = torch.sigmoid(z2).reshape(B, H, 4, N, N)

    g_and, g_or, g_not, g_ch = g[:, :, 0], g[:, :, 1], g[:, :, 2], g[:, :, 3]
    Smix = S[0] + g_and * U + g_or * O - g_not * (beta_not / Vd) * U + g_ch * Lf
    Am = torch.softmax(Smix, -1)
    w = torch.sigmoid(logit)
    yh = Am @ v1 + w * (Fc @ vL)  # [B,H,N,dk]
    y = yh.permute(0, 2, 1, 3).contiguous()
    if dy is None:
        return y

    # ---------------- backward ----------------
    dY = dy.permute(0, 2, 1, 3)  # [B,H,N,dk]
    dV1 = Am.transpose(-1, -2) @ dY
    dA = dY @ v1.transpose(-1, -2)
    D = Am * (dA - (dA * Am).sum(-1, keepdim=True))
    dVL = w * (Fc.transpose(-1, -2) @ dY)
    dF = w * (dY @ vL.transpose(-1, -2))
    dlogit = (1 - w) * (Fc * dF).sum()
    # gate pre-activation grads
    dg = torch.stack([D * U, D * O, -(beta_not / Vd) * D * U, D * Lf], 2)
    dG = dg * g * (1 - g)  # [B,H,4,N,N]
    pis = [torch.exp(m - L) for m in S]
    dS = [D * (1 + g_or * (pis[0] - 1))]
    for i in range(1, V):
        dS.append(D * (g_and + g_or * pis[i] - g_not * (beta_not / Vd)))
    dLf = D * g_ch
    dLr = torch.zeros_like(D)
    grads: Dict[str, torch.Tensor] = {}
    if gate_mode == "lowrank":
        da = torch.einsum("bhtij,bhtkj->bhtki", dG, b4).reshape(B, H, 4 * r, N)
        db = torch.einsum("bhtij,bhtki->bhtkj", dG, a4).reshape(B, H, 4 * r, N)
        grads["row_proj.weight"] = torch.einsum("bhqn,bhcn->qc", da, rho).reshape(head["row_proj.weight"].shape)
        grads["col_proj.weight"] = torch.einsum("bhqn,bhcn->qc", db, kap).reshape(head["col_proj.weight"].shape)
        grads["row_proj.bias"] = da.sum((0, 1, 3))
        grads["col_proj.bias"] = db.sum((0, 1, 3))
        drho = torch.einsum("qc,bhqn->bhcn", Wr, da) / N  # [B,H,C,N], indexed by row i
        dkap = torch.einsum("qc,bhqn->bhcn", Wc, db) / N  # indexed by column j
        # dfeat_c[i,j] = drho_c[i] + dkap_c[j]
        for i in range(V):
            dS[i] = dS[i] + drho[:, :, i, :, None] + dkap[:, :, i, None, :]
            # channel V+i is S_i^T: dfeat_{V+i}[i',j'] lands on S_i[j',i']
            dS[i] = dS[i] + drho[:, :, V + i, None, :] + dkap[:, :, V + i, :, None]
        dLf = dLf + drho[:, :, 2 * V, :, None] + dkap[:, :, 2 * V, None, :]
        dLr = dLr + drho[:, :, 2 * V + 1, :, None] + dkap[:, :, 2 * V + 1, None, :]
    else:
        dz2 = dG.reshape(B * H, 4, N, N)
        grads["conv2.weight"] = torch.einsum("goij,gcij->oc", dz2, h2).reshape(head["conv2.weight"].shape)
        grads["conv2.bias"] = dz2.sum((0, 2, 3))
        dh2 = torch.einsum("oc,goij->gcij", W2, dz2)
        if W3 is not None:
            grads["mid3.bias"] = dh2.sum((0, 2, 3))
            # dW3[o,c,u,v] = sum dh2[o,i,j] * h1g[c,i+u-1,j+v-1]
            pad = F.pad(h1g, (1, 1, 1, 1))
            dW3 = torch.zeros_like(W3)
            for u in range(3):
                for v_ in range(3):
                    dW3[:, :, u, v_] = torch.einsum("goij,gcij->oc", dh2, pad[:, :, u:u + N, v_:v_ + N])
            grads["mid3.weight"] = dW3
            dh1g = F.conv_transpose2d(dh2, W3, None, padding=1)
            dh1 = dh1g * _dgelu_tanh(h1)
        else:
            dh1 = dh2
        dz1 = dh1 * _dgelu_tanh(z1)
        grads["conv1.weight"] = torch.einsum("goij,gcij->oc", dz1, fst).reshape(head["conv1.weight"].shape)
        grads["conv1.bias"] = dz1.sum((0, 2, 3))
        dfeat = torch.einsum("oc,goij->gcij", W1, dz1).reshape(B, H, C, N, N)
        for i in range(V):
            dS[i] = dS[i] + dfeat[:, :, i] + dfeat[:, :, V + i].transpose(-1, -2)
        dLf = dLf + dfeat[:, :, 2 * V]
        dLr = dLr + dfeat[:, :, 2 * V + 1]
    dF = dF + dLf / (Fc + eps)
    dR = dLr / (Rc + eps)
    # chains
    dAk = [torch.zeros_like(D) for _ in range(V)]
    X = dF
    for k in range(V - 1, -1, -1):  # F = A_1..A_V
        dAk[k] = dAk[k] + (X if k == 0 else P[k - 1].transpose(-1, -2) @ X)
        if k > 0:
            X = X @ A[k].transpose(-1, -2)
    X = dR
    for k in range(0, V):  # R = A_V..A_1
        dAk[k] = dAk[k] + (X if k == V - 1 else Rs[k + 1].transpose(-1, -2) @ X)
        if k < V - 1:
            X = X @ A[k].transpose(-1, -2)
    for k in range(V):
        dS[k] = dS[k] + A[k] * (dAk[k] - (dAk[k] * A[k]).sum(-1, keepdim=True))
    # projections
    dqkv = torch.zeros_like(qkv)
    qb = lambda i: qkv[:, :, 0 if Vp == 1 else i, 0].permute(0, 2, 1, 3)
    kb = lambda i: qkv[:, :, 0 if Vp == 1 else i, 1].permute(0, 2, 1, 3)
    vb = lambda i: qkv[:, :, 0 if Vp == 1 else i, 2].permute(0, 2, 1, 3)
    if q_scale is not None:
        dqs = torch.zeros_like(q_scale); dks = torch.zeros_like(k_scale); dvs = torch.zeros_like(v_scale)
    for i in range(V):
        T = s * dS[i] @ kb(i)                      # [B,H,N,dk] (unscaled base K)
        Ut = s * dS[i].transpose(-1, -2) @ qb(i)
        vi = 0 if Vp == 1 else i
        if q_scale is not None:
            c = (q_scale[i] * k_scale[i])[None, :, None, :]
            dqkv[:, :, vi, 0] += (T * c).permute(0, 2, 1, 3)
            dqkv[:, :, vi, 1] += (Ut * c).permute(0, 2, 1, 3)
            Z = (T * qb(i)).sum((0, 2))  # [H,dk]
            dqs[i] = k_scale[i] * Z
            dks[i] = q_scale[i] * Z
        else:
            dqkv[:, :, vi, 0] += T.permute(0, 2, 1, 3)
            dqkv[:, :, vi, 1] += Ut.permute(0, 2, 1, 3)
    v0i, vLi = 0, (0 if Vp == 1 else V - 1)
    if v_scale is not None:
        dqkv[:, :, v0i, 2] += (dV1 * v_scale[0][None, :, None, :]).permute(0, 2, 1, 3)
        dqkv[:, :, vLi, 2] += (dVL * v_scale[V - 1][None, :, None, :]).permute(0, 2, 1, 3)
        dvs[0] = (dV1 * vb(0)).sum((0, 2))
        dvs[V - 1] = (dVL * vb(V - 1)).sum((0, 2))
    else:
        dqkv[:, :, v0i, 2] += dV1.permute(0, 2, 1, 3)
        dqkv[:, :, vLi, 2] += dVL.permute(0, 2, 1, 3)
    grads["qkv"] = dqkv
    if q_scale is not None:
        grads["q_scale"], grads["k_scale"], grads["v_scale"] = dqs, dks, dvs
    grads["logit"] = dlogit
    return y, grads

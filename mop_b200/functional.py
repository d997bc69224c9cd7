"""Autograd-aware Python entry points over the C ABI (libmop_b200.so).

Each function takes the tensors that sit between the reference module's input
projection(s) and its output projection, launches the CUDA kernels on the
current stream through ``ctypes`` and returns the attention output in the
merge-heads layout ``[B, N, H, dk]``.  Nothing here computes on the CPU: a
non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch

from . import _lib

_DT = {torch.float32: _lib.MOP_F32, torch.bfloat16: _lib.MOP_BF16}
_IMPL = {None: _lib.MOP_IMPL_AUTO, "auto": _lib.MOP_IMPL_AUTO, "simt": _lib.MOP_IMPL_SIMT, "tcgen05": _lib.MOP_IMPL_TCGEN05}

# what the most recent launch of each op used ("simt" | "tcgen05"); read by tests / bench
last_impl: Dict[str, str] = {}
# number of kernel-launching ABI calls made so far (bench.py reports launches from this)
abi_calls: Dict[str, int] = {"edgewise_fwd": 0, "edgewise_bwd": 0, "sdpa_fwd": 0, "sdpa_bwd": 0, "quartet_fwd": 0, "quartet_bwd": 0}


# Test hook: when `keep_partials` is True the un-reduced per-(batch, head) contributions to scalar parameter gradients that the
# kernels write are kept here (cloned) after each backward: "edgewise_dlogit" [B*H], "quartet_dscalar" [B*H, 2]
keep_partials = False
last_partials: Dict[str, torch.Tensor] = {}

# Optional per-launch device timing: when `kernel_timing` is True every ABI call is bracketed by
# CUDA events on the launching stream; bench.py reads `kernel_events` after a synchronize.
kernel_timing = False
kernel_events: Dict[str, list] = {}


class _Timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if kernel_timing:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if kernel_timing:
            self.e1.record()
            kernel_events.setdefault(self.name, []).append((self.e0, self.e1))
        return False


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _need_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"mop_b200: {name} must be a CUDA tensor (there is no CPU path; the CPU oracle lives in oracle/)")


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype not in _DT:
        raise RuntimeError(f"mop_b200: unsupported activation dtype {t.dtype} (float32 or bfloat16)")
    return _DT[t.dtype]


def _f32c(t: Optional[torch.Tensor]):
    return None if t is None else t.detach().float().contiguous()


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ----------------------------------------------------------------------------
# Edgewise
# ----------------------------------------------------------------------------
_LOWRANK_KEYS = ("row_proj.weight", "row_proj.bias", "col_proj.weight", "col_proj.bias")


def _dense_keys(use_k3: bool):
    return ("conv1.weight", "conv1.bias") + (("mid3.weight", "mid3.bias") if use_k3 else ()) + ("conv2.weight", "conv2.bias")


def _fill_edgewise(p, qkv, scales, logit, head, cfg):
    B, N, Vp, _, H, dk = qkv.shape
    p.dtype = _dtype_code(qkv)
    p.impl = _IMPL[cfg["impl"]]
    p.B, p.H, p.N, p.dk, p.V, p.Vp = B, H, N, dk, cfg["n_views"], Vp
    p.gate_mode = {"lowrank": _lib.MOP_GATE_LOWRANK, "dense": _lib.MOP_GATE_DENSE, "const": _lib.MOP_GATE_CONST}[cfg["gate_mode"]]
    if cfg["gate_mode"] == "const":
        for i, gv in enumerate(cfg["const_gates"]):
            p.const_gates[i] = float(gv)
        p.hops = int(cfg["hops"])
    p.gate_rank = cfg["gate_rank"]
    p.hidden = cfg["hidden"]
    p.use_k3 = int(cfg["use_k3"])
    p.beta_not = cfg["beta_not"]
    p.eps = 1e-6
    p.qkv = _ptr(qkv)
    if scales is not None:
        p.q_scale, p.k_scale, p.v_scale = (_ptr(s) for s in scales)
    p.chain_value_logit = _ptr(logit)
    lens = cfg.get("lens_dilations", ())
    if lens:   # S lens bank: the stacked [L,V,3,3] weights travel as the last "head" tensor
        p.lens_n = len(lens)
        for i, d in enumerate(lens):
            p.lens_dil[i] = int(d)
        p.lens_w = _ptr(head[-1])
        head = head[:-1]
    if cfg["gate_mode"] == "lowrank":
        p.row_w, p.row_b, p.col_w, p.col_b = (_ptr(h) for h in head)
    elif cfg["gate_mode"] == "dense":
        it = iter(head)
        p.conv1_w, p.conv1_b = _ptr(next(it)), _ptr(next(it))
        if cfg["use_k3"]:
            p.mid3_w, p.mid3_b = _ptr(next(it)), _ptr(next(it))
        p.conv2_w, p.conv2_b = _ptr(next(it)), _ptr(next(it))


class _Edgewise(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, qkv, q_scale, k_scale, v_scale, logit, *head):
        lib = _lib.load()
        _need_cuda(qkv, "qkv")
        qkv_c = qkv.detach().contiguous()
        B, N, Vp, _, H, dk = qkv_c.shape
        scales = None if q_scale is None else tuple(_f32c(s).reshape(cfg["n_views"], H, dk) for s in (q_scale, k_scale, v_scale))
        logit32 = _f32c(logit).reshape(1)
        head32 = tuple(_f32c(h) for h in head)
        y = torch.empty(B, N, H, dk, dtype=qkv_c.dtype, device=qkv_c.device)
        with torch.cuda.device(qkv_c.device):
            p = _lib.new_params(_lib.EdgewiseParams)
            _fill_edgewise(p, qkv_c, scales, logit32, head32, cfg)
            p.y = _ptr(y)
            # the tcgen05 kernels for N != 64 hand the row statistics of the mixed map and A V_1 (fp32) to their backward
            stats = ybase = None
            if lib.mop_edgewise_needs_row_stats(C.byref(p)) and any(ctx.needs_input_grad):
                stats = torch.empty(B, H, N, 2, dtype=torch.float32, device=qkv_c.device)
                ybase = torch.empty(B, N, H, dk, dtype=torch.float32, device=qkv_c.device)
                p.row_stats, p.y_base = _ptr(stats), _ptr(ybase)
            # the N = 64 tcgen05 kernels hand their small per-(b,h) vectors (softmax statistics, feature means, gate factors) on
            aux = None
            n_aux = lib.mop_edgewise_aux_floats(C.byref(p))
            if n_aux and any(ctx.needs_input_grad):
                aux = torch.empty(n_aux, dtype=torch.float32, device=qkv_c.device)
                p.aux = _ptr(aux)
            nbytes = lib.mop_edgewise_workspace_bytes(C.byref(p), 0)
            ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=qkv_c.device)
            p.workspace, p.workspace_bytes = _ptr(ws), nbytes
            with _Timed("edgewise_fwd"):
                _lib.check(lib.mop_edgewise_fwd(C.byref(p), _stream()), "mop_edgewise_fwd")
        last_impl["edgewise_fwd"] = _lib.IMPL_NAMES.get(p.impl_used, "?")
        abi_calls["edgewise_fwd"] += 1
        ctx.cfg = cfg
        ctx.has_scales = scales is not None
        ctx.head_shapes = [h.shape for h in head]
        ctx.in_dtypes = [None if t is None else t.dtype for t in (q_scale, k_scale, v_scale, logit, *head)]
        ctx.scale_shape = None if q_scale is None else q_scale.shape
        ctx.has_stats = stats is not None
        ctx.has_aux = aux is not None
        ctx.save_for_backward(qkv_c, logit32, *(scales or ()), *head32, *((ybase, stats) if stats is not None else ()),
                              *((aux,) if aux is not None else ()))
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        cfg = ctx.cfg
        saved = ctx.saved_tensors
        ybase = stats = aux = None
        if ctx.has_aux:
            aux, saved = saved[-1], saved[:-1]
        if ctx.has_stats:
            ybase, stats = saved[-2], saved[-1]
            saved = saved[:-2]
        qkv_c, logit32 = saved[0], saved[1]
        scales = tuple(saved[2:5]) if ctx.has_scales else None
        head32 = tuple(saved[5:] if ctx.has_scales else saved[2:])
        B, N, Vp, _, H, dk = qkv_c.shape
        V = cfg["n_views"]
        dy_c = dy.detach().to(qkv_c.dtype).contiguous()
        dev = qkv_c.device
        dqkv = torch.empty_like(qkv_c)
        with torch.cuda.device(dev):
            p = _lib.new_params(_lib.EdgewiseParams)
            _fill_edgewise(p, qkv_c, scales, logit32, head32, cfg)
            p.y = _ptr(dqkv)  # the forward output is not needed by the backward kernels
            if stats is not None:
                p.row_stats, p.y_base = _ptr(stats), _ptr(ybase)
            p.aux = _ptr(aux)
            nhead = lib.mop_edgewise_head_param_count(C.byref(p))
            G = B * H
            R = lib.mop_edgewise_partial_rows(C.byref(p))   # rows of the partial-gradient buffers: B*H or one per CTA
            dhead_part = torch.empty(R, nhead, dtype=torch.float32, device=dev) if nhead else None
            dlogit_part = torch.empty(G, dtype=torch.float32, device=dev)
            dscale_part = torch.empty(R, 3, V, dk, dtype=torch.float32, device=dev) if scales is not None else None
            p.dy, p.dqkv = _ptr(dy_c), _ptr(dqkv)
            p.dhead_part, p.dlogit_part, p.dscale_part = _ptr(dhead_part), _ptr(dlogit_part), _ptr(dscale_part)
            dlens_part = None
            if cfg.get("lens_dilations", ()):
                dlens_part = torch.empty(G, head32[-1].numel(), dtype=torch.float32, device=dev)   # fp32-mode kernel: one row per problem
                p.dlens_part = _ptr(dlens_part)
            nbytes = lib.mop_edgewise_workspace_bytes(C.byref(p), 1)
            ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
            p.workspace, p.workspace_bytes = _ptr(ws), nbytes
            with _Timed("edgewise_bwd"):
                _lib.check(lib.mop_edgewise_bwd(C.byref(p), _stream()), "mop_edgewise_bwd")
        last_impl["edgewise_bwd"] = _lib.IMPL_NAMES.get(p.impl_used, "?")
        abi_calls["edgewise_bwd"] += 1
        dts = ctx.in_dtypes
        # the sums of the per-CTA / per-problem partials (row i of the partials belongs to head i % H; the scale sums come out in
        # the [3, V, H, dk] layout of the parameters, so the three gradients are contiguous views).  Few rows (one per CTA: the
        # N = 64 kernels): one launch for all three.  One row per problem (thousands): three tree reductions.
        ds = torch.empty(3, V, H, dk, dtype=torch.float32, device=dev) if scales is not None else None
        if R <= 1024:
            flat = torch.empty(nhead, dtype=torch.float32, device=dev) if dhead_part is not None else None
            dlogit32 = torch.empty(1, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.mop_edgewise_reduce_partials(_ptr(dscale_part), _ptr(dhead_part), _ptr(dlogit_part), R, H, V, dk, nhead, G,
                                                            _ptr(ds), _ptr(flat), _ptr(dlogit32), _stream()), "mop_edgewise_reduce_partials")
        else:
            if scales is not None:
                torch.sum(dscale_part.view(R // H, H, 3, V, dk).permute(0, 2, 3, 1, 4), dim=0, out=ds)
            flat = dhead_part.sum(0) if dhead_part is not None else None
            dlogit32 = dlogit_part.sum()
        if scales is not None:
            dq_s, dk_s, dv_s = (ds[i].reshape(ctx.scale_shape).to(dts[i]) for i in range(3))
        else:
            dq_s = dk_s = dv_s = None
        if keep_partials:
            last_partials["edgewise_dlogit"] = dlogit_part.detach().clone()
        dlogit = dlogit32.reshape(()).to(dts[3])
        dheads, off = [], 0
        n_gate = len(ctx.head_shapes) - (1 if dlens_part is not None else 0)
        for shp, dt in zip(ctx.head_shapes[:n_gate], dts[4:]):
            n = math.prod(shp)
            dheads.append(flat[off:off + n].reshape(shp).to(dt))
            off += n
        if dlens_part is not None:
            dheads.append(dlens_part.sum(0).reshape(ctx.head_shapes[-1]).to(dts[-1]))
        return (None, dqkv, dq_s, dk_s, dv_s, dlogit, *dheads)


def edgewise_attention(qkv: torch.Tensor, q_scale, k_scale, v_scale, chain_value_logit: torch.Tensor,
                       head: Dict[str, torch.Tensor], *, n_views: int, beta_not: float, gate_mode: str,
                       gate_rank: int = 4, use_k3: bool = False, impl: Optional[str] = None,
                       const_gates=None, hops: Optional[int] = None, lens_w: Optional[torch.Tensor] = None,
                       lens_dilations=()) -> torch.Tensor:
    """Edgewise Mixture-of-Products attention core (reference attention_variants.py:500-562).

    qkv   ``[B, N, Vp, 3, H, dk]``: output of the shared qkv Linear (Vp=1, with
          per-view ``q/k/v_scale [V,H,1,dk]``) or of one Linear per view (Vp=V,
          scales None).
    head  gate-head tensors under the reference names (``row_proj.weight`` ... or
          ``conv1.weight`` ...).
    Returns ``y [B, N, H, dk]``; differentiable w.r.t. qkv, the scales,
    ``chain_value_logit`` and every head tensor.
    """
    if gate_mode not in ("lowrank", "dense", "const"):
        raise ValueError(f"gate_mode {gate_mode!r}")
    if gate_mode == "const":
        # variant D (MultiHopMSA, reference :163-231): fixed scalar gates (and_, or_, not_, chain), chain A_1 A_2^(hops-1)
        if const_gates is None or len(const_gates) != 4 or hops is None or hops < 2:
            raise ValueError("gate_mode='const' needs const_gates=(and_, or_, not_, chain) and hops >= 2")
        cfg = dict(n_views=int(n_views), beta_not=float(beta_not), gate_mode="const", gate_rank=1, use_k3=False, hidden=16, impl=impl,
                   const_gates=tuple(float(g) for g in const_gates), hops=int(hops))
        return _Edgewise.apply(cfg, qkv, q_scale, k_scale, v_scale, chain_value_logit)
    if qkv.dim() != 6 or qkv.shape[3] != 3:
        raise ValueError("qkv must be [B,N,Vp,3,H,dk]")
    dense_k3 = bool(use_k3) and gate_mode == "dense"  # use_k3 has no effect on the low-rank head (:273-278)
    keys = _LOWRANK_KEYS if gate_mode == "lowrank" else _dense_keys(dense_k3)
    hidden = head["conv1.weight"].shape[0] if gate_mode == "dense" else 16
    cfg = dict(n_views=int(n_views), beta_not=float(beta_not), gate_mode=gate_mode, gate_rank=int(gate_rank),
               use_k3=dense_k3, hidden=int(hidden), impl=impl)
    tensors = [head[k] for k in keys]
    if lens_w is not None and len(lens_dilations):
        # S lens bank (reference :427-442, :523-533): lens_w [L, V, 3, 3] = the depthwise 3x3 weights per dilation, stacked
        if lens_w.shape != (len(lens_dilations), int(n_views), 3, 3):
            raise ValueError(f"lens_w must be [L={len(lens_dilations)}, V={n_views}, 3, 3], got {tuple(lens_w.shape)}")
        cfg["lens_dilations"] = tuple(int(d) for d in lens_dilations)
        tensors.append(lens_w)
    return _Edgewise.apply(cfg, qkv, q_scale, k_scale, v_scale, chain_value_logit, *tensors)


# ----------------------------------------------------------------------------
# attention dropout
# ----------------------------------------------------------------------------
# (seed, offset) of the most recent dropout-enabled forward (tests evaluate the oracle under the kernels' own mask)
last_dropout: Dict[str, tuple] = {}


def _draw_dropout(p: float):
    """(p, seed, offset) for one attention call.  The pair comes from torch's CPU generator, so `torch.manual_seed` makes a run
    reproducible; every kernel of the call (forward, dQ, dK/dV) regenerates the same keep mask from it.  Under CUDA-graph
    capture the pair is baked into the captured launches (one fixed mask per replay)."""
    if not (0.0 <= p <= 1.0):
        raise ValueError(f"dropout probability has to be between 0 and 1, but got {p}")
    if p == 0.0:
        return (0.0, 0, 0)
    seed, offset = torch.randint(0, 2 ** 62, (2,), dtype=torch.int64).tolist()
    return (float(p), int(seed), int(offset))


def dropout_mask(BH: int, Nq: int, Nk: int, p: float, seed: int, offset: int, device="cuda") -> torch.Tensor:
    """The factor (0 or 1/(1-p)) the attention kernels apply to P[bh, i, j] for (p, seed, offset): fp32 ``[BH, Nq, Nk]``."""
    lib = _lib.load()
    out = torch.empty(BH, Nq, Nk, dtype=torch.float32, device=device)
    with torch.cuda.device(out.device):
        _lib.check(lib.mop_dropout_mask(_ptr(out), BH, Nq, Nk, float(p), int(seed), int(offset), _stream()), "mop_dropout_mask")
    return out


# ----------------------------------------------------------------------------
# SDPA
# ----------------------------------------------------------------------------
def _bstrides(t: torch.Tensor):
    """Element strides of a 4-d tensor broadcastable to [B,H,Nq,Nk] (0 on size-1 dims)."""
    return [0 if t.shape[i] == 1 else t.stride(i) for i in range(4)]


def _as4(t: torch.Tensor, B, H, Nq, Nk, name):
    while t.dim() < 4:
        t = t.unsqueeze(0)
    for have, want in zip(t.shape, (B, H, Nq, Nk)):
        if have not in (1, want):
            raise ValueError(f"{name} shape {tuple(t.shape)} does not broadcast to {(B, H, Nq, Nk)}")
    return t


def _fill_sdpa(p, q, k, v, causal, bias, zero_mask, cfg):
    B, Nq, H, dk = q.shape
    Nk = k.shape[1]
    p.dtype = _dtype_code(q)
    p.impl = _IMPL[cfg.get("impl")]
    p.B, p.H, p.Nq, p.Nk, p.dk, p.causal = B, H, Nq, Nk, dk, int(causal)
    p.scale = 1.0 / math.sqrt(dk)
    for name, t in (("q", q), ("k", k), ("v", v)):
        if t.stride(3) != 1:
            raise ValueError(f"{name}: last dim must be contiguous")
        setattr(p, name, _ptr(t))
        setattr(p, f"{name}_sb", t.stride(0)); setattr(p, f"{name}_sn", t.stride(1)); setattr(p, f"{name}_sh", t.stride(2))
    if bias is not None:
        p.bias = _ptr(bias)
        p.bias_sb, p.bias_sh, p.bias_sq, p.bias_sk = _bstrides(bias)
    if zero_mask is not None:
        p.zero_mask = _ptr(zero_mask)
        p.zm_sb, p.zm_sh, p.zm_sq, p.zm_sk = _bstrides(zero_mask)
    p.dropout_p, p.dropout_seed, p.dropout_offset = cfg.get("dropout", (0.0, 0, 0))


class _Sdpa(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, q, k, v, bias, zero_mask):
        lib = _lib.load()
        for n, t in (("q", q), ("k", k), ("v", v)):
            _need_cuda(t, n)
        q, k, v = q.detach(), k.detach(), v.detach()
        B, Nq, H, dk = q.shape
        y = torch.empty(B, Nq, H, dk, dtype=q.dtype, device=q.device)
        lse = torch.empty(B, H, Nq, dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            p = _lib.new_params(_lib.SdpaParams)
            _fill_sdpa(p, q, k, v, cfg["causal"], bias, zero_mask, cfg)
            p.y, p.lse = _ptr(y), _ptr(lse)
            with _Timed("sdpa_fwd"):
                _lib.check(lib.mop_sdpa_fwd(C.byref(p), _stream()), "mop_sdpa_fwd")
        last_impl["sdpa_fwd"] = _lib.IMPL_NAMES.get(p.impl_used, "?")
        abi_calls["sdpa_fwd"] += 1
        ctx.cfg = cfg
        ctx.save_for_backward(q, k, v, y, lse, *[t for t in (bias, zero_mask) if t is not None])
        ctx.has = (bias is not None, zero_mask is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        q, k, v, y, lse, *rest = ctx.saved_tensors
        rest = list(rest)
        bias = rest.pop(0) if ctx.has[0] else None
        zero_mask = rest.pop(0) if ctx.has[1] else None
        B, Nq, H, dk = q.shape
        Nk = k.shape[1]
        dy_c = dy.detach().to(q.dtype).contiguous()
        dq = torch.empty(B, Nq, H, dk, dtype=q.dtype, device=q.device)
        dk_ = torch.empty(B, Nk, H, dk, dtype=q.dtype, device=q.device)
        dv = torch.empty(B, Nk, H, dk, dtype=q.dtype, device=q.device)
        with torch.cuda.device(q.device):
            p = _lib.new_params(_lib.SdpaParams)
            _fill_sdpa(p, q, k, v, ctx.cfg["causal"], bias, zero_mask, ctx.cfg)
            p.y, p.lse, p.dy = _ptr(y), _ptr(lse), _ptr(dy_c)
            p.dq, p.dk_, p.dv = _ptr(dq), _ptr(dk_), _ptr(dv)
            nbytes = lib.mop_sdpa_workspace_bytes(C.byref(p), 1)
            ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=q.device)
            p.workspace, p.workspace_bytes = _ptr(ws), nbytes
            with _Timed("sdpa_bwd"):
                _lib.check(lib.mop_sdpa_bwd(C.byref(p), _stream()), "mop_sdpa_bwd")
        last_impl["sdpa_bwd"] = _lib.IMPL_NAMES.get(p.impl_used, "?")
        abi_calls["sdpa_bwd"] += 1
        return None, dq, dk_, dv, None, None


def sdpa(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, causal: bool = False,
         bias: Optional[torch.Tensor] = None, zero_mask: Optional[torch.Tensor] = None,
         dropout_p: float = 0.0, impl: Optional[str] = None) -> torch.Tensor:
    """dropout_p(softmax(q k^T / sqrt(dk) [zero_mask==0 -> -inf] [causal] [+ bias])) v.

    q ``[B,Nq,H,dk]``, k/v ``[B,Nk,H,dk]`` (any batch/token/head strides, last
    dim contiguous - e.g. slices of a fused qkv projection).  ``bias`` and
    ``zero_mask`` broadcast to ``[B,H,Nq,Nk]`` and are not differentiated.
    Returns ``[B,Nq,H,dk]``.
    """
    B, Nq, H, dk = q.shape
    Nk = k.shape[1]
    if bias is not None:
        if bias.requires_grad:
            raise NotImplementedError("gradient w.r.t. the additive attention bias is not provided")
        bias = _as4(bias.detach().float(), B, H, Nq, Nk, "bias")
    if zero_mask is not None:
        zero_mask = _as4(zero_mask.detach().float(), B, H, Nq, Nk, "zero_mask")
    drop = _draw_dropout(float(dropout_p))
    if drop[0] > 0.0:
        last_dropout["sdpa"] = drop
    return _Sdpa.apply(dict(causal=bool(causal), impl=impl, dropout=drop), q, k, v, bias, zero_mask)


# ----------------------------------------------------------------------------
# Quartet
# ----------------------------------------------------------------------------
def _fill_quartet(p, q, k, v, q2, k2, mixture, gamma, add_mask, cfg):
    B, T, H, dk = q.shape
    p.dtype = _dtype_code(q)
    p.impl = _IMPL[cfg.get("impl")]
    p.B, p.H, p.T, p.dk = B, H, T, dk
    p.use_quartet = int(q2 is not None)
    p.scale = 1.0 / math.sqrt(dk)
    p.eps = cfg["eps"]
    p.q, p.k, p.v = _ptr(q), _ptr(k), _ptr(v)
    if q2 is not None:
        p.q2, p.k2, p.mixture, p.quartet_scale = _ptr(q2), _ptr(k2), _ptr(mixture), _ptr(gamma)
    if add_mask is not None:
        p.add_mask = _ptr(add_mask)
        p.am_sb, p.am_sh, p.am_sq, p.am_sk = _bstrides(add_mask)
    p.dropout_p, p.dropout_seed, p.dropout_offset = cfg.get("dropout", (0.0, 0, 0))


quartet_reuse_prep = True   # keep the forward workspace for the backward (ABI v9 `fwd_workspace`); False: the backward redoes the key preparation


class _Quartet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, q, k, v, q2, k2, mixture, gamma, add_mask):
        lib = _lib.load()
        _need_cuda(q, "q")
        quart = q2 is not None
        q, k, v = (t.detach().contiguous() for t in (q, k, v))
        if quart:
            q2, k2 = q2.detach().contiguous(), k2.detach().contiguous()
            mix32, gam32 = _f32c(mixture).reshape(1), _f32c(gamma).reshape(1)
        else:
            mix32 = gam32 = None
        B, T, H, dk = q.shape
        y = torch.empty_like(q)
        stats = torch.empty(B, H, T, 3, dtype=torch.float32, device=q.device)
        # bf16 storage: the backward forms delta = dO . y from an fp32 copy of y (the scalar gradients cancel heavily)
        y32 = torch.empty(B, T, H, dk, dtype=torch.float32, device=q.device) if (q.dtype != torch.float32 and quart and any(ctx.needs_input_grad)) else None
        with torch.cuda.device(q.device):
            p = _lib.new_params(_lib.QuartetParams)
            _fill_quartet(p, q, k, v, q2, k2, mix32, gam32, add_mask, cfg)
            p.y, p.stats, p.y_f32 = _ptr(y), _ptr(stats), _ptr(y32)
            nbytes = lib.mop_quartet_workspace_bytes(C.byref(p), 0)
            ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=q.device)
            p.workspace, p.workspace_bytes = _ptr(ws), nbytes
            with _Timed("quartet_fwd"):
                _lib.check(lib.mop_quartet_fwd(C.byref(p), _stream()), "mop_quartet_fwd")
        last_impl["quartet_fwd"] = _lib.IMPL_NAMES.get(p.impl_used, "?")
        abi_calls["quartet_fwd"] += 1
        ctx.cfg, ctx.quart, ctx.has_mask = cfg, quart, add_mask is not None
        ctx.scalar_dtypes = (None, None) if not quart else (mixture.dtype, gamma.dtype)
        saved = [q, k, v, y, stats] + ([q2, k2, mix32, gam32] if quart else []) + ([add_mask] if add_mask is not None else [])
        ctx.has_y32 = y32 is not None
        if y32 is not None:
            saved.append(y32)
        # the tensor-core backward reads the forward's key preparation (centred keys, Gram tiles) from the forward workspace
        ctx.has_ws = quartet_reuse_prep and p.impl_used == _lib.MOP_IMPL_TCGEN05 and any(ctx.needs_input_grad)
        if ctx.has_ws:
            saved.append(ws)
        ctx.save_for_backward(*saved)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        saved = list(ctx.saved_tensors)
        fws = saved.pop() if ctx.has_ws else None
        y32 = saved.pop() if ctx.has_y32 else None
        q, k, v, y, stats = saved[:5]
        rest = saved[5:]
        q2 = k2 = mix32 = gam32 = add_mask = None
        if ctx.quart:
            q2, k2, mix32, gam32 = rest[:4]
            rest = rest[4:]
        if ctx.has_mask:
            add_mask = rest[0]
        B, T, H, dk = q.shape
        dev = q.device
        dy_c = dy.detach().to(q.dtype).contiguous()
        dq, dk_, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
        dq2 = dk2 = dsc = None
        if ctx.quart:
            dq2, dk2 = torch.empty_like(q), torch.empty_like(q)
            dsc = torch.empty(B * H, 2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            p = _lib.new_params(_lib.QuartetParams)
            _fill_quartet(p, q, k, v, q2, k2, mix32, gam32, add_mask, ctx.cfg)
            p.y, p.stats, p.dy, p.y_f32 = _ptr(y), _ptr(stats), _ptr(dy_c), _ptr(y32)
            p.dq, p.dk_, p.dv, p.dq2, p.dk2, p.dscalar_part = _ptr(dq), _ptr(dk_), _ptr(dv), _ptr(dq2), _ptr(dk2), _ptr(dsc)
            nbytes = lib.mop_quartet_workspace_bytes(C.byref(p), 1)
            ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
            p.workspace, p.workspace_bytes = _ptr(ws), nbytes
            if fws is not None:
                p.fwd_workspace, p.fwd_workspace_bytes = _ptr(fws), fws.numel()
            with _Timed("quartet_bwd"):
                _lib.check(lib.mop_quartet_bwd(C.byref(p), _stream()), "mop_quartet_bwd")
        last_impl["quartet_bwd"] = _lib.IMPL_NAMES.get(p.impl_used, "?")
        abi_calls["quartet_bwd"] += 1
        dmix = dgam = None
        if ctx.quart:
            if keep_partials:
                last_partials["quartet_dscalar"] = dsc.detach().clone()
            tot = dsc.sum(0)
            dmix = tot[0].reshape(1).to(ctx.scalar_dtypes[0])
            dgam = tot[1].reshape(1).to(ctx.scalar_dtypes[1])
        return None, dq, dk_, dv, dq2, dk2, dmix, dgam, None


def quartet_attention(q, k, v, q2=None, k2=None, mixture=None, quartet_scale=None, *, eps: float = 1e-5,
                      add_mask: Optional[torch.Tensor] = None, dropout_p: float = 0.0, impl: Optional[str] = None) -> torch.Tensor:
    """GPT Quartet causal attention core (reference quartet_attn_patch.py:88-121).

    q,k,v,(q2,k2) ``[B,T,H,dk]`` (the Linear outputs viewed as heads).  ``q2 is None`` selects the
    ``use_quartet=False`` branch (single z-scored map).  ``add_mask`` broadcasts to ``[B,H,T,T]`` and is
    added after the causal fill.  Returns ``[B,T,H,dk]``.
    """
    B, T, H, dk = q.shape
    if add_mask is not None:
        if add_mask.requires_grad:
            raise NotImplementedError("gradient w.r.t. the additive attention mask is not provided")
        add_mask = _as4(add_mask.detach().float(), B, H, T, T, "attention_mask")
    drop = _draw_dropout(float(dropout_p))
    if drop[0] > 0.0:
        last_dropout["quartet"] = drop
    return _Quartet.apply(dict(eps=float(eps), impl=impl, dropout=drop), q, k, v, q2, k2, mixture, quartet_scale, add_mask)


# ----------------------------------------------------------------------------
# Fused residual-add + DropPath scale + LayerNorm
# ----------------------------------------------------------------------------
def _ln_params(x2, r2, scale, gamma, beta, eps, rows_per_sample, y_dtype):
    p = _lib.new_params(_lib.LnParams)
    p.rows, p.D = x2.shape
    p.r_dtype = _DT[r2.dtype] if r2 is not None else _lib.MOP_F32
    p.y_dtype = _DT[y_dtype]
    p.rows_per_sample = int(rows_per_sample)
    p.eps = float(eps)
    p.x, p.r, p.scale, p.gamma, p.beta = _ptr(x2), _ptr(r2), _ptr(scale), _ptr(gamma), _ptr(beta)
    return p


class _AddLayerNorm(torch.autograd.Function):
    """(x, r, scale) -> (x_new = x + scale[b] r, y = LayerNorm(x_new)); r may be None (then x_new is not produced)."""

    @staticmethod
    def forward(ctx, x, r, scale, gamma, beta, eps, rows_per_sample, y_dtype):
        lib = _lib.load()
        _need_cuda(x, "x")
        D = x.shape[-1]
        x2 = x.detach().reshape(-1, D)
        if x2.dtype != torch.float32 or not x2.is_contiguous():
            x2 = x2.float().contiguous()
        r2 = None
        if r is not None:
            r2 = r.detach().reshape(-1, D)
            if r2.dtype not in _DT or not r2.is_contiguous():
                r2 = r2.contiguous() if r2.dtype in _DT else r2.float().contiguous()
        g32, b32 = _f32c(gamma), _f32c(beta)
        sc32 = None if scale is None else scale.detach().float().contiguous()
        rows = x2.shape[0]
        dev = x2.device
        y = torch.empty(rows, D, dtype=y_dtype, device=dev)
        x_new = torch.empty(rows, D, dtype=torch.float32, device=dev) if r2 is not None else None
        mean = torch.empty(rows, dtype=torch.float32, device=dev)
        rstd = torch.empty(rows, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            p = _ln_params(x2, r2, sc32, g32, b32, eps, rows_per_sample, y_dtype)
            p.x_new, p.y, p.mean, p.rstd = _ptr(x_new), _ptr(y), _ptr(mean), _ptr(rstd)
            _lib.check(lib.mop_ln_fwd(C.byref(p), _stream()), "mop_ln_fwd")
        abi_calls["ln_fwd"] = abi_calls.get("ln_fwd", 0) + 1
        ctx.save_for_backward(x_new if x_new is not None else x2, mean, rstd, g32, sc32 if sc32 is not None else mean)
        ctx.meta = (eps, rows_per_sample, y_dtype, r2 is not None, None if r2 is None else r2.dtype, sc32 is not None,
                    x.shape, None if r is None else r.dtype, gamma.dtype, beta.dtype)
        shp = x.shape
        if x_new is None:
            ctx.mark_non_differentiable()
            return None, y.view(shp)
        return x_new.view(shp), y.view(shp)

    @staticmethod
    def backward(ctx, dx_new, dy):
        lib = _lib.load()
        xs, mean, rstd, g32, sc = ctx.saved_tensors
        eps, rps, y_dtype, has_r, r_dtype, has_scale, shp, r_in_dtype, g_dt, b_dt = ctx.meta
        rows, D = xs.shape
        dev = xs.device
        if dy is None:
            dy = torch.zeros(rows, D, dtype=y_dtype, device=dev)
        dy2 = dy.detach().reshape(rows, D).to(y_dtype).contiguous()
        dxn = None
        if has_r and dx_new is not None:
            dxn = dx_new.detach().reshape(rows, D).float().contiguous()
        dx = torch.empty(rows, D, dtype=torch.float32, device=dev)
        dr = torch.empty(rows, D, dtype=r_dtype, device=dev) if has_r else None
        with torch.cuda.device(dev):
            nparts = lib.mop_ln_partial_rows_d(rows, D)
            parts = torch.empty(2, nparts, D, dtype=torch.float32, device=dev)
            p = _ln_params(xs, dr, sc if has_scale else None, g32, g32, eps, rps, y_dtype)
            p.r = _ptr(dr)          # only its presence / dtype matters to the backward
            p.x_new = _ptr(xs) if has_r else None
            p.mean, p.rstd = _ptr(mean), _ptr(rstd)
            p.dy, p.dx_new, p.dx, p.dr = _ptr(dy2), _ptr(dxn), _ptr(dx), _ptr(dr)
            p.dgamma_part, p.dbeta_part, p.nparts = _ptr(parts[0]), _ptr(parts[1]), nparts
            _lib.check(lib.mop_ln_bwd(C.byref(p), _stream()), "mop_ln_bwd")
        abi_calls["ln_bwd"] = abi_calls.get("ln_bwd", 0) + 1
        sums = parts.sum(1)
        d_r = None if not has_r else dr.view(shp).to(r_in_dtype)
        return dx.view(shp), d_r, None, sums[0].to(g_dt), sums[1].to(b_dt), None, None, None


def add_layer_norm(x: torch.Tensor, r: Optional[torch.Tensor], scale: Optional[torch.Tensor], gamma: torch.Tensor,
                   beta: torch.Tensor, eps: float = 1e-5, *, out_dtype: Optional[torch.dtype] = None):
    """``x_new = x + scale[b] * r`` (per-sample DropPath factor, optional) and ``y = LayerNorm(x_new)`` in one pass.

    x ``[B, N, D]`` fp32 residual stream; r ``[B, N, D]`` branch output (bf16 / fp32) or None; scale ``[B]`` or None.
    Returns ``(x_new, y)`` (``x_new is x`` when r is None).  ``out_dtype`` defaults to the CUDA autocast dtype when autocast
    is on (what the Linear that consumes y would cast to), else fp32.
    """
    if out_dtype is None:
        out_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else torch.float32
    rows_per_sample = x.shape[-2] if x.dim() >= 2 else 1
    x_new, y = _AddLayerNorm.apply(x, r, scale, gamma, beta, float(eps), rows_per_sample, out_dtype)
    return (x if r is None else x_new), y


# ----------------------------------------------------------------------------
# ViT-MoP token gate
# ----------------------------------------------------------------------------
def token_gate_supported(tok: torch.Tensor, grid, n_views: int, n_kernels: int, hid: int) -> bool:
    """Shapes the fused token-gate kernel takes (others run the reference's own PyTorch composition)."""
    B, T, D = tok.shape
    return (tok.is_cuda and tok.dtype in (torch.float32, torch.bfloat16) and grid[0] * grid[1] == T and T <= 256 and D % 8 == 0
            and D <= 2048 and 1 <= n_views <= 8 and 1 <= n_kernels <= 8 and 1 <= hid <= 16)


def _fill_token_gate(p, tok, grid, w):
    B, T, D = tok.shape
    p.dtype = _lib.MOP_BF16 if tok.dtype == torch.bfloat16 else _lib.MOP_F32
    p.B, p.T, p.D, p.Gh, p.Gw = B, T, D, int(grid[0]), int(grid[1])
    p.V, p.K, p.hid = w[0].shape[0], w[2].shape[0], w[3].shape[0]
    p.x = _ptr(tok)
    p.views_w, p.k3_w, p.k1_w, p.f1_w, p.f2_w, p.f2_b, p.a_pos, p.a_neg = (_ptr(t) for t in w)


class _TokenGate(torch.autograd.Function):
    """out = tok * gate(tok).  Inputs after `grid`: views_w [V,D], k3_w [16,V,3,3], k1_w [K,16], f1_w [hid,V+K], f2_w [2,hid],
    f2_b [2], a_pos [1], a_neg [1]."""

    @staticmethod
    def forward(ctx, tok, grid, max_ctas, *weights):
        lib = _lib.load()
        _need_cuda(tok, "tok")
        x = tok.detach().contiguous()
        w = [_f32c(t).reshape(s) for t, s in zip(weights, (
            (weights[0].shape[0], -1), (16, weights[0].shape[0], 3, 3), (weights[2].shape[0], 16), (weights[3].shape[0], -1),
            (2, -1), (2,), (1,), (1,)))]
        B, T, D = x.shape
        V = w[0].shape[0]
        out = torch.empty_like(x)
        views = torch.empty(B, T, V, dtype=torch.float32, device=x.device)
        gate = torch.empty(B, T, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            p = _lib.new_params(_lib.TokenGateParams)
            _fill_token_gate(p, x, grid, w)
            p.out, p.views, p.gate = _ptr(out), _ptr(views), _ptr(gate)
            _lib.check(lib.mop_token_gate_fwd(C.byref(p), _stream()), "mop_token_gate_fwd")
        abi_calls["token_gate_fwd"] = abi_calls.get("token_gate_fwd", 0) + 1
        ctx.save_for_backward(x, views, gate, *w)
        ctx.grid, ctx.max_ctas = (int(grid[0]), int(grid[1])), max_ctas
        ctx.wmeta = [(t.shape, t.dtype) for t in weights]
        ctx.tok_dtype = tok.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        x, views, gate, *w = ctx.saved_tensors
        B, T, D = x.shape
        V = w[0].shape[0]
        dy = dout.detach().to(x.dtype).contiguous()
        dx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            p = _lib.new_params(_lib.TokenGateParams)
            _fill_token_gate(p, x, ctx.grid, w)
            p.views, p.gate, p.dout, p.dx = _ptr(views), _ptr(gate), _ptr(dy), _ptr(dx)
            nparts = lib.mop_token_gate_partial_rows(B)
            if ctx.max_ctas:
                nparts = max(1, min(nparts, int(ctx.max_ctas)))
            nnet = lib.mop_token_gate_net_params(C.byref(p))
            groups = lib.mop_token_gate_wv_groups(D)
            dwv = torch.empty(nparts * groups, V, D, dtype=torch.float32, device=x.device)
            dnet = torch.empty(nparts, nnet + 2, dtype=torch.float32, device=x.device)
            p.nparts, p.dwv_part, p.dnet_part = nparts, _ptr(dwv), _ptr(dnet)
            _lib.check(lib.mop_token_gate_bwd(C.byref(p), _stream()), "mop_token_gate_bwd")
        abi_calls["token_gate_bwd"] = abi_calls.get("token_gate_bwd", 0) + 1
        tot = dnet.sum(0)
        sizes = [w[1].numel(), w[2].numel(), w[3].numel(), w[4].numel(), 2, 1, 1]
        pieces = torch.split(tot, sizes)
        grads = [dwv.sum(0)] + list(pieces)
        grads = [g.reshape(shape).to(dt) for g, (shape, dt) in zip(grads, ctx.wmeta)]
        return (dx.to(ctx.tok_dtype), None, None, *grads)


def token_gate(tok: torch.Tensor, grid, views_w, k3_w, k1_w, f1_w, f2_w, f2_b, a_pos, a_neg, *, _max_ctas: int = 0) -> torch.Tensor:
    """``tok * gate`` of the ViT-MoP post-encoder gate (reference vit_mop.py:95-114, components.py:255-303), one fused kernel per
    direction.  ``tok [B, T, D]`` (fp32 or bf16), ``grid = (Gh, Gw)``; the weights are the module parameters
    (``ViewsLinear.proj.weight``, ``Kernels3.k[0].weight``, ``Kernels3.k[2].weight``, ``FuseExcInh.fuse[0].weight``,
    ``FuseExcInh.fuse[2].weight / .bias``) and ``a_pos / a_neg = softplus(alpha)`` as one-element tensors."""
    return _TokenGate.apply(tok, tuple(grid), _max_ctas, views_w, k3_w, k1_w, f1_w, f2_w, f2_b, a_pos.reshape(1), a_neg.reshape(1))


class _TokenGate1d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, views_w, w_eff, max_ctas):
        lib = _lib.load()
        _need_cuda(x, "x")
        xc = x.detach().contiguous()
        wv, we = _f32c(views_w), _f32c(w_eff)
        B, T, D = xc.shape
        V = wv.shape[0]
        out = torch.empty_like(xc)
        views = torch.empty(B, T, V, dtype=torch.float32, device=xc.device)
        gate = torch.empty(B, T, dtype=torch.float32, device=xc.device)
        with torch.cuda.device(xc.device):
            p = _lib.new_params(_lib.TokenGate1dParams)
            p.dtype = _lib.MOP_BF16 if xc.dtype == torch.bfloat16 else _lib.MOP_F32
            p.B, p.T, p.D, p.V = B, T, D, V
            p.x, p.views_w, p.w_eff, p.out, p.views, p.gate = _ptr(xc), _ptr(wv), _ptr(we), _ptr(out), _ptr(views), _ptr(gate)
            _lib.check(lib.mop_token_gate1d_fwd(C.byref(p), _stream()), "mop_token_gate1d_fwd")
        abi_calls["token_gate1d_fwd"] = abi_calls.get("token_gate1d_fwd", 0) + 1
        ctx.save_for_backward(xc, views, gate, wv, we)
        ctx.meta = (x.dtype, views_w.dtype, w_eff.dtype, views_w.shape, max_ctas)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        xc, views, gate, wv, we = ctx.saved_tensors
        x_dt, wv_dt, we_dt, wv_shape, max_ctas = ctx.meta
        B, T, D = xc.shape
        V = wv.shape[0]
        dy = dout.detach().to(xc.dtype).contiguous()
        dx = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            p = _lib.new_params(_lib.TokenGate1dParams)
            p.dtype = _lib.MOP_BF16 if xc.dtype == torch.bfloat16 else _lib.MOP_F32
            p.B, p.T, p.D, p.V = B, T, D, V
            nparts = lib.mop_token_gate1d_partial_rows(B, T)
            if max_ctas:
                nparts = max(1, min(nparts, int(max_ctas)))
            groups = lib.mop_token_gate_wv_groups(D)
            dwv = torch.empty(nparts * groups, V, D, dtype=torch.float32, device=xc.device)
            dwe = torch.empty(nparts, 3 * V, dtype=torch.float32, device=xc.device)
            p.nparts = nparts
            p.x, p.views_w, p.w_eff, p.views, p.gate = _ptr(xc), _ptr(wv), _ptr(we), _ptr(views), _ptr(gate)
            p.dout, p.dx, p.dwv_part, p.dweff_part = _ptr(dy), _ptr(dx), _ptr(dwv), _ptr(dwe)
            _lib.check(lib.mop_token_gate1d_bwd(C.byref(p), _stream()), "mop_token_gate1d_bwd")
        abi_calls["token_gate1d_bwd"] = abi_calls.get("token_gate1d_bwd", 0) + 1
        return dx.to(x_dt), dwv.sum(0).reshape(wv_shape).to(wv_dt), dwe.sum(0).reshape(3, V).to(we_dt), None


def token_gate_1d_supported(x: torch.Tensor, n_views: int, kernel_size: int) -> bool:
    return (x.is_cuda and x.dim() == 3 and x.dtype in (torch.float32, torch.bfloat16) and x.shape[-1] % 8 == 0 and x.shape[-1] <= 2048
            and 1 <= n_views <= 8 and kernel_size == 3)


def token_gate_1d(x: torch.Tensor, views_w, kernels_w, fuse_w, alpha, *, _max_ctas: int = 0) -> torch.Tensor:
    """``x * gate`` of ``MoPBlock.apply_mop`` (reference gpt_mop.py:102-123), one fused kernel per direction.

    x ``[B, T, D]``; ``views_w [V, D]`` (``ViewsLinear1D.proj.weight``), ``kernels_w [K, V, 3]`` (``Kernels1D.conv.weight``),
    ``fuse_w [2, V+K, 1]`` (``FuseExcInh1D.conv.weight``), ``alpha [2]``.  The module is linear and bias-free, so the last three fold
    into one 3-tap filter of the views (differentiable PyTorch, a few hundred flops)."""
    V = views_w.shape[0]
    fw = fuse_w.reshape(2, -1).float()
    we = torch.einsum("ck,kvt->ctv", fw[:, V:], kernels_w.float())      # through Kernels1D: [2, 3, V]
    centre = torch.zeros_like(we)
    centre[:, 1] = fw[:, :V]                                             # the views themselves (1x1 path)
    we = we + centre
    w_eff = alpha[0].float() * we[0] - alpha[1].float() * we[1]          # [3, V]
    return _TokenGate1d.apply(x, views_w, w_eff, _max_ctas)


# ----------------------------------------------------------------------------
# Whisper-MoP 2D gate
# ----------------------------------------------------------------------------
class _MoP2DGate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mel, He):
        lib = _lib.load()
        _need_cuda(mel, "mel")
        if mel.requires_grad:
            raise NotImplementedError("MoP2D gate: gradient w.r.t. the mel spectrogram is not provided (it is input data)")
        B, T, Fb = mel.shape
        ks = He.shape[-1]
        mel32 = mel.detach().float().contiguous()
        he32 = He.detach().float().contiguous()
        R = torch.empty(B, T, ks, dtype=torch.float32, device=mel.device)
        gate = torch.empty(B, T, dtype=torch.float32, device=mel.device)
        with torch.cuda.device(mel.device):
            _lib.check(lib.mop_mop2d_fwd(_ptr(mel32), _ptr(he32), _ptr(R), _ptr(gate), B, T, Fb, ks, _stream()), "mop_mop2d_fwd")
        abi_calls["mop2d_fwd"] = abi_calls.get("mop2d_fwd", 0) + 1
        ctx.save_for_backward(R)
        ctx.meta = (B, T, Fb, ks, He.dtype)
        return gate

    @staticmethod
    def backward(ctx, dgate):
        lib = _lib.load()
        (R,) = ctx.saved_tensors
        B, T, Fb, ks, dt = ctx.meta
        dg = dgate.detach().float().contiguous()
        with torch.cuda.device(R.device):
            nparts = lib.mop_mop2d_partial_rows()
            parts = torch.empty(nparts, ks * ks, dtype=torch.float32, device=R.device)
            _lib.check(lib.mop_mop2d_bwd(_ptr(R), _ptr(dg), _ptr(parts), nparts, B, T, Fb, ks, _stream()), "mop_mop2d_bwd")
        abi_calls["mop2d_bwd"] = abi_calls.get("mop2d_bwd", 0) + 1
        return None, parts.sum(0).reshape(ks, ks).to(dt)


def mop2d_gate(mel: torch.Tensor, views_w: torch.Tensor, kernels_w: torch.Tensor, fuse_w: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    """Per-time-step gate of ``MoP2D`` (reference whisper_mop.py:91-124) without the [B, V+K, T, F] intermediates.

    mel ``[B, T, F]`` (the ``mel2d[:, 0]`` map); ``views_w [V,1,1,1]``, ``kernels_w [K,V,k,k]``, ``fuse_w [2,V+K,1,1]``,
    ``alpha [2]`` are the module's parameters.  The module is linear, so the weights fold into one k x k filter
    (differentiable PyTorch, a few hundred flops); the kernel applies it fused with the mean over the mel bins.
    Returns ``gate [B, T]`` (fp32)."""
    V = views_w.shape[0]
    ks = kernels_w.shape[-1]
    wv = views_w.reshape(V).float()
    fw = fuse_w.reshape(2, -1).float()
    H = torch.einsum("ck,kvuw,v->cuw", fw[:, V:], kernels_w.float(), wv)       # through Kernels2D
    centre = torch.zeros_like(H)
    centre[:, ks // 2, ks // 2] = fw[:, :V] @ wv                                 # the views themselves (1x1 path)
    H = H + centre
    He = alpha[0].float() * H[0] - alpha[1].float() * H[1]
    return _MoP2DGate.apply(mel, He)
